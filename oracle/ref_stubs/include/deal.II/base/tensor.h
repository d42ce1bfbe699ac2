// Stand-in for dealii::Tensor<1,dim> (test infrastructure, see ../../../README.md).
#pragma once
#include <cmath>
#include <cstddef>
namespace dealii {
template <int rank, int dim>
class Tensor;
template <int dim>
class Tensor<1, dim> {
    double v_[dim];
public:
    Tensor() { for (int i = 0; i < dim; i++) v_[i] = 0.0; }
    double& operator[](std::size_t i) { return v_[i]; }
    const double& operator[](std::size_t i) const { return v_[i]; }
    Tensor& operator*=(double a) { for (int i = 0; i < dim; i++) v_[i] *= a; return *this; }
    double operator*(const Tensor& o) const { double s = 0; for (int i = 0; i < dim; i++) s += v_[i] * o.v_[i]; return s; }
    double norm() const { return std::sqrt((*this) * (*this)); }
};
}  // namespace dealii
