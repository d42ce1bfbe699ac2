// Stand-in for L/utilities/BasicNames.h (test infrastructure, see ../../README.md): the names the reference's
// collision and stencil sources use, without deal.II / Trilinos / Boost / MPI behind them.
#pragma once
#include <cassert>
#include <cmath>
#include <cstddef>
#include <fstream>
#include <iostream>
#include <map>
#include <math.h>
#include <memory>
#include <string>
#include <vector>

#include "deal.II/base/index_set.h"
#include "deal.II/base/tensor.h"
#include "Logging.h"

namespace boost {
using std::shared_ptr;
using std::make_shared;
}

namespace natrium {

using std::vector;
using std::map;
using std::cout;
using std::cerr;
using std::endl;
using std::size_t;
using std::string;

// dealii::Vector<double>
class numeric_vector {
    std::vector<double> v_;
public:
    numeric_vector() = default;
    explicit numeric_vector(size_t n) : v_(n, 0.0) {}
    size_t size() const { return v_.size(); }
    void reinit(size_t n) { v_.assign(n, 0.0); }
    double& operator()(size_t i) { return v_[i]; }
    const double& operator()(size_t i) const { return v_[i]; }
    double& operator[](size_t i) { return v_[i]; }
    const double& operator[](size_t i) const { return v_[i]; }
    double operator*(const numeric_vector& o) const { double s = 0.0; for (size_t i = 0; i < v_.size(); i++) s += v_[i] * o.v_[i]; return s; }
    numeric_vector& operator*=(double a) { for (auto& x : v_) x *= a; return *this; }
    numeric_vector& operator+=(const numeric_vector& o) { for (size_t i = 0; i < v_.size(); i++) v_[i] += o.v_[i]; return *this; }
    numeric_vector& operator-=(const numeric_vector& o) { for (size_t i = 0; i < v_.size(); i++) v_[i] -= o.v_[i]; return *this; }
    double l2_norm() const { return std::sqrt((*this) * (*this)); }
};

// dealii::FullMatrix<double>
class numeric_matrix {
    size_t n_ = 0, m_ = 0;
    std::vector<double> a_;
public:
    numeric_matrix() = default;
    explicit numeric_matrix(size_t n) : n_(n), m_(n), a_(n * n, 0.0) {}
    numeric_matrix(size_t n, size_t m) : n_(n), m_(m), a_(n * m, 0.0) {}
    size_t n() const { return m_; }
    size_t m() const { return n_; }
    double& operator()(size_t i, size_t j) { return a_[i * m_ + j]; }
    const double& operator()(size_t i, size_t j) const { return a_[i * m_ + j]; }
    // in-place inverse with partial pivoting (the stencils only keep the result for getInverseMomentBasis)
    void gauss_jordan()
    {
        const size_t n = n_;
        std::vector<double> inv(n * n, 0.0);
        for (size_t i = 0; i < n; i++) inv[i * n + i] = 1.0;
        for (size_t c = 0; c < n; c++) {
            size_t p = c;
            for (size_t r = c + 1; r < n; r++) if (std::fabs(a_[r * n + c]) > std::fabs(a_[p * n + c])) p = r;
            if (a_[p * n + c] == 0.0) return;     // singular: leave as is (not used on the collision path)
            for (size_t j = 0; j < n; j++) { std::swap(a_[c * n + j], a_[p * n + j]); std::swap(inv[c * n + j], inv[p * n + j]); }
            const double d = 1.0 / a_[c * n + c];
            for (size_t j = 0; j < n; j++) { a_[c * n + j] *= d; inv[c * n + j] *= d; }
            for (size_t r = 0; r < n; r++) {
                if (r == c) continue;
                const double f = a_[r * n + c];
                if (f == 0.0) continue;
                for (size_t j = 0; j < n; j++) { a_[r * n + j] -= f * a_[c * n + j]; inv[r * n + j] -= f * inv[c * n + j]; }
            }
        }
        a_.swap(inv);
    }
};

// dealii::TrilinosWrappers::MPI::Vector on one rank: a view of a caller-owned array.  trilinos_vector().ExtractView()
// returns the raw pointer and the local length like Epetra_MultiVector::ExtractView (CollisionOperator.h:38-48);
// operator()(i) is the global-index access the legacy models use (one rank: global = local).
class distributed_vector {
public:
    struct EpetraView {
        double* p = nullptr;
        int n = 0;
        int ExtractView(double** v, int* len) const { *v = p; *len = n; return 0; }
    };
private:
    EpetraView view_;
public:
    distributed_vector() = default;
    distributed_vector(double* p, size_t n) { view_.p = p; view_.n = (int)n; }
    const EpetraView& trilinos_vector() const { return view_; }
    EpetraView& trilinos_vector() { return view_; }
    double& operator()(size_t i) { return view_.p[i]; }
    const double& operator()(size_t i) const { return view_.p[i]; }
    size_t size() const { return (size_t)view_.n; }
};

}  // namespace natrium
