// Stand-in for L/utilities/BasicNames.h as far as L/smoothing/ExponentialFilter.{h,cpp} needs it (test infrastructure).
#pragma once
#include <cassert>
#include <cmath>
#include <cstddef>
#include <math.h>
#include <vector>

#include "deal.II/filter_stubs.h"

namespace natrium {
using std::vector;
using std::size_t;
typedef dealii::Vector<double> numeric_vector;
typedef dealii::FullMatrix<double> numeric_matrix;
// dealii::TrilinosWrappers::MPI::Vector on one rank: operator()(global index) on a caller-owned array
class distributed_vector {
    double* p_;
public:
    explicit distributed_vector(double* p) : p_(p) {}
    double& operator()(size_t i) { return p_[i]; }
    const double& operator()(size_t i) const { return p_[i]; }
};
}  // namespace natrium
