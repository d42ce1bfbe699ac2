// Stand-in for L/utilities/Math.h (test infrastructure, see ../../README.md): the helpers CollisionModel.h and the
// stencil classes call.
#pragma once
#include <functional>
#include <math.h>
#include "BasicNames.h"
namespace natrium {
namespace Math {
const double PI = 3.141592653589793238462;
inline double scalar_product(const numeric_vector& x, const numeric_vector& y) { return x * y; }
inline void scale_vector(double a, numeric_vector& x) { x *= a; }
inline numeric_vector scalar_vector(double a, const numeric_vector& x) { numeric_vector y(x); y *= a; return y; }
inline void add_vector(numeric_vector& x, const numeric_vector& y) { x += y; }
inline void subtract_vector(numeric_vector& x, const numeric_vector& y) { x -= y; }
inline double euclidean_norm(numeric_vector& x) { return x.l2_norm(); }
}  // namespace Math
}  // namespace natrium
