// ref_filter_driver.cpp -- C entry points over the REFERENCE's own ExponentialFilter (test infrastructure, part of oracle/).
//
// Compiled by oracle/Makefile.ref together with L/smoothing/ExponentialFilter.cpp where it lies under /root/reference (scratch
// tree of symbolic links + the stand-ins of oracle/ref_stubs_filter/) into oracle/_ref/libnatrium_ref_filter.so.  Nothing of the
// reference is restated here: this file builds what the solver hands the filter (the operator's Gauss-Lobatto quadrature and
// Lagrange element on Gauss-Lobatto nodes, L/advection/AdvectionOperator.cpp:34-39, and a cell -> DoF list) and calls
//   natrium::ExponentialFilter<dim>::ExponentialFilter   (makeProjectionMatrices, makeDegreeVectors)   ExponentialFilter.cpp:15-137
//   natrium::ExponentialFilter<dim>::applyFilter                                                       ExponentialFilter.cpp:139-199
// Used by oracle/ref.py -> tests/test_oracle_vs_ref.py (oracle/filter.py + orc_exponential_filter == reference code).
#include <cstdint>
#include <cstring>

#include "smoothing/ExponentialFilter.h"

using namespace natrium;

namespace {

// Gauss-Lobatto nodes / weights on [0, 1]: roots of P_n' by Newton from the Chebyshev-Gauss-Lobatto points
void gll(int npts, std::vector<double>& x, std::vector<double>& w)
{
    const int n = npts - 1;
    x.assign(npts, 0.0); w.assign(npts, 0.0);
    if (n == 0) { x[0] = 0.5; w[0] = 1.0; return; }
    for (int i = 0; i <= n; i++) {
        double t = -std::cos(M_PI * i / n);
        for (int it = 0; it < 100; it++) {
            double p0 = 1.0, p1 = t;
            for (int k = 2; k <= n; k++) { const double p2 = ((2.0 * k - 1.0) * t * p1 - (k - 1.0) * p0) / k; p0 = p1; p1 = p2; }
            // q(t) = (1 - t^2) P_n'(t) = n (P_{n-1} - t P_n); q'(t) = -n (n + 1) P_n
            const double q = n * (p0 - t * p1), dq = -n * (n + 1.0) * p1;
            if (i == 0 || i == n) break;
            const double dt = q / dq;
            t -= dt;
            if (std::fabs(dt) < 1e-16) break;
        }
        double p0 = 1.0, p1 = t;
        for (int k = 2; k <= n; k++) { const double p2 = ((2.0 * k - 1.0) * t * p1 - (k - 1.0) * p0) / k; p0 = p1; p1 = p2; }
        if (n == 1) p1 = t;
        x[i] = 0.5 * (t + 1.0);
        w[i] = 1.0 / (n * (n + 1.0) * p1 * p1);        // 2 / (n (n+1) P_n^2) on [-1, 1], halved for [0, 1]
    }
}

template <int dim>
dealii::Quadrature<dim> tensor_gll(const std::vector<double>& x, const std::vector<double>& w)
{
    const int n1 = (int)x.size();
    int n = 1;
    for (int d = 0; d < dim; d++) n *= n1;
    std::vector<dealii::Point<dim>> pts((size_t)n);
    std::vector<double> ws((size_t)n, 1.0);
    for (int q = 0; q < n; q++) {
        int r = q;
        for (int d = 0; d < dim; d++) { pts[(size_t)q](d) = x[(size_t)(r % n1)]; ws[(size_t)q] *= w[(size_t)(r % n1)]; r /= n1; }   // first coordinate fastest
    }
    return dealii::Quadrature<dim>(pts, ws);
}

template <int dim>
int run(int p, double alpha, double s, int Nc, int by_sum, double* to, double* from, int64_t n_cells, const int32_t* cell_dofs, double* v)
{
    std::vector<double> x, w;
    gll(p + 1, x, w);
    const dealii::Quadrature<dim> quad = tensor_gll<dim>(x, w);
    const dealii::FiniteElement<dim> fe(x);
    ExponentialFilter<dim> filter(alpha, s, (size_t)Nc, by_sum != 0, quad, fe);
    const size_t n = fe.dofs_per_cell;
    if (to && from)
        for (size_t i = 0; i < n; i++)
            for (size_t j = 0; j < n; j++) {
                to[i * n + j] = filter.getProjectToLegendre()(i, j);
                from[i * n + j] = filter.getProjectFromLegendre()(i, j);
            }
    if (n_cells > 0 && cell_dofs && v) {
        dealii::DoFHandler<dim> dh;
        dh.cell_dofs = cell_dofs;
        dh.n_cells = (size_t)n_cells;
        dh.dofs_per_cell = n;
        distributed_vector vec(v);
        filter.applyFilter(dh, vec);
    }
    return 0;
}

}  // namespace

// to / from: [n][n] row-major or null; cell_dofs [n_cells][n] / v: the vector filtered in place, or null / 0 cells
extern "C" int ref_exponential_filter(int dim, int p, double alpha, double s, int Nc, int by_sum, double* to, double* from,
                                      int64_t n_cells, const int32_t* cell_dofs, double* v)
{
    if (dim == 1) return run<1>(p, alpha, s, Nc, by_sum, to, from, n_cells, cell_dofs, v);
    if (dim == 2) return run<2>(p, alpha, s, Nc, by_sum, to, from, n_cells, cell_dofs, v);
    if (dim == 3) return run<3>(p, alpha, s, Nc, by_sum, to, from, n_cells, cell_dofs, v);
    return -1;
}
