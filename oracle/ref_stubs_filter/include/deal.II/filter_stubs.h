// Stand-ins for the deal.II classes L/smoothing/ExponentialFilter.{h,cpp} uses (test infrastructure, part of oracle/; see
// ../../README.md): just the members that file touches, with deal.II's names and call signatures, so that the reference's own
// ExponentialFilter.cpp compiles where it lies and its cell loop, quadrature sums, degree vectors and damping formula can be
// run here.  What deal.II itself would supply -- the Lagrange element on Gauss-Lobatto nodes, the Gauss-Lobatto quadrature,
// the Legendre basis (orthonormal on [0,1]) and dense linear algebra -- is written out in the plainest form.
#pragma once
#include <cassert>
#include <cmath>
#include <cstddef>
#include <vector>

namespace dealii {

namespace types { typedef unsigned int global_dof_index; }

template <int dim>
class Point {
    double c_[dim > 0 ? dim : 1];
public:
    Point() { for (int i = 0; i < dim; i++) c_[i] = 0.0; }
    double& operator()(unsigned i) { return c_[i]; }
    double operator()(unsigned i) const { return c_[i]; }
};

template <int dim>
class Quadrature {
protected:
    std::vector<Point<dim>> pts_;
    std::vector<double> w_;
public:
    Quadrature() = default;
    Quadrature(const std::vector<Point<dim>>& p, const std::vector<double>& w) : pts_(p), w_(w) {}
    const std::vector<Point<dim>>& get_points() const { return pts_; }
    const std::vector<double>& get_weights() const { return w_; }
    unsigned int size() const { return (unsigned int)pts_.size(); }
};

// FE_Q / FE_DGQArbitraryNodes(QGaussLobatto<1>(degree + 1)): tensor Lagrange polynomials on the given 1-d nodes,
// shape function i at node (i % n1, (i / n1) % n1, i / n1^2) (lexicographic, x fastest)
template <int dim>
class FiniteElement {
    std::vector<double> nodes_;
public:
    const unsigned int degree;
    const unsigned int dofs_per_cell;
    FiniteElement(const std::vector<double>& nodes_1d)
        : nodes_(nodes_1d), degree((unsigned int)nodes_1d.size() - 1), dofs_per_cell(ipow((unsigned int)nodes_1d.size(), dim)) {}
    static unsigned int ipow(unsigned int b, int e) { unsigned int r = 1; for (int i = 0; i < e; i++) r *= b; return r; }
    double lagrange(unsigned j, double x) const
    {
        double v = 1.0;
        for (unsigned m = 0; m < nodes_.size(); m++) if (m != j) v *= (x - nodes_[m]) / (nodes_[j] - nodes_[m]);
        return v;
    }
    double shape_value(unsigned int i, const Point<dim>& p) const
    {
        const unsigned n1 = degree + 1;
        double v = 1.0;
        for (int d = 0; d < dim; d++) { v *= lagrange(i % n1, p(d)); i /= n1; }
        return v;
    }
};

namespace Polynomials {
template <typename number>
class Polynomial {
    std::vector<number> c_;     // c_[k] x^k
public:
    Polynomial() = default;
    explicit Polynomial(const std::vector<number>& c) : c_(c) {}
    number value(number x) const { number v = 0; for (size_t k = c_.size(); k-- > 0;) v = v * x + c_[k]; return v; }
    unsigned int degree() const { return (unsigned int)c_.size() - 1; }
};
// Legendre polynomials on [0, 1], orthonormal: L_k(x) = sqrt(2k + 1) P_k(2x - 1); coefficients from the three-term recurrence
class Legendre {
public:
    static std::vector<Polynomial<double>> generate_complete_basis(unsigned int degree)
    {
        std::vector<std::vector<double>> P(degree + 1);
        P[0] = {1.0};
        if (degree >= 1) P[1] = {-1.0, 2.0};                       // t = 2x - 1
        for (unsigned k = 2; k <= degree; k++) {                  // k P_k = (2k-1) t P_{k-1} - (k-1) P_{k-2}
            std::vector<double> a(k + 1, 0.0);
            for (size_t i = 0; i < P[k - 1].size(); i++) { a[i] += -(2.0 * k - 1.0) * P[k - 1][i]; a[i + 1] += 2.0 * (2.0 * k - 1.0) * P[k - 1][i]; }
            for (size_t i = 0; i < P[k - 2].size(); i++) a[i] -= (k - 1.0) * P[k - 2][i];
            for (auto& v : a) v /= (double)k;
            P[k] = a;
        }
        std::vector<Polynomial<double>> out;
        for (unsigned k = 0; k <= degree; k++) {
            std::vector<double> c = P[k];
            for (auto& v : c) v *= std::sqrt(2.0 * k + 1.0);
            out.emplace_back(c);
        }
        return out;
    }
};
}  // namespace Polynomials

template <typename number>
class Vector {
    std::vector<number> v_;
public:
    Vector() = default;
    explicit Vector(size_t n) : v_(n, number(0)) {}
    size_t size() const { return v_.size(); }
    number& operator()(size_t i) { return v_[i]; }
    const number& operator()(size_t i) const { return v_[i]; }
};

template <typename number>
class FullMatrix {
    size_t n_ = 0;              // square matrices only
    std::vector<number> a_;
public:
    FullMatrix() = default;
    explicit FullMatrix(size_t n) : n_(n), a_(n * n, number(0)) {}
    size_t n() const { return n_; }
    size_t m() const { return n_; }
    number& operator()(size_t i, size_t j) { return a_[i * n_ + j]; }
    const number& operator()(size_t i, size_t j) const { return a_[i * n_ + j]; }
    // dst = A src, j ascending (FullMatrix::vmult)
    void vmult(Vector<number>& dst, const Vector<number>& src) const
    {
        for (size_t i = 0; i < n_; i++) {
            number s = 0;
            for (size_t j = 0; j < n_; j++) s += a_[i * n_ + j] * src(j);
            dst(i) = s;
        }
    }
    // C = A B
    void mmult(FullMatrix& C, const FullMatrix& B) const
    {
        for (size_t i = 0; i < n_; i++)
            for (size_t j = 0; j < n_; j++) {
                number s = 0;
                for (size_t k = 0; k < n_; k++) s += a_[i * n_ + k] * B(k, j);
                C(i, j) = s;
            }
    }
    // this = M^-1 (Gauss-Jordan with partial pivoting)
    void invert(const FullMatrix& M)
    {
        const size_t n = M.n();
        std::vector<number> a(M.a_), inv(n * n, number(0));
        for (size_t i = 0; i < n; i++) inv[i * n + i] = 1;
        for (size_t c = 0; c < n; c++) {
            size_t p = c;
            for (size_t r = c + 1; r < n; r++) if (std::fabs(a[r * n + c]) > std::fabs(a[p * n + c])) p = r;
            for (size_t j = 0; j < n; j++) { std::swap(a[c * n + j], a[p * n + j]); std::swap(inv[c * n + j], inv[p * n + j]); }
            const number d = number(1) / a[c * n + c];
            for (size_t j = 0; j < n; j++) { a[c * n + j] *= d; inv[c * n + j] *= d; }
            for (size_t r = 0; r < n; r++) {
                if (r == c) continue;
                const number f = a[r * n + c];
                if (f == number(0)) continue;
                for (size_t j = 0; j < n; j++) { a[r * n + j] -= f * a[c * n + j]; inv[r * n + j] -= f * inv[c * n + j]; }
            }
        }
        n_ = n;
        a_.swap(inv);
    }
};

// DoFHandler: the active-cell loop of applyFilter over a list of cells given as DoF index lists
template <int dim>
class DoFHandler {
public:
    struct Cell {
        const DoFHandler* h;
        size_t k;
        bool is_locally_owned() const { return true; }
        void get_dof_indices(std::vector<types::global_dof_index>& idx) const
        {
            for (size_t i = 0; i < idx.size(); i++) idx[i] = (types::global_dof_index)h->cell_dofs[k * h->dofs_per_cell + i];
        }
    };
    struct active_cell_iterator {
        Cell c;
        active_cell_iterator() : c{nullptr, 0} {}
        active_cell_iterator(const DoFHandler* h, size_t k) : c{h, k} {}
        const Cell* operator->() const { return &c; }
        active_cell_iterator& operator++() { ++c.k; return *this; }
        bool operator!=(const active_cell_iterator& o) const { return c.k != o.c.k; }
    };
    const int* cell_dofs = nullptr;
    size_t n_cells = 0, dofs_per_cell = 0;
    active_cell_iterator begin_active() const { return active_cell_iterator(this, 0); }
    active_cell_iterator end() const { return active_cell_iterator(this, n_cells); }
};

}  // namespace dealii
