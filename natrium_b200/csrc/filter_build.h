// filter_build.h -- host-side level schedule of the exponential filter (nb200_set_filter).
//
// ExponentialFilter<dim>::applyFilter (L/smoothing/ExponentialFilter.cpp:139-199) visits the cells one after the other and
// writes back into the vector it reads; cells of a continuous element share face DoFs, so a cell sees what earlier cells
// wrote and the result depends on the cell order.  The device keeps that order: a cell's level is one above the highest level
// of any EARLIER cell it shares a DoF with.  Cells of one level share nothing (one launch handles them in parallel), and for
// every pair of cells that share a DoF the earlier one sits in a lower level (levels run in order) -- so every read sees
// exactly the value the sequential loop would have seen, and the last writer of every DoF is the same.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace nbfilter {

struct Levels {
    std::vector<int64_t> level_off;   // [#levels + 1] into `cells`
    std::vector<int32_t> cells;       // cell ids sorted by level, original order inside a level
};

// dofs: [n_cells][n] local DoF indices in [0, n_loc).  Returns 0, or 1 + the cell that lists a DoF twice.
static inline int64_t build_levels(int64_t n_cells, int n, const int32_t* dofs, int64_t n_loc, Levels& L)
{
    std::vector<int32_t> last_level((size_t)std::max<int64_t>(1, n_loc), 0), level((size_t)n_cells, 0);
    int32_t n_levels = 0;
    for (int64_t cell = 0; cell < n_cells; cell++) {
        int32_t lv = 0;
        for (int i = 0; i < n; i++) lv = std::max(lv, last_level[(size_t)dofs[cell * n + i]]);
        lv += 1;
        for (int i = 0; i < n; i++) {
            int32_t& ll = last_level[(size_t)dofs[cell * n + i]];
            if (ll == lv) return cell + 1;
            ll = lv;
        }
        level[(size_t)cell] = lv;
        n_levels = std::max(n_levels, lv);
    }
    L.level_off.assign((size_t)n_levels + 1, 0);
    for (int64_t cell = 0; cell < n_cells; cell++) L.level_off[(size_t)level[(size_t)cell]]++;
    for (int32_t l = 0; l < n_levels; l++) L.level_off[(size_t)l + 1] += L.level_off[(size_t)l];
    L.cells.assign((size_t)n_cells, 0);
    std::vector<int64_t> cur(L.level_off.begin(), L.level_off.end() - 1);
    for (int64_t cell = 0; cell < n_cells; cell++) L.cells[(size_t)cur[(size_t)level[(size_t)cell] - 1]++] = (int32_t)cell;
    return 0;
}

}  // namespace nbfilter
