// grid_build.h -- host-side construction of the driving tables of the grid (TMA box) kernels, NB_FMT_GRID.
//
// When the host declares that its DoFs sit on a tensor-product grid (nb200_set_dof_grid: integer grid coordinates
// of every local DoF -- what a continuous FE_Q(p) space on any (stretched) hyper-rectangle mesh has, whatever the
// DoF numbering), the library keeps a second, lexicographic copy of the populations, f_grid[q][z][y][x], and the
// fused kernel fetches the support values of a whole tile of rows with TMA tensor copies (cp.async.bulk.tensor):
// the (p+1)^k support points of a row are a small box of that grid, and the boxes of the rows of a tile overlap
// into one (or, across a periodic boundary, a few) box per direction.  No per-element gather instructions, no
// staging index lists, and the copy of pass n+1 runs while the rows of pass n are multiplied.
//
// Input: the finished dictionary (dict_build.h; rows already sorted by grid offset at upload time, so that the k-th
// entry of every full row of a direction has the same offset relative to the row's first entry).
//   tile      = up to NB_CTA_ROWS grid points handled by one CTA: two half-tiles of <= 64 points (boxes of whole
//               cells where possible) that are neighbours in x, so that threads t and t+64 hold the same position in
//               neighbouring cells and share their weights (row pairing of the staged kernels).
//   box       = one TMA copy: origin (x, y, z) in the grid of direction a's population, fixed dims per direction.
//   pass      = consecutive directions whose boxes fit one staging buffer.
//   row desc  = offset of the row's first support value in the pass's buffer (class 0 rows: the row length and
//               offset template almost every row of the direction has), or the "generic" mark: the row is then
//               multiplied straight from its dictionary list (rows with bounce-back blocks, truncated rows, ...).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "dict_build.h"

namespace nbgrid {

struct PassHost {       // = NbGridPass
    int32_t box_begin;
    int16_t n_box, a0, a1, pad;
    int32_t bytes;      // expected TMA bytes of the pass (one distribution)
};
struct BoxHost {        // = NbGridBox
    int16_t x, y, z, dir;
    int32_t smem_off;   // doubles from the start of the staging buffer
};

struct Grid {
    int dim = 0;
    int32_t n[3] = {1, 1, 1};      // grid points per axis
    int64_t nxp = 0;               // padded x pitch (even, so that rows are 16-byte aligned)
    int64_t G = 0;                 // nxp * n[1] * n[2]
    int fe_order = 0;
    int xshift = 0;                // the host's x coordinates were shifted by this much when the grid was declared (cell faces sit
                                   // at xshift + multiples of p): 1 makes the x origin of every whole-cell half-tile even, which a
                                   // TMA store of the half-tile into the grid copy needs (16-byte aligned inner coordinate)
    std::vector<int32_t> gidx_of_int;   // [n_owned + n_ghost] internal canonical index -> flat grid index
    int64_t flat(int x, int y, int z) const { return ((int64_t)z * n[1] + y) * nxp + x; }
    void unflat(int64_t g, int c[3]) const
    {
        c[0] = (int)(g % nxp);
        c[1] = (int)((g / nxp) % n[1]);
        c[2] = (int)(g / (nxp * n[1]));
    }
};

// Sorts the entries of every CSR row into ascending grid position of their columns (z, y, x).  col: internal canonical
// column indices.  Returns false when the rows were already in that order (nothing copied); otherwise out_col / out_val
// hold the reordered arrays.
static inline bool sort_rows_by_grid(const std::vector<int32_t>& gidx_of_int, int64_t n_rows, const int64_t* rowptr, const int32_t* col,
                                     const double* val, std::vector<int32_t>& out_col, std::vector<double>& out_val)
{
    const int64_t nnz = rowptr[n_rows];
    bool sorted = true;
    for (int64_t i = 0; i < n_rows && sorted; i++)
        for (int64_t j = rowptr[i] + 1; j < rowptr[i + 1]; j++)
            if (gidx_of_int[(size_t)col[j]] < gidx_of_int[(size_t)col[j - 1]]) { sorted = false; break; }
    if (sorted) return false;
    if (out_col.data() != col) out_col.assign(col, col + nnz);
    if (out_val.data() != val) out_val.assign(val, val + nnz);
    unsigned nt = std::thread::hardware_concurrency();
    nt = nt == 0 ? 4 : (nt > 32 ? 32 : nt);
    if (n_rows < 20000) nt = 1;
    auto work = [&](int64_t lo, int64_t hi) {
        std::vector<std::pair<int32_t, int32_t>> key;
        std::vector<int32_t> tc;
        std::vector<double> tv;
        for (int64_t i = lo; i < hi; i++) {
            const int64_t b = rowptr[i], K = rowptr[i + 1] - b;
            key.resize((size_t)K); tc.resize((size_t)K); tv.resize((size_t)K);
            for (int64_t k = 0; k < K; k++) key[(size_t)k] = std::make_pair(gidx_of_int[(size_t)out_col[(size_t)(b + k)]], (int32_t)k);
            std::stable_sort(key.begin(), key.end());
            for (int64_t k = 0; k < K; k++) { tc[(size_t)k] = out_col[(size_t)(b + key[(size_t)k].second)]; tv[(size_t)k] = out_val[(size_t)(b + key[(size_t)k].second)]; }
            for (int64_t k = 0; k < K; k++) { out_col[(size_t)(b + k)] = tc[(size_t)k]; out_val[(size_t)(b + k)] = tv[(size_t)k]; }
        }
    };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nt; t++) th.emplace_back(work, n_rows * t / nt, n_rows * (t + 1) / nt);
    work(0, n_rows / nt);
    for (auto& t : th) t.join();
    return true;
}

struct Tables {
    int64_t n_tiles = 0, desc_stride = 0;
    std::vector<int32_t> tile_row, tile_gidx;     // [n_tiles * rows] canonical internal row / flat grid index, -1 = idle thread
    std::vector<int32_t> desc_x, desc_y;          // [n_dirs][desc_stride]
    std::vector<PassHost> passes;
    std::vector<int32_t> tile_pass;               // [n_tiles + 1]
    std::vector<BoxHost> boxes;
    std::vector<int32_t> box_dims;                // [n_dirs][3] TMA box dims per direction (x even)
    std::vector<int32_t> box_vol;                 // [n_dirs] doubles per box, rounded up to 16 (128-byte smem alignment)
    std::vector<int16_t> off_table;               // [n_dirs][max_k] offset of entry k relative to entry 0 inside a box
    std::vector<uint8_t> tile_reads_ghost;        // [n_tiles]
    std::vector<int16_t> tile_store;              // [n_tiles][4]: grid origin (x, y, z) of the tile's first half, bit h of [3] = half h
                                                  // may be written to the grid copy as ONE box (TMA store) instead of per thread
    int half_dims[3] = {1, 1, 1};                 // box of a half-tile
    std::vector<double> pair_hist;                // diagnostics (GRID_BUILD_PAIR_HIST): largest entry difference of every mismatched pair
    int64_t pairs = 0, pairs_same = 0, pairs_unified = 0;   // row pairs of class-0 box rows: with equal pattern ids / made equal (pair_tol)
    int64_t generic_rows = 0, grid_rows = 0, total_boxes = 0, max_pass_doubles = 0;
};

static inline void half_tile_dims(int dim, int p, int h[3])
{
    // largest box of whole cells (multiples of p per axis) with <= 64 points, x longest
    h[0] = h[1] = h[2] = 1;
    const int pp = p > 0 ? p : 4;
    int best = 0;
    const int zmax = dim == 3 ? 64 : 1;
    for (int a = pp; a <= 64; a += pp)
        for (int b = pp; b <= 64; b += pp)
            for (int c = (dim == 3 ? pp : 1); c <= zmax; c += pp) {
                const int v = a * b * c;
                if (v > 64) break;
                // prefer volume, then a cube-like shape (smaller source boxes), then long x
                const int score = v * 1000 - (a - b) * (a - b) - (dim == 3 ? (b - c) * (b - c) + (a - c) * (a - c) : 0);
                if (score > best) { best = score; h[0] = a; h[1] = b; h[2] = dim == 3 ? c : 1; }
                if (dim != 3) break;
            }
    if (best == 0) {    // p^dim > 64: no whole cell fits; fall back to a plain box
        h[0] = dim == 3 ? 4 : 8; h[1] = dim == 3 ? 4 : 8; h[2] = dim == 3 ? 4 : 1;
    }
}

// Builds the tables.  Returns false (tables unusable, keep the staged dictionary kernels) when a direction's box does not
// fit the staging buffer or its row length exceeds max_k.
//   dirs         finished dictionary, lists still on the host, rows sorted by grid offset
//   n_owned      rows; stride: population pitch of the canonical arrays (flat column = beta * stride + col)
//   cap          staging buffer capacity in doubles (per distribution)
//   pair_tol     > 0: rows t and t + rows/2 of a tile whose class-0 patterns got different ids but agree entry by entry to this
//                tolerance (the dictionary's value tolerance) both take the first one's pattern in the grid descriptors, so that
//                the kernels' paired product is taken by whole warps (see the note at the end of the function); 0 = ids as they are
static bool build(const std::vector<nbdict::DirBuild>& dirs, const Grid& g, int64_t n_owned, int64_t stride, int rows_per_tile,
                  int cap, int max_k, int empty_cls, Tables& T, double pair_tol = 0.0)
{
    const int nd = (int)dirs.size();
    const int p = g.fe_order > 0 ? g.fe_order : 4;
    int h[3];
    half_tile_dims(g.dim, g.fe_order, h);
    const int half_rows = rows_per_tile / 2;
    // tiles are aligned to cells: cell c of an axis owns coordinates 1 + c p .. (c + 1) p (coordinate 0 belongs to the
    // first cell), so tiles start at coordinate 1 - tile length (a thin first tile that only holds coordinate 0)
    const int org[3] = {(g.fe_order > 0 ? 1 : 0) + g.xshift, g.fe_order > 0 ? 1 : 0, g.fe_order > 0 ? 1 : 0};
    const int tl[3] = {2 * h[0], h[1], h[2]};
    int nt[3];
    for (int a = 0; a < 3; a++) {
        if (a >= g.dim) { nt[a] = 1; continue; }
        nt[a] = (g.n[a] - org[a] + tl[a] - 1) / tl[a] + (org[a] ? 1 : 0);
    }
    for (int a = 0; a < 3; a++) T.half_dims[a] = a < g.dim ? h[a] : 1;
    // grid points that hold a local DoF at all (owned or ghost): a box store may run over the others (padding), never over a
    // ghost or over a row of another tile
    std::vector<uint8_t> has_dof((size_t)g.G, 0);
    for (int32_t f : g.gidx_of_int) if (f >= 0) has_dof[(size_t)f] = 1;
    const bool store_shape_ok = (h[0] * 8) % 16 == 0 && half_rows >= h[0] * h[1] * h[2];
    T.tile_store.clear();
    // grid point -> canonical row (owned only)
    std::vector<int32_t> row_of_g((size_t)g.G, -1);
    for (int64_t i = 0; i < n_owned; i++) row_of_g[(size_t)g.gidx_of_int[(size_t)i]] = (int32_t)i;
    // ---- tiles (skip tiles without owned points)
    struct TileGeo { int o[3]; };
    std::vector<TileGeo> geo;
    T.tile_row.clear(); T.tile_gidx.clear();
    for (int tz = 0; tz < nt[2]; tz++)
        for (int ty = 0; ty < nt[1]; ty++)
            for (int tx = 0; tx < nt[0]; tx++) {
                int o[3] = {org[0] + (tx - (org[0] ? 1 : 0)) * tl[0], g.dim > 1 ? org[1] + (ty - (org[1] ? 1 : 0)) * tl[1] : 0,
                            g.dim > 2 ? org[2] + (tz - (org[2] ? 1 : 0)) * tl[2] : 0};
                int32_t rows[256], gi[256];
                bool any = false;
                for (int t = 0; t < rows_per_tile; t++) {
                    const int hf = t / half_rows, tt = t % half_rows;
                    int c[3] = {-1, -1, -1};
                    if (tt < h[0] * h[1] * h[2]) {
                        c[0] = o[0] + hf * h[0] + tt % h[0];
                        c[1] = o[1] + (tt / h[0]) % h[1];
                        c[2] = o[2] + tt / (h[0] * h[1]);
                    }
                    rows[t] = gi[t] = -1;
                    if (c[0] >= 0 && c[0] < g.n[0] && c[1] >= 0 && c[1] < g.n[1] && c[2] >= 0 && c[2] < g.n[2]) {
                        const int64_t f = g.flat(c[0], c[1], c[2]);
                        const int32_t r = row_of_g[(size_t)f];
                        if (r >= 0) { rows[t] = r; gi[t] = (int32_t)f; any = true; }
                    }
                }
                if (!any) continue;
                {   // which halves can go to the grid copy as one box
                    int flags = 0;
                    for (int hf = 0; hf < 2 && store_shape_ok; hf++) {
                        const int bx0 = o[0] + hf * h[0];
                        if (bx0 & 1) continue;                               // TMA: even inner coordinate
                        // boxes that stick out of the grid copy are left to the threads (the copy would clip them, but only whole
                        // boxes are worth the special case: they are all but a surface layer of the tiles)
                        if (bx0 < 0 || bx0 + h[0] > g.nxp || o[1] < 0 || o[1] + h[1] > g.n[1] || o[2] < 0 || o[2] + h[2] > g.n[2]) continue;
                        bool ok = true, some = false;
                        for (int tt = 0; tt < h[0] * h[1] * h[2] && ok; tt++) {
                            const int c[3] = {bx0 + tt % h[0], o[1] + (tt / h[0]) % h[1], o[2] + tt / (h[0] * h[1])};
                            if (c[0] < 0 || c[0] >= g.n[0] || c[1] < 0 || c[1] >= g.n[1] || c[2] < 0 || c[2] >= g.n[2]) continue;   // clipped by the copy
                            if (c[0] >= g.nxp) { ok = false; break; }
                            const int64_t f = g.flat(c[0], c[1], c[2]);
                            if (rows[hf * half_rows + tt] >= 0) some = true;
                            else if (has_dof[(size_t)f]) ok = false;        // a ghost (or nothing of ours): leave it alone
                        }
                        // points of the box between n[0] and the padded pitch are written too (harmless: never read as values)
                        if (ok && some) flags |= 1 << hf;
                    }
                    const int16_t rec[4] = {(int16_t)o[0], (int16_t)o[1], (int16_t)o[2], (int16_t)flags};
                    T.tile_store.insert(T.tile_store.end(), rec, rec + 4);
                }
                T.tile_row.insert(T.tile_row.end(), rows, rows + rows_per_tile);
                T.tile_gidx.insert(T.tile_gidx.end(), gi, gi + rows_per_tile);
                TileGeo tg; tg.o[0] = o[0]; tg.o[1] = o[1]; tg.o[2] = o[2];
                geo.push_back(tg);
            }
    const int64_t n_tiles = (int64_t)geo.size();
    T.n_tiles = n_tiles;
    T.desc_stride = std::max<int64_t>(32, n_tiles * rows_per_tile);
    {   // every owned row must sit in exactly one tile
        int64_t cnt = 0;
        for (int32_t r : T.tile_row) cnt += r >= 0;
        if (cnt != n_owned) return false;
    }
    // ---- per direction: class 0 template (relative grid offsets of a full row) and per-row source origin
    // origin[a][row] = flat grid index of the row's first support point, or -1 when the row is generic
    std::vector<std::vector<int32_t>> origin((size_t)nd);
    std::vector<std::vector<int>> tmpl((size_t)nd);        // [K][3] relative coordinates
    std::vector<int> ext((size_t)nd * 3, 1);
    for (int a = 0; a < nd; a++) {
        const nbdict::DirBuild& d = dirs[(size_t)a];
        origin[(size_t)a].assign((size_t)n_owned, -1);
        if (d.cls.empty()) continue;
        const nbdict::ClassBuild& C0 = d.cls[0];
        const int K = C0.K;
        if (K > max_k) return false;
        const int64_t lo = (int64_t)(a + 1) * stride, hi = lo + stride;      // columns of population a + 1 only
        std::vector<int>& tp = tmpl[(size_t)a];
        // relative grid coordinates of the K entries of every class-0 list; the template of the direction is the pattern
        // most lists have (a first-come choice could pick an oddity, e.g. a list that straddles a periodic seam)
        const int64_t nl = C0.n_lists();
        std::vector<int16_t> rel((size_t)nl * K * 3, 0);
        std::vector<uint8_t> lok((size_t)nl, 0), lused((size_t)nl, 0);
        for (int64_t r = 0; r < n_owned; r++) if (d.row_cls[(size_t)r] == 0) lused[(size_t)d.row_lst[(size_t)r]] = 1;
        std::vector<uint64_t> lhash((size_t)nl, 0);
        for (int64_t li = 0; li < nl; li++) {
            if (!lused[(size_t)li]) continue;
            const int32_t* L = C0.lists.data() + (size_t)li * K;
            bool ok = true;
            int c0[3] = {0, 0, 0};
            uint64_t hsh = 0x9e3779b97f4a7c15ull;
            for (int k = 0; k < K && ok; k++) {
                if (L[k] < lo || L[k] >= hi) { ok = false; break; }
                const int64_t gi = g.gidx_of_int[(size_t)(L[k] - lo)];
                if (gi < 0) { ok = false; break; }
                int c[3];
                g.unflat(gi, c);
                if (k == 0) { c0[0] = c[0]; c0[1] = c[1]; c0[2] = c[2]; }
                for (int j = 0; j < 3; j++) {
                    const int rj = c[j] - c0[j];
                    if (rj < 0 || rj > 64) ok = false;
                    rel[((size_t)li * K + k) * 3 + j] = (int16_t)rj;
                    hsh = nbdict::mix64(hsh, (uint64_t)(rj + 1));
                }
            }
            lok[(size_t)li] = ok ? 1 : 0;
            lhash[(size_t)li] = hsh;
        }
        // majority pattern
        int64_t best = -1;
        {
            std::vector<std::pair<uint64_t, int64_t>> hs;
            for (int64_t li = 0; li < nl; li++) if (lok[(size_t)li]) hs.emplace_back(lhash[(size_t)li], li);
            std::sort(hs.begin(), hs.end());
            int64_t best_cnt = 0;
            for (size_t i0 = 0; i0 < hs.size();) {
                size_t i1 = i0;
                while (i1 < hs.size() && hs[i1].first == hs[i0].first) i1++;
                if ((int64_t)(i1 - i0) > best_cnt) { best_cnt = (int64_t)(i1 - i0); best = hs[i0].second; }
                i0 = i1;
            }
        }
        const bool have = best >= 0;
        if (have) {
            tp.resize((size_t)K * 3);
            for (int k = 0; k < K * 3; k++) tp[(size_t)k] = rel[(size_t)best * K * 3 + k];
        }
        std::vector<int32_t> list_origin((size_t)nl, -1);
        for (int64_t li = 0; li < nl && have; li++) {
            if (!lok[(size_t)li]) continue;
            bool same = true;
            for (int k = 0; k < K * 3 && same; k++) same = rel[(size_t)li * K * 3 + k] == rel[(size_t)best * K * 3 + k];
            if (same) list_origin[(size_t)li] = (int32_t)g.gidx_of_int[(size_t)(C0.lists[(size_t)li * K] - lo)];
        }
        for (int64_t r = 0; r < n_owned; r++)
            if (d.row_cls[(size_t)r] == 0) origin[(size_t)a][(size_t)r] = list_origin[(size_t)d.row_lst[(size_t)r]];
        if (have)
            for (int k = 0; k < K; k++)
                for (int j = 0; j < 3; j++) ext[(size_t)a * 3 + j] = std::max(ext[(size_t)a * 3 + j], tp[(size_t)k * 3 + j] + 1);
    }
    // ---- per (tile, direction): cluster the source intervals per axis, boxes = products of clusters in use
    struct TB { int lo[3], len[3]; };
    std::vector<std::vector<TB>> tboxes((size_t)n_tiles * nd);
    std::vector<int16_t> row_box((size_t)nd * T.desc_stride, -1);      // box slot (within tile, direction) of every tile thread
    T.box_dims.assign((size_t)nd * 3, 1);
    const int merge_gap = p + 1;
    unsigned nthr = std::thread::hardware_concurrency();
    nthr = nthr == 0 ? 4 : (nthr > 32 ? 32 : nthr);
    if (n_tiles < 64) nthr = 1;
    std::vector<std::vector<int32_t>> dims_part(nthr, std::vector<int32_t>((size_t)nd * 3, 1));
    auto work = [&](unsigned th) {
        const int64_t b0 = n_tiles * th / nthr, b1 = n_tiles * (th + 1) / nthr;
        std::vector<int32_t>& dp = dims_part[th];
        for (int64_t b = b0; b < b1; b++)
            for (int a = 0; a < nd; a++) {
                std::vector<int> los[3];
                for (int t = 0; t < rows_per_tile; t++) {
                    const int32_t r = T.tile_row[(size_t)(b * rows_per_tile + t)];
                    if (r < 0) continue;
                    const int32_t og = origin[(size_t)a][(size_t)r];
                    if (og < 0) continue;
                    int c[3];
                    g.unflat(og, c);
                    for (int j = 0; j < 3; j++) los[j].push_back(c[j]);
                }
                if (los[0].empty()) continue;
                // clusters per axis: sorted distinct origins, merged while the intervals touch (or nearly do)
                std::vector<std::pair<int, int>> cl[3];     // (lo, hi) inclusive-exclusive
                for (int j = 0; j < 3; j++) {
                    std::sort(los[j].begin(), los[j].end());
                    los[j].erase(std::unique(los[j].begin(), los[j].end()), los[j].end());
                    const int e = ext[(size_t)a * 3 + j];
                    // a tile's rows read their own cells and one neighbour: longer clusters only arise where periodic images
                    // meet on a tiny mesh; cut them so that the box of a direction stays small
                    const int max_len = std::max(e, (j < g.dim ? tl[j] : 1) + p + 1);
                    for (int v : los[j]) {
                        if (cl[j].empty() || v > cl[j].back().second + merge_gap || v + e - cl[j].back().first > max_len) cl[j].emplace_back(v, v + e);
                        else cl[j].back().second = std::max(cl[j].back().second, v + e);
                    }
                    // TMA: the innermost coordinate of a box must be 16-byte aligned (even for doubles; an odd origin is an
                    // illegal instruction on sm_100): boxes start at the even coordinate at or below the cluster
                    if (j == 0) for (auto& c : cl[j]) c.first &= ~1;
                    for (auto& c : cl[j]) dp[(size_t)a * 3 + j] = std::max(dp[(size_t)a * 3 + j], c.second - c.first);
                }
                std::vector<TB>& tb = tboxes[(size_t)b * nd + a];
                for (int t = 0; t < rows_per_tile; t++) {
                    const int32_t r = T.tile_row[(size_t)(b * rows_per_tile + t)];
                    if (r < 0) continue;
                    const int32_t og = origin[(size_t)a][(size_t)r];
                    if (og < 0) continue;
                    int c[3], lo3[3];
                    g.unflat(og, c);
                    for (int j = 0; j < 3; j++) {
                        // the cluster that holds the row's whole interval [c, c + e)
                        const int e = ext[(size_t)a * 3 + j];
                        size_t ci = 0;
                        while (!(c[j] >= cl[j][ci].first && c[j] + e <= cl[j][ci].second)) ci++;
                        lo3[j] = cl[j][ci].first;
                    }
                    int slot = -1;
                    for (size_t s = 0; s < tb.size(); s++)
                        if (tb[s].lo[0] == lo3[0] && tb[s].lo[1] == lo3[1] && tb[s].lo[2] == lo3[2]) { slot = (int)s; break; }
                    if (slot < 0) {
                        TB nb;
                        for (int j = 0; j < 3; j++) { nb.lo[j] = lo3[j]; nb.len[j] = 0; }
                        tb.push_back(nb);
                        slot = (int)tb.size() - 1;
                    }
                    row_box[(size_t)a * T.desc_stride + (size_t)(b * rows_per_tile + t)] = (int16_t)slot;
                }
            }
    };
    if (nthr == 1) work(0);
    else {
        std::vector<std::thread> thv;
        for (unsigned t = 0; t < nthr; t++) thv.emplace_back(work, t);
        for (auto& t : thv) t.join();
    }
    for (auto& dp : dims_part)
        for (size_t i = 0; i < dp.size(); i++) T.box_dims[i] = std::max(T.box_dims[i], dp[i]);
    T.box_vol.assign((size_t)nd, 0);
    T.off_table.assign((size_t)nd * max_k, 0);
    for (int a = 0; a < nd; a++) {
        int32_t* bd = T.box_dims.data() + (size_t)a * 3;
        bd[0] = (bd[0] + 1) & ~1;                                 // 16-byte multiples in x
        if (bd[0] > 256 || bd[1] > 256 || bd[2] > 256) return false;
        const int64_t vol = (int64_t)bd[0] * bd[1] * bd[2];
        const int64_t volp = (vol + 15) / 16 * 16;
        if (volp > cap) return false;
        T.box_vol[(size_t)a] = (int32_t)volp;
        const std::vector<int>& tp = tmpl[(size_t)a];
        for (size_t k = 0; k * 3 < tp.size(); k++) {
            const int64_t off = ((int64_t)tp[k * 3 + 2] * bd[1] + tp[k * 3 + 1]) * bd[0] + tp[k * 3];
            if (off > 32767) return false;
            T.off_table[(size_t)a * max_k + k] = (int16_t)off;
        }
    }
    // ---- passes, boxes, descriptors
    T.desc_x.assign((size_t)nd * T.desc_stride, (int32_t)((uint32_t)empty_cls << 16));
    T.desc_y.assign((size_t)nd * T.desc_stride, 0);
    T.tile_pass.assign((size_t)n_tiles + 1, 0);
    T.passes.clear(); T.boxes.clear();
    T.tile_reads_ghost.assign((size_t)n_tiles, 0);
    T.generic_rows = T.grid_rows = 0;
    const bool any_ghost = (int64_t)g.gidx_of_int.size() > n_owned;
    for (int64_t b = 0; b < n_tiles; b++) {
        T.tile_pass[(size_t)b] = (int32_t)T.passes.size();
        PassHost cur{(int32_t)T.boxes.size(), 0, 0, 0, 0, 0};
        int32_t fill = 0, tx = 0;       // fill: doubles of the buffer in use (boxes start 128-byte aligned); tx: bytes the copies deliver
        for (int a = 0; a < nd; a++) {
            const nbdict::DirBuild& d = dirs[(size_t)a];
            const std::vector<TB>& tb = tboxes[(size_t)b * nd + a];
            // boxes beyond the buffer capacity (corner tiles of a periodic mesh need up to 2^dim of them) are dropped:
            // their rows are taken from the dictionary lists instead
            const size_t n_keep = std::min<size_t>(tb.size(), (size_t)(cap / std::max(1, T.box_vol[(size_t)a])));
            const int32_t need = (int32_t)n_keep * T.box_vol[(size_t)a];
            if (fill + need > cap) {
                cur.a1 = (int16_t)a;
                cur.bytes = tx;
                T.max_pass_doubles = std::max<int64_t>(T.max_pass_doubles, fill);
                T.passes.push_back(cur);
                cur = PassHost{(int32_t)T.boxes.size(), 0, (int16_t)a, (int16_t)a, 0, 0};
                fill = 0;
                tx = 0;
            }
            const int32_t base = fill;
            const int32_t* bd = T.box_dims.data() + (size_t)a * 3;
            for (size_t s = 0; s < n_keep; s++) {
                BoxHost bx;
                bx.x = (int16_t)tb[s].lo[0]; bx.y = (int16_t)tb[s].lo[1]; bx.z = (int16_t)tb[s].lo[2];
                bx.dir = (int16_t)a;
                bx.smem_off = base + (int32_t)s * T.box_vol[(size_t)a];
                T.boxes.push_back(bx);
                cur.n_box++;
                tx += bd[0] * bd[1] * bd[2] * 8;
            }
            fill += need;
            for (int t = 0; t < rows_per_tile; t++) {
                const size_t slot = (size_t)(b * rows_per_tile + t);
                const int32_t r = T.tile_row[slot];
                if (r < 0) continue;
                const int ci = d.row_cls[(size_t)r];
                if (ci < 0) continue;                                   // empty row: K = 0 class
                const size_t di = (size_t)a * T.desc_stride + slot;
                const int32_t og = origin[(size_t)a][(size_t)r];
                // does the row read a ghost slot?  (lists hold flat canonical indices)
                if (any_ghost && !T.tile_reads_ghost[(size_t)b]) {
                    const nbdict::ClassBuild& C = d.cls[(size_t)ci];
                    const int32_t* L = C.lists.data() + (size_t)d.row_lst[(size_t)r] * C.K;
                    for (int k = 0; k < C.K; k++) if (L[k] % stride >= n_owned) { T.tile_reads_ghost[(size_t)b] = 1; break; }
                }
                if (og < 0 || (size_t)row_box[di] >= n_keep) {
                    T.desc_x[di] = (int32_t)0x80000000u;                // generic: the kernel takes the dictionary descriptor
                    T.generic_rows++;
                    continue;
                }
                int c[3];
                g.unflat(og, c);
                const TB& bx = tb[(size_t)row_box[di]];
                const int32_t off = base + (int32_t)row_box[di] * T.box_vol[(size_t)a]
                    + ((c[2] - bx.lo[2]) * bd[1] + (c[1] - bx.lo[1])) * bd[0] + (c[0] - bx.lo[0]);
                if (off > 0xffff) return false;
                T.desc_x[di] = off;                                     // class 0
                T.desc_y[di] = d.row_pat[(size_t)r];
                T.grid_rows++;
            }
        }
        cur.a1 = (int16_t)nd;
        cur.bytes = tx;
        T.max_pass_doubles = std::max<int64_t>(T.max_pass_doubles, fill);
        T.passes.push_back(cur);
    }
    T.tile_pass[(size_t)n_tiles] = (int32_t)T.passes.size();
    T.total_boxes = (int64_t)T.boxes.size();
    // Pairing.  The kernels multiply rows t and t + rows/2 of a tile together when both are class-0 box rows with the SAME pattern
    // id (one weight load feeds both); a pair with different ids takes two separate products -- and drags its whole warp through
    // both code paths.  With a value tolerance the dictionary keeps one representative per bucket of round-off variants, and which
    // bucket a row falls into is decided by its noise: on the uniform 32^3-cell mesh 2.5 % of the pairs (same position in
    // x-neighbour cells, weights equal up to round-off) ended up with different ids, sprinkled over half of all warps (ncu: 24 of
    // 32 threads active per instruction in the product loop; thread-level instruction counts equal to a graded mesh's, warp-level
    // 13 % higher).  Such pairs are unified here: if the two patterns agree entry by entry to pair_tol, the second row takes the
    // first one's id.  Its stored weights then differ from its own by at most 2 pair_tol instead of pair_tol.
    T.pairs = T.pairs_same = T.pairs_unified = 0;
    for (int a = 0; a < nd; a++) {
        const nbdict::DirBuild& d = dirs[(size_t)a];
        if (d.cls.empty()) continue;
        const nbdict::ClassBuild& C0 = d.cls[0];
        const int K = C0.K;
        for (int64_t b = 0; b < n_tiles; b++)
            for (int t = 0; t < half_rows; t++) {
                const size_t s0 = (size_t)a * T.desc_stride + (size_t)(b * rows_per_tile + t), s1 = s0 + (size_t)half_rows;
                const uint32_t x0 = (uint32_t)T.desc_x[s0], x1 = (uint32_t)T.desc_x[s1];
                if (((x0 | x1) >> 16) != 0) continue;                   // idle, empty or generic on either side
                T.pairs++;
                if (T.desc_y[s0] == T.desc_y[s1]) { T.pairs_same++; continue; }
                if (!(pair_tol > 0.0) || C0.pats.empty()) continue;
                const double* p0 = C0.pats.data() + (size_t)T.desc_y[s0] * K;
                const double* p1 = C0.pats.data() + (size_t)T.desc_y[s1] * K;
                bool close = true;
                double worst = 0.0;
                for (int k = 0; k < K; k++) worst = std::max(worst, std::fabs(p0[k] - p1[k]));
                if (getenv("GRID_BUILD_PAIR_HIST")) T.pair_hist.push_back(worst);
                close = worst <= pair_tol;
                if (close) { T.desc_y[s1] = T.desc_y[s0]; T.pairs_unified++; }
            }
    }
    return true;
}

}  // namespace nbgrid
