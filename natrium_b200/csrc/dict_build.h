// dict_build.h -- host-side construction of the dictionary streaming format (NB_FMT_DICT).
//
// Input: the CSR blocks of getSystemMatrix() as they arrive through nb200_upload_block_csr.
// For every block-row alpha each row becomes (list id, pattern id):
//   list    = the row's column indices (already flattened to population indices beta*stride+col);
//             rows whose departure point lies in the same source cell have identical lists
//             (all DoFs of that cell that carry a non-zero shape value, SemiLagrangian.cpp:476-501);
//   pattern = the row's values (shape function values at the departure point).  Two rows share a
//             pattern if every entry agrees to within `tol` (tol = 0: bitwise).  The reference
//             itself drops entries below 1e-10 (SemiLagrangian.cpp:483), so a tolerance several
//             orders below that does not change the operator beyond its own noise floor.
// Lists and patterns are pooled per row length K.  Ids are handed out in first-occurrence order, so
// without any sharing the pools are simply the rows in row order (a column-major ELL).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <thread>
#include <utility>
#include <vector>

namespace nbdict {

static inline uint64_t mix64(uint64_t h, uint64_t v)
{
    h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
    h *= 0xff51afd7ed558ccdull;
    h ^= h >> 32;
    return h;
}

// open-addressing index over ids whose content lives elsewhere; equality is decided by the caller
struct HashIndex {
    std::vector<int32_t> slot;     // -1 = empty
    std::vector<uint64_t> hash_of; // per id
    size_t mask = 0;

    void reserve_pow2(size_t n)
    {
        size_t cap = 64;
        while (cap < n * 2) cap <<= 1;
        slot.assign(cap, -1);
        mask = cap - 1;
        for (size_t id = 0; id < hash_of.size(); id++) place((int32_t)id, hash_of[id]);
    }
    void place(int32_t id, uint64_t h)
    {
        size_t p = (size_t)h & mask;
        while (slot[p] >= 0) p = (p + 1) & mask;
        slot[p] = id;
    }
    // returns the id of an equal entry, or -1 after which the caller must call insert()
    template <class Eq>
    int32_t find(uint64_t h, Eq eq) const
    {
        if (slot.empty()) return -1;
        size_t p = (size_t)h & mask;
        while (slot[p] >= 0) {
            const int32_t id = slot[p];
            if (hash_of[id] == h && eq(id)) return id;
            p = (p + 1) & mask;
        }
        return -1;
    }
    void insert(uint64_t h)
    {
        const int32_t id = (int32_t)hash_of.size();
        hash_of.push_back(h);
        if (hash_of.size() * 2 > slot.size()) reserve_pow2(hash_of.size() * 2);
        else place(id, h);
    }
};

struct ClassBuild {
    int K = 0;
    std::vector<int32_t> lists;   // [n_lists][K]
    std::vector<double> pats;     // [n_pats][K]
    HashIndex hl, hp;
    int64_t n_lists() const { return (int64_t)hl.hash_of.size(); }
    int64_t n_pats() const { return (int64_t)hp.hash_of.size(); }
};

struct DirBuild {
    std::vector<ClassBuild> cls;
    std::vector<int32_t> row_lst;   // per row; -1 = empty row
    std::vector<int32_t> row_pat;
    std::vector<int8_t> row_cls;
    int64_t nnz = 0;

    void init(int64_t n_rows)
    {
        cls.clear();
        row_lst.assign((size_t)n_rows, -1);
        row_pat.assign((size_t)n_rows, 0);
        row_cls.assign((size_t)n_rows, -1);
        nnz = 0;
    }
    // Class 0 is the one the kernels treat best (its parameters travel in the kernel arguments, the grid kernels take its rows
    // from TMA boxes): make it the class that holds the most rows.  Classes are numbered as row lengths first appear, and the
    // first row of a direction can be an oddity -- the corner DoF of a walled mesh whose path is stuck (one entry) made all
    // nine-entry rows of 8 of D2Q25H's 24 directions second class.  Call once all blocks are in.
    void majority_class_first()
    {
        if (cls.size() < 2) return;
        std::vector<int64_t> cnt(cls.size(), 0);
        for (int8_t c : row_cls) if (c >= 0) cnt[(size_t)c]++;
        size_t best = 0;
        for (size_t c = 1; c < cls.size(); c++) if (cnt[c] > cnt[best]) best = c;
        if (best == 0) return;
        std::swap(cls[0], cls[best]);
        for (auto& c : row_cls) {
            if (c == 0) c = (int8_t)best;
            else if (c == (int8_t)best) c = 0;
        }
    }
    // Class holding rows of length K.  The first `exact_limit` distinct lengths get a class of their own;
    // later ones share power-of-two classes (rows are padded with zero weights), so the class count is bounded.
    int class_of(int K, int exact_limit, int* padded_K)
    {
        for (size_t c = 0; c < cls.size(); c++) if (cls[c].K == K) { *padded_K = K; return (int)c; }
        int Kp = K;
        if ((int)cls.size() >= exact_limit) {
            Kp = 1;
            while (Kp < K) Kp <<= 1;
            for (size_t c = 0; c < cls.size(); c++) if (cls[c].K == Kp) { *padded_K = Kp; return (int)c; }
        }
        cls.emplace_back();
        cls.back().K = Kp;
        *padded_K = Kp;
        return (int)cls.size() - 1;
    }
};

static inline uint64_t hash_list(const int32_t* c, int K)
{
    uint64_t h = 0x243f6a8885a308d3ull ^ (uint64_t)K;
    for (int k = 0; k < K; k++) h = mix64(h, (uint64_t)(uint32_t)c[k]);
    return h;
}

// tol > 0: hash a coarse quantisation (2^-32) so that values that differ by round-off almost always
// land in the same bucket; equality is then decided against the bucket's representatives.
static inline uint64_t hash_pattern(const double* v, int K, double tol)
{
    uint64_t h = 0x13198a2e03707344ull ^ (uint64_t)K;
    if (tol > 0.0) {
        for (int k = 0; k < K; k++) h = mix64(h, (uint64_t)(int64_t)std::nearbyint(v[k] * 4294967296.0));
    } else {
        for (int k = 0; k < K; k++) {
            uint64_t b;
            double x = v[k] == 0.0 ? 0.0 : v[k];   // -0.0 == +0.0
            memcpy(&b, &x, 8);
            h = mix64(h, b);
        }
    }
    return h;
}

// Adds the rows of one CSR block (flat column offset col_base = (bj+1)*stride) to the direction.
// Rows that already hold entries from an earlier block of the same block-row are concatenated.
// Returns false if a limit of the device format is exceeded (msg says which).
static bool add_block(DirBuild& d, int64_t n_rows, const int64_t* rowptr, const int32_t* col, const double* val,
                      int64_t col_base, double tol, int max_cls, int64_t max_pat, const char** msg)
{
    // phase 1 (parallel): per-row hashes
    std::vector<uint64_t> hl((size_t)n_rows), hp((size_t)n_rows);
    std::vector<int32_t> flat((size_t)(rowptr[n_rows]));
    unsigned nt = std::thread::hardware_concurrency();
    nt = nt == 0 ? 4 : (nt > 32 ? 32 : nt);
    if (n_rows < 20000) nt = 1;
    auto work = [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; i++) {
            const int64_t b = rowptr[i];
            const int K = (int)(rowptr[i + 1] - b);
            for (int k = 0; k < K; k++) flat[(size_t)(b + k)] = (int32_t)(col_base + col[b + k]);
            hl[(size_t)i] = hash_list(flat.data() + b, K);
            hp[(size_t)i] = hash_pattern(val + b, K, tol);
        }
    };
    if (nt == 1) work(0, n_rows);
    else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; t++) th.emplace_back(work, n_rows * t / nt, n_rows * (t + 1) / nt);
        for (auto& t : th) t.join();
    }
    // phase 2 (sequential): dictionary lookups in row order
    std::vector<int32_t> ctmp;
    std::vector<double> vtmp;
    for (int64_t i = 0; i < n_rows; i++) {
        const int64_t b = rowptr[i];
        int K = (int)(rowptr[i + 1] - b);
        if (K == 0) continue;
        const int32_t* c = flat.data() + b;
        const double* v = val + b;
        uint64_t hli = hl[(size_t)i], hpi = hp[(size_t)i];
        if (d.row_lst[(size_t)i] >= 0) {   // concatenate with what the row already holds
            const ClassBuild& o = d.cls[(size_t)d.row_cls[(size_t)i]];
            ctmp.assign(o.lists.begin() + (size_t)d.row_lst[(size_t)i] * o.K, o.lists.begin() + (size_t)(d.row_lst[(size_t)i] + 1) * o.K);
            vtmp.assign(o.pats.begin() + (size_t)d.row_pat[(size_t)i] * o.K, o.pats.begin() + (size_t)(d.row_pat[(size_t)i] + 1) * o.K);
            ctmp.insert(ctmp.end(), c, c + K);
            vtmp.insert(vtmp.end(), v, v + K);
            K = (int)ctmp.size();
            c = ctmp.data();
            v = vtmp.data();
            hli = hash_list(c, K);
            hpi = hash_pattern(v, K, tol);
        }
        int Kp = K;
        const int ci = d.class_of(K, max_cls - 20, &Kp);
        if (ci >= max_cls) { *msg = "more distinct row lengths in one direction than the dictionary format supports"; return false; }
        if (Kp != K) {   // pad: zero weight on the row's last column
            if (c != ctmp.data()) { ctmp.assign(c, c + K); vtmp.assign(v, v + K); }
            ctmp.resize((size_t)Kp, ctmp.back());
            vtmp.resize((size_t)Kp, 0.0);
            K = Kp;
            c = ctmp.data();
            v = vtmp.data();
            hli = hash_list(c, K);
            hpi = hash_pattern(v, K, tol);
        }
        ClassBuild& C = d.cls[(size_t)ci];
        int32_t lid = C.hl.find(hli, [&](int32_t id) { return memcmp(C.lists.data() + (size_t)id * K, c, (size_t)K * 4) == 0; });
        if (lid < 0) {
            lid = (int32_t)C.n_lists();
            C.lists.insert(C.lists.end(), c, c + K);
            C.hl.insert(hli);
        }
        int32_t pid = C.hp.find(hpi, [&](int32_t id) {
            const double* r = C.pats.data() + (size_t)id * K;
            for (int k = 0; k < K; k++) if (!(std::fabs(r[k] - v[k]) <= tol)) return false;
            return true;
        });
        if (pid < 0) {
            pid = (int32_t)C.n_pats();
            if (pid >= max_pat) { *msg = "more weight patterns in one direction than the dictionary format supports"; return false; }
            C.pats.insert(C.pats.end(), v, v + K);
            C.hp.insert(hpi);
        }
        d.nnz += rowptr[i + 1] - b;
        d.row_lst[(size_t)i] = lid;
        d.row_pat[(size_t)i] = pid;
        d.row_cls[(size_t)i] = (int8_t)ci;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// NB_FMT_STAGED tables (see stream_common.cuh).  Input: the finished dictionary (row -> class, list, pattern for
// every direction, lists still on the host).  For every CTA (cta_rows consecutive rows) and direction the distinct
// column lists of its rows are appended to stage_col in first-occurrence order; directions are packed into passes
// of at most `cap` staged values.  Returns false when a single direction of some CTA needs more than `cap` values
// (the caller then keeps the plain dictionary kernel).
// ---------------------------------------------------------------------------------------------
struct StagePassHost {
    int64_t begin;
    int32_t count;
    int16_t a0, a1;
};

struct StagingBuild {
    std::vector<int32_t> stage_col;
    std::vector<StagePassHost> passes;
    std::vector<int32_t> cta_ptr;          // [n_cta + 1]
    std::vector<int32_t> sdesc_x;          // [n_dirs][desc_stride]  offset | class << 16
    int64_t max_pass_count = 0;
};

static bool build_staging(const std::vector<DirBuild>& dirs, int64_t n_rows, int64_t desc_stride, int cta_rows, int cap,
                          int empty_cls, StagingBuild& out)
{
    const int nd = (int)dirs.size();
    const int64_t n_cta = (n_rows + cta_rows - 1) / cta_rows;
    out.cta_ptr.assign((size_t)n_cta + 1, 0);
    out.sdesc_x.assign((size_t)nd * desc_stride, (int32_t)((uint32_t)empty_cls << 16));
    unsigned nt = std::thread::hardware_concurrency();
    nt = nt == 0 ? 4 : (nt > 32 ? 32 : nt);
    if (n_cta < 64) nt = 1;
    // each worker builds the passes / stage_col of a contiguous range of CTAs; stitched together afterwards
    struct Part { std::vector<int32_t> col; std::vector<StagePassHost> passes; std::vector<int32_t> n_pass; bool ok = true; int64_t max_count = 0; };
    std::vector<Part> parts(nt);
    auto work = [&](unsigned t) {
        Part& P = parts[t];
        const int64_t b0 = n_cta * t / nt, b1 = n_cta * (t + 1) / nt;
        // scratch: open-addressing table over (class, list) of one CTA and direction
        const int TS = 1024;                      // > 2 * cta_rows
        std::vector<int64_t> key(TS, (int64_t)-1);
        std::vector<int32_t> val(TS);
        std::vector<int> used;
        for (int64_t b = b0; b < b1; b++) {
            const int64_t r0 = b * cta_rows, r1 = std::min<int64_t>(n_rows, r0 + cta_rows);
            StagePassHost cur{(int64_t)P.col.size(), 0, 0, 0};
            int np = 0;
            for (int a = 0; a < nd; a++) {
                const DirBuild& d = dirs[(size_t)a];
                // distinct lists of this (CTA, direction), first-occurrence order
                for (int u : used) key[(size_t)u] = -1;
                used.clear();
                int32_t need = 0;
                std::vector<std::pair<int, int32_t>> order;   // (class, list) in staging order
                for (int64_t r = r0; r < r1; r++) {
                    const int ci = d.row_cls[(size_t)r];
                    if (ci < 0) continue;
                    const int64_t kk = ((int64_t)ci << 32) | (uint32_t)d.row_lst[(size_t)r];
                    size_t h = (size_t)(mix64(0x9e3779b97f4a7c15ull, (uint64_t)kk)) & (TS - 1);
                    while (key[h] >= 0 && key[h] != kk) h = (h + 1) & (TS - 1);
                    if (key[h] < 0) {
                        key[h] = kk;
                        val[h] = need;
                        used.push_back((int)h);
                        order.emplace_back(ci, d.row_lst[(size_t)r]);
                        need += d.cls[(size_t)ci].K;
                    }
                }
                if (need > cap) { P.ok = false; return; }
                if (cur.count + need > cap) {       // close the pass before this direction
                    cur.a1 = (int16_t)a;
                    P.passes.push_back(cur);
                    P.max_count = std::max<int64_t>(P.max_count, cur.count);
                    np++;
                    cur = StagePassHost{(int64_t)P.col.size(), 0, (int16_t)a, (int16_t)a};
                }
                const int32_t base = cur.count;
                for (auto& cl : order) {
                    const ClassBuild& C = d.cls[(size_t)cl.first];
                    const int32_t* L = C.lists.data() + (size_t)cl.second * C.K;
                    P.col.insert(P.col.end(), L, L + C.K);
                }
                int32_t* sx = out.sdesc_x.data() + (size_t)a * desc_stride;
                for (int64_t r = r0; r < r1; r++) {
                    const int ci = d.row_cls[(size_t)r];
                    if (ci < 0) continue;
                    const int64_t kk = ((int64_t)ci << 32) | (uint32_t)d.row_lst[(size_t)r];
                    size_t h = (size_t)(mix64(0x9e3779b97f4a7c15ull, (uint64_t)kk)) & (TS - 1);
                    while (key[h] != kk) h = (h + 1) & (TS - 1);
                    sx[r] = (int32_t)(((uint32_t)ci << 16) | (uint32_t)(base + val[h]));
                }
                cur.count += need;
            }
            cur.a1 = (int16_t)nd;
            P.passes.push_back(cur);
            P.max_count = std::max<int64_t>(P.max_count, cur.count);
            np++;
            P.n_pass.push_back(np);
        }
    };
    if (nt == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; t++) th.emplace_back(work, t);
        for (auto& t : th) t.join();
    }
    for (auto& P : parts) if (!P.ok) return false;
    size_t tot_col = 0, tot_pass = 0;
    for (auto& P : parts) { tot_col += P.col.size(); tot_pass += P.passes.size(); }
    if (tot_pass >= (size_t)INT32_MAX) return false;
    out.stage_col.resize(tot_col);
    out.passes.resize(tot_pass);
    size_t oc = 0, op = 0;
    int64_t b = 0;
    for (auto& P : parts) {
        memcpy(out.stage_col.data() + oc, P.col.data(), P.col.size() * sizeof(int32_t));
        for (size_t i = 0; i < P.passes.size(); i++) {
            out.passes[op + i] = P.passes[i];
            out.passes[op + i].begin += (int64_t)oc;
        }
        size_t pp = op;
        for (int npass : P.n_pass) {
            out.cta_ptr[(size_t)b] = (int32_t)pp;
            pp += (size_t)npass;
            b++;
        }
        oc += P.col.size();
        op += P.passes.size();
        out.max_pass_count = std::max(out.max_pass_count, P.max_count);
        std::vector<int32_t>().swap(P.col);
    }
    out.cta_ptr[(size_t)n_cta] = (int32_t)op;
    return true;
}

}  // namespace nbdict
