// knn.cuh — device-side exact k-nearest-neighbour search over the uniform-grid spatial hash
// (shared by the bulk kernel in grid.cu and the planner engines).
//
// Replaces CGAL::Orthogonal_k_neighbor_search (reference CGAL/Orthogonal_k_neighbor_search.h:51-79,
// 121-150; call sites RTRRT.cpp:96,265,304,333,368 etc.): the min(k,N) points of smallest squared
// distance (qx-px)^2 + (qy-py)^2 evaluated in binary64 on the float coordinates widened exactly
// (CGAL/Euclidean_distance.h:72-83), ascending.  Ties: CGAL's order among equal distances is
// unspecified; ours is (d2, insertion id) ascending — the oracle's rule.
//
// One warp answers one query.  The grid has power-of-two cells (exact binning: cell = (int)x >> s).
// The warp walks square rings of cells around the query cell; a row of cells is one contiguous,
// coalesced range of the cell-sorted point array.  Lane j holds the j-th best (d2,id) so far
// (sorted across lanes); candidates that beat the current k-th are inserted with one ballot and
// one shfl_up (chunks with many winners are merged in with a bitonic sort + merge).  The walk stops when the k-th distance is strictly below the squared gap between
// the query and every still-open side of the scanned square (computed rounding down), which
// proves that no unscanned point can enter or tie into the result.  Points appended since the
// last re-sort (the "tail") are scanned linearly by every query.
#pragma once
#include <math_constants.h>

#include "common.cuh"

// (no __restrict__/read-only qualifiers: the planner engines append nodes and re-sort the hash
// inside the same kernel that queries it, so these loads must stay coherent)
struct GridView {
    const float2 *sxy;                    // cell-sorted coordinates            [n_sorted]
    const int *sid;                       // cell-sorted position -> node id    [n_sorted]
    const int *cell_start;                // first sorted position of each cell [gw*gh + 1]
    const float2 *xy;                     // coordinates by node id             [n]
    int gw, gh, shift;                    // grid of (1<<shift)-pixel cells
    int n_sorted, n;                      // ids >= n_sorted are the unsorted tail
    const unsigned char *gap = nullptr;   // optional [gw*gh]: rings 0 .. gap-1 around the cell hold no sorted node (saturates at 255)
};

struct KnnEntry {
    double d2;
    int id;
};

__device__ __forceinline__ bool knn_less(double a, int ia, double b, int ib) { return a < b || (a == b && ia < ib); }

// squared distance exactly as CGAL evaluates it without FMA contraction
__device__ __forceinline__ double knn_d2(double qx, double qy, float px, float py) {
    const double dx = __dsub_rn(qx, (double)px), dy = __dsub_rn(qy, (double)py);
    return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
}

__device__ __forceinline__ int grid_cell_of(double v, int shift, int gmax) {
    // floor for v >= 0; everything below 0 (or NaN) lands in cell 0, everything beyond in gmax
    int c = (v > 0.0) ? (__double2int_rd(fmin(v, 1073741823.0)) >> shift) : 0;
    return min(c, gmax);
}

// Warp-collective.  All 32 lanes must call with the same q and k (1 <= k <= 32).
// On return lane j (j < count) holds the j-th neighbour; returns count = min(k, n).
// found0 < 0: a fresh query.  found0 >= 0: `mine` already holds the exact result over the node ids [0, scan_from)
// (lanes >= found0 hold +inf) and only the ids [scan_from, n) are scanned, linearly — the planner engines cache the
// neighbour list of a tree node and extend it by the nodes appended since (nodes never move or disappear).
// kWide: the variant for queries that may lie far from every node (bulk queries at arbitrary points: a third of uniform
// samples fall inside obstacles at C4): a ring's side cells are looked up one per lane and the walk starts at the first
// non-empty ring (GridView::gap).  The engines and the all-nodes table keep the plain walk: their queries sit among the
// nodes, and the extra code costs the all-nodes kernel registers (0.72 -> 0.77 ms when it was shared).
template <bool kWide = false>
__device__ __forceinline__ int knn_warp_impl(const GridView &g, double qx, double qy, int k, KnnEntry &mine, int found0, int scan_from) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    if (found0 < 0) {
        mine.d2 = CUDART_INF;
        mine.id = 0x7fffffff;
    }
    int found = found0 < 0 ? 0 : found0;
    double kd2 = __shfl_sync(full, mine.d2, k - 1); // current k-th entry (broadcast copy)
    int kid = __shfl_sync(full, mine.id, k - 1);

    // compare-exchange with the lane `stride` away: keep the smaller (keep_min) or larger pair
    auto cmpx = [&](double &d, int &i, int stride, bool keep_min) {
        const double od = __shfl_xor_sync(full, d, stride);
        const int oi = __shfl_xor_sync(full, i, stride);
        const bool other_less = knn_less(od, oi, d, i);
        if (other_less == keep_min) d = od, i = oi;
    };
    auto offer = [&](bool valid, double d2, int id) {
        const bool acc = valid && knn_less(d2, id, kd2, kid);
        unsigned m = __ballot_sync(full, acc);
        const int cnt = __popc(m);
        if (cnt == 0) return;
        if (cnt >= 6) {
            // many candidates beat the k-th (early chunks): sort the chunk across the lanes (bitonic,
            // 15 compare-exchange stages), keep the 32 smallest of {best list, chunk} and re-sort
            // them with one bitonic merge (the pointwise minimum of an ascending and a reversed
            // ascending sequence is bitonic)
            double cd = acc ? d2 : CUDART_INF;
            int ci = acc ? id : 0x7fffffff;
#pragma unroll
            for (int size = 2; size <= 32; size <<= 1)
#pragma unroll
                for (int stride = size >> 1; stride > 0; stride >>= 1)
                    cmpx(cd, ci, stride, ((lane & stride) == 0) == ((lane & size) == 0));
            const double rd = __shfl_sync(full, cd, 31 - lane);
            const int ri = __shfl_sync(full, ci, 31 - lane);
            if (knn_less(rd, ri, mine.d2, mine.id)) mine.d2 = rd, mine.id = ri;
#pragma unroll
            for (int stride = 16; stride > 0; stride >>= 1) cmpx(mine.d2, mine.id, stride, (lane & stride) == 0);
            if (lane >= k) {
                mine.d2 = CUDART_INF;
                mine.id = 0x7fffffff;
            }
            found = min(found + cnt, k);
            kd2 = __shfl_sync(full, mine.d2, k - 1);
            kid = __shfl_sync(full, mine.id, k - 1);
            return;
        }
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const double cd2 = __shfl_sync(full, d2, src);
            const int cid = __shfl_sync(full, id, src);
            if (!knn_less(cd2, cid, kd2, kid)) continue; // the k-th moved since the ballot
            const int pos = __popc(__ballot_sync(full, knn_less(mine.d2, mine.id, cd2, cid)));
            const double ud2 = __shfl_up_sync(full, mine.d2, 1);
            const int uid = __shfl_up_sync(full, mine.id, 1);
            if (lane == pos) {
                mine.d2 = cd2;
                mine.id = cid;
            } else if (lane > pos) {
                mine.d2 = ud2;
                mine.id = uid;
            }
            if (lane >= k) {
                mine.d2 = CUDART_INF;
                mine.id = 0x7fffffff;
            }
            found = min(found + 1, k);
            kd2 = __shfl_sync(full, mine.d2, k - 1);
            kid = __shfl_sync(full, mine.id, k - 1);
        }
    };

    // unsorted tail first (usually empty or a few points); a resumed query scans [scan_from, n) by id instead
    for (int i0 = found0 < 0 ? g.n_sorted : scan_from; i0 < g.n; i0 += 32) {
        const int i = i0 + lane;
        const bool v = i < g.n;
        float2 p = make_float2(0.f, 0.f);
        if (v) p = g.xy[i];
        offer(v, knn_d2(qx, qy, p.x, p.y), i);
    }

    if (found0 < 0 && g.n_sorted > 0) {
        const int cx = grid_cell_of(qx, g.shift, g.gw - 1), cy = grid_cell_of(qy, g.shift, g.gh - 1);
        const int want = min(k, g.n);
        const int rmax = max(max(cx, g.gw - 1 - cx), max(cy, g.gh - 1 - cy));
        auto scan_range = [&](int s, int e) { // positions [s, e) of the cell-sorted arrays
            for (int i0 = s; i0 < e; i0 += 32) {
                const int i = i0 + lane;
                const bool v = i < e;
                float2 p = make_float2(0.f, 0.f);
                int id = 0;
                if (v) {
                    p = g.sxy[i];
                    id = g.sid[i];
                }
                offer(v, knn_d2(qx, qy, p.x, p.y), id);
            }
        };
        auto scan_row = [&](int y, int x0, int x1) {
            if (y < 0 || y >= g.gh) return;
            x0 = max(x0, 0), x1 = min(x1, g.gw - 1);
            if (x0 > x1) return;
            scan_range(g.cell_start[y * g.gw + x0], g.cell_start[y * g.gw + x1 + 1]);
        };
        // rings known to be empty (a query deep inside an obstacle) are skipped: exact, the table describes the sorted part
        for (int r = (kWide && g.gap) ? min((int)g.gap[cy * g.gw + cx], rmax) : 0; r <= rmax; r++) {
            if (r == 0)
                scan_row(cy, cx, cx);
            else {
                scan_row(cy - r, cx - r, cx + r);
                scan_row(cy + r, cx - r, cx + r);
                // the 2(2r - 1) side cells of the ring: one lane looks up one cell, only the non-empty ones are scanned (a
                // query far from every node — a sample inside an obstacle — walks many rings of empty cells; the result does
                // not depend on the order in which candidates are offered)
                const int nside = kWide ? 2 * (2 * r - 1) : 0;
                if (!kWide)
                    for (int y = cy - r + 1; y <= cy + r - 1; y++) {
                        scan_row(y, cx - r, cx - r);
                        scan_row(y, cx + r, cx + r);
                    }
                for (int c0 = 0; c0 < nside; c0 += 32) {
                    const int c = c0 + lane;
                    int cs = 0, ce = 0;
                    if (c < nside) {
                        const int y = cy - r + 1 + (c >> 1), x = (c & 1) ? cx + r : cx - r;
                        if (y >= 0 && y < g.gh && x >= 0 && x < g.gw) cs = g.cell_start[y * g.gw + x], ce = g.cell_start[y * g.gw + x + 1];
                    }
                    unsigned ne = __ballot_sync(full, ce > cs);
                    while (ne) {
                        const int src = __ffs(ne) - 1;
                        ne &= ne - 1;
                        scan_range(__shfl_sync(full, cs, src), __shfl_sync(full, ce, src));
                    }
                }
            }
            if (found >= want) {
                // squared gap to the nearest open side, rounded toward zero => a valid lower bound
                double gap = CUDART_INF;
                const double cell = (double)(1 << g.shift);
                if (cx - r > 0) gap = fmin(gap, __dsub_rd(qx, (double)(cx - r) * cell));
                if (cx + r < g.gw - 1) gap = fmin(gap, __dsub_rd((double)(cx + r + 1) * cell, qx));
                if (cy - r > 0) gap = fmin(gap, __dsub_rd(qy, (double)(cy - r) * cell));
                if (cy + r < g.gh - 1) gap = fmin(gap, __dsub_rd((double)(cy + r + 1) * cell, qy));
                if (gap == CUDART_INF) break;
                if (gap > 0.0 && kd2 < __dmul_rd(gap, gap)) break;
            }
        }
    }
    return found;
}
__device__ __forceinline__ int knn_warp(const GridView &g, double qx, double qy, int k, KnnEntry &mine) {
    return knn_warp_impl(g, qx, qy, k, mine, -1, 0);
}
__device__ __forceinline__ int knn_warp_wide(const GridView &g, double qx, double qy, int k, KnnEntry &mine) {
    return knn_warp_impl<true>(g, qx, qy, k, mine, -1, 0);
}

// ---------------------------------------------------------------------------------------------
// Fast path of the bulk kernels: radius filter + one warp sort.
// The 3x3 block of cells around the query is three contiguous ranges of the cell-sorted array; the warp reads
// them as one list (<= 32*KNN_FAST_CHUNKS candidates, d2 kept in registers), finds a threshold T with
// k <= |{d2 <= T}| <= 32 by a guarded secant iteration on the counts (first guess from the block's density),
// compacts the survivors to one per lane through shared memory and sorts them once (bitonic, 64-bit key).
// The k smallest of the block are the answer iff the k-th distance is strictly inside the block (the same
// certificate as the ring walk at r = 1); anything else — a tail of unsorted nodes, k > 28, a sparse or very dense
// block, many equal distances, an uncertified result — returns -1 and the caller runs knn_warp.
// Same result as knn_warp by construction: the k smallest (d2, id) of a superset of the certified disc.
// ---------------------------------------------------------------------------------------------
#define KNN_FAST_CHUNKS 10
struct KnnScratch { // per warp
    double d2[32];
    int id[32];
};
__device__ __forceinline__ int knn_warp_block3(const GridView &g, double qx, double qy, int k, KnnEntry &mine, KnnScratch &sm) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    if (g.n_sorted != g.n || k > 28 || g.n < k) return -1;
    const int cx = grid_cell_of(qx, g.shift, g.gw - 1), cy = grid_cell_of(qy, g.shift, g.gh - 1);
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.gw - 1), y0 = max(cy - 1, 0), y1 = min(cy + 1, g.gh - 1);
    // the three rows (a missing row is an empty range)
    const int ra = y0 * g.gw, rb = min(y0 + 1, y1) * g.gw, rc = y1 * g.gw;
    const int s0 = g.cell_start[ra + x0], e0 = g.cell_start[ra + x1 + 1];
    const int s1 = g.cell_start[rb + x0], e1 = (y1 - y0 >= 1) ? g.cell_start[rb + x1 + 1] : s1;
    const int s2 = g.cell_start[rc + x0], e2 = (y1 - y0 >= 2) ? g.cell_start[rc + x1 + 1] : s2;
    const int n0 = e0 - s0, n01 = n0 + (e1 - s1), L = n01 + (e2 - s2);
    if (L < k || L > 32 * KNN_FAST_CHUNKS) return -1;
    double d[KNN_FAST_CHUNKS];
    int id[KNN_FAST_CHUNKS];
#pragma unroll
    for (int c = 0; c < KNN_FAST_CHUNKS; c++) d[c] = CUDART_INF, id[c] = 0x7fffffff;
    {
        // lane-private cursor over the concatenated list: position of candidate j = lane, lane + 32, ...
        const int nch = (L + 31) >> 5;
#pragma unroll
        for (int c = 0; c < KNN_FAST_CHUNKS; c++) {
            if (c >= nch) break;
            const int j = 32 * c + lane;
            if (j < L) {
                const int pos = j < n0 ? s0 + j : (j < n01 ? s1 + (j - n0) : s2 + (j - n01));
                const float2 p = g.sxy[pos];
                id[c] = g.sid[pos];
                d[c] = knn_d2(qx, qy, p.x, p.y);
            }
        }
    }
    // threshold with k <= count <= 32
    const float tgt = 0.5f * (float)(k + 32);
    const float area = (float)((x1 - x0 + 1) * (y1 - y0 + 1)) * (float)(1 << g.shift) * (float)(1 << g.shift);
    double T = (double)(tgt * 0.31830987f * area / (float)L), lo = 0.0, hi = CUDART_INF;
    int cnt = 0, mycnt = 0;
    bool ok = false;
    for (int it = 0; it < 6; it++) {
        mycnt = 0;
#pragma unroll
        for (int c = 0; c < KNN_FAST_CHUNKS; c++) mycnt += d[c] <= T ? 1 : 0;
        cnt = __reduce_add_sync(full, mycnt);
        if (cnt >= k && cnt <= 32) {
            ok = true;
            break;
        }
        if (cnt < k) lo = T;
        else hi = T;
        double Tn = T * (double)(tgt / (float)max(cnt, 1));
        if (cnt == 0) Tn = T * 4.0;
        if (!(Tn > lo && Tn < hi)) Tn = hi == CUDART_INF ? lo * 2.0 : 0.5 * (lo + hi);
        T = Tn;
    }
    if (!ok) return -1;
    // compaction: one survivor per slot of the scratch list
    int base = mycnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(full, base, o);
        if (lane >= o) base += t;
    }
    base -= mycnt;
    __syncwarp();
#pragma unroll
    for (int c = 0; c < KNN_FAST_CHUNKS; c++)
        if (d[c] <= T) sm.d2[base] = d[c], sm.id[base] = id[c], base++;
    __syncwarp();
    // one bitonic sort across the lanes on a single 64-bit key: the distance bits (non-negative binary64 values order
    // like unsigned integers) with the low 5 bits replaced by the slot.  Exact unless two survivors agree in all but
    // those 5 bits (equal or nearly equal distances): then the (d2, id) order is not decided by the key -> slow path.
    unsigned long long key = ~0ull;
    if (lane < cnt) key = ((unsigned long long)__double_as_longlong(sm.d2[lane]) & ~31ull) | (unsigned)lane;
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1)
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const unsigned long long other = __shfl_xor_sync(full, key, stride);
            const bool keep_min = ((lane & stride) == 0) == ((lane & size) == 0);
            if ((other < key) == keep_min) key = other;
        }
    const unsigned long long prev = __shfl_up_sync(full, key, 1);
    const bool near_tie = lane > 0 && lane < cnt && ((prev ^ key) >> 5) == 0;
    if (__any_sync(full, near_tie)) return -1;
    mine.d2 = CUDART_INF, mine.id = 0x7fffffff;
    if (lane < k) mine.d2 = sm.d2[key & 31], mine.id = sm.id[key & 31];
    __syncwarp();
    // certificate: the k-th distance lies strictly inside the block (open sides only), as in knn_warp at r = 1
    const double kd2 = __shfl_sync(full, mine.d2, k - 1);
    double gap = CUDART_INF;
    const double cell = (double)(1 << g.shift);
    if (cx - 1 > 0) gap = fmin(gap, __dsub_rd(qx, (double)(cx - 1) * cell));
    if (cx + 1 < g.gw - 1) gap = fmin(gap, __dsub_rd((double)(cx + 2) * cell, qx));
    if (cy - 1 > 0) gap = fmin(gap, __dsub_rd(qy, (double)(cy - 1) * cell));
    if (cy + 1 < g.gh - 1) gap = fmin(gap, __dsub_rd((double)(cy + 2) * cell, qy));
    if (gap != CUDART_INF && !(gap > 0.0 && kd2 < __dmul_rd(gap, gap))) return -1;
    return k;
}

// ---------------------------------------------------------------------------------------------
// All-nodes table: consecutive cell-sorted queries share their cell, hence their 3x3 block.  The warp keeps the
// block's candidates (coordinates + ids) in shared memory and refills them only when the query cell changes
// (about once per 20 queries at C4), so a query costs no dependent global load beyond its own coordinates:
// distances from shared memory, then the same threshold / compaction / sort / certificate as knn_warp_block3.
// ---------------------------------------------------------------------------------------------
struct KnnBlockCache { // per warp
    double2 xy[32 * KNN_FAST_CHUNKS]; // candidate coordinates widened once per fill
    int id[32 * KNN_FAST_CHUNKS];
    double2 surv[32]; // survivors of the radius filter: (d2, candidate index as bits)
};
struct KnnBlockState { // warp-uniform registers
    int cx, cy, L;            // cached cell (-1: none) and its candidate count (-1: block unusable for the fast path)
    float T0;                 // first threshold guess: the radius^2 that holds (k + 32) / 2 points at the block's mean density
    float xlo, xhi, ylo, yhi; // the block's open sides (-inf / +inf where the block reaches the grid border); exact in binary32
};
// cell of a binary32 coordinate; equals grid_cell_of((double)v, ...) for every v
__device__ __forceinline__ int grid_cell_of_f32(float v, int shift, int gmax) {
    int c = (v > 0.f) ? (__float2int_rd(fminf(v, 1073741760.f)) >> shift) : 0;
    return min(c, gmax);
}
// (re)fills the cache for cell (cx, cy); warp-collective
__device__ __forceinline__ void knn_block_fill(const GridView &g, int cx, int cy, int k, KnnBlockCache &cache, KnnBlockState &st) {
    const int lane = threadIdx.x & 31;
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.gw - 1), y0 = max(cy - 1, 0), y1 = min(cy + 1, g.gh - 1);
    const int ra = y0 * g.gw, rb = min(y0 + 1, y1) * g.gw, rc = y1 * g.gw;
    const int s0 = g.cell_start[ra + x0], e0 = g.cell_start[ra + x1 + 1];
    const int s1 = g.cell_start[rb + x0], e1 = (y1 - y0 >= 1) ? g.cell_start[rb + x1 + 1] : s1;
    const int s2 = g.cell_start[rc + x0], e2 = (y1 - y0 >= 2) ? g.cell_start[rc + x1 + 1] : s2;
    const int n0 = e0 - s0, n01 = n0 + (e1 - s1), L = n01 + (e2 - s2);
    st.cx = cx, st.cy = cy;
    st.L = (L < k || L > 32 * KNN_FAST_CHUNKS || g.shift > 20) ? -1 : L;
    const float cell = (float)(1 << min(g.shift, 20));
    const float area = (float)((x1 - x0 + 1) * (y1 - y0 + 1)) * cell * cell;
    st.T0 = 0.5f * (float)(k + 32) * 0.31830987f * area / (float)max(L, 1);
    st.xlo = cx - 1 > 0 ? (float)(cx - 1) * cell : -CUDART_INF_F;
    st.xhi = cx + 1 < g.gw - 1 ? (float)(cx + 2) * cell : CUDART_INF_F;
    st.ylo = cy - 1 > 0 ? (float)(cy - 1) * cell : -CUDART_INF_F;
    st.yhi = cy + 1 < g.gh - 1 ? (float)(cy + 2) * cell : CUDART_INF_F;
    __syncwarp();
    if (st.L > 0)
        for (int j = lane; j < L; j += 32) {
            const int pos = j < n0 ? s0 + j : (j < n01 ? s1 + (j - n0) : s2 + (j - n01));
            const float2 p = g.sxy[pos];
            cache.xy[j] = make_double2((double)p.x, (double)p.y);
            cache.id[j] = g.sid[pos];
        }
    __syncwarp();
}
// the binary32 query must lie in the cached cell and st.L must be > 0; returns k or -1 (caller runs knn_warp)
__device__ __forceinline__ int knn_warp_block3_cached(float qxf, float qyf, int k, KnnEntry &mine, KnnBlockCache &cache, const KnnBlockState &st) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int L = st.L, nch = (L + 31) >> 5;
    const double qx = (double)qxf, qy = (double)qyf;
    double d[KNN_FAST_CHUNKS];
#pragma unroll
    for (int c = 0; c < KNN_FAST_CHUNKS; c++) d[c] = CUDART_INF;
#pragma unroll
    for (int c = 0; c < KNN_FAST_CHUNKS; c++) {
        if (c >= nch) break;
        const int j = 32 * c + lane;
        if (j < L) {
            const double2 p = cache.xy[j];
            const double dx = __dsub_rn(qx, p.x), dy = __dsub_rn(qy, p.y); // = knn_d2
            d[c] = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
        }
    }
    // threshold with k <= count <= 32 (guarded secant on the counts; the guesses are heuristics, binary32 is enough)
    const float tgt = 0.5f * (float)(k + 32);
    float Tf = st.T0, lo = 0.f, hi = CUDART_INF_F;
    double T = 0.0;
    int cnt = 0, mycnt = 0;
    bool ok = false;
    for (int it = 0; it < 6; it++) {
        T = (double)Tf;
        mycnt = 0;
#pragma unroll
        for (int c = 0; c < KNN_FAST_CHUNKS; c++) mycnt += d[c] <= T ? 1 : 0;
        cnt = __reduce_add_sync(full, mycnt);
        if (cnt >= k && cnt <= 32) {
            ok = true;
            break;
        }
        if (cnt < k) lo = Tf;
        else hi = Tf;
        float Tn = cnt == 0 ? Tf * 4.f : Tf * __fdividef(tgt, (float)cnt);
        if (!(Tn > lo && Tn < hi)) Tn = hi == CUDART_INF_F ? lo * 2.f : 0.5f * (lo + hi);
        Tf = Tn;
    }
    if (!ok) return -1;
    // compaction: survivor -> its own slot (exclusive prefix of the per-lane counts)
    int base = mycnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(full, base, o);
        if (lane >= o) base += t;
    }
    double2 *dst = cache.surv + (base - mycnt);
    __syncwarp();
#pragma unroll
    for (int c = 0; c < KNN_FAST_CHUNKS; c++) {
        if (c >= nch) break;
        if (d[c] <= T) {
            *dst = make_double2(d[c], __longlong_as_double((long long)(32 * c + lane)));
            dst++;
        }
    }
    __syncwarp();
    // one bitonic sort across the lanes on a single 64-bit key: the distance bits (non-negative binary64 values order
    // like unsigned integers) with the low 5 bits replaced by the slot.  Exact unless two survivors agree in all but
    // those 5 bits (equal or nearly equal distances): then the (d2, id) order is not decided by the key -> slow path.
    unsigned long long key = ~0ull;
    if (lane < cnt) key = ((unsigned long long)__double_as_longlong(cache.surv[lane].x) & ~31ull) | (unsigned)lane;
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1)
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const unsigned long long other = __shfl_xor_sync(full, key, stride);
            const bool keep_min = ((lane & stride) == 0) == ((lane & size) == 0);
            if ((other < key) == keep_min) key = other;
        }
    const unsigned long long prev = __shfl_up_sync(full, key, 1);
    const bool near_tie = lane > 0 && lane < cnt && ((prev ^ key) >> 5) == 0;
    if (__any_sync(full, near_tie)) return -1;
    mine.d2 = CUDART_INF, mine.id = 0x7fffffff;
    if (lane < k) {
        const double2 sv = cache.surv[key & 31];
        mine.d2 = sv.x, mine.id = cache.id[(int)__double_as_longlong(sv.y)];
    }
    __syncwarp();
    // certificate (sufficient, in binary32 rounded toward zero): the k-th distance lies strictly inside the block.
    // x - (-inf) = +inf, so closed sides drop out and a block that covers the grid certifies everything.
    const float kd2 = __double2float_ru(__shfl_sync(full, mine.d2, k - 1));
    const float gap = fminf(fminf(__fsub_rd(qxf, st.xlo), __fsub_rd(st.xhi, qxf)), fminf(__fsub_rd(qyf, st.ylo), __fsub_rd(st.yhi, qyf)));
    if (!(gap > 0.f && kd2 < __fmul_rd(gap, gap))) return -1;
    return k;
}
