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
        auto scan_row = [&](int y, int x0, int x1) {
            if (y < 0 || y >= g.gh) return;
            x0 = max(x0, 0), x1 = min(x1, g.gw - 1);
            if (x0 > x1) return;
            const int s = g.cell_start[y * g.gw + x0], e = g.cell_start[y * g.gw + x1 + 1];
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
        for (int r = 0; r <= rmax; r++) {
            if (r == 0)
                scan_row(cy, cx, cx);
            else {
                scan_row(cy - r, cx - r, cx + r);
                scan_row(cy + r, cx - r, cx + r);
                for (int y = cy - r + 1; y <= cy + r - 1; y++) {
                    scan_row(y, cx - r, cx - r);
                    scan_row(y, cx + r, cx + r);
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
