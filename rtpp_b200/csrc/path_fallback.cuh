// path_fallback.cuh — the second branch of GetPath (reference RTRRT.cpp:401-437, RTFMT.cpp:322-391): when no neighbour of
// the goal gives a collision-free entry, the reference walks the tree breadth-first from the root through collision-free
// children down to depth 5, keeps the node nearest the goal (strict <, so the FIRST in BFS order among equals) and returns
// its parent chain root..node — "head for the reachable node closest to the goal".  RT-RRT* and RT-FMT share it.
//
// On the device the set "within 5 collision-free levels of the root" is a property of each node's parent chain (<= 6 loads),
// so one warp scans the instance's nodes instead of keeping children sets current through every rewire.  BFS order, needed
// only among nodes at exactly the same binary32 distance, is the lexicographic order of the chains' points (the children
// of a node are a std::set<RTpoint>, ordered by x then y: RTRRT.h:36-42).
#pragma once
#include "nodes.cuh"

// hops from x up to the root through collision-free nodes; -1 when x is not within 5 levels (the root itself is level 0
// whatever its flag: it is pushed unconditionally, RTRRT.cpp:406)
__device__ __forceinline__ int fb_depth(const cpword *cp, int x, int root) {
    int y = x, d = 0;
    while (y != root) {
        if (y < 0 || d == 5) return -1;
        const cpword w = cp[y];
        if (cost_col(cp_hi(w))) return -1; // "if (!Node_Map[subPoint].cost.collision) Qpk.push"
        y = cp_parent(w);
        d++;
    }
    return d;
}

// true when a (level da) is dequeued before b (level db)
__device__ inline bool fb_before(const float2 *xy, const cpword *cp, int a, int da, int b, int db) {
    if (da != db) return da < db;
    int ca[5], cb[5];
    for (int i = da - 1, x = a, y = b; i >= 0; i--) ca[i] = x, cb[i] = y, x = cp_parent(cp[x]), y = cp_parent(cp[y]);
    for (int i = 0; i < da; i++) {
        if (ca[i] == cb[i]) continue;
        const float2 p = xy[ca[i]], q = xy[cb[i]];
        if (p.x != q.x) return p.x < q.x;
        if (p.y != q.y) return p.y < q.y;
        return ca[i] < cb[i];
    }
    return false;
}

// warp-collective.  Writes node ids x_min, parent(x_min), ..., root into path[] (the callers' read-back reverses them) and
// returns their count; 0 = no node qualified; -1 = the chain exceeded 2000 nodes (the reference clears the path).
__device__ inline int path_fallback_warp(const float2 *xy, const cpword *cp, int n, int root, float2 goal, int *path, int path_cap) {
    const int lane = threadIdx.x & 31;
    const uint32_t none = cost_pack(true, 3.3e38f); // c_min = {1, 3.3e38f}
    auto key_of = [&](int x) -> uint32_t {
        if (x >= n || fb_depth(cp, x, root) < 0) return 0xffffffffu;
        const float2 p = xy[x];
        return cost_pack(cost_col(cp_hi(cp[x])), rt_dist(p.x, p.y, goal.x, goal.y)); // {collision, x % goal}
    };
    uint32_t best = none;
    for (int b0 = 0; b0 < n; b0 += 32) best = min(best, key_of(b0 + lane));
    for (int off = 16; off; off >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, off));
    if (best >= none) return 0;
    int xm = -1, dm = 0;
    for (int b0 = 0; b0 < n; b0 += 32) {
        unsigned m = __ballot_sync(0xffffffffu, key_of(b0 + lane) == best);
        if (lane == 0)
            for (; m; m &= m - 1) {
                const int x = b0 + __ffs(m) - 1, d = fb_depth(cp, x, root);
                if (xm < 0 || fb_before(xy, cp, x, d, xm, dm)) xm = x, dm = d;
            }
    }
    int len = 0;
    if (lane == 0) {
        for (int x = xm; x >= 0; x = cp_parent(cp[x])) {
            if (len < path_cap) path[len] = x;
            if (++len > 2000) { // "if (path.size() > 2000) { path.clear(); return ...; }"
                len = -1;
                break;
            }
        }
    }
    return __shfl_sync(0xffffffffu, len, 0);
}
