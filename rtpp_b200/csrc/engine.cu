// engine.cu — sequential-exact planner engines: one warp runs one planner instance.
//
// The reference's replanning loops are sequential queues of tiny work items (pop a node, 25-NN,
// 25 edge checks, 25 relaxations: RTRRT.cpp:226-355).  Batching them changes intermediate states,
// so for drop-in parity the loops keep their order and the GPU parallelises *inside* a work item
// (lanes = the 25 neighbours / the k-NN scan / the ray march) and *across* planner instances
// (map x start/goal x seed scenarios: thousands of warps).  Results are bit-identical to the
// reference run under the same virtual clock and injected samples (tests/test_engine_gpu.py).
//
// RT-RRT* engine (planner_kind 1):
//   SetAgent/SetGoal   RTRRT.cpp:11-55        -> rtpp_engine_init_instances / rtpp_engine_set_agents
//   UpdateTree         RTRRT.cpp:92-136       -> update_tree()   (re-root + BFS re-cost, R3)
//   ExpandRewiring     RTRRT.cpp:226-355      -> expand_rewiring() (R1 choose-parent, R2 rewires)
//   GetPath            RTRRT.cpp:356-400      -> get_path()      (P1, goal-entry branch)
// The uniform/informed sampler (RTRRT.cpp:234-262) stays on the host (glibc transcendentals); the
// engine consumes pre-drawn sample points.  Wall-clock gates (RTRRT.cpp:228-230,300,329) become
// the virtual clock "k-th read returns k*dt" shared with the oracle.
#include <float.h>
#include <math_constants.h>

#include <algorithm>
#include <vector>

#include "nodes.cuh"

int rtpp_raster_into(rtpp_ctx *ctx, uint32_t *plane, int W, int H, const int32_t *circles, int n_circles, const int32_t *poly_xy,
                     const int32_t *poly_sizes, int n_polys); // raster.cu

#include "glibc_math.cuh"
#include "path_fallback.cuh"

#define FULL 0xffffffffu

struct EngInst {
    int map_slot;
    int n, n_sorted;
    int shift, gw, gh; // the instance's current hash resolution (chosen at every re-sort from the node count)
    int root;
    float2 agent, goal;
    float c_base;
    int qr_head, qr_tail, qs_head, qs_tail, qs_epoch;
    long long clk_n;
    int samp_pos, samp_n;
    long long iters, pops_r, pops_s;
    int status; // 0 ok; 1 node store full; 2 queue overflow; 3 RewireGoal needed (goal is a tree node)
    int path_len, path_col;
    float path_cost;
    // AM-RRT* (planner_kind 2): injected rand() stream position and the informed-sampler ellipse
    int rand_pos, rand_n, rand_exhausted, outpath_nonempty;
    int goal_node;                          // id of the tree node that IS the goal point (Node_Map.find(goal), AMRRT.cpp:606), -1 = none
    int gs_top, gq_head, gq_tail, g_epoch;  // RewireGoal: Sgoal (stack), Qgoal (FIFO), seen_goal (epoch marks)
    float ell_a, ell_b, ell_cos, ell_sin, ell_mx, ell_my;
    int ell_host; // 1: the caller supplied the ellipse for the next tick (rtpp_engine_set_am_inputs), 0: computed on the device
};

struct EngView {
    int n_inst, cap, qr_cap, max_samples, path_cap;
    int gw, gh, shift, ncells;
    EngInst *inst;
    float2 *xy, *sxy;
    cpword *cp;
    int *sid, *cell_of, *cell_start, *cursor;
    int *qr, *qs, *qs_mark;
    int *child_start, *children, *frontier;
    int *nn_ids, *nn_n; // cached 25-NN list of every tree node (ids) and the node count it was computed for (0 = none)
    int *nn_blk, *nn_blk_ep; // cached verdict bits of the edges node <-> its cached neighbours, and the plane epoch they hold for
    int plane_epoch[16];     // bumped whenever a plane of the map slot changes
    int *gs, *gq, *g_mark; // AM-RRT* RewireGoal: Sgoal stack, Qgoal queue, seen_goal marks [cap each]
    float2 *samples;
    int *path;
    PlaneView planes[16]; // merged S|D plane per map slot
    int divisor;
    int kind; // 1 RT-RRT*, 2 AM-RRT*
    // injected rand() stream (AM-RRT*: always; RT-RRT*: when set with rtpp_engine_set_rand, the samples are then drawn on the
    // device), the libm build whose sinf / cosf bits the samplers reproduce (glibc_math.cuh), and AM-RRT*'s diffusion-map table
    // (AMRRT.cpp:58-116)
    const int *rand_stream;
    int glibc_fma;
    int rand_cap;
    const int *diff_cell; // [diff_gw * diff_gh] row index or -1
    const float *diff_rows;
    int diff_gw, diff_gh, diff_dim;
    float diff_ratio;
};

struct rtpp_engine {
    rtpp_ctx *ctx;
    EngView v;
    uint32_t *plane_s[16], *plane_d[16], *plane_sd[16];
    int W[16], H[16];
    size_t plane_words[16];
    MarchThresholds mt;
    int *d_rand = nullptr, *d_diff_cell = nullptr;
    float *d_diff_rows = nullptr;
};

// ---------------------------------------------------------------------------------------------
// per-instance views
// ---------------------------------------------------------------------------------------------
struct Inst {
    EngInst s; // register copy, uniform across the warp
    float2 *xy, *sxy;
    cpword *cp;
    int *sid, *cell_of, *cell_start, *cursor, *qr, *qs, *qs_mark, *child_start, *children, *frontier, *nn_ids, *nn_n, *gs, *gq, *g_mark;
    int *nn_blk, *nn_blk_ep;
    int pl_epoch;
    const float2 *samples;
    int *path;
    PlaneView pl;
    const int *rand_stream; // injected rand() values of this instance (nullptr: RT-RRT* with injected samples)
    int glibc_fma;
};

__device__ __forceinline__ GridView inst_grid(const EngView &ev, const Inst &in) {
    GridView g;
    g.sxy = in.sxy, g.sid = in.sid, g.cell_start = in.cell_start, g.xy = in.xy;
    g.gw = in.s.gw, g.gh = in.s.gh, g.shift = in.s.shift, g.n_sorted = in.s.n_sorted, g.n = in.s.n;
    return g;
}

// counting sort of all nodes by cell, done by the instance's own warp
__device__ void grid_rebuild_warp(const EngView &ev, Inst &in) {
    const int lane = threadIdx.x & 31;
    const int n = in.s.n;
    // cell size for about two nodes per cell (a sparse tree on a fine grid makes every query walk hundreds of empty
    // cells; the result of a query does not depend on the resolution), never finer than the engine's cell_log2
    int shift = ev.shift;
    while (shift < 14 && (long long)(((ev.gw << ev.shift) - 1 >> (shift + 1)) + 1) * (((ev.gh << ev.shift) - 1 >> (shift + 1)) + 1) * 2 >= n) shift++;
    const int gw = (((ev.gw << ev.shift) - 1) >> shift) + 1, gh = (((ev.gh << ev.shift) - 1) >> shift) + 1, m = gw * gh;
    in.s.shift = shift, in.s.gw = gw, in.s.gh = gh;
    for (int c = lane; c <= m; c += 32) in.cursor[c] = 0;
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
        const float2 p = in.xy[i];
        const int c = grid_cell_of((double)p.y, shift, gh - 1) * gw + grid_cell_of((double)p.x, shift, gw - 1);
        in.cell_of[i] = c;
        atomicAdd(&in.cursor[c], 1);
    }
    __syncwarp();
    int carry = 0;
    for (int base = 0; base <= m; base += 32) {
        const int c = base + lane;
        const int v = c <= m ? in.cursor[c] : 0;
        int x = v;
        for (int off = 1; off < 32; off <<= 1) {
            const int y = __shfl_up_sync(FULL, x, off);
            if (lane >= off) x += y;
        }
        if (c <= m) {
            in.cell_start[c] = carry + x - v;
            in.cursor[c] = carry + x - v;
        }
        carry += __shfl_sync(FULL, x, 31);
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
        const int pos = atomicAdd(&in.cursor[in.cell_of[i]], 1);
        in.sxy[pos] = in.xy[i];
        in.sid[pos] = i;
    }
    __syncwarp();
    in.s.n_sorted = n;
}

__device__ __forceinline__ void maybe_rebuild(const EngView &ev, Inst &in) {
    // trees below 128 nodes are scanned linearly (4 chunks); afterwards the hash is re-sorted, at a resolution that
    // follows the node count, whenever the unsorted tail reaches max(64, n/8)
    const int tail = in.s.n - in.s.n_sorted;
    if (in.s.n_sorted == 0 ? tail >= 128 : (tail >= 64 && tail * 8 >= in.s.n_sorted)) grid_rebuild_warp(ev, in);
}

// 25-NN of tree node x (the query of every queue pop: RTRRT.cpp:304,333; AMRRT.cpp:496).  The list found at the last pop
// of x is kept (ids only: the squared distances are recomputed with the same operations) and extended by the nodes
// appended since — exact, because nodes never move or disappear — instead of searching the whole tree again.
__device__ __forceinline__ int knn_at_node(const EngView &ev, Inst &in, int x, KnnEntry &e) {
    const int lane = threadIdx.x & 31;
    const float2 px = in.xy[x];
    const int n0 = in.nn_n[x];
    int found;
    __syncwarp();
    if (n0 > 0 && in.s.n - n0 <= 512) {
        const int cnt = min(n0, 25);
        e.d2 = CUDART_INF, e.id = 0x7fffffff;
        if (lane < cnt) e.id = in.nn_ids[(size_t)x * 25 + lane];
        if (in.s.n == n0) return cnt; // nothing appended since: the callers use the ids only
        if (lane < cnt) {
            const float2 q = in.xy[e.id];
            e.d2 = knn_d2((double)px.x, (double)px.y, q.x, q.y);
        }
        const int old_id = e.id;
        found = knn_warp_impl(inst_grid(ev, in), (double)px.x, (double)px.y, 25, e, cnt, n0);
        const unsigned diff = __ballot_sync(FULL, e.id != old_id);
        if (diff == 0) { // no appended node entered the list: the cached edge verdicts stay valid
            if (lane == 0) in.nn_n[x] = in.s.n;
            __syncwarp();
            return found;
        }
        // ONE appended node entered a full list (the usual case: the root rewire pops every node between two accepted
        // samples): the old entries moved down by one behind it, and the verdict of the new edge (x, nid) is already
        // known if x sits in nid's own cached list — CrashDetection is symmetric in its end points (t and 1 - t are
        // both exact dyadic fractions, RTRRT.cpp:214-217) — so the cached bits are shifted instead of marching 25 edges.
        if (cnt == 25 && found == 25 && in.nn_blk_ep[x] == in.pl_epoch) {
            const int pos = __ffs(diff) - 1;
            const int prev_old = __shfl_up_sync(FULL, old_id, 1);
            if (__all_sync(FULL, lane <= pos || lane >= 25 || e.id == prev_old)) {
                const int nid = __shfl_sync(FULL, e.id, pos);
                if (in.nn_blk_ep[nid] == in.pl_epoch) {
                    const int cn = min(in.nn_n[nid], 25);
                    int other = -1;
                    if (lane < cn) other = in.nn_ids[(size_t)nid * 25 + lane];
                    const unsigned m = __ballot_sync(FULL, other == x);
                    if (m) {
                        const unsigned v = ((unsigned)in.nn_blk[nid] >> (__ffs(m) - 1)) & 1u;
                        const unsigned old = (unsigned)in.nn_blk[x];
                        const unsigned nb = ((old & ((1u << pos) - 1u)) | (v << pos) | ((old >> pos) << (pos + 1))) & 0x1ffffffu;
                        __syncwarp();
                        if (lane < found) in.nn_ids[(size_t)x * 25 + lane] = e.id;
                        if (lane == 0) in.nn_n[x] = in.s.n, in.nn_blk[x] = (int)nb;
                        __syncwarp();
                        return found;
                    }
                }
            }
        }
    } else
        found = knn_warp(inst_grid(ev, in), (double)px.x, (double)px.y, 25, e);
    if (lane < found) in.nn_ids[(size_t)x * 25 + lane] = e.id;
    if (lane == 0) in.nn_n[x] = in.s.n, in.nn_blk_ep[x] = -1;
    __syncwarp();
    return found;
}

// Verdicts of the edges between tree node x and its cached neighbours (lane j = j-th neighbour), from_x = true for
// CrashDetection(x, neighbour) (RT-RRT*, RTRRT.cpp:310,339), false for CrashDetection(neighbour, x) (AM-RRT*,
// AMRRT.cpp:500,556; an engine is one kind, so the cached bits have one orientation).  A verdict is a pure function of
// the planes and the two points: it is kept per node until a plane of the map slot changes (epoch) or the node's
// neighbour list does — a node is popped many times per tick, so most pops find their 25 verdicts ready.
__device__ __forceinline__ bool edges_at_node(const EngView &ev, Inst &in, const MarchThresholds &mt, int x, const KnnEntry &e, int found,
                                              bool from_x) {
    const int lane = threadIdx.x & 31;
    if (in.nn_blk_ep[x] == in.pl_epoch) return (in.nn_blk[x] >> lane) & 1;
    bool blk = false;
    if (lane < found) {
        const float2 px = in.xy[x], q = in.xy[e.id];
        blk = from_x ? rt_edge_blocked(in.pl, px.x, px.y, q.x, q.y, mt) : rt_edge_blocked(in.pl, q.x, q.y, px.x, px.y, mt);
    }
    const unsigned bits = __ballot_sync(FULL, blk);
    if (lane == 0) in.nn_blk[x] = (int)bits, in.nn_blk_ep[x] = in.pl_epoch;
    __syncwarp();
    return blk;
}

// virtual clock: the k-th read returns k*dt (orc::VClock)
__device__ __forceinline__ long long clk_now(Inst &in, long long dt) { return dt * (in.s.clk_n++); }

// ---------------------------------------------------------------------------------------------
// UpdateTree (RTRRT.cpp:92-136)
// ---------------------------------------------------------------------------------------------
__device__ void update_tree(const EngView &ev, Inst &in, const MarchThresholds &mt) {
    const int lane = threadIdx.x & 31;
    const int n = in.s.n;
    const float2 rootp = in.xy[in.s.root];
    if (in.s.agent.x != rootp.x || in.s.agent.y != rootp.y) {
        KnnEntry e;
        knn_warp(inst_grid(ev, in), (double)in.s.agent.x, (double)in.s.agent.y, 1, e);
        const int near = __shfl_sync(FULL, e.id, 0);
        if (near != in.s.root) { // reverse the parent chain near -> ... -> root
            if (lane == 0) {
                int x = near, x_old = -1, guard = 0;
                while (x != in.s.root && x >= 0 && guard++ < n) {
                    const cpword w = in.cp[x];
                    in.cp[x] = cp_make(cp_hi(w), x_old);
                    x_old = x;
                    x = cp_parent(w);
                }
                if (x >= 0) in.cp[x] = cp_make(cp_hi(in.cp[x]), x_old);
                in.cp[near] = cp_make(cost_pack(false, 0.0f), -1);
            }
            __syncwarp();
            in.s.root = near;
        }
    }
    // children lists (counting sort on the parent id)
    for (int i = lane; i <= n; i += 32) in.child_start[i] = 0;
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
        const int p = cp_parent(in.cp[i]);
        if (i != in.s.root && p >= 0) atomicAdd(&in.child_start[p], 1);
    }
    __syncwarp();
    int carry = 0;
    for (int base = 0; base <= n; base += 32) {
        const int c = base + lane;
        const int v = c <= n ? in.child_start[c] : 0;
        int x = v;
        for (int off = 1; off < 32; off <<= 1) {
            const int y = __shfl_up_sync(FULL, x, off);
            if (lane >= off) x += y;
        }
        if (c <= n) in.child_start[c] = carry + x - v;
        carry += __shfl_sync(FULL, x, 31);
    }
    __syncwarp();
    int *fill = in.cell_of; // per-parent fill counters (scratch; rebuilt by grid_rebuild_warp when needed)
    for (int i = lane; i < n; i += 32) fill[i] = 0;
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
        const int p = cp_parent(in.cp[i]);
        if (i != in.s.root && p >= 0) in.children[in.child_start[p] + atomicAdd(&fill[p], 1)] = i;
    }
    __syncwarp();
    // BFS by levels: cost[child] = cost[parent] + {E(parent,child), |parent-child|}
    int *cur = in.frontier, *nxt = in.frontier + ev.cap;
    if (lane == 0) cur[0] = in.s.root;
    int cur_n = 1;
    __syncwarp();
    while (cur_n > 0) {
        int nxt_n = 0;
        for (int f0 = 0; f0 < cur_n; f0 += 32) {
            const int f = f0 + lane;
            int s = 0, e = 0, p = -1;
            uint32_t hp = 0;
            float2 pp = make_float2(0.f, 0.f);
            if (f < cur_n) {
                p = cur[f];
                s = in.child_start[p], e = in.child_start[p + 1];
                hp = cp_hi(in.cp[p]);
                pp = in.xy[p];
            }
            // exclusive prefix of the child counts gives each lane its slot range in the next frontier
            const int cnt = e - s;
            int x = cnt;
            for (int off = 1; off < 32; off <<= 1) {
                const int y = __shfl_up_sync(FULL, x, off);
                if (lane >= off) x += y;
            }
            const int slot = nxt_n + x - cnt;
            for (int j = 0; j < cnt; j++) {
                const int c = in.children[s + j];
                const float2 pc = in.xy[c];
                const uint32_t w = cost_pack(rt_edge_blocked(in.pl, pp.x, pp.y, pc.x, pc.y, mt), rt_dist(pp.x, pp.y, pc.x, pc.y));
                in.cp[c] = cp_make(cost_add(hp, w), p);
                nxt[slot + j] = c;
            }
            nxt_n += __shfl_sync(FULL, x, 31);
        }
        __syncwarp();
        int *t = cur;
        cur = nxt, nxt = t;
        cur_n = nxt_n;
    }
}

// ---------------------------------------------------------------------------------------------
// ExpandRewiring (RTRRT.cpp:226-355)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void qr_push(const EngView &ev, Inst &in, bool want, int id) {
    const int lane = threadIdx.x & 31;
    const unsigned m = __ballot_sync(FULL, want);
    if (want) in.qr[(unsigned)(in.s.qr_tail + __popc(m & ((1u << lane) - 1))) % (unsigned)ev.qr_cap] = id;
    in.s.qr_tail += __popc(m);
    if (in.s.qr_tail - in.s.qr_head > ev.qr_cap) in.s.status = 2;
    __syncwarp();
}

// One sample: RTRRT.cpp:263-298 (R1 choose-parent)
__device__ void expand_one(const EngView &ev, Inst &in, const MarchThresholds &mt, float2 p) {
    const int lane = threadIdx.x & 31;
    in.s.iters++;
    KnnEntry e;
    const int found = knn_warp(inst_grid(ev, in), (double)p.x, (double)p.y, 25, e);
    const bool have = lane < found;
    float2 q = make_float2(0.f, 0.f);
    uint32_t hs = 0;
    bool blk = false;
    if (have) {
        q = in.xy[e.id];
        hs = cp_hi(in.cp[e.id]);
        blk = rt_edge_blocked(in.pl, p.x, p.y, q.x, q.y, mt); // CrashDetection(randpoint, x_near)
    }
    const bool blocked0 = __shfl_sync(FULL, blk, 0);
    if (blocked0) return;
    const int closest = __shfl_sync(FULL, e.id, 0);
    const double d2_0 = __shfl_sync(FULL, e.d2, 0);
    if (__fsqrt_rn(__double2float_rn(d2_0)) > 8.0f) { // RnearR, RTRRT.cpp:270
        // c_new = cost(x_near) + {E, |rand - x_near|}; strict <, first neighbour wins ties
        unsigned long long key = ~0ull;
        if (have) key = ((unsigned long long)cost_add(hs, cost_pack(blk, rt_dist(p.x, p.y, q.x, q.y))) << 32) | (unsigned)lane;
        for (int off = 16; off; off >>= 1) {
            const unsigned long long o = __shfl_xor_sync(FULL, key, off);
            key = o < key ? o : key;
        }
        const int jmin = (int)(key & 31u);
        const int x_min = __shfl_sync(FULL, e.id, jmin);
        if (in.s.n >= ev.cap) {
            in.s.status = 1;
            return;
        }
        const int id = in.s.n;
        if (lane == 0) {
            in.xy[id] = p;
            in.cp[id] = cp_make((uint32_t)(key >> 32), x_min);
        }
        in.s.n = id + 1;
        // the new node's own neighbour list and edge verdicts are known already: itself (d2 = 0; the nearest other node is
        // more than RnearR away), then the neighbours of the sample in the same (d2, id) order; the self edge is free (the
        // sample's pixel was probed with every edge above) and CrashDetection(sample, x_near) = CrashDetection(x_near,
        // sample).  Its first pop then needs neither a search nor a march, and its neighbours can take the verdict of
        // their new edge from here (knn_at_node).
        {
            const unsigned bits = __ballot_sync(FULL, blk);
            if (lane == 0) in.nn_ids[(size_t)id * 25] = id;
            if (have && lane < 24) in.nn_ids[(size_t)id * 25 + lane + 1] = e.id;
            if (lane == 0) in.nn_n[id] = id + 1, in.nn_blk[id] = (int)((bits << 1) & 0x1ffffffu), in.nn_blk_ep[id] = in.pl_epoch;
        }
        __syncwarp();
        qr_push(ev, in, lane == 0, id);
        maybe_rebuild(ev, in);
    } else
        qr_push(ev, in, lane == 0, closest);
}

// relax the <=25 neighbours of x (RTRRT.cpp:304-319 / 333-346); returns per-lane "improved"
__device__ __forceinline__ bool relax_neighbours(const EngView &ev, Inst &in, const MarchThresholds &mt, int x, KnnEntry &e,
                                                 int &found) {
    const int lane = threadIdx.x & 31;
    const float2 px = in.xy[x];
    found = knn_at_node(ev, in, x, e);
    const bool blk = edges_at_node(ev, in, mt, x, e, found, true);
    bool improved = false;
    if (lane < found) {
        const float2 q = in.xy[e.id];
        const uint32_t w = cost_pack(blk, rt_dist(px.x, px.y, q.x, q.y));
        const uint32_t c_new = cost_add(cp_hi(in.cp[x]), w);
        if (c_new < cp_hi(in.cp[e.id])) {
            in.cp[e.id] = cp_make(c_new, x);
            improved = true;
        }
    }
    __syncwarp();
    return improved;
}

// glibc's cosf / sinf of one angle, in the build (fused or not) of the libm the host runs (glibc_math.cuh)
__device__ __forceinline__ float2 glibc_cos_sin(float th, int fma) {
    return fma ? make_float2(glibc_math::cosf_<true>(th), glibc_math::sinf_<true>(th)) : make_float2(glibc_math::cosf_<false>(th), glibc_math::sinf_<false>(th));
}
// "random_theta = rand() / float(RAND_MAX) * 2 * 3.141592954f" (RTRRT.cpp:248, AMRRT.cpp:423): (cosf, sinf) of it for raw draw r
__device__ __forceinline__ float2 theta_cos_sin(int r, int fma) {
    const float u = __fdiv_rn((float)r, 2147483648.0f);
    return glibc_cos_sin(__fmul_rn(__fmul_rn(u, 2.0f), 3.141592954f), fma);
}
// the informed sampler's ellipse (RTRRT.cpp:243-246, AMRRT.cpp:418-421) from c_base, root and goal: binary32 throughout, glibc's
// atan2f / cosf / sinf.  {a, b, cos(angle), sin(angle), mid.x, mid.y}
__device__ __forceinline__ void informed_ellipse(Inst &in) {
    const float2 root = in.xy[in.s.root], goal = in.s.goal;
    const float c_min = rt_dist(goal.x, goal.y, root.x, root.y), cb = in.s.c_base;
    in.s.ell_a = __fdiv_rn(cb, 2.0f);
    in.s.ell_b = __fdiv_rn(__fsqrt_rn(__fsub_rn(__fmul_rn(cb, cb), __fmul_rn(c_min, c_min))), 2.0f);
    const float angle = glibc_math::atan2f_(__fsub_rn(goal.y, root.y), __fsub_rn(goal.x, root.x));
    const float2 cs = glibc_cos_sin(angle, in.glibc_fma);
    in.s.ell_cos = cs.x, in.s.ell_sin = cs.y;
    in.s.ell_mx = __fdiv_rn(__fadd_rn(goal.x, root.x), 2.0f), in.s.ell_my = __fdiv_rn(__fadd_rn(goal.y, root.y), 2.0f);
}

// SamplePoint of RT-RRT* (RTRRT.cpp:234-262) from the injected rand() stream, on the device: one draw decides between the uniform
// sample and the informed ellipse around (root, goal) of the path cost c_base the last GetPath found; glibc's atan2f / cosf / sinf
// bits (glibc_math.cuh).  Defined behaviour where the reference loops forever (c_base < |goal - root|: b = sqrt(negative)):
// uniform, as the oracle and the CPU planners do.  Returns false when the stream ran out inside the sample (the position is
// restored: the caller refills the stream from there).
__device__ bool rt_sample_point(Inst &in, float2 &out) {
    const int pos0 = in.s.rand_pos;
    const float W = (float)in.pl.W, H = (float)in.pl.H;
    auto u = [&]() -> float {
        if (in.s.rand_pos >= in.s.rand_n) {
            in.s.rand_exhausted = 1;
            return 0.0f;
        }
        return __fdiv_rn((float)in.rand_stream[in.s.rand_pos++], 2147483648.0f);
    };
    const float p0 = u();
    const float2 root = in.xy[in.s.root], goal = in.s.goal;
    const float cb = in.s.c_base, c_min = rt_dist(goal.x, goal.y, root.x, root.y);
    float rx, ry;
    if (p0 < 0.1f || cb < 0.0f || cb < c_min) {
        rx = __fmul_rn(u(), W);
        ry = __fmul_rn(u(), H);
    } else {
        informed_ellipse(in);
        for (int guard = 0;; guard++) {
            const int r_th = in.s.rand_pos < in.s.rand_n ? in.rand_stream[in.s.rand_pos] : 0;
            u();
            const float2 cs = theta_cos_sin(r_th, in.glibc_fma);
            const float rr = __fsqrt_rn(u());
            const float x = __fmul_rn(__fmul_rn(in.s.ell_a, cs.x), rr), y = __fmul_rn(__fmul_rn(in.s.ell_b, cs.y), rr);
            rx = __fadd_rn(__fsub_rn(__fmul_rn(in.s.ell_cos, x), __fmul_rn(in.s.ell_sin, y)), in.s.ell_mx);
            ry = __fadd_rn(__fadd_rn(__fmul_rn(in.s.ell_sin, x), __fmul_rn(in.s.ell_cos, y)), in.s.ell_my);
            if (rx >= 0.0f && rx < W && ry >= 0.0f && ry < H) break;
            if (in.s.rand_exhausted || guard >= 100000) break;
        }
    }
    if (in.s.rand_exhausted) {
        in.s.rand_pos = pos0;
        return false;
    }
    out = make_float2(rx, ry);
    return true;
}

__device__ void expand_rewiring(const EngView &ev, Inst &in, const MarchThresholds &mt, long long run_ns, long long dt) {
    const int lane = threadIdx.x & 31;
    const long long t0 = clk_now(in, dt);
    while (clk_now(in, dt) - t0 < run_ns) {
        if (in.s.status) break;
        float2 p;
        if (in.rand_stream) { // samples drawn here from the injected rand() stream (rtpp_engine_set_rand)
            if (in.s.rand_exhausted || !rt_sample_point(in, p)) break;
            in.s.samp_pos++;
        } else { // samples injected by the caller (rtpp_engine_set_inputs)
            if (in.s.samp_pos >= in.s.samp_n) break;
            p = in.samples[in.s.samp_pos++];
        }
        expand_one(ev, in, mt, p);
        // rewire from the random queue
        while (in.s.qr_tail != in.s.qr_head && (clk_now(in, dt) - t0 < run_ns)) {
            const int xr = in.qr[(unsigned)in.s.qr_head % (unsigned)ev.qr_cap];
            in.s.qr_head++;
            in.s.pops_r++;
            KnnEntry e;
            int found;
            const bool imp = relax_neighbours(ev, in, mt, xr, e, found);
            qr_push(ev, in, imp, e.id);
        }
        // rewire from the root
        if (in.s.qs_tail == in.s.qs_head) {
            in.s.qs_epoch++;
            in.s.qs_head = in.s.qs_tail = 0;
            if (lane == 0) {
                in.qs[0] = in.s.root;
                in.qs_mark[in.s.root] = in.s.qs_epoch;
            }
            in.s.qs_tail = 1;
            __syncwarp();
        }
        while (in.s.qs_tail != in.s.qs_head && (clk_now(in, dt) - t0 < run_ns)) {
            const int xs = in.qs[in.s.qs_head];
            in.s.qs_head++;
            in.s.pops_s++;
            KnnEntry e;
            int found;
            relax_neighbours(ev, in, mt, xs, e, found);
            const bool fresh = lane < found && in.qs_mark[e.id] != in.s.qs_epoch;
            const unsigned m = __ballot_sync(FULL, fresh);
            if (fresh) {
                in.qs[in.s.qs_tail + __popc(m & ((1u << lane) - 1))] = e.id;
                in.qs_mark[e.id] = in.s.qs_epoch;
            }
            in.s.qs_tail += __popc(m);
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// GetPath (RTRRT.cpp:356-400): direct goal-root shortcut, else best entry among the 25-NN of the
// goal, then the parent walk accumulating the binary32 length c_base.  The depth-5 fallback of
// :401-437 (no collision-free entry) is path_fallback_warp: path root..nearest reachable node, cost {1, 3.4e38}.
// ---------------------------------------------------------------------------------------------
__device__ void get_path(const EngView &ev, Inst &in, const MarchThresholds &mt) {
    const int lane = threadIdx.x & 31;
    in.s.c_base = -1.0f;
    in.s.path_len = 0, in.s.path_col = 1, in.s.path_cost = 3.4e38f;
    if (!(in.s.goal.x > 0.0f)) return; // "goal.x <= 0"
    const float2 g = in.s.goal, r = in.xy[in.s.root];
    if (!rt_edge_blocked_warp(in.pl, g.x, g.y, r.x, r.y, mt)) { // CrashDetection(goal, root)
        if (lane == 0) in.path[0] = in.s.root, in.path[1] = -2; // -2 = the goal point itself
        in.s.path_len = 2, in.s.path_col = 0, in.s.path_cost = rt_dist(r.x, r.y, g.x, g.y);
        return;
    }
    KnnEntry e;
    const int found = knn_warp(inst_grid(ev, in), (double)g.x, (double)g.y, 25, e);
    unsigned long long key = ((unsigned long long)cost_pack(true, FLT_MAX) << 32) | 63u; // c_min = {1, FLT_MAX}
    if (lane < found) {
        const float2 q = in.xy[e.id];
        const uint32_t w = cost_pack(rt_edge_blocked(in.pl, g.x, g.y, q.x, q.y, mt), rt_dist(g.x, g.y, q.x, q.y));
        const uint32_t c = cost_add(cp_hi(in.cp[e.id]), w);
        if (c < cost_pack(true, FLT_MAX)) key = ((unsigned long long)c << 32) | (unsigned)lane;
    }
    for (int off = 16; off; off >>= 1) {
        const unsigned long long o = __shfl_xor_sync(FULL, key, off);
        key = o < key ? o : key;
    }
    const uint32_t cmin = (uint32_t)(key >> 32);
    if (cmin < cost_pack(true, 0.0f)) { // c_min < {1, 0}: a collision-free entry exists
        const int x_min = __shfl_sync(FULL, e.id, (int)(key & 31u));
        int len = 0, bad = 0;
        float cb = 0.0f;
        if (lane == 0) {
            // path.push_front(goal); walk parents, c_base += |x_i - path.front()|
            float2 front = g;
            int x = x_min;
            in.path[len++] = -2;
            while (x >= 0) {
                const float2 px = in.xy[x];
                cb = __fadd_rn(cb, rt_dist(px.x, px.y, front.x, front.y));
                if (len < ev.path_cap) in.path[len] = x;
                len++;
                front = px;
                x = cp_parent(in.cp[x]);
                if (len > 2000) {
                    bad = 1;
                    break;
                }
            }
        }
        len = __shfl_sync(FULL, len, 0), bad = __shfl_sync(FULL, bad, 0), cb = __shfl_sync(FULL, cb, 0);
        if (bad) {
            in.s.path_len = 0, in.s.path_col = 1, in.s.path_cost = 2.4e38f;
            in.s.c_base = cb; // the reference leaves the partial sum in c_base (RTRRT.cpp:389-396)
        } else {
            in.s.path_len = len, in.s.path_col = 0, in.s.path_cost = cb, in.s.c_base = cb;
        }
    } else { // RTRRT.cpp:401-437: no collision-free entry — the chain of the reachable node nearest the goal, cost {1, 3.4e38}
        const int len = path_fallback_warp(in.xy, in.cp, in.s.n, in.s.root, g, in.path, ev.path_cap);
        if (len < 0) in.s.path_cost = 2.4e38f;
        else in.s.path_len = len; // path_col stays 1: rtpp_engine_read_path emits root..node, no goal point
    }
}


// =============================================================================================
// AM-RRT* (planner_kind 2): AMRRT/lib/AMrrt/AMRRT.cpp
//   Expand :395-481 (sampler, Steer :354-390, choose-parent), RewireRoot :483-519, MainTask loop
//   :598-613, DiffusivityMetric :94-116, CrashDetectionStart :287-308, GetPath :310-351.
// rand() is consumed on the device from the injected stream (its consumption depends on the tree);
// cosf32/sinf32 of the drawn angles are host-tabulated per stream position so that the glibc
// results are used bit for bit.  RewireGoal (:521-585) runs only while the goal point itself is a
// collision-free tree node — Expand cannot put it there (the goal-bias branch :402-403 never breaks), a caller
// who sets the goal onto an existing node can.
// =============================================================================================
__device__ __forceinline__ int am_rand(Inst &in) { // orc::RandStream::next
    if (in.s.rand_pos >= in.s.rand_n) {
        in.s.rand_exhausted = 1;
        return 0;
    }
    return in.rand_stream[in.s.rand_pos++];
}
__device__ __forceinline__ float am_u(Inst &in) { return __fdiv_rn((float)am_rand(in), 2147483648.0f); }
// draws theta and returns the (cosf, sinf) of it
__device__ __forceinline__ float2 am_theta(Inst &in) {
    const int pos = in.s.rand_pos;
    const int r = am_rand(in);
    return pos < in.s.rand_n ? theta_cos_sin(r, in.glibc_fma) : make_float2(1.0f, 0.0f);
}
// DiffusivityMetric (AMRRT.cpp:94-116)
__device__ float am_diffusivity(const EngView &ev, float ax, float ay, float bx, float by) {
    const int is = __float2int_rz(__fmul_rn(ax, ev.diff_ratio)), js = __float2int_rz(__fmul_rn(ay, ev.diff_ratio));
    if ((unsigned)is >= (unsigned)ev.diff_gw || (unsigned)js >= (unsigned)ev.diff_gh) return 3.4e38f;
    const int ra = ev.diff_cell[js * ev.diff_gw + is];
    if (ra < 0) return 3.4e38f;
    const int ie = __float2int_rz(__fmul_rn(bx, ev.diff_ratio)), je = __float2int_rz(__fmul_rn(by, ev.diff_ratio));
    if ((unsigned)ie >= (unsigned)ev.diff_gw || (unsigned)je >= (unsigned)ev.diff_gh) return 3.4e38f;
    const int rb = ev.diff_cell[je * ev.diff_gw + ie];
    if (rb < 0) return 3.4e38f;
    const float *pa = ev.diff_rows + (size_t)ra * ev.diff_dim, *pb = ev.diff_rows + (size_t)rb * ev.diff_dim;
    float len = 0.0f;
    for (int i = 0; i < ev.diff_dim; i++) {
        const float d = __fsub_rn(pa[i], pb[i]);
        len = __fadd_rn(len, __fmul_rn(d, d));
    }
    return __fdiv_rn(__fsqrt_rn(len), ev.diff_ratio);
}

// CrashDetectionStart (AMRRT.cpp:287-308) for one edge by the whole warp
__device__ bool am_first_hit_warp(const PlaneView &pl, float2 ps, float2 pe, float2 &pstart) {
    const int lane = threadIdx.x & 31;
    pstart = ps;
    if (rt_probe(pl, __float2int_rz(ps.y), __float2int_rz(ps.x)) | rt_probe(pl, __float2int_rz(pe.y), __float2int_rz(pe.x))) return true;
    const int M = rt_march_M(rt_dist(ps.x, ps.y, pe.x, pe.y), 2.0f);
    const float Mf = (float)M;
    auto point = [&](int i) {
        const float t = __fdiv_rn((float)i, Mf), omt = __fsub_rn(1.0f, t);
        return make_float2(__fadd_rn(__fmul_rn(pe.x, t), __fmul_rn(omt, ps.x)), __fadd_rn(__fmul_rn(pe.y, t), __fmul_rn(omt, ps.y)));
    };
    for (int i0 = 1; i0 < M; i0 += 32) {
        const int i = i0 + lane;
        bool hit = false;
        if (i < M) {
            const float2 q = point(i);
            hit = rt_probe(pl, __float2int_rz(q.y), __float2int_rz(q.x));
        }
        const unsigned hm = __ballot_sync(FULL, hit);
        if (hm) {
            const int f = i0 + __ffs(hm) - 1;
            if (f > 1) pstart = point(f - 1);
            return true;
        }
    }
    pstart = pe;
    return false;
}

// Steer (AMRRT.cpp:354-390)
__device__ float2 am_steer(const EngView &ev, Inst &in, const MarchThresholds &mt, float2 xi, float2 xe, long long dt) {
    const int lane = threadIdx.x & 31;
    float2 x_start;
    if (!am_first_hit_warp(in.pl, xi, xe, x_start)) {
        const float len = rt_dist(xi.x, xi.y, xe.x, xe.y);
        const float p = __fdiv_rn(100.0f, len);
        if (p < 1.0f) {
            const float omp = __fsub_rn(1.0f, p);
            return make_float2(__fadd_rn(__fmul_rn(omp, xi.x), __fmul_rn(p, xe.x)), __fadd_rn(__fmul_rn(omp, xi.y), __fmul_rn(p, xe.y)));
        }
        return xe;
    }
    float2 x_min = x_start;
    float c_min = am_diffusivity(ev, x_start.x, x_start.y, xe.x, xe.y);
    const float r_sam = fminf(100.0f, rt_dist(xi.x, xi.y, xe.x, xe.y));
    // the random local search runs while (now() - t0 < 5 ms): count the iterations first, then
    // evaluate the candidates 32 at a time (they are independent; only the running minimum is
    // sequential, and "strictly smaller, not blocked" = first minimum among the unblocked ones)
    const long long t0 = clk_now(in, dt);
    int K = 0;
    while (clk_now(in, dt) - t0 < 5000000LL) K++;
    for (int k0 = 0; k0 < K; k0 += 32) {
        const int k = k0 + lane;
        float2 cand = make_float2(0.f, 0.f);
        float c_new = CUDART_INF_F;
        if (k < K) {
            const int pos = in.s.rand_pos + 2 * k;
            float2 cs = make_float2(1.0f, 0.0f);
            float rr = 0.0f;
            if (pos < in.s.rand_n) cs = theta_cos_sin(in.rand_stream[pos], in.glibc_fma);
            if (pos + 1 < in.s.rand_n) rr = __fdiv_rn((float)in.rand_stream[pos + 1], 2147483648.0f);
            rr = __fmul_rn(__fsqrt_rn(rr), r_sam);
            cand = make_float2(__fadd_rn(__fmul_rn(cs.x, rr), xi.x), __fadd_rn(__fmul_rn(cs.y, rr), xi.y));
            c_new = am_diffusivity(ev, cand.x, cand.y, xe.x, xe.y);
        }
        bool ok = k < K && c_new < c_min;
        if (ok) ok = !rt_edge_blocked(in.pl, xi.x, xi.y, cand.x, cand.y, mt);
        unsigned long long key = ok ? (((unsigned long long)__float_as_uint(c_new) << 32) | (unsigned)lane) : ~0ull;
        for (int off = 16; off; off >>= 1) {
            const unsigned long long o = __shfl_xor_sync(FULL, key, off);
            key = o < key ? o : key;
        }
        if (key != ~0ull) {
            const int j = (int)(key & 31u);
            c_min = __uint_as_float((uint32_t)(key >> 32));
            x_min = make_float2(__shfl_sync(FULL, cand.x, j), __shfl_sync(FULL, cand.y, j));
        }
    }
    const int used = 2 * K;
    if (in.s.rand_pos + used > in.s.rand_n) {
        in.s.rand_exhausted = 1;
        in.s.rand_pos = in.s.rand_n;
    } else
        in.s.rand_pos += used;
    return x_min;
}

// Expand (AMRRT.cpp:395-481)
__device__ void am_expand(const EngView &ev, Inst &in, const MarchThresholds &mt, long long dt) {
    const int lane = threadIdx.x & 31;
    in.s.iters++;
    float2 x_rand = make_float2(0.f, 0.f);
    const float W = (float)in.pl.W, H = (float)in.pl.H;
    while (!in.s.rand_exhausted) { // SampleState
        const float p = am_u(in);
        if (p > 0.9f && !in.s.outpath_nonempty && in.s.goal.x >= 0.0f) {
            x_rand = in.s.goal; // no break in the reference: draws again
        } else if (p < (0.9f / 4.0f) || !in.s.outpath_nonempty) {
            const float rx = __fmul_rn(am_u(in), W), ry = __fmul_rn(am_u(in), H);
            if (!rt_probe(in.pl, __float2int_rz(ry), __float2int_rz(rx))) {
                x_rand = make_float2(rx, ry);
                break;
            }
        } else if (in.s.goal.x >= 0.0f) {
            const float2 cs = am_theta(in); // theta draw: only its cosf/sinf are needed
            const float rr = __fsqrt_rn(am_u(in));
            const float x = __fmul_rn(__fmul_rn(in.s.ell_a, cs.x), rr), y = __fmul_rn(__fmul_rn(in.s.ell_b, cs.y), rr);
            float rx = __fsub_rn(__fmul_rn(in.s.ell_cos, x), __fmul_rn(in.s.ell_sin, y));
            float ry = __fadd_rn(__fmul_rn(in.s.ell_sin, x), __fmul_rn(in.s.ell_cos, y));
            rx = __fadd_rn(rx, in.s.ell_mx), ry = __fadd_rn(ry, in.s.ell_my);
            if (rx >= 0.0f && rx < W && ry >= 0.0f && ry < H && !rt_probe(in.pl, __float2int_rz(ry), __float2int_rz(rx))) {
                x_rand = make_float2(rx, ry);
                break;
            }
        }
    }
    if (in.s.rand_exhausted) return;
    KnnEntry e;
    knn_warp(inst_grid(ev, in), (double)x_rand.x, (double)x_rand.y, 25, e);
    const int x_nrst = __shfl_sync(FULL, e.id, 0);
    const float2 p_nrst = in.xy[x_nrst];
    const float2 x_new = am_steer(ev, in, mt, p_nrst, x_rand, dt);
    const int found = knn_warp(inst_grid(ev, in), (double)x_new.x, (double)x_new.y, 25, e);
    const int nearest = __shfl_sync(FULL, e.id, 0);
    const float2 p_near = in.xy[nearest];
    if (!(rt_dist(p_near.x, p_near.y, x_new.x, x_new.y) > 4.0f)) return; // k_max
    // c_min starts at x_nrst WITHOUT an edge check (:455-456); neighbours must be strictly better and free
    const uint32_t c0 = cost_add(cp_hi(in.cp[x_nrst]), cost_pack(false, rt_dist(p_nrst.x, p_nrst.y, x_new.x, x_new.y)));
    unsigned long long key = ((unsigned long long)c0 << 32) | 63u;
    if (lane < found) {
        const float2 q = in.xy[e.id];
        const uint32_t w = cost_pack(rt_edge_blocked(in.pl, q.x, q.y, x_new.x, x_new.y, mt), rt_dist(q.x, q.y, x_new.x, x_new.y));
        const uint32_t c = cost_add(cp_hi(in.cp[e.id]), w);
        if (c < c0 && !cost_col(c)) key = ((unsigned long long)c << 32) | (unsigned)lane;
    }
    for (int off = 16; off; off >>= 1) {
        const unsigned long long o = __shfl_xor_sync(FULL, key, off);
        key = o < key ? o : key;
    }
    const uint32_t c_min = (uint32_t)(key >> 32);
    if (cost_col(c_min)) return;
    const int x_min = (key & 63u) == 63u ? x_nrst : __shfl_sync(FULL, e.id, (int)(key & 31u));
    if (in.s.n >= ev.cap) {
        in.s.status = 1;
        return;
    }
    const int id = in.s.n;
    if (lane == 0) {
        in.xy[id] = x_new;
        in.cp[id] = cp_make(c_min, x_min);
    }
    in.s.n = id + 1;
    __syncwarp();
    maybe_rebuild(ev, in);
    if (x_new.x == in.s.goal.x && x_new.y == in.s.goal.y) in.s.goal_node = id; // Node_Map.find(goal) now succeeds (AMRRT.cpp:606)
}

// RewireRoot (AMRRT.cpp:483-519): Qroot / rewired_root reuse the Qs arrays and epoch marks
__device__ void am_rewire_root(const EngView &ev, Inst &in, const MarchThresholds &mt, long long dt) {
    const int lane = threadIdx.x & 31;
    if (in.s.qs_tail == in.s.qs_head) {
        in.s.qs_epoch++;
        in.s.qs_head = 0, in.s.qs_tail = 1;
        if (lane == 0) {
            in.qs[0] = in.s.root;
            in.qs_mark[in.s.root] = in.s.qs_epoch;
        }
        __syncwarp();
    }
    const long long t0 = clk_now(in, dt);
    while (clk_now(in, dt) - t0 < 20000000LL && in.s.qs_tail != in.s.qs_head) {
        const int xr = in.qs[in.s.qs_head];
        in.s.qs_head++;
        in.s.pops_s++;
        const float2 pr = in.xy[xr];
        KnnEntry e;
        const int found = knn_at_node(ev, in, xr, e);
        const bool blk = edges_at_node(ev, in, mt, xr, e, found, false); // CrashDetection(x_near, x_r)
        bool free_edge = false;
        if (lane < found) {
            const float2 q = in.xy[e.id];
            free_edge = !blk;
            if (free_edge) {
                const uint32_t c_new = cost_add(cp_hi(in.cp[xr]), cost_pack(false, rt_dist(q.x, q.y, pr.x, pr.y)));
                if (c_new < cp_hi(in.cp[e.id]) && !cost_col(c_new)) in.cp[e.id] = cp_make(c_new, xr);
            }
        }
        __syncwarp();
        const bool fresh = free_edge && in.qs_mark[e.id] != in.s.qs_epoch;
        const unsigned m = __ballot_sync(FULL, fresh);
        if (fresh) {
            in.qs[in.s.qs_tail + __popc(m & ((1u << lane) - 1))] = e.id;
            in.qs_mark[e.id] = in.s.qs_epoch;
        }
        in.s.qs_tail += __popc(m);
        __syncwarp();
    }
}

// RewireGoal (AMRRT.cpp:521-585): depth-first from the root through the informed ellipse, each node's 25 neighbours
// visited in ascending DiffusivityMetric(goal, .) order (std::priority_queue of (float, RTpoint) pairs: ties by point).
// Sgoal is a stack (push_front / pop_front), Qgoal a FIFO, seen_goal an epoch mark.  Runs only while the goal point
// itself is a collision-free tree node.
__device__ void am_rewire_goal(const EngView &ev, Inst &in, const MarchThresholds &mt, long long dt) {
    const int lane = threadIdx.x & 31;
    if (in.s.gq_head == in.s.gq_tail && in.s.gs_top == 0) {
        in.s.g_epoch++; // seen_goal.clear()
        in.s.gq_head = in.s.gq_tail = 0;
        if (lane == 0) in.gs[0] = in.s.root;
        in.s.gs_top = 1;
        __syncwarp();
    }
    const float2 g = in.s.goal, rt = in.xy[in.s.root];
    const long long t0 = clk_now(in, dt);
    while (clk_now(in, dt) - t0 < 100000000LL && (in.s.gs_top > 0 || in.s.gq_head != in.s.gq_tail)) {
        int xr;
        if (in.s.gs_top > 0) xr = in.gs[--in.s.gs_top];
        else xr = in.gq[in.s.gq_head++];
        in.s.pops_r++; // (AM-RRT* has no Qr: the slot counts RewireGoal pops)
        const float2 pr = in.xy[xr];
        if (__fadd_rn(rt_dist(rt.x, rt.y, pr.x, pr.y), rt_dist(g.x, g.y, pr.x, pr.y)) <= in.s.c_base) {
            KnnEntry e;
            const int found = knn_at_node(ev, in, xr, e);
            const bool blk = edges_at_node(ev, in, mt, xr, e, found, false); // CrashDetection(x_near, x_r)
            float dk = 0.f;
            float2 q = make_float2(0.f, 0.f);
            bool free_edge = false;
            if (lane < found) {
                q = in.xy[e.id];
                dk = am_diffusivity(ev, g.x, g.y, q.x, q.y);
                free_edge = !blk;
                if (free_edge) {
                    const uint32_t c_new = cost_add(cp_hi(in.cp[xr]), cost_pack(false, rt_dist(q.x, q.y, pr.x, pr.y)));
                    if (c_new < cp_hi(in.cp[e.id]) && !cost_col(c_new)) in.cp[e.id] = cp_make(c_new, xr);
                }
            }
            __syncwarp();
            const bool fresh = free_edge && in.g_mark[e.id] != in.s.g_epoch;
            // position of a fresh neighbour in X_next = number of fresh neighbours that sort before it
            int before = 0;
            for (int m = 0; m < found; m++) {
                const float odk = __shfl_sync(FULL, dk, m), ox = __shfl_sync(FULL, q.x, m), oy = __shfl_sync(FULL, q.y, m);
                const bool of = __shfl_sync(FULL, fresh, m);
                const bool less = odk < dk || (odk == dk && (ox != q.x ? ox < q.x : oy < q.y));
                before += of && less;
            }
            const int cnt = __popc(__ballot_sync(FULL, fresh));
            if (fresh) {
                in.g_mark[e.id] = in.s.g_epoch;
                in.gs[in.s.gs_top + before] = e.id; // push_front in X_next order: the last one ends on top
                in.gq[in.s.gq_tail + before] = e.id;
            }
            in.s.gs_top += cnt, in.s.gq_tail += cnt;
            __syncwarp();
        }
        if (in.s.gs_top > 0) {
            const float2 f = in.xy[in.gs[in.s.gs_top - 1]];
            if (am_diffusivity(ev, f.x, f.y, g.x, g.y) > am_diffusivity(ev, pr.x, pr.y, g.x, g.y)) in.s.gs_top = 0; // Sgoal.clear()
        }
    }
}

// MainTask loop (AMRRT.cpp:598-613) after UpdateTree
__device__ void am_main_loop(const EngView &ev, Inst &in, const MarchThresholds &mt, long long dt) {
    // Node_Map.find(goal): the goal may have been moved onto (or off) a tree node since the last tick
    {
        const int lane = threadIdx.x & 31;
        int hit = -1;
        if (in.s.goal.x >= 0.0f)
            for (int i = lane; i < in.s.n; i += 32) {
                const float2 p = in.xy[i];
                if (p.x == in.s.goal.x && p.y == in.s.goal.y) hit = i;
            }
        in.s.goal_node = __reduce_max_sync(FULL, hit);
    }
    // the informed sampler's constants for this MainTask (AMRRT.cpp:418-421), from the root UpdateTree has just chosen
    if (!in.s.ell_host) informed_ellipse(in);
    in.s.ell_host = 0;
    const long long t0 = clk_now(in, dt);
    while (clk_now(in, dt) - t0 < 80000000LL && !in.s.rand_exhausted && !in.s.status) {
        am_expand(ev, in, mt, dt);
        am_rewire_root(ev, in, mt, dt);
        if (in.s.goal.x >= 0.0f && in.s.goal_node >= 0 && !cost_col(cp_hi(in.cp[in.s.goal_node]))) am_rewire_goal(ev, in, mt, dt);
    }
}

// GetPath (AMRRT.cpp:310-351): like RT-RRT*'s goal-entry branch, without the direct shortcut
__device__ void am_get_path(const EngView &ev, Inst &in, const MarchThresholds &mt) {
    const int lane = threadIdx.x & 31;
    in.s.c_base = -1.0f;
    in.s.path_len = 0, in.s.path_col = 1, in.s.path_cost = 3.4e38f;
    if (!(in.s.goal.x > 0.0f)) return;
    const float2 g = in.s.goal;
    KnnEntry e;
    const int found = knn_warp(inst_grid(ev, in), (double)g.x, (double)g.y, 25, e);
    unsigned long long key = ((unsigned long long)cost_pack(true, FLT_MAX) << 32) | 63u;
    if (lane < found) {
        const float2 q = in.xy[e.id];
        const uint32_t w = cost_pack(rt_edge_blocked(in.pl, g.x, g.y, q.x, q.y, mt), rt_dist(g.x, g.y, q.x, q.y));
        const uint32_t c = cost_add(cp_hi(in.cp[e.id]), w);
        if (c < cost_pack(true, FLT_MAX)) key = ((unsigned long long)c << 32) | (unsigned)lane;
    }
    for (int off = 16; off; off >>= 1) {
        const unsigned long long o = __shfl_xor_sync(FULL, key, off);
        key = o < key ? o : key;
    }
    if ((uint32_t)(key >> 32) < cost_pack(true, 0.0f)) {
        const int x_min = __shfl_sync(FULL, e.id, (int)(key & 31u));
        int len = 0, bad = 0;
        float cb = 0.0f;
        if (lane == 0) {
            float2 front = g;
            int x = x_min;
            in.path[len++] = -2;
            while (x >= 0) {
                const float2 px = in.xy[x];
                cb = __fadd_rn(cb, rt_dist(px.x, px.y, front.x, front.y));
                if (len < ev.path_cap) in.path[len] = x;
                len++;
                front = px;
                x = cp_parent(in.cp[x]);
                if (len > 2000) {
                    bad = 1;
                    break;
                }
            }
        }
        len = __shfl_sync(FULL, len, 0), bad = __shfl_sync(FULL, bad, 0), cb = __shfl_sync(FULL, cb, 0);
        if (bad) in.s.path_len = 0, in.s.path_col = 1, in.s.path_cost = 2.4e38f, in.s.c_base = -1.0f;
        else in.s.path_len = len, in.s.path_col = 0, in.s.path_cost = cb, in.s.c_base = cb;
    }
}

// ---------------------------------------------------------------------------------------------
// kernels: one warp per instance
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool load_inst(const EngView &ev, int w, Inst &in) {
    if (w >= ev.n_inst) return false;
    in.s = ev.inst[w];
    const size_t c = (size_t)w * ev.cap, g = (size_t)w * (ev.ncells + 1);
    in.xy = ev.xy + c, in.sxy = ev.sxy + c, in.cp = ev.cp + c, in.sid = ev.sid + c, in.cell_of = ev.cell_of + c;
    in.cell_start = ev.cell_start + g, in.cursor = ev.cursor + g;
    in.qr = ev.qr + (size_t)w * ev.qr_cap, in.qs = ev.qs + c, in.qs_mark = ev.qs_mark + c;
    in.child_start = ev.child_start + (size_t)w * (ev.cap + 1), in.children = ev.children + c;
    in.frontier = ev.frontier + 2 * c;
    in.nn_ids = ev.nn_ids + 25 * c, in.nn_n = ev.nn_n + c;
    in.nn_blk = ev.nn_blk + c, in.nn_blk_ep = ev.nn_blk_ep + c;
    in.gs = ev.gs ? ev.gs + c : nullptr, in.gq = ev.gq ? ev.gq + c : nullptr, in.g_mark = ev.g_mark ? ev.g_mark + c : nullptr;
    in.samples = ev.samples + (size_t)w * ev.max_samples;
    in.path = ev.path + (size_t)w * ev.path_cap;
    in.pl = ev.planes[in.s.map_slot & 15];
    in.pl_epoch = ev.plane_epoch[in.s.map_slot & 15];
    in.rand_stream = ev.rand_stream ? ev.rand_stream + (size_t)w * ev.rand_cap : nullptr;
    in.glibc_fma = ev.glibc_fma;
    return true;
}

// one instantiation per planner kind: the register allocation of one planner's loop is not paid by the other
template <int KIND>
__global__ void __launch_bounds__(128, KIND == 1 ? 4 : 1) engine_tick_kernel(EngView ev, MarchThresholds mt, int do_update, long long run_ns,
                                                                          long long dt, int do_path) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    Inst in;
    if (!load_inst(ev, w, in)) return;
    if (in.s.n > 0) {
        if (do_update) update_tree(ev, in, mt);
        if (KIND == 2) {
            if (run_ns > 0) am_main_loop(ev, in, mt, dt);
            if (do_path) am_get_path(ev, in, mt);
        } else {
            if (run_ns > 0) expand_rewiring(ev, in, mt, run_ns, dt);
            if (do_path) get_path(ev, in, mt);
        }
    }
    if ((threadIdx.x & 31) == 0) ev.inst[w] = in.s;
}

// SetAgent for fresh instances (RTRRT.cpp:23-31): the first agent position becomes the root node
__global__ void engine_init_kernel(EngView ev, const int *__restrict__ slots, const float2 *__restrict__ agent,
                                   const float2 *__restrict__ goal) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= ev.n_inst) return;
    EngInst s;
    memset(&s, 0, sizeof(s));
    s.map_slot = slots[w];
    s.agent = agent[w], s.goal = goal[w];
    s.c_base = -1.0f;
    s.n = 1, s.n_sorted = 0, s.root = 0;
    s.shift = ev.shift, s.gw = ev.gw, s.gh = ev.gh;
    s.goal_node = -1;
    s.path_col = 1, s.path_cost = 3.4e38f;
    ev.xy[(size_t)w * ev.cap] = agent[w];
    ev.cp[(size_t)w * ev.cap] = cp_make(cost_pack(false, 0.0f), -1);
    ev.inst[w] = s;
}

__global__ void engine_set_agents_kernel(EngView ev, const float2 *__restrict__ agent, const float2 *__restrict__ goal,
                                         const int *__restrict__ samp_n) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= ev.n_inst) return;
    if (agent) ev.inst[w].agent = agent[w];
    if (goal) ev.inst[w].goal = goal[w];
    if (samp_n) ev.inst[w].samp_n = samp_n[w], ev.inst[w].samp_pos = 0;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
#define ENG_CK(e, call) RTPP_CK((e)->ctx, call)

template <class T> static int eng_alloc(rtpp_engine *e, T **p, size_t count) {
    ENG_CK(e, cudaMalloc((void **)p, count * sizeof(T)));
    ENG_CK(e, cudaMemsetAsync(*p, 0, count * sizeof(T), e->ctx->stream));
    return RTPP_OK;
}

// host staging for the read-back helpers: malloc, not std::string — a bad_alloc must not unwind through the C ABI
struct HostBuf {
    void *p = nullptr;
    explicit HostBuf(size_t bytes) { p = malloc(bytes ? bytes : 1); }
    ~HostBuf() { free(p); }
    HostBuf(const HostBuf &) = delete;
    HostBuf &operator=(const HostBuf &) = delete;
};

extern "C" {

int rtpp_engine_create(rtpp_ctx *ctx, int planner_kind, int n_instances, int node_capacity, int cell_log2, int domain_w,
                       int domain_h, int max_samples, rtpp_engine **out) {
    if (!ctx || !out) return RTPP_ERR_INVALID;
    *out = nullptr;
    if (planner_kind != 1 && planner_kind != 2) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "planner_kind: 1 = RT-RRT*, 2 = AM-RRT*");
    if (n_instances < 1 || node_capacity < 32 || cell_log2 < 2 || cell_log2 > 10 || domain_w < 1 || domain_h < 1 || max_samples < 1)
        RTPP_FAIL(ctx, RTPP_ERR_INVALID, "bad engine geometry");
    RTPP_CK(ctx, cudaSetDevice(ctx->device));
    rtpp_engine *e = new (std::nothrow) rtpp_engine();
    if (!e) return RTPP_ERR_NOMEM;
    memset(&e->v, 0, sizeof(e->v));
    for (int i = 0; i < 16; i++) e->plane_s[i] = e->plane_d[i] = e->plane_sd[i] = nullptr, e->W[i] = e->H[i] = 0, e->plane_words[i] = 0;
    e->ctx = ctx;
    EngView &v = e->v;
    v.n_inst = n_instances, v.cap = node_capacity, v.qr_cap = 8 * node_capacity, v.max_samples = max_samples, v.path_cap = 2048;
    v.shift = cell_log2;
    v.gw = ((domain_w - 1) >> cell_log2) + 1, v.gh = ((domain_h - 1) >> cell_log2) + 1, v.ncells = v.gw * v.gh;
    v.kind = planner_kind;
    v.glibc_fma = rtpp_glibc_host_fma();
    v.divisor = planner_kind == 2 ? 2 : 5; // AMRRT.cpp:273 vs RTRRT.cpp:209
    rtpp_march_thresholds(v.divisor, &e->mt);
    const size_t N = (size_t)n_instances, C = N * node_capacity, G = N * (v.ncells + 1);
    int rc = 0;
    rc |= eng_alloc(e, &v.inst, N);
    rc |= eng_alloc(e, &v.xy, C);
    rc |= eng_alloc(e, &v.sxy, C);
    rc |= eng_alloc(e, &v.cp, C);
    rc |= eng_alloc(e, &v.sid, C);
    rc |= eng_alloc(e, &v.cell_of, C);
    rc |= eng_alloc(e, &v.cell_start, G);
    rc |= eng_alloc(e, &v.cursor, G);
    rc |= eng_alloc(e, &v.qr, N * v.qr_cap);
    rc |= eng_alloc(e, &v.qs, C);
    rc |= eng_alloc(e, &v.qs_mark, C);
    rc |= eng_alloc(e, &v.child_start, N * (node_capacity + 1));
    rc |= eng_alloc(e, &v.children, C);
    rc |= eng_alloc(e, &v.frontier, 2 * C);
    rc |= eng_alloc(e, &v.nn_ids, 25 * C);
    rc |= eng_alloc(e, &v.nn_n, C);
    rc |= eng_alloc(e, &v.nn_blk, C), rc |= eng_alloc(e, &v.nn_blk_ep, C);
    if (planner_kind == 2) rc |= eng_alloc(e, &v.gs, C), rc |= eng_alloc(e, &v.gq, C), rc |= eng_alloc(e, &v.g_mark, C);
    rc |= eng_alloc(e, &v.samples, planner_kind == 1 ? N * max_samples : 1);
    if (planner_kind == 2) { // max_samples = rand() draws available per instance and call
        v.rand_cap = max_samples;
        rc |= eng_alloc(e, &e->d_rand, N * max_samples);
        v.rand_stream = e->d_rand;
    }
    rc |= eng_alloc(e, &v.path, N * v.path_cap);
    if (rc) {
        rtpp_engine_destroy(e); // frees what was allocated (cudaFree(nullptr) is a no-op)
        RTPP_FAIL(ctx, RTPP_ERR_NOMEM, "engine buffers do not fit the device");
    }
    *out = e;
    return RTPP_OK;
}

int rtpp_engine_destroy(rtpp_engine *e) {
    if (!e) return RTPP_ERR_INVALID;
    cudaStreamSynchronize(e->ctx->stream);
    EngView &v = e->v;
    cudaFree(v.inst), cudaFree(v.xy), cudaFree(v.sxy), cudaFree(v.cp), cudaFree(v.sid), cudaFree(v.cell_of);
    cudaFree(v.cell_start), cudaFree(v.cursor), cudaFree(v.qr), cudaFree(v.qs), cudaFree(v.qs_mark);
    cudaFree(v.child_start), cudaFree(v.children), cudaFree(v.frontier), cudaFree(v.samples), cudaFree(v.path);
    cudaFree(v.nn_ids), cudaFree(v.nn_n), cudaFree(v.nn_blk), cudaFree(v.nn_blk_ep), cudaFree(v.gs), cudaFree(v.gq), cudaFree(v.g_mark);
    for (int i = 0; i < 16; i++) cudaFree(e->plane_s[i]), cudaFree(e->plane_d[i]), cudaFree(e->plane_sd[i]);
    cudaFree(e->d_rand), cudaFree(e->d_diff_cell), cudaFree(e->d_diff_rows);
    delete e;
    return RTPP_OK;
}

// static (plane = 1) or dynamic (plane = 2; src NULL clears it) byte plane of a map slot
int rtpp_engine_set_map(rtpp_engine *e, int slot, int plane, const uint8_t *src, int W, int H) {
    if (!e || slot < 0 || slot >= 16 || (plane != 1 && plane != 2)) return RTPP_ERR_INVALID;
    rtpp_ctx *ctx = e->ctx;
    if (plane == 1 && !src) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "static plane needs data");
    if (plane == 2 && !e->plane_s[slot]) RTPP_FAIL(ctx, RTPP_ERR_STATE, "set the static plane of the slot first");
    if (plane == 2) W = e->W[slot], H = e->H[slot];
    if (W < 1 || H < 1 || (long long)W * H >= (1ll << 31)) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "plane shape: 1 <= W, H and W*H < 2^31 (as rtpp_map_set_plane)");
    const int tilesX = (W + 31) >> 5, tilesY = (H + 31) >> 5;
    const size_t words = (size_t)tilesX * tilesY * 32;
    if (plane == 1) {
        cudaFree(e->plane_s[slot]), cudaFree(e->plane_d[slot]), cudaFree(e->plane_sd[slot]);
        e->plane_s[slot] = e->plane_d[slot] = e->plane_sd[slot] = nullptr;
        ENG_CK(e, cudaMalloc(&e->plane_s[slot], words * 4));
        ENG_CK(e, cudaMalloc(&e->plane_d[slot], words * 4));
        ENG_CK(e, cudaMalloc(&e->plane_sd[slot], words * 4));
        ENG_CK(e, cudaMemsetAsync(e->plane_d[slot], 0, words * 4, ctx->stream));
        e->W[slot] = W, e->H[slot] = H, e->plane_words[slot] = words;
    }
    uint32_t *dst = plane == 1 ? e->plane_s[slot] : e->plane_d[slot];
    if (src) {
        void *d = nullptr;
        int rc = rtpp_scratch(ctx, 0, (size_t)W * H, &d);
        if (rc) return rc;
        ENG_CK(e, cudaMemcpyAsync(d, src, (size_t)W * H, cudaMemcpyHostToDevice, ctx->stream));
        rc = rtpp_plane_pack(ctx, (const uint8_t *)d, W, H, dst);
        if (rc) return rc;
    } else
        ENG_CK(e, cudaMemsetAsync(dst, 0, words * 4, ctx->stream));
    {
        const int rcm = rtpp_plane_merge(ctx, e->plane_s[slot], e->plane_d[slot], e->plane_sd[slot], words);
        if (rcm) return rcm;
    }
    ENG_CK(e, cudaStreamSynchronize(ctx->stream));
    e->v.plane_epoch[slot]++; // cached edge verdicts of this slot's instances are stale now
    e->v.planes[slot].w = e->plane_sd[slot];
    e->v.planes[slot].W = W, e->v.planes[slot].H = H, e->v.planes[slot].tilesX = tilesX;
    return RTPP_OK;
}

// DrawDObsmap of a map slot on the device (RTRRT.cpp:138-169, AMRRT.cpp:202-233): clears the slot's dynamic plane, draws the
// circles ((cx, cy, r) int triples) and polygons straight into its bit tiles and refreshes S|D — a tick ships the obstacle
// poses instead of an H*W byte plane.  Stream-ordered, no synchronisation.
int rtpp_engine_draw_dynamic(rtpp_engine *e, int slot, const int32_t *circles, int n_circles, const int32_t *poly_xy,
                             const int32_t *poly_sizes, int n_polys) {
    if (!e || slot < 0 || slot >= 16 || n_circles < 0 || n_polys < 0 || (n_circles && !circles) || (n_polys && (!poly_xy || !poly_sizes)))
        return RTPP_ERR_INVALID;
    rtpp_ctx *ctx = e->ctx;
    if (!e->plane_s[slot]) RTPP_FAIL(ctx, RTPP_ERR_STATE, "set the static plane of the slot first");
    const int W = e->W[slot], H = e->H[slot];
    const size_t words = e->plane_words[slot];
    int rc = rtpp_raster_into(ctx, e->plane_d[slot], W, H, circles, n_circles, poly_xy, poly_sizes, n_polys); // clears first
    if (rc) return rc;
    {
        const int rcm = rtpp_plane_merge(ctx, e->plane_s[slot], e->plane_d[slot], e->plane_sd[slot], words);
        if (rcm) return rcm;
    }
    e->v.plane_epoch[slot]++; // cached edge verdicts of this slot's instances are stale now
    return RTPP_OK;
}

static int upload(rtpp_engine *e, int slot, const void *src, size_t bytes, void **d) {
    *d = nullptr;
    if (!src) return RTPP_OK;
    int rc = rtpp_scratch(e->ctx, slot, bytes, d);
    if (rc) return rc;
    ENG_CK(e, cudaMemcpyAsync(*d, src, bytes, cudaMemcpyHostToDevice, e->ctx->stream));
    return RTPP_OK;
}

// SetAgent on empty planners (creates the root, RTRRT.cpp:23-31) + SetGoal; goal (-1,-1) = none
int rtpp_engine_init_instances(rtpp_engine *e, const int32_t *map_slot, const float *agent_xy, const float *goal_xy) {
    if (!e || !map_slot || !agent_xy || !goal_xy) return RTPP_ERR_INVALID;
    rtpp_ctx *ctx = e->ctx;
    const int n = e->v.n_inst;
    for (int i = 0; i < n; i++)
        if (map_slot[i] < 0 || map_slot[i] >= 16 || !e->plane_s[map_slot[i]])
            RTPP_FAIL(ctx, RTPP_ERR_INVALID, "map_slot: 0..15, and the slot's static plane must have been set (rtpp_engine_set_map)");
    void *ds, *da, *dg;
    int rc = upload(e, 0, map_slot, (size_t)n * 4, &ds);
    rc |= upload(e, 1, agent_xy, (size_t)n * 8, &da);
    rc |= upload(e, 2, goal_xy, (size_t)n * 8, &dg);
    if (rc) return rc;
    ENG_CK(e, cudaMemsetAsync(e->v.nn_n, 0, (size_t)n * e->v.cap * sizeof(int), ctx->stream)); // no cached neighbour lists
    engine_init_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(e->v, (const int *)ds, (const float2 *)da, (const float2 *)dg);
    RTPP_LAUNCHED(ctx);
    ENG_CK(e, cudaStreamSynchronize(ctx->stream));
    return RTPP_OK;
}

// per-tick inputs: agent positions (SetAgent on a non-empty planner just moves the agent,
// RTRRT.cpp:32-35), goals, and the pre-drawn sample points for the coming ExpandRewiring call
// (samples_xy: n_inst rows of max_samples points; n_samples: points valid per row).  Any pointer
// may be NULL = unchanged.
int rtpp_engine_set_inputs(rtpp_engine *e, const float *agent_xy, const float *goal_xy, const float *samples_xy,
                           const int32_t *n_samples) {
    if (!e) return RTPP_ERR_INVALID;
    rtpp_ctx *ctx = e->ctx;
    const int n = e->v.n_inst;
    if (n_samples)
        for (int i = 0; i < n; i++)
            if (n_samples[i] < 0 || n_samples[i] > e->v.max_samples)
                RTPP_FAIL(ctx, RTPP_ERR_INVALID, "n_samples: 0 .. max_samples of rtpp_engine_create");
    void *da, *dg, *dn;
    int rc = upload(e, 0, agent_xy, (size_t)n * 8, &da);
    rc |= upload(e, 1, goal_xy, (size_t)n * 8, &dg);
    rc |= upload(e, 2, n_samples, (size_t)n * 4, &dn);
    if (rc) return rc;
    if (samples_xy)
        ENG_CK(e, cudaMemcpyAsync(e->v.samples, samples_xy, (size_t)n * e->v.max_samples * 8, cudaMemcpyHostToDevice, ctx->stream));
    engine_set_agents_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(e->v, (const float2 *)da, (const float2 *)dg, (const int *)dn);
    RTPP_LAUNCHED(ctx);
    ENG_CK(e, cudaStreamSynchronize(ctx->stream));
    return RTPP_OK;
}

// AM-RRT*: InitializeDiffusion (AMRRT.cpp:58-93): cells (i,j) -> row of `matrix` (ncells x dim floats)
int rtpp_engine_set_diffusion(rtpp_engine *e, const int32_t *cells_ij, int ncells, const float *matrix, int dim, float ratio) {
    if (!e || !cells_ij || !matrix || ncells < 1 || dim < 1) return RTPP_ERR_INVALID;
    rtpp_ctx *ctx = e->ctx;
    int gw = 1, gh = 1;
    for (int i = 0; i < ncells; i++) {
        if (cells_ij[2 * i] < 0 || cells_ij[2 * i + 1] < 0) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "negative diffusion cell");
        gw = std::max(gw, cells_ij[2 * i] + 1), gh = std::max(gh, cells_ij[2 * i + 1] + 1);
    }
    std::vector<int> grid((size_t)gw * gh, -1);
    for (int i = 0; i < ncells; i++) grid[(size_t)cells_ij[2 * i + 1] * gw + cells_ij[2 * i]] = i; // later rows win, like std::map[]=
    cudaFree(e->d_diff_cell), cudaFree(e->d_diff_rows);
    ENG_CK(e, cudaMalloc(&e->d_diff_cell, grid.size() * 4));
    ENG_CK(e, cudaMalloc(&e->d_diff_rows, (size_t)ncells * dim * 4));
    ENG_CK(e, cudaMemcpyAsync(e->d_diff_cell, grid.data(), grid.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    ENG_CK(e, cudaMemcpyAsync(e->d_diff_rows, matrix, (size_t)ncells * dim * 4, cudaMemcpyHostToDevice, ctx->stream));
    ENG_CK(e, cudaStreamSynchronize(ctx->stream));
    e->v.diff_cell = e->d_diff_cell, e->v.diff_rows = e->d_diff_rows;
    e->v.diff_gw = gw, e->v.diff_gh = gh, e->v.diff_dim = dim, e->v.diff_ratio = ratio;
    return RTPP_OK;
}

__global__ void engine_set_am_kernel(EngView ev, const int *__restrict__ n_rand, const float *__restrict__ ell, const int *__restrict__ outpath) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= ev.n_inst) return;
    EngInst &s = ev.inst[w];
    if (n_rand) s.rand_n = n_rand[w], s.rand_pos = 0, s.rand_exhausted = 0, s.samp_pos = 0; // samp_pos: samples drawn from this stream (RT-RRT*)
    if (ell) {
        const float *p = ell + 6 * w;
        s.ell_a = p[0], s.ell_b = p[1], s.ell_cos = p[2], s.ell_sin = p[3], s.ell_mx = p[4], s.ell_my = p[5];
        s.ell_host = 1;
    }
    if (outpath) s.outpath_nonempty = outpath[w];
}

// RT-RRT*: hands the engine the rand() stream of the next tick(s) — n_inst rows of max_samples raw rand() values, n_rand[i] valid —
// and switches it to drawing its samples on the device (SamplePoint, RTRRT.cpp:234-262) instead of taking injected sample points;
// rand == NULL switches back.  rtpp_engine_read_stats reports the draws consumed (slot 14) and whether the stream ran out (15).
int rtpp_engine_set_rand(rtpp_engine *e, const int32_t *rand, const int32_t *n_rand) {
    if (!e || e->v.kind != 1) return RTPP_ERR_INVALID;
    rtpp_ctx *ctx = e->ctx;
    const int n = e->v.n_inst, cap = e->v.max_samples;
    if (!rand) {
        e->v.rand_stream = nullptr;
        return RTPP_OK;
    }
    if (!n_rand) return RTPP_ERR_INVALID;
    for (int w = 0; w < n; w++)
        if (n_rand[w] < 0 || n_rand[w] > cap) RTPP_FAIL(ctx, RTPP_ERR_CAPACITY, "more rand draws than the engine's capacity");
    if (!e->d_rand && eng_alloc(e, &e->d_rand, (size_t)n * cap)) return RTPP_ERR_NOMEM;
    e->v.rand_stream = e->d_rand, e->v.rand_cap = cap;
    ENG_CK(e, cudaMemcpyAsync(e->d_rand, rand, (size_t)n * cap * 4, cudaMemcpyHostToDevice, ctx->stream));
    void *dn;
    int rc = upload(e, 0, n_rand, (size_t)n * 4, &dn);
    if (rc) return rc;
    engine_set_am_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(e->v, (const int *)dn, nullptr, nullptr);
    RTPP_LAUNCHED(ctx);
    ENG_CK(e, cudaStreamSynchronize(ctx->stream));
    return RTPP_OK;
}

// AM-RRT* per-call inputs.  rand: n_inst rows of rand capacity raw rand() values (n_rand[i] valid); cosf / sinf of the angles
// they stand for are evaluated on the device with glibc's bits (glibc_math.cuh).
// ellipse: NULL = the informed sampler's constants (AMRRT.cpp:418-432) are computed on the device from c_base, the root and the
// goal at the start of the next tick's MainTask loop; or 6 floats per instance {a, b, cos(angle), sin(angle), (goal.x+root.x)/2,
// (goal.y+root.y)/2} computed by the caller after UpdateTree, used for the next tick only.
// outpath_nonempty: OutPath.size() != 0 (AMRRT.cpp:402,404).  Any pointer may be NULL = unchanged.
int rtpp_engine_set_am_inputs(rtpp_engine *e, const int32_t *rand, const int32_t *n_rand, const float *ellipse,
                              const int32_t *outpath_nonempty) {
    if (!e || e->v.kind != 2) return RTPP_ERR_INVALID;
    rtpp_ctx *ctx = e->ctx;
    const int n = e->v.n_inst, cap = e->v.rand_cap;
    if (rand) {
        if (!n_rand) return RTPP_ERR_INVALID;
        for (int w = 0; w < n; w++)
            if (n_rand[w] < 0 || n_rand[w] > cap) RTPP_FAIL(ctx, RTPP_ERR_CAPACITY, "more rand draws than the engine's capacity");
        ENG_CK(e, cudaMemcpyAsync(e->d_rand, rand, (size_t)n * cap * 4, cudaMemcpyHostToDevice, ctx->stream));
        ENG_CK(e, cudaStreamSynchronize(ctx->stream));
    }
    void *dn, *de, *dout;
    int rc = upload(e, 0, n_rand, (size_t)n * 4, &dn);
    rc |= upload(e, 1, ellipse, (size_t)n * 24, &de);
    rc |= upload(e, 2, outpath_nonempty, (size_t)n * 4, &dout);
    if (rc) return rc;
    engine_set_am_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(e->v, (const int *)dn, (const float *)de, (const int *)dout);
    RTPP_LAUNCHED(ctx);
    ENG_CK(e, cudaStreamSynchronize(ctx->stream));
    return RTPP_OK;
}

// One MainTask minus DrawDObsmap for every instance (RTRRT.cpp:483-493): optional UpdateTree,
// ExpandRewiring(run_ms) under the virtual clock (k-th read = k*dt_ns), optional GetPath.
// Stream-ordered; no synchronisation.
int rtpp_engine_tick(rtpp_engine *e, int do_update_tree, long run_ms, long long dt_ns, int do_get_path) {
    if (!e || run_ms < 0 || dt_ns < 1) return RTPP_ERR_INVALID;
    rtpp_ctx *ctx = e->ctx;
    const int warps_per_block = e->v.n_inst >= 4 * ctx->sm_count ? 4 : 1;
    const int blocks = (e->v.n_inst + warps_per_block - 1) / warps_per_block;
    if (e->v.kind == 2)
        engine_tick_kernel<2><<<blocks, 32 * warps_per_block, 0, ctx->stream>>>(e->v, e->mt, do_update_tree, (long long)run_ms * 1000000LL, dt_ns,
                                                                                do_get_path);
    else
        engine_tick_kernel<1><<<blocks, 32 * warps_per_block, 0, ctx->stream>>>(e->v, e->mt, do_update_tree, (long long)run_ms * 1000000LL, dt_ns,
                                                                                do_get_path);
    RTPP_LAUNCHED(ctx);
    return RTPP_OK;
}

// per-instance statistics, 16 int64 each: n_nodes, root, samples_used, clock_reads, iters, pops_r,
// pops_s, qr_size, qs_size, status, path_len, path_collision, path_cost bits, c_base bits, 0, 0
int rtpp_engine_read_stats(rtpp_engine *e, long long *out16) {
    if (!e || !out16) return RTPP_ERR_INVALID;
    rtpp_ctx *ctx = e->ctx;
    const int n = e->v.n_inst;
    HostBuf buf((size_t)n * sizeof(EngInst));
    if (!buf.p) return RTPP_ERR_NOMEM;
    ENG_CK(e, cudaMemcpyAsync(buf.p, e->v.inst, (size_t)n * sizeof(EngInst), cudaMemcpyDeviceToHost, ctx->stream));
    ENG_CK(e, cudaStreamSynchronize(ctx->stream));
    const EngInst *s = (const EngInst *)buf.p;
    for (int i = 0; i < n; i++) {
        long long *o = out16 + (size_t)i * 16;
        uint32_t pc, cb;
        memcpy(&pc, &s[i].path_cost, 4), memcpy(&cb, &s[i].c_base, 4);
        o[0] = s[i].n, o[1] = s[i].root, o[2] = s[i].samp_pos, o[3] = s[i].clk_n, o[4] = s[i].iters, o[5] = s[i].pops_r;
        o[6] = s[i].pops_s, o[7] = e->v.kind == 2 ? s[i].gs_top + (s[i].gq_tail - s[i].gq_head) : s[i].qr_tail - s[i].qr_head;
        o[8] = s[i].qs_tail - s[i].qs_head, o[9] = s[i].status;
        o[10] = s[i].path_len, o[11] = s[i].path_col, o[12] = pc, o[13] = cb, o[14] = s[i].rand_pos, o[15] = s[i].rand_exhausted;
    }
    return RTPP_OK;
}

// nodes of one instance in insertion order (xy, parent, cost, collision): Node_Map read-out
int rtpp_engine_read_tree(rtpp_engine *e, int inst, int max_nodes, float *xy, int32_t *parent, float *cost, uint8_t *collision,
                          int *n_out) {
    if (!e || inst < 0 || inst >= e->v.n_inst || !n_out) return RTPP_ERR_INVALID;
    rtpp_ctx *ctx = e->ctx;
    EngInst s;
    ENG_CK(e, cudaMemcpyAsync(&s, e->v.inst + inst, sizeof(s), cudaMemcpyDeviceToHost, ctx->stream));
    ENG_CK(e, cudaStreamSynchronize(ctx->stream));
    const int n = s.n < max_nodes ? s.n : max_nodes;
    *n_out = s.n;
    if (n <= 0 || !xy) return RTPP_OK;
    HostBuf words((size_t)n * 8);
    if (!words.p) return RTPP_ERR_NOMEM;
    ENG_CK(e, cudaMemcpyAsync(xy, e->v.xy + (size_t)inst * e->v.cap, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    ENG_CK(e, cudaMemcpyAsync(words.p, e->v.cp + (size_t)inst * e->v.cap, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    ENG_CK(e, cudaStreamSynchronize(ctx->stream));
    const cpword *w = (const cpword *)words.p;
    for (int i = 0; i < n; i++) {
        const uint32_t hi = cp_hi(w[i]), cb = hi & 0x7fffffffu;
        if (parent) parent[i] = cp_parent(w[i]);
        if (cost) memcpy(&cost[i], &cb, 4);
        if (collision) collision[i] = hi >> 31;
    }
    return RTPP_OK;
}

// path of one instance after a tick with do_get_path: node ids root..entry, then -2 = the goal
int rtpp_engine_read_path(rtpp_engine *e, int inst, int max_len, int32_t *ids, int *len_out) {
    if (!e || inst < 0 || inst >= e->v.n_inst || !ids || !len_out) return RTPP_ERR_INVALID;
    rtpp_ctx *ctx = e->ctx;
    EngInst s;
    ENG_CK(e, cudaMemcpyAsync(&s, e->v.inst + inst, sizeof(s), cudaMemcpyDeviceToHost, ctx->stream));
    ENG_CK(e, cudaStreamSynchronize(ctx->stream));
    int n = s.path_len < e->v.path_cap ? s.path_len : e->v.path_cap;
    *len_out = s.path_len;
    if (n > max_len) n = max_len;
    if (n <= 0) return RTPP_OK;
    HostBuf buf((size_t)n * 4);
    if (!buf.p) return RTPP_ERR_NOMEM;
    ENG_CK(e, cudaMemcpyAsync(buf.p, e->v.path + (size_t)inst * e->v.path_cap, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    ENG_CK(e, cudaStreamSynchronize(ctx->stream));
    const int *p = (const int *)buf.p;
    // the device stores goal-first (push_front order reversed): emit root..goal like the std::list
    // (the direct root-goal shortcut is stored in order; the fallback chain, path_col = 1, is stored node-first as well)
    if ((n >= 2 && p[0] == -2) || s.path_col) for (int i = 0; i < n; i++) ids[i] = p[n - 1 - i];
    else memcpy(ids, p, (size_t)n * 4);
    return RTPP_OK;
}

} // extern "C"
