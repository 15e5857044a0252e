// tree.cu — K4/K5: cost/parent words, tree re-rooting, level-synchronous cost propagation and the
// wave-synchronous rewire relaxation.
//
// Replaces (reference file:line)
//   R3  UpdateTree          RTRRT.cpp:92-136, RTFMT.cpp:92-176 (+N_b capture :167-171),
//                           AMRRT.cpp:156-200, CDRT.cpp:764-796 (larger/smaller sets)
//   R2  rewire relaxations  RTRRT.cpp:300-320,323-353; AMRRT.cpp:483-519 (free edges only)
//   C1  RTrrt_Cost algebra  RTRRT.h:55-91
//
// propagate: (A) every tree edge (parent -> child) gets its weight {E(parent,child), |parent-child|}
// in one fully parallel pass (verdicts do not depend on costs); (B) children lists are built by a
// counting sort on the parent ids; (C) one cooperative, persistent kernel walks the tree level by
// level, cost[child] = cost[parent] + w[child] strictly parent-before-child, so the binary32 sums
// follow the tree order exactly as the reference's BFS does.
// relax: edge weights of the k-NN graph once, then Bellman-Ford sweeps with one 64-bit atomicMin
// per improving edge until no word changes.
#include <cooperative_groups.h>
#include <math_constants.h>

#include "nodes.cuh"
#include "scan.cuh"
namespace cg = cooperative_groups;

static int grid_for(rtpp_ctx *ctx, size_t n, int per_block = 256) {
    size_t need = (n + per_block - 1) / per_block;
    size_t cap = (size_t)ctx->sm_count * 8;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

__global__ void __launch_bounds__(256) tree_pack_kernel(const int *__restrict__ parent, const float *__restrict__ cost,
                                                       const uint8_t *__restrict__ col, int n, cpword *__restrict__ cp) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        cp[i] = cp_make(cost_pack(col[i] != 0, cost[i]), parent[i]);
}
__global__ void __launch_bounds__(256) tree_unpack_kernel(const cpword *__restrict__ cp, int n, int *__restrict__ parent,
                                                         float *__restrict__ cost, uint8_t *__restrict__ col) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const cpword w = cp[i];
        parent[i] = cp_parent(w);
        cost[i] = cost_val(cp_hi(w));
        col[i] = cost_col(cp_hi(w));
    }
}

// (A) + child counting.  w[i] = {E(parent,i), |parent - i|} packed like a cost.  A lane owns a tree edge; edges of 32
// march steps or more (a node that found no near parent hangs far away) are marched by the whole warp afterwards, and
// the children of the ROOT — the one hub a tree may have: thousands of nodes in the synthetic C4 tree — are counted per
// block, so that neither a lane-serial 2000-probe march nor 40 k atomics on one address sit on the critical path
// (270 -> 33 us at C4).
__global__ void __launch_bounds__(256)
tree_edge_weights_kernel(PlaneView pl, const float2 *__restrict__ xy, const cpword *__restrict__ cp, int n, int root,
                         MarchThresholds divisor, uint32_t *__restrict__ w, int *__restrict__ child_count) {
    __shared__ int root_children;
    if (threadIdx.x == 0) root_children = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (int i0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31); i0 < n; i0 += gridDim.x * blockDim.x) {
        const int i = i0 + lane;
        int p = -1;
        if (i < n) p = cp_parent(cp[i]);
        const bool edge = i < n && i != root && p >= 0 && p < n;
        float2 a = make_float2(0.f, 0.f), b = a;
        bool blk = false, longer = false;
        if (edge) {
            a = xy[p], b = xy[i];
            longer = rt_march_P(divisor, rt_sqlen(a.x, a.y, b.x, b.y)) >= 32;
            if (!longer) blk = rt_edge_blocked(pl, a.x, a.y, b.x, b.y, divisor);
        }
        unsigned todo = __ballot_sync(0xffffffffu, longer);
        while (todo) {
            const int k = __ffs(todo) - 1;
            todo &= todo - 1;
            const bool bk = rt_edge_blocked_warp(pl, __shfl_sync(0xffffffffu, a.x, k), __shfl_sync(0xffffffffu, a.y, k),
                                                 __shfl_sync(0xffffffffu, b.x, k), __shfl_sync(0xffffffffu, b.y, k), divisor);
            if (lane == k) blk = bk;
        }
        if (i < n) w[i] = edge ? cost_pack(blk, rt_dist(a.x, a.y, b.x, b.y)) : 0u;
        const unsigned to_root = __ballot_sync(0xffffffffu, edge && p == root);
        if (lane == 0 && to_root) atomicAdd(&root_children, __popc(to_root));
        if (edge && p != root) atomicAdd(&child_count[p], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0 && root_children) atomicAdd(&child_count[root], root_children);
}

// (B) children lists by parent (counting sort: cursor = exclusive scan of the counts).  The root's children take their
// slots block by block (one atomic per 256 nodes) for the same reason as above.
__global__ void __launch_bounds__(256) tree_children_kernel(const cpword *__restrict__ cp, int n, int root,
                                                           int *__restrict__ cursor, int *__restrict__ children) {
    __shared__ int blk_cnt, blk_base;
    const int lane = threadIdx.x & 31;
    for (int i0 = blockIdx.x * blockDim.x; i0 < n; i0 += gridDim.x * blockDim.x) {
        if (threadIdx.x == 0) blk_cnt = 0;
        __syncthreads();
        const int i = i0 + threadIdx.x;
        int p = -1;
        if (i < n) p = cp_parent(cp[i]);
        const bool edge = i < n && i != root && p >= 0 && p < n;
        const bool to_root = edge && p == root;
        const unsigned m = __ballot_sync(0xffffffffu, to_root);
        int warp_base = 0;
        if (lane == 0 && m) warp_base = atomicAdd(&blk_cnt, __popc(m));
        warp_base = __shfl_sync(0xffffffffu, warp_base, 0);
        __syncthreads();
        if (threadIdx.x == 0 && blk_cnt) blk_base = atomicAdd(&cursor[root], blk_cnt);
        __syncthreads();
        if (to_root) children[blk_base + warp_base + __popc(m & ((1u << lane) - 1u))] = i;
        else if (edge) children[atomicAdd(&cursor[p], 1)] = i;
    }
}

// (C) persistent level-synchronous walk.  counters[0..2] rotate: level L reads [L%3], appends to
// [(L+1)%3] and clears [(L+2)%3].  flags (optional): bit0 = collision went 0 -> 1 (RT-FMT's N_b,
// RTFMT.cpp:170-171), bit1 = cost grew, bit2 = cost shrank (CDRT's larger/smaller sets,
// CDRT.cpp:781-790).  counters[3] = number of nodes reached.
// Fan-out decides who expands a frontier node: its own thread (<= 8 children), its warp (<= kTreeBig: the
// children are spread over the lanes), or — a root that thousands of nodes hang from — the whole grid after the
// level's barrier (counters[4 + (L&1)] = how many such nodes level L met, listed behind the counters).
static constexpr int kTreeBig = 2048;
__device__ __forceinline__ uint32_t tree_visit(int c, int p, uint32_t hp, const uint32_t *__restrict__ w, cpword *__restrict__ cp,
                                               uint8_t *__restrict__ flags) {
    const uint32_t hn = cost_add(hp, w[c]);
    if (flags) {
        const uint32_t ho = cp_hi(cp[c]);
        flags[c] = (uint8_t)(((hn >> 31) & ~(ho >> 31) & 1u) | ((hn > ho) ? 2u : 0u) | ((hn < ho) ? 4u : 0u));
    }
    cp[c] = cp_make(hn, p);
    return hn;
}
// One warp step of the walk: lane l expands parent p (cost word hp, children [s, e)); the warp reserves the slots of
// all its children at once — in the block's shared-memory queue `qo` when they fit (the block then keeps walking
// down these subtrees without a grid barrier), else in the global next frontier.
constexpr int kTreeQ = 2048; // entries per shared-memory queue (node, cost word)
struct TreeBlock {
    int2 q[2][kTreeQ];
    int qn[2];
};
__device__ __forceinline__ void tree_expand(TreeBlock &sm, int qo, int p, uint32_t hp, int s, int e, const int *__restrict__ children,
                                            const uint32_t *__restrict__ w, cpword *__restrict__ cp, uint8_t *__restrict__ flags, int *nxt,
                                            int *cnt_next, int *reached, int *big, int *big_count, int big_cap) {
    const int lane = threadIdx.x & 31;
    int cnt = e - s;
    if (cnt > kTreeBig) { // expanded by the grid after the barrier
        const int k = atomicAdd(big_count, 1);
        if (k < big_cap) big[k] = p;
        cnt = 0;
    }
    int x = cnt;
    for (int off = 1; off < 32; off <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, off);
        if (lane >= off) x += y;
    }
    const int total = __shfl_sync(0xffffffffu, x, 31);
    if (total == 0) return;
    int base = -1; // >= 0: slot in the shared-memory queue
    if (lane == 31) {
        int old = sm.qn[qo], assumed;
        do {
            assumed = old;
            if (assumed + total > kTreeQ) break;
            old = atomicCAS(&sm.qn[qo], assumed, assumed + total);
        } while (old != assumed);
        if (assumed + total <= kTreeQ) base = assumed;
        atomicAdd(reached, total);
    }
    base = __shfl_sync(0xffffffffu, base, 31);
    const bool local = base >= 0;
    if (!local) {
        if (lane == 31) base = atomicAdd(cnt_next, total);
        base = __shfl_sync(0xffffffffu, base, 31);
    }
    base += x - cnt;
    if (cnt <= 8)
        for (int j = 0; j < cnt; j++) {
            const int c = children[s + j];
            const uint32_t hn = tree_visit(c, p, hp, w, cp, flags);
            if (local) sm.q[qo][base + j] = make_int2(c, (int)hn);
            else nxt[base + j] = c;
        }
    unsigned wide = __ballot_sync(0xffffffffu, cnt > 8);
    while (wide) {
        const int o = __ffs(wide) - 1;
        wide &= wide - 1;
        const int po = __shfl_sync(0xffffffffu, p, o), so = __shfl_sync(0xffffffffu, s, o), co = __shfl_sync(0xffffffffu, cnt, o);
        const int bo = __shfl_sync(0xffffffffu, base, o);
        const uint32_t ho = __shfl_sync(0xffffffffu, hp, o);
        for (int j = lane; j < co; j += 32) {
            const int c = children[so + j];
            const uint32_t hn = tree_visit(c, po, ho, w, cp, flags);
            if (local) sm.q[qo][bo + j] = make_int2(c, (int)hn);
            else nxt[bo + j] = c;
        }
    }
}

__global__ void __launch_bounds__(256)
tree_bfs_kernel(int n, int root, const int *__restrict__ child_start, const int *__restrict__ children,
                const uint32_t *__restrict__ w, cpword *__restrict__ cp, int *__restrict__ frontier, int capacity,
                int *__restrict__ counters, uint8_t *__restrict__ flags, int big_cap) {
    cg::grid_group grid = cg::this_grid();
    __shared__ TreeBlock sm;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x, lane = tid & 31;
    int *big_list = counters + 8; // [2][big_cap]
    if (tid == 0) {
        frontier[0] = root;
        counters[0] = 1, counters[1] = 0, counters[2] = 0, counters[3] = 1, counters[4] = 0, counters[5] = 0;
    }
    grid.sync();
    for (int level = 0;; level++) {
        const int cur_n = counters[level % 3];
        if (cur_n == 0) break;
        int *cnt_next = &counters[(level + 1) % 3];
        const int *cur = frontier + (size_t)(level & 1) * capacity;
        int *nxt = frontier + (size_t)((level + 1) & 1) * capacity;
        int *big = big_list + (size_t)(level & 1) * big_cap;
        int *big_count = &counters[4 + (level & 1)];
        if (tid == 0) counters[(level + 2) % 3] = 0;
        if (threadIdx.x < 2) sm.qn[threadIdx.x] = 0;
        __syncthreads();
        // this block's share of the global frontier, one node per thread
        for (int f0 = (tid & ~31); f0 < cur_n; f0 += nth) {
            const int f = f0 + lane;
            int p = -1, s = 0, e = 0;
            uint32_t hp = 0;
            if (f < cur_n) {
                p = cur[f];
                hp = cp_hi(cp[p]);
                s = child_start[p], e = child_start[p + 1];
            }
            tree_expand(sm, 0, p, hp, s, e, children, w, cp, flags, nxt, cnt_next, &counters[3], big, big_count, big_cap);
        }
        // ... then down the subtrees it holds, level by level inside the block (a __syncthreads per level instead of a
        // grid barrier; the cost word travels with the node, so a step is children -> weight -> store)
        for (int pq = 0;; pq ^= 1) {
            __syncthreads();
            const int m = sm.qn[pq];
            if (m == 0) break;
            if (threadIdx.x == 0) sm.qn[pq ^ 1] = 0; // (nobody reads the drained queue's count again)
            __syncthreads();
            for (int f0 = (threadIdx.x & ~31); f0 < m; f0 += blockDim.x) {
                const int f = f0 + lane;
                int p = -1, s = 0, e = 0;
                uint32_t hp = 0;
                if (f < m) {
                    const int2 ent = sm.q[pq][f];
                    p = ent.x, hp = (uint32_t)ent.y;
                    s = child_start[p], e = child_start[p + 1];
                }
                tree_expand(sm, pq ^ 1, p, hp, s, e, children, w, cp, flags, nxt, cnt_next, &counters[3], big, big_count, big_cap);
            }
        }
        grid.sync();
        const int nb = min(counters[4 + (level & 1)], big_cap);
        if (nb) { // uniform: every thread reads the same count after the barrier
            int off = *cnt_next;
            for (int b = 0; b < nb; b++) {
                const int p = big[b];
                const uint32_t hp = cp_hi(cp[p]);
                const int s = child_start[p], cnt = child_start[p + 1] - s;
                for (int j = tid; j < cnt; j += nth) {
                    const int c = children[s + j];
                    tree_visit(c, p, hp, w, cp, flags);
                    nxt[off + j] = c;
                }
                off += cnt;
            }
            grid.sync(); // every thread has read *cnt_next and written its share
            if (tid == 0) counters[3] += off - *cnt_next, *cnt_next = off, counters[4 + (level & 1)] = 0;
            grid.sync();
        }
    }
}

// re-root (RTRRT.cpp:98-119): reverse the parent chain new_root -> ... -> old_root; cost[new_root] = {0,0}
__global__ void tree_reroot_kernel(cpword *__restrict__ cp, int n, int old_root, int new_root) {
    if (threadIdx.x || blockIdx.x) return;
    int x = new_root, x_old = -1;
    int guard = 0;
    while (x != old_root && x >= 0 && guard++ < n) {
        const cpword w = cp[x];
        const int par = cp_parent(w);
        cp[x] = cp_make(cp_hi(w), x_old);
        x_old = x;
        x = par;
    }
    if (x >= 0) cp[x] = cp_make(cp_hi(cp[x]), x_old);
    cp[new_root] = cp_make(cost_pack(false, 0.0f), -1);
}

// relax phase 1: weights of the directed k-NN edges x -> knn[x][j]; 0xffffffff = no edge
__global__ void __launch_bounds__(256)
relax_weights_kernel(PlaneView pl, const float2 *__restrict__ xy, const int *__restrict__ knn, size_t n_edges, int k,
                     MarchThresholds divisor, uint32_t *__restrict__ w) {
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(e / k), s = knn[e];
        uint32_t we = 0xffffffffu;
        if (s >= 0 && s != x) {
            const float2 a = xy[x], b = xy[s];
            we = cost_pack(rt_edge_blocked(pl, a.x, a.y, b.x, b.y, divisor), rt_dist(a.x, a.y, b.x, b.y));
        }
        w[e] = we;
    }
}

// relax phase 2: c_new = cost[x] + w; if c_new < cost[s] (strict, lexicographic: RTRRT.cpp:311) the word {c_new, x}
// replaces cost|parent of s by atomicMin (equal costs: the smaller parent id, so the result does not depend on the
// schedule).  free_only: AM-RRT* RewireRoot relaxes only over free edges and never to a colliding cost
// (AMRRT.cpp:500-504).
//
// Schedule.  Only the edges out of a node whose cost fell can improve anything, so the iteration is a work list of
// such nodes, and the least fixpoint does not depend on the order in which it is drained (every update is an
// atomicMin).  Two measured facts shape the order (C4, 1e6 nodes, 25 M edges, longest improving chain ~740 hops):
// a wave-synchronous drain pays one grid barrier per hop (~25 us each with the counter and frontier round trips
// through L2: 20 ms), and draining without any order lets provisional costs run ahead and be corrected over and over
// (28 ms at 32 unordered hops per barrier, 95 ms at 128).  So the drain is ordered by cost bucket, near-far /
// delta-stepping style: an outer round (one grid barrier) has a threshold T = (smallest queued cost) + delta; a block
// relaxes its share of the queued nodes with cost <= T and CHASES the nodes it improves — those with a new cost <= T go
// to a queue in its shared memory and are relaxed in the next inner round after a __syncthreads (a hop costs the L2
// latency chain of one node, not a barrier), everything else waits in the global list for a later bucket.
// Whether a node sits in some list is bit 0 of its word WHILE THE KERNEL RUNS (cost | parent << 1 | idle): the
// improving atomicMin carries idle = 0 and its return value says whether the node was idle before (then the improver
// queues it), the warp that takes a node sets the bit and reads the cost with one atomicOr — one L2 round trip each,
// no separate flag, no fence — so an improvement that lands after the take re-queues the node (asynchronous
// work-list protocol).  Order of a word: cost, parent id, idle bit, so equal costs still resolve to the smaller parent.  The packed cost word is the bucket key: colliding
// costs (bit 31) sort after every free cost and are drained last, in one bucket.
// counters: [0..2] list sizes in rotation (current, next, being cleared), [3] outer rounds done, [4..6] smallest key of
// the list, same rotation
constexpr int kRelaxThreads = 512;
constexpr int kRelaxQ = 1024;     // entries per shared-memory queue
constexpr int kRelaxSpill = 2048; // block-local staging of the entries bound for the global list

struct RelaxBlock { // shared memory of one block
    int q[2][kRelaxQ];
    int spill[kRelaxSpill];
    int qn[2];
    int spill_n;
    uint32_t spill_min;
    int base;
};

// in-kernel word: cost(32) | parent(31) | idle(1); parent "none" (-1) = 0x7fffffff
__device__ __forceinline__ cpword rq_make(uint32_t hi, int parent, uint32_t idle) {
    return ((cpword)hi << 32) | (((uint32_t)parent & 0x7fffffffu) << 1) | idle;
}
__device__ __forceinline__ cpword rq_to_plain(cpword wq) {
    const uint32_t p31 = ((uint32_t)wq) >> 1;
    return cp_make(cp_hi(wq), p31 == 0x7fffffffu ? -1 : (int)p31);
}

__device__ __forceinline__ uint32_t relax_threshold(uint32_t min_key, float delta) {
    if (min_key >> 31) return 0xffffffffu; // only colliding costs are left: one last bucket
    return __float_as_uint(__fadd_rn(__uint_as_float(min_key), delta));
}

// lanes with `go`: stage node sn (bucket key `key`) for the global list
__device__ __forceinline__ void relax_defer(RelaxBlock &sm, bool go, int sn, uint32_t key, int *nxt, int *nxt_count, uint32_t *nxt_min) {
    const int lane = threadIdx.x & 31;
    const unsigned bal = __ballot_sync(0xffffffffu, go);
    if (!bal) return;
    const uint32_t mk = __reduce_min_sync(0xffffffffu, go ? key : 0xffffffffu);
    int base = 0;
    if (lane == 0) {
        base = atomicAdd(&sm.spill_n, __popc(bal));
        atomicMin(&sm.spill_min, mk);
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    if (go) {
        const int pos = base + __popc(bal & ((1u << lane) - 1u));
        if (pos < kRelaxSpill) sm.spill[pos] = sn;
        else nxt[atomicAdd(nxt_count, 1)] = sn; // staging full: straight to the list (its key is in spill_min)
    }
}

// block-collective: move the staged entries to the global list
__device__ __forceinline__ void relax_flush(RelaxBlock &sm, int *nxt, int *nxt_count) {
    __syncthreads();
    const int cnt = min(sm.spill_n, kRelaxSpill);
    if (cnt > 0) {
        if (threadIdx.x == 0) sm.base = atomicAdd(nxt_count, cnt);
        __syncthreads();
        for (int f = threadIdx.x; f < cnt; f += blockDim.x) nxt[sm.base + f] = sm.spill[f];
        __syncthreads();
        if (threadIdx.x == 0) sm.spill_n = 0;
    }
    __syncthreads();
}

// warp-collective: relax the k edges out of x (already taken: hx is the cost relax_take returned); improved, idle
// neighbours go to the shared-memory queue qo when their new cost is in the bucket, else to the staging
__device__ __forceinline__ void relax_node(RelaxBlock &sm, int x, uint32_t hx, uint32_t T, const int *__restrict__ knn,
                                           const uint32_t *__restrict__ w, int k, int free_only, cpword *cp, int qo,
                                           int *nxt, int *nxt_count, uint32_t *nxt_min) {
    const int lane = threadIdx.x & 31;
    for (int j0 = 0; j0 < k; j0 += 32) {
        const int j = j0 + lane;
        bool improved = false;
        int sn = -1;
        uint32_t hn = 0xffffffffu;
        if (j < k) {
            const size_t e = (size_t)x * k + j;
            const uint32_t we = w[e];
            if (we != 0xffffffffu && !(free_only && (we >> 31))) {
                sn = knn[e];
                hn = cost_add(hx, we);
                const cpword neww = rq_make(hn, x, 0);
                // the test before the atomic also keeps a re-relaxation at unchanged cost from re-queueing a whole subtree
                if (!(free_only && (hn >> 31)) && neww < (((volatile cpword *)cp)[sn] & ~1ull)) {
                    if (hn > hx) {
                        const cpword prev = atomicMin(&cp[sn], neww);
                        improved = neww < prev && (prev & 1ull); // the word was replaced and the node was idle: queue it
                    } else {
                        // hn == hx: the edge's weight is absorbed by the rounding (coincident nodes).  Such an edge may carry a
                        // strict cost improvement — the reference's sequential loop takes it too, RTRRT.cpp:311 — but not the
                        // equal-cost parent-id tie-break of the word compare: two coincident nodes would adopt each other as
                        // parents and drop out of the tree.  Rare, so a compare-and-swap loop instead of the atomicMin.
                        cpword cur = ((volatile cpword *)cp)[sn];
                        while (hn < cp_hi(cur)) {
                            const cpword prev = atomicCAS(&cp[sn], cur, neww);
                            if (prev == cur) {
                                improved = prev & 1ull;
                                break;
                            }
                            cur = prev;
                        }
                    }
                }
            }
        }
        if (improved) { // sn will be taken within a few inner rounds: start its two edge rows towards L2 now
            const size_t r = (size_t)sn * k;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(w + r));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(w + r + k - 1));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(knn + r));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(knn + r + k - 1));
        }
        bool near = improved && hn <= T;
        const unsigned bal = __ballot_sync(0xffffffffu, near);
        if (bal) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&sm.qn[qo], __popc(bal));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (near) {
                const int pos = base + __popc(bal & ((1u << lane) - 1u));
                if (pos < kRelaxQ) sm.q[qo][pos] = sn;
                else near = false; // queue full: to the global list
            }
        }
        relax_defer(sm, improved && !near, sn, hn, nxt, nxt_count, nxt_min);
    }
}

// warp-collective: take x off its list: mark it idle and read its cost in one atomic
__device__ __forceinline__ uint32_t relax_take(int x, cpword *cp) {
    uint32_t h = 0;
    if ((threadIdx.x & 31) == 0) h = cp_hi(atomicOr(&cp[x], 1ull));
    return __shfl_sync(0xffffffffu, h, 0);
}

__global__ void __launch_bounds__(kRelaxThreads)
relax_chase_kernel(const int *__restrict__ knn, const uint32_t *__restrict__ w, int n, int k, int free_only, cpword *cp, int *fr_a, int *fr_b,
                   int *counters, int max_rounds, int chase, float delta) {
    cg::grid_group grid = cg::this_grid();
    __shared__ RelaxBlock sm;
    const int wib = threadIdx.x >> 5, wpb = blockDim.x >> 5, lane = threadIdx.x & 31;
    uint32_t *mins = (uint32_t *)counters + 4;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        counters[0] = counters[1] = counters[2] = counters[3] = 0;
        mins[0] = mins[1] = mins[2] = 0xffffffffu;
    }
    if (threadIdx.x == 0) sm.spill_n = 0, sm.spill_min = 0xffffffffu;
    grid.sync();
    // list 0: every node that can relax an edge at all
    for (int i0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31; i0 < n; i0 += gridDim.x * blockDim.x) {
        const int i = i0 + lane;
        uint32_t h = 0xffffffffu;
        cpword wd = 0;
        if (i < n) wd = cp[i], h = cp_hi(wd);
        const bool go = i < n && h != 0xffffffffu && !(free_only && (h >> 31));
        if (i < n) cp[i] = rq_make(h, cp_parent(wd), go ? 0u : 1u);
        relax_defer(sm, go, i, h, fr_a, &counters[0], &mins[0]);
    }
    relax_flush(sm, fr_a, &counters[0]);
    if (threadIdx.x == 0 && sm.spill_min != 0xffffffffu) atomicMin(&mins[0], sm.spill_min);
    grid.sync();
    int *cur = fr_a, *nxt = fr_b;
    for (int round = 0; round < max_rounds; round++) {
        const int m = ((volatile int *)counters)[round % 3];
        if (m == 0) break;
        const uint32_t T = relax_threshold(((volatile uint32_t *)mins)[round % 3], delta);
        int *nxt_count = &counters[(round + 1) % 3];
        uint32_t *nxt_min = &mins[(round + 1) % 3];
        if (blockIdx.x == 0 && threadIdx.x == 0) counters[(round + 2) % 3] = 0, mins[(round + 2) % 3] = 0xffffffffu, counters[3] = round + 1;
        if (threadIdx.x < 2) sm.qn[threadIdx.x] = 0;
        if (threadIdx.x == 0) sm.spill_min = 0xffffffffu;
        __syncthreads();
        // this block's share of the list: entries of the bucket go to the block's queue, the others wait.  A short
        // list is dealt one entry per warp, blocks first, so that every block gets some of it to chase.
        {
            const int total_warps = wpb * gridDim.x, gw = wib * gridDim.x + blockIdx.x;
            const int g = min(32, (m + total_warps - 1) / total_warps);
            for (int f0 = gw * g; f0 < m; f0 += total_warps * g) {
                const int f = f0 + lane;
                const bool have = lane < g && f < m;
                int x = -1;
                uint32_t h = 0xffffffffu;
                if (have) x = cur[f], h = cp_hi(((volatile cpword *)cp)[x]);
                bool near = have && h <= T;
                const unsigned bal = __ballot_sync(0xffffffffu, near);
                if (bal) {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&sm.qn[0], __popc(bal));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (near) {
                        const int pos = base + __popc(bal & ((1u << lane) - 1u));
                        if (pos < kRelaxQ) sm.q[0][pos] = x;
                        else near = false;
                    }
                }
                relax_defer(sm, have && !near, x, h, nxt, nxt_count, nxt_min);
            }
        }
        int p = 0;
        for (int it = 0;; it++) {
            relax_flush(sm, nxt, nxt_count); // (also the barrier that completes q[p])
            const int cnt = min(sm.qn[p], kRelaxQ);
            if (cnt == 0) break;
            if (it == chase) { // hand the rest to the next outer round (entries stay marked as queued)
                for (int f0 = wib * 32; f0 < cnt; f0 += blockDim.x) {
                    const int f = f0 + lane;
                    int x = -1;
                    uint32_t h = 0xffffffffu;
                    if (f < cnt) x = sm.q[p][f], h = cp_hi(((volatile cpword *)cp)[x]);
                    relax_defer(sm, f < cnt, x, h, nxt, nxt_count, nxt_min);
                }
                relax_flush(sm, nxt, nxt_count);
                break;
            }
            if (threadIdx.x == 0) sm.qn[p ^ 1] = 0;
            __syncthreads();
            for (int f = wib; f < cnt; f += wpb) {
                const int x = sm.q[p][f];
                const uint32_t hx = relax_take(x, cp);
                relax_node(sm, x, hx, T, knn, w, k, free_only, cp, p ^ 1, nxt, nxt_count, nxt_min);
            }
            p ^= 1;
        }
        if (threadIdx.x == 0 && sm.spill_min != 0xffffffffu) atomicMin(nxt_min, sm.spill_min);
        int *t = cur;
        cur = nxt, nxt = t;
        grid.sync();
    }
    // back to the resident layout (cost | parent)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) cp[i] = rq_to_plain(cp[i]);
}

static int ensure_tree_scratch(rtpp_ctx *ctx, NodeStore *s) {
    if (s->edge_w) return RTPP_OK;
    RTPP_CK(ctx, cudaMalloc(&s->edge_w, s->capacity * 4));
    RTPP_CK(ctx, cudaMalloc(&s->child_start, (s->capacity + 2) * 4));
    RTPP_CK(ctx, cudaMalloc(&s->children, s->capacity * 4));
    RTPP_CK(ctx, cudaMalloc(&s->frontier, 2 * s->capacity * 4));
    RTPP_CK(ctx, cudaMalloc(&s->counters, (8 + 2 * (s->capacity / kTreeBig + 1)) * 4)); // + the two lists of wide nodes
    return RTPP_OK;
}

extern "C" {

int rtpp_tree_set(rtpp_ctx *ctx, const int32_t *parent, const float *cost, const uint8_t *collision, size_t n) {
    if (!ctx || (n && (!parent || !cost || !collision))) return RTPP_ERR_INVALID;
    NodeStore *s = ctx->nodes;
    if (!s) RTPP_FAIL(ctx, RTPP_ERR_STATE, "rtpp_nodes_reset has not been called");
    if (n > (size_t)s->n) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "more tree entries than nodes");
    if (n == 0) return RTPP_OK;
    void *dp = nullptr, *dc = nullptr, *dl = nullptr;
    int rc = rtpp_scratch(ctx, 0, n * 4, &dp);
    if (rc) return rc;
    rc = rtpp_scratch(ctx, 1, n * 4, &dc);
    if (rc) return rc;
    rc = rtpp_scratch(ctx, 2, n, &dl);
    if (rc) return rc;
    RTPP_CK(ctx, cudaMemcpyAsync(dp, parent, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    RTPP_CK(ctx, cudaMemcpyAsync(dc, cost, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    RTPP_CK(ctx, cudaMemcpyAsync(dl, collision, n, cudaMemcpyHostToDevice, ctx->stream));
    tree_pack_kernel<<<grid_for(ctx, n), 256, 0, ctx->stream>>>((const int *)dp, (const float *)dc, (const uint8_t *)dl, (int)n, s->cp);
    RTPP_LAUNCHED(ctx);
    RTPP_CK(ctx, cudaStreamSynchronize(ctx->stream));
    return RTPP_OK;
}

int rtpp_tree_get(rtpp_ctx *ctx, int32_t *parent, float *cost, uint8_t *collision, size_t n) {
    if (!ctx || (n && (!parent || !cost || !collision))) return RTPP_ERR_INVALID;
    NodeStore *s = ctx->nodes;
    if (!s) RTPP_FAIL(ctx, RTPP_ERR_STATE, "rtpp_nodes_reset has not been called");
    if (n > (size_t)s->n) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "more tree entries than nodes");
    if (n == 0) return RTPP_OK;
    void *dp = nullptr, *dc = nullptr, *dl = nullptr;
    int rc = rtpp_scratch(ctx, 0, n * 4, &dp);
    if (rc) return rc;
    rc = rtpp_scratch(ctx, 1, n * 4, &dc);
    if (rc) return rc;
    rc = rtpp_scratch(ctx, 2, n, &dl);
    if (rc) return rc;
    tree_unpack_kernel<<<grid_for(ctx, n), 256, 0, ctx->stream>>>(s->cp, (int)n, (int *)dp, (float *)dc, (uint8_t *)dl);
    RTPP_LAUNCHED(ctx);
    RTPP_CK(ctx, cudaMemcpyAsync(parent, dp, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    RTPP_CK(ctx, cudaMemcpyAsync(cost, dc, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    RTPP_CK(ctx, cudaMemcpyAsync(collision, dl, n, cudaMemcpyDeviceToHost, ctx->stream));
    RTPP_CK(ctx, cudaStreamSynchronize(ctx->stream));
    return RTPP_OK;
}

int rtpp_tree_words_dev(rtpp_ctx *ctx, unsigned long long **d_words) {
    if (!ctx || !d_words) return RTPP_ERR_INVALID;
    if (!ctx->nodes) RTPP_FAIL(ctx, RTPP_ERR_STATE, "rtpp_nodes_reset has not been called");
    *d_words = ctx->nodes->cp;
    return RTPP_OK;
}

int rtpp_tree_reroot(rtpp_ctx *ctx, int old_root, int new_root) {
    if (!ctx) return RTPP_ERR_INVALID;
    NodeStore *s = ctx->nodes;
    if (!s) RTPP_FAIL(ctx, RTPP_ERR_STATE, "rtpp_nodes_reset has not been called");
    if (old_root < 0 || old_root >= s->n || new_root < 0 || new_root >= s->n) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "bad root id");
    if (old_root == new_root) return RTPP_OK;
    tree_reroot_kernel<<<1, 32, 0, ctx->stream>>>(s->cp, s->n, old_root, new_root);
    RTPP_LAUNCHED(ctx);
    return RTPP_OK;
}

int rtpp_tree_propagate_dev(rtpp_ctx *ctx, int root, int divisor, int plane_mask, uint8_t *d_flags, int *reached) {
    if (!ctx) return RTPP_ERR_INVALID;
    NodeStore *s = ctx->nodes;
    if (!s) RTPP_FAIL(ctx, RTPP_ERR_STATE, "rtpp_nodes_reset has not been called");
    if (root < 0 || root >= s->n) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "bad root id");
    if (divisor <= 0) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "divisor must be positive");
    PlaneView pv;
    int rc = rtpp_select_plane(ctx, plane_mask, &pv);
    if (rc) return rc;
    rc = ensure_tree_scratch(ctx, s);
    if (rc) return rc;
    const int n = s->n;
    MarchThresholds mt;
    rtpp_march_thresholds(divisor, &mt);
    // child_start doubles as the count array, then the scan turns it into starts; children cursor = frontier tail
    int *cursor = s->frontier; // scratch before the BFS uses it
    RTPP_CK(ctx, cudaMemsetAsync(s->child_start, 0, (size_t)(n + 2) * 4, ctx->stream));
    if (d_flags) RTPP_CK(ctx, cudaMemsetAsync(d_flags, 0, n, ctx->stream));
    tree_edge_weights_kernel<<<grid_for(ctx, n), 256, 0, ctx->stream>>>(pv, s->xy, s->cp, n, root, mt, s->edge_w, s->child_start);
    RTPP_LAUNCHED(ctx);
    rc = rtpp_exclusive_scan(ctx, s->child_start, n, s->child_start, cursor, s->scan_totals);
    if (rc) return rc;
    tree_children_kernel<<<grid_for(ctx, n), 256, 0, ctx->stream>>>(s->cp, n, root, cursor, s->children);
    RTPP_LAUNCHED(ctx);
    // persistent cooperative walk: one block for small trees (cheap syncs), else all SMs
    int blocks = 1;
    if (n > 16384) {
        int per_sm = 0;
        RTPP_CK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tree_bfs_kernel, 256, 0));
        blocks = ctx->sm_count * (per_sm > 2 ? 2 : (per_sm < 1 ? 1 : per_sm));
    }
    int cap = (int)s->capacity;
    const int *cs = s->child_start, *ch = s->children;
    const uint32_t *w = s->edge_w;
    cpword *cp = s->cp;
    int *fr = s->frontier, *cn = s->counters;
    int nn = n, rr = root, big_cap = (int)(s->capacity / kTreeBig + 1);
    void *args[] = {&nn, &rr, &cs, &ch, &w, &cp, &fr, &cap, &cn, &d_flags, &big_cap};
    RTPP_CK(ctx, cudaLaunchCooperativeKernel((void *)tree_bfs_kernel, dim3(blocks), dim3(256), args, 0, ctx->stream));
    ctx->launches++;
    if (reached) {
        RTPP_CK(ctx, cudaMemcpyAsync(reached, s->counters + 3, 4, cudaMemcpyDeviceToHost, ctx->stream));
        RTPP_CK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return RTPP_OK;
}

int rtpp_tree_propagate(rtpp_ctx *ctx, int root, int divisor, int plane_mask, uint8_t *flags, int *reached) {
    if (!ctx) return RTPP_ERR_INVALID;
    NodeStore *s = ctx->nodes;
    if (!s) RTPP_FAIL(ctx, RTPP_ERR_STATE, "rtpp_nodes_reset has not been called");
    void *df = nullptr;
    if (flags) {
        int rc = rtpp_scratch(ctx, 3, (size_t)s->n + 1, &df);
        if (rc) return rc;
    }
    int r = 0;
    int rc = rtpp_tree_propagate_dev(ctx, root, divisor, plane_mask, (uint8_t *)df, &r);
    if (rc) return rc;
    if (reached) *reached = r;
    if (flags) {
        RTPP_CK(ctx, cudaMemcpyAsync(flags, df, s->n, cudaMemcpyDeviceToHost, ctx->stream));
        RTPP_CK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return RTPP_OK;
}

int rtpp_relax_dev(rtpp_ctx *ctx, const int32_t *d_knn_idx, int k, int divisor, int plane_mask, int free_only,
                   int max_sweeps, int *sweeps_done) {
    if (!ctx || !d_knn_idx) return RTPP_ERR_INVALID;
    NodeStore *s = ctx->nodes;
    if (!s) RTPP_FAIL(ctx, RTPP_ERR_STATE, "rtpp_nodes_reset has not been called");
    if (k < 1 || divisor <= 0 || max_sweeps < 1) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "bad relax arguments");
    PlaneView pv;
    int rc = rtpp_select_plane(ctx, plane_mask, &pv);
    if (rc) return rc;
    rc = ensure_tree_scratch(ctx, s);
    if (rc) return rc;
    const size_t ne = (size_t)s->n * k;
    MarchThresholds mt;
    rtpp_march_thresholds(divisor, &mt);
    void *dw = nullptr;
    rc = rtpp_scratch(ctx, 3, ne * 4, &dw);
    if (rc) return rc;
    relax_weights_kernel<<<grid_for(ctx, ne), 256, 0, ctx->stream>>>(pv, s->xy, d_knn_idx, ne, k, mt, (uint32_t *)dw);
    RTPP_LAUNCHED(ctx);
    int sweeps = 0;
    {
        int per_sm = 0;
        RTPP_CK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, relax_chase_kernel, kRelaxThreads, 0));
        // all warp slots of an SM (4 blocks of 512): a hop is a chain of L2 round trips, more warps hide more of it
        const int cap_bps = 4;
        int chase = 64;
        // bucket width: 3 k-NN radii (r_k = sqrt(k / (pi * density)) for uniformly spread nodes); the measured optimum
        // at C4 is flat between 2 and 3.5 r_k (wider: provisional costs run ahead; narrower: more barriers)
        float delta = 3.0f * sqrtf((float)k * (float)ctx->W * (float)ctx->H / (3.14159265f * (float)(s->n > 0 ? s->n : 1)));
        int blocks = ctx->sm_count * (per_sm < 1 ? 1 : (per_sm > cap_bps ? cap_bps : per_sm));
        const int need = (int)(((size_t)s->n * 32 + kRelaxThreads - 1) / kRelaxThreads);
        if (blocks > need) blocks = need < 1 ? 1 : need;
        const int *knn = d_knn_idx;
        const uint32_t *ww = (const uint32_t *)dw;
        cpword *cp = s->cp;
        int nn = s->n, kk = k, fo = free_only, mr = max_sweeps;
        int *fa = s->frontier, *fb = s->frontier + s->capacity, *cn = s->counters;
        void *args[] = {&knn, &ww, &nn, &kk, &fo, &cp, &fa, &fb, &cn, &mr, &chase, &delta};
        RTPP_CK(ctx, cudaLaunchCooperativeKernel((void *)relax_chase_kernel, dim3(blocks), dim3(kRelaxThreads), args, 0, ctx->stream));
        ctx->launches++;
        RTPP_CK(ctx, cudaMemcpyAsync(&sweeps, s->counters + 3, 4, cudaMemcpyDeviceToHost, ctx->stream));
        RTPP_CK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (sweeps_done) *sweeps_done = sweeps;
    return RTPP_OK;
}

} // extern "C"
