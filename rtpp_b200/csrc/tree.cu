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

// (A) + child counting.  w[i] = {E(parent,i), |parent - i|} packed like a cost.
__global__ void __launch_bounds__(256)
tree_edge_weights_kernel(PlaneView pl, const float2 *__restrict__ xy, const cpword *__restrict__ cp, int n, int root,
                         MarchThresholds divisor, uint32_t *__restrict__ w, int *__restrict__ child_count) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int p = cp_parent(cp[i]);
        uint32_t wi = 0;
        if (i != root && p >= 0 && p < n) {
            const float2 a = xy[p], b = xy[i];
            wi = cost_pack(rt_edge_blocked(pl, a.x, a.y, b.x, b.y, divisor), rt_dist(a.x, a.y, b.x, b.y));
            atomicAdd(&child_count[p], 1);
        }
        w[i] = wi;
    }
}

__global__ void __launch_bounds__(256) tree_children_kernel(const cpword *__restrict__ cp, int n, int root,
                                                           int *__restrict__ cursor, int *__restrict__ children) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int p = cp_parent(cp[i]);
        if (i != root && p >= 0 && p < n) children[atomicAdd(&cursor[p], 1)] = i;
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
__device__ __forceinline__ void tree_visit(int c, int p, uint32_t hp, const uint32_t *__restrict__ w, cpword *__restrict__ cp,
                                           uint8_t *__restrict__ flags) {
    const uint32_t hn = cost_add(hp, w[c]);
    if (flags) {
        const uint32_t ho = cp_hi(cp[c]);
        flags[c] = (uint8_t)(((hn >> 31) & ~(ho >> 31) & 1u) | ((hn > ho) ? 2u : 0u) | ((hn < ho) ? 4u : 0u));
    }
    cp[c] = cp_make(hn, p);
}
__global__ void __launch_bounds__(256)
tree_bfs_kernel(int n, int root, const int *__restrict__ child_start, const int *__restrict__ children,
                const uint32_t *__restrict__ w, cpword *__restrict__ cp, int *__restrict__ frontier, int capacity,
                int *__restrict__ counters, uint8_t *__restrict__ flags, int big_cap) {
    cg::grid_group grid = cg::this_grid();
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
        if (tid == 0) counters[(level + 2) % 3] = 0;
        // one frontier node per thread; the warp reserves its children's slots with ONE atomicAdd
        for (int f0 = (tid & ~31); f0 < cur_n; f0 += nth) {
            const int f = f0 + lane;
            int p = -1, s = 0, e = 0;
            uint32_t hp = 0;
            if (f < cur_n) {
                p = cur[f];
                hp = cp_hi(cp[p]);
                s = child_start[p], e = child_start[p + 1];
            }
            int cnt = e - s;
            if (cnt > kTreeBig) { // expanded by the grid after the barrier
                const int k = atomicAdd(&counters[4 + (level & 1)], 1);
                if (k < big_cap) big[k] = p;
                cnt = 0;
            }
            int x = cnt;
            for (int off = 1; off < 32; off <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, x, off);
                if (lane >= off) x += y;
            }
            const int total = __shfl_sync(0xffffffffu, x, 31);
            int base = 0;
            if (lane == 31 && total) base = atomicAdd(cnt_next, total);
            base = __shfl_sync(0xffffffffu, base, 31) + x - cnt;
            if (cnt <= 8)
                for (int j = 0; j < cnt; j++) {
                    const int c = children[s + j];
                    tree_visit(c, p, hp, w, cp, flags);
                    nxt[base + j] = c;
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
                    tree_visit(c, po, ho, w, cp, flags);
                    nxt[bo + j] = c;
                }
            }
            if (lane == 31 && total) atomicAdd(&counters[3], total);
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
// replaces cost|parent of s by atomicMin.  free_only: AM-RRT* RewireRoot relaxes only over free edges and never to a
// colliding cost (AMRRT.cpp:500-504).
// relax phase 2 as a work list: only the edges out of nodes whose cost fell in the previous round can improve
// anything, so a round handles the current frontier (one warp per node, lanes = its k edges), collects the nodes it
// improved (deduplicated by a round stamp) and the whole iteration to quiescence is ONE persistent cooperative kernel
// with a grid-wide barrier per round — the late rounds of a long improving chain touch a handful of nodes and cost a
// barrier instead of a 25 M-edge pass and a host round trip.  Same least fixpoint as the sweeps.
// counters: [0..2] frontier sizes in rotation (current, next, being cleared), [3] rounds done
__global__ void __launch_bounds__(256)
relax_worklist_kernel(const int *__restrict__ knn, const uint32_t *__restrict__ w, int n, int k, int free_only, cpword *cp, int *fr_a, int *fr_b,
                      int *stamp, int *counters, int max_rounds) {
    cg::grid_group grid = cg::this_grid();
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    // round 0: every node that has a cost
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) fr_a[i] = i, stamp[i] = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) counters[0] = n, counters[1] = 0, counters[2] = 0, counters[3] = 0;
    grid.sync();
    int *cur = fr_a, *nxt = fr_b;
    int round = 0;
    for (; round < max_rounds; round++) {
        const int m = ((volatile int *)counters)[round % 3];
        if (m == 0) break;
        if (blockIdx.x == 0 && threadIdx.x == 0) counters[(round + 2) % 3] = 0, counters[3] = round + 1; // next round's target
        for (int f = warp; f < m; f += nwarps) {
            const int x = cur[f];
            const uint32_t hx = cp_hi(((volatile cpword *)cp)[x]);
            if (hx == 0xffffffffu) continue; // x has no cost yet
            for (int j0 = 0; j0 < k; j0 += 32) {
                const int j = j0 + lane;
                bool improved = false;
                int sn = -1;
                if (j < k) {
                    const size_t e = (size_t)x * k + j;
                    const uint32_t we = w[e];
                    if (we != 0xffffffffu && !(free_only && (we >> 31))) {
                        sn = knn[e];
                        const uint32_t hn = cost_add(hx, we);
                        if (!(free_only && (hn >> 31)) && hn < cp_hi(((volatile cpword *)cp)[sn])) {
                            const cpword neww = cp_make(hn, x);
                            const cpword prev = atomicMin(&cp[sn], neww);
                            improved = prev > neww && cp_hi(prev) > hn;
                        }
                    }
                }
                // a node enters the next frontier once per round
                if (improved) improved = atomicExch(&stamp[sn], round + 1) != round + 1;
                const unsigned bal = __ballot_sync(0xffffffffu, improved);
                if (bal) {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&counters[(round + 1) % 3], __popc(bal));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (improved) nxt[base + __popc(bal & ((1u << lane) - 1u))] = sn;
                }
            }
        }
        int *t = cur;
        cur = nxt, nxt = t;
        grid.sync();
    }
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
        RTPP_CK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, relax_worklist_kernel, 256, 0));
        int blocks = ctx->sm_count * (per_sm < 1 ? 1 : (per_sm > 2 ? 2 : per_sm)); // few blocks: the barrier is the cost of a late round
        const int need = (int)(((size_t)s->n * 32 + 255) / 256);
        if (blocks > need) blocks = need < 1 ? 1 : need;
        const int *knn = d_knn_idx;
        const uint32_t *ww = (const uint32_t *)dw;
        cpword *cp = s->cp;
        int nn = s->n, kk = k, fo = free_only, mr = max_sweeps;
        int *fa = s->frontier, *fb = s->frontier + s->capacity, *st = s->children, *cn = s->counters;
        void *args[] = {&knn, &ww, &nn, &kk, &fo, &cp, &fa, &fb, &st, &cn, &mr};
        RTPP_CK(ctx, cudaLaunchCooperativeKernel((void *)relax_worklist_kernel, dim3(blocks), dim3(256), args, 0, ctx->stream));
        ctx->launches++;
        RTPP_CK(ctx, cudaMemcpyAsync(&sweeps, s->counters + 3, 4, cudaMemcpyDeviceToHost, ctx->stream));
        RTPP_CK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (sweeps_done) *sweeps_done = sweeps;
    return RTPP_OK;
}

} // extern "C"
