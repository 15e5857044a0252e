// grid.cu — K2/K3: GPU-resident node store, uniform-grid spatial hash, exact k-NN.
//
// Replaces the reference's CGAL kd-tree: Kd_tree::insert (CGAL/Kd_tree.h:461-466; call sites
// RTRRT.cpp:29,291, RTFMT.cpp:29,417, AMRRT.cpp:31,474, CDRT.cpp:918) — which drops the whole tree
// and rebuilds it O(N log N) on the next query — and Orthogonal_k_neighbor_search (see knn.cuh).
//
// Node store: float2 xy[capacity], node id = insertion index.  The hash is a counting sort of the
// nodes by power-of-two cell (cell_start[] + cell-sorted copies of xy and id).  Appends land in an
// unsorted tail that every query scans; the hash is re-sorted lazily when the tail exceeds
// kTailMax (a 1e6-point counting sort is tens of microseconds on B200).
#include <math_constants.h>

#include "knn.cuh"
#include "nodes.cuh"
#include "scan.cuh"

static constexpr int kTailMax = 512;

__global__ void __launch_bounds__(256) grid_count_kernel(const float2 *__restrict__ xy, int n, int shift, int gw, int gh,
                                                        int *__restrict__ cell_of, int *__restrict__ count) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float2 p = xy[i];
        const int cx = grid_cell_of((double)p.x, shift, gw - 1), cy = grid_cell_of((double)p.y, shift, gh - 1);
        const int c = cy * gw + cx;
        cell_of[i] = c;
        atomicAdd(&count[c], 1);
    }
}

__global__ void __launch_bounds__(256) grid_scatter_kernel(const float2 *__restrict__ xy, const int *__restrict__ cell_of,
                                                          int n, int *__restrict__ cursor, float2 *__restrict__ sxy,
                                                          int *__restrict__ sid) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int pos = atomicAdd(&cursor[cell_of[i]], 1);
        sxy[pos] = xy[i];
        sid[pos] = i;
    }
}

// one warp per query; rows of idx/d2 have k entries (idx -1 / d2 +inf padded), count = min(k, n)
__global__ void __launch_bounds__(256) knn_query_kernel(GridView g, const double2 *__restrict__ q, size_t nq, int k,
                                                       int *__restrict__ idx, double *__restrict__ d2,
                                                       int *__restrict__ count) {
    const int lane = threadIdx.x & 31;
    const size_t warp0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    __shared__ KnnScratch scratch[8];
    for (size_t w = warp0; w < nq; w += nwarps) {
        const double2 qq = q[w];
        KnnEntry e;
        int found = knn_warp_block3(g, qq.x, qq.y, k, e, scratch[threadIdx.x >> 5]);
        if (found < 0) found = knn_warp_wide(g, qq.x, qq.y, k, e);
        if (lane < k) {
            idx[w * k + lane] = lane < found ? e.id : -1;
            if (d2) d2[w * k + lane] = lane < found ? e.d2 : CUDART_INF;
        }
        if (count && lane == 0) count[w] = found;
    }
}

// all-kNN of the node set itself (RT-FMT's static table, C4's rewire sweep): a warp takes batches of
// KNN_NODES_BATCH consecutive nodes in CELL-SORTED order, so its queries share their 3x3 block of cells
// (candidates cached in shared memory, knn.cuh), and writes row sid[w] of the table.
#define KNN_NODES_BATCH 16
#define KNN_NODES_WARPS 4
__global__ void __launch_bounds__(32 * KNN_NODES_WARPS) knn_nodes_kernel(GridView g, int k, int *__restrict__ idx, double *__restrict__ d2) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const size_t warp0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    __shared__ KnnBlockCache cache[KNN_NODES_WARPS];
    const bool fast = g.n_sorted == g.n && k <= 28 && g.n >= k;
    KnnBlockState st;
    st.cx = st.cy = -1, st.L = -1, st.T0 = 0.f;
    st.xlo = st.xhi = st.ylo = st.yhi = 0.f;
    const size_t nbatch = ((size_t)g.n + KNN_NODES_BATCH - 1) / KNN_NODES_BATCH;
    for (size_t b = warp0; b < nbatch; b += nwarps) {
        const size_t wend = min((size_t)g.n, (b + 1) * KNN_NODES_BATCH);
        for (size_t w = b * KNN_NODES_BATCH; w < wend; w++) {
            int node;
            float2 p;
            if (w < (size_t)g.n_sorted) {
                node = g.sid[w];
                p = g.sxy[w];
            } else {
                node = (int)w;
                p = g.xy[w];
            }
            KnnEntry e;
            int found = -1;
            if (fast) {
                const int cx = grid_cell_of_f32(p.x, g.shift, g.gw - 1), cy = grid_cell_of_f32(p.y, g.shift, g.gh - 1);
                if (cx != st.cx || cy != st.cy) knn_block_fill(g, cx, cy, k, cache[wib], st);
                if (st.L > 0) found = knn_warp_block3_cached(p.x, p.y, k, e, cache[wib], st);
            }
            if (found < 0) found = knn_warp(g, (double)p.x, (double)p.y, k, e);
            if (lane < k) {
                idx[(size_t)node * k + lane] = lane < found ? e.id : -1;
                if (d2) d2[(size_t)node * k + lane] = lane < found ? e.d2 : CUDART_INF;
            }
        }
    }
}

void rtpp_nodes_destroy(rtpp_ctx *ctx) {
    NodeStore *s = ctx->nodes;
    if (!s) return;
    rtpp_tick_plan_free(s);
    cudaFree(s->xy), cudaFree(s->sxy), cudaFree(s->sid), cudaFree(s->cell_of), cudaFree(s->cell_start), cudaFree(s->cursor), cudaFree(s->gap);
    cudaFree(s->knn_table), cudaFree(s->knn_pos), cudaFree(s->pos_of), cudaFree(s->edge_list), cudaFree(s->scan_totals), cudaFree(s->cp), cudaFree(s->edge_w), cudaFree(s->child_start), cudaFree(s->children), cudaFree(s->frontier), cudaFree(s->counters);
    delete s;
    ctx->nodes = nullptr;
}

static int grid_for(rtpp_ctx *ctx, size_t n) {
    size_t need = (n + 255) / 256;
    size_t cap = (size_t)ctx->sm_count * 8;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

// (re)sort the hash so that it covers every node appended so far
// gap[c] = the first ring around cell c that holds a sorted node (0: the cell itself), saturating at 255: the ring walk of
// a query starts there.  Ring r = rows cy-r and cy+r (each ONE range of the cell-sorted array) + 2(2r-1) side cells.
__global__ void __launch_bounds__(256) grid_gap_kernel(const int *__restrict__ cell_start, int gw, int gh, unsigned char *__restrict__ gap) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= gw * gh) return;
    const int cx = c % gw, cy = c / gw;
    const int rmax = min(255, max(max(cx, gw - 1 - cx), max(cy, gh - 1 - cy)));
    int r = 0;
    for (; r < rmax; r++) {
        const int x0 = max(cx - r, 0), x1 = min(cx + r, gw - 1);
        bool any = false;
        if (cy - r >= 0) any |= cell_start[(cy - r) * gw + x1 + 1] > cell_start[(cy - r) * gw + x0];
        if (r > 0 && cy + r < gh) any |= cell_start[(cy + r) * gw + x1 + 1] > cell_start[(cy + r) * gw + x0];
        for (int y = max(cy - r + 1, 0); y <= min(cy + r - 1, gh - 1) && !any; y++) {
            if (cx - r >= 0) any |= cell_start[y * gw + cx - r + 1] > cell_start[y * gw + cx - r];
            if (cx + r < gw) any |= cell_start[y * gw + cx + r + 1] > cell_start[y * gw + cx + r];
        }
        if (any) break;
    }
    gap[c] = (unsigned char)r;
}

int rtpp_nodes_build_internal(rtpp_ctx *ctx) {
    NodeStore *s = ctx->nodes;
    if (!s) RTPP_FAIL(ctx, RTPP_ERR_STATE, "rtpp_nodes_reset has not been called");
    if (s->n_sorted == s->n) return RTPP_OK;
    const int m = s->gw * s->gh;
    RTPP_CK(ctx, cudaMemsetAsync(s->cursor, 0, (size_t)(m + 1) * 4, ctx->stream));
    grid_count_kernel<<<grid_for(ctx, s->n), 256, 0, ctx->stream>>>(s->xy, s->n, s->shift, s->gw, s->gh, s->cell_of, s->cursor);
    RTPP_LAUNCHED(ctx);
    // cursor currently holds the counts; scan them into cell_start and reset cursor to the starts
    {
        int rc = rtpp_exclusive_scan(ctx, s->cursor, m, s->cell_start, s->cursor, s->scan_totals);
        if (rc) return rc;
    }
    grid_scatter_kernel<<<grid_for(ctx, s->n), 256, 0, ctx->stream>>>(s->xy, s->cell_of, s->n, s->cursor, s->sxy, s->sid);
    RTPP_LAUNCHED(ctx);
    grid_gap_kernel<<<(m + 255) / 256, 256, 0, ctx->stream>>>(s->cell_start, s->gw, s->gh, s->gap);
    RTPP_LAUNCHED(ctx);
    s->n_sorted = s->n;
    return RTPP_OK;
}

int rtpp_grid_view(rtpp_ctx *ctx, GridView *g) {
    NodeStore *s = ctx->nodes;
    if (!s) RTPP_FAIL(ctx, RTPP_ERR_STATE, "rtpp_nodes_reset has not been called");
    if (s->n - s->n_sorted > kTailMax) {
        int rc = rtpp_nodes_build_internal(ctx);
        if (rc) return rc;
    }
    g->sxy = s->sxy, g->sid = s->sid, g->cell_start = s->cell_start, g->xy = s->xy;
    g->gw = s->gw, g->gh = s->gh, g->shift = s->shift, g->n_sorted = s->n_sorted, g->n = s->n;
    g->gap = s->n_sorted > 0 ? s->gap : nullptr;
    return RTPP_OK;
}

extern "C" {

int rtpp_knn_nodes_dev(rtpp_ctx *ctx, int k, int32_t *d_idx, double *d_d2);

int rtpp_nodes_reset(rtpp_ctx *ctx, int cell_log2, size_t capacity, int domain_w, int domain_h) {
    if (!ctx) return RTPP_ERR_INVALID;
    if (cell_log2 < 0 || cell_log2 > 12 || capacity == 0 || capacity > (1u << 30) || domain_w <= 0 || domain_h <= 0)
        RTPP_FAIL(ctx, RTPP_ERR_INVALID, "bad node store geometry");
    RTPP_CK(ctx, cudaSetDevice(ctx->device));
    RTPP_CK(ctx, cudaStreamSynchronize(ctx->stream));
    rtpp_nodes_destroy(ctx);
    NodeStore *s = new (std::nothrow) NodeStore();
    if (!s) return RTPP_ERR_NOMEM;
    ctx->nodes = s;
    s->shift = cell_log2;
    s->gw = ((domain_w - 1) >> cell_log2) + 1;
    s->gh = ((domain_h - 1) >> cell_log2) + 1;
    s->capacity = capacity;
    const size_t m = (size_t)s->gw * s->gh + 1;
    // a failed allocation must not leave a half-built store behind: collect the first error, then tear down
    cudaError_t ce = cudaSuccess;
    auto keep = [&](cudaError_t r) {
        if (ce == cudaSuccess) ce = r;
    };
    keep(cudaMalloc(&s->xy, capacity * sizeof(float2)));
    keep(cudaMalloc(&s->sxy, capacity * sizeof(float2)));
    keep(cudaMalloc(&s->sid, capacity * sizeof(int)));
    keep(cudaMalloc(&s->cell_of, capacity * sizeof(int)));
    keep(cudaMalloc(&s->cell_start, m * sizeof(int)));
    keep(cudaMalloc(&s->cursor, m * sizeof(int)));
    keep(cudaMalloc(&s->gap, m));
    keep(cudaMalloc(&s->scan_totals, ((m > capacity ? m : capacity) / 2048 + 4) * sizeof(int)));
    keep(cudaMalloc(&s->cp, capacity * sizeof(unsigned long long)));
    if (ce == cudaSuccess) keep(cudaMemsetAsync(s->cp, 0xff, capacity * sizeof(unsigned long long), ctx->stream));
    if (ce == cudaSuccess) keep(cudaMemsetAsync(s->cell_start, 0, m * sizeof(int), ctx->stream));
    if (ce != cudaSuccess) {
        rtpp_nodes_destroy(ctx);
        RTPP_CK(ctx, ce);
    }
    return RTPP_OK;
}

int rtpp_nodes_append_dev(rtpp_ctx *ctx, const float *d_xy, size_t n, int32_t *first_id) {
    if (!ctx || (n && !d_xy)) return RTPP_ERR_INVALID;
    NodeStore *s = ctx->nodes;
    if (!s) RTPP_FAIL(ctx, RTPP_ERR_STATE, "rtpp_nodes_reset has not been called");
    if ((size_t)s->n + n > s->capacity) RTPP_FAIL(ctx, RTPP_ERR_CAPACITY, "node store is full");
    if (first_id) *first_id = s->n;
    if (n) RTPP_CK(ctx, cudaMemcpyAsync(s->xy + s->n, d_xy, n * sizeof(float2), cudaMemcpyDeviceToDevice, ctx->stream));
    s->n += (int)n;
    return RTPP_OK;
}

int rtpp_nodes_append(rtpp_ctx *ctx, const float *xy, size_t n, int32_t *first_id) {
    if (!ctx || (n && !xy)) return RTPP_ERR_INVALID;
    NodeStore *s = ctx->nodes;
    if (!s) RTPP_FAIL(ctx, RTPP_ERR_STATE, "rtpp_nodes_reset has not been called");
    if ((size_t)s->n + n > s->capacity) RTPP_FAIL(ctx, RTPP_ERR_CAPACITY, "node store is full");
    if (first_id) *first_id = s->n;
    if (n) {
        RTPP_CK(ctx, cudaMemcpyAsync(s->xy + s->n, xy, n * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
        RTPP_CK(ctx, cudaStreamSynchronize(ctx->stream)); // the host buffer may be reused by the caller
    }
    s->n += (int)n;
    return RTPP_OK;
}

int rtpp_nodes_count(rtpp_ctx *ctx, size_t *n) {
    if (!ctx || !n) return RTPP_ERR_INVALID;
    *n = ctx->nodes ? (size_t)ctx->nodes->n : 0;
    return RTPP_OK;
}

int rtpp_nodes_build(rtpp_ctx *ctx) {
    if (!ctx) return RTPP_ERR_INVALID;
    return rtpp_nodes_build_internal(ctx);
}

int rtpp_nodes_xy_dev(rtpp_ctx *ctx, const float **d_xy) {
    if (!ctx || !d_xy) return RTPP_ERR_INVALID;
    if (!ctx->nodes) RTPP_FAIL(ctx, RTPP_ERR_STATE, "rtpp_nodes_reset has not been called");
    *d_xy = (const float *)ctx->nodes->xy;
    return RTPP_OK;
}

int rtpp_knn_dev(rtpp_ctx *ctx, const double *d_qxy, size_t nq, int k, int32_t *d_idx, double *d_d2, int32_t *d_count) {
    if (!ctx || (nq && (!d_qxy || !d_idx))) return RTPP_ERR_INVALID;
    if (k < 1 || k > 32) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "k must be in 1..32");
    GridView g;
    int rc = rtpp_grid_view(ctx, &g);
    if (rc) return rc;
    if (nq == 0) return RTPP_OK;
    size_t blocks = (nq + 7) / 8;
    size_t cap = (size_t)ctx->sm_count * 8;
    knn_query_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, ctx->stream>>>(g, (const double2 *)d_qxy, nq, k, d_idx, d_d2, d_count);
    RTPP_LAUNCHED(ctx);
    return RTPP_OK;
}

int rtpp_knn(rtpp_ctx *ctx, const double *qxy, size_t nq, int k, int32_t *idx, double *d2, int32_t *count) {
    if (!ctx || (nq && (!qxy || !idx))) return RTPP_ERR_INVALID;
    if (k < 1 || k > 32) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "k must be in 1..32");
    if (nq == 0) return RTPP_OK;
    void *d_q = nullptr, *d_i = nullptr, *d_d = nullptr, *d_c = nullptr;
    int rc = rtpp_scratch(ctx, 0, nq * 16, &d_q);
    if (rc) return rc;
    rc = rtpp_scratch(ctx, 1, nq * k * 4, &d_i);
    if (rc) return rc;
    rc = rtpp_scratch(ctx, 2, nq * k * 8, &d_d);
    if (rc) return rc;
    rc = rtpp_scratch(ctx, 3, nq * 4, &d_c);
    if (rc) return rc;
    RTPP_CK(ctx, cudaMemcpyAsync(d_q, qxy, nq * 16, cudaMemcpyHostToDevice, ctx->stream));
    rc = rtpp_knn_dev(ctx, (const double *)d_q, nq, k, (int32_t *)d_i, (double *)d_d, (int32_t *)d_c);
    if (rc) return rc;
    RTPP_CK(ctx, cudaMemcpyAsync(idx, d_i, nq * k * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (d2) RTPP_CK(ctx, cudaMemcpyAsync(d2, d_d, nq * k * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (count) RTPP_CK(ctx, cudaMemcpyAsync(count, d_c, nq * 4, cudaMemcpyDeviceToHost, ctx->stream));
    RTPP_CK(ctx, cudaStreamSynchronize(ctx->stream));
    return RTPP_OK;
}

__global__ void __launch_bounds__(256) grid_pos_of_kernel(const int *__restrict__ sid, int n, int *__restrict__ pos_of) {
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) pos_of[sid[p]] = p;
}
__global__ void __launch_bounds__(256) grid_knn_pos_kernel(const int *__restrict__ knn, size_t m, const int *__restrict__ pos_of,
                                                          int *__restrict__ knn_pos) {
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < m; e += (size_t)gridDim.x * blockDim.x) {
        const int b = knn[e];
        knn_pos[e] = b >= 0 ? pos_of[b] : -1;
    }
}

// resident all-k-NN table of the node set, kept inside the ctx for the tick-level calls
int rtpp_knn_table_build(rtpp_ctx *ctx, int k) {
    if (!ctx) return RTPP_ERR_INVALID;
    NodeStore *s = ctx->nodes;
    if (!s) RTPP_FAIL(ctx, RTPP_ERR_STATE, "rtpp_nodes_reset has not been called");
    if (k < 1 || k > 32) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "k must be in 1..32");
    if (s->knn_table && (s->knn_k != k || s->knn_n < s->n)) {
        RTPP_CK(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(s->knn_table), cudaFree(s->knn_pos);
        s->knn_table = s->knn_pos = nullptr;
    }
    if (!s->knn_table) {
        RTPP_CK(ctx, cudaMalloc(&s->knn_table, (size_t)s->capacity * k * sizeof(int)));
        RTPP_CK(ctx, cudaMalloc(&s->knn_pos, (size_t)s->capacity * k * sizeof(int)));
        if (!s->pos_of) RTPP_CK(ctx, cudaMalloc(&s->pos_of, (size_t)s->capacity * sizeof(int)));
        s->knn_k = k;
    }
    s->knn_n = s->n;
    s->edge_list_n = -1;
    rtpp_tick_plan_invalidate(s);
    int rc = rtpp_knn_nodes_dev(ctx, k, s->knn_table, nullptr); // also re-sorts the hash: sid covers every node
    if (rc) return rc;
    // the same table in cell-sorted positions: neighbours of a node are neighbours in sxy[], so the
    // tick-level edge kernel gathers their coordinates from a few sectors instead of 32
    grid_pos_of_kernel<<<grid_for(ctx, s->n), 256, 0, ctx->stream>>>(s->sid, s->n, s->pos_of);
    RTPP_LAUNCHED(ctx);
    const size_t m = (size_t)s->n * k;
    grid_knn_pos_kernel<<<grid_for(ctx, m), 256, 0, ctx->stream>>>(s->knn_table, m, s->pos_of, s->knn_pos);
    RTPP_LAUNCHED(ctx);
    return RTPP_OK;
}

int rtpp_knn_table_dev(rtpp_ctx *ctx, const int32_t **d_idx, int *k) {
    if (!ctx || !d_idx) return RTPP_ERR_INVALID;
    NodeStore *s = ctx->nodes;
    if (!s || !s->knn_table) RTPP_FAIL(ctx, RTPP_ERR_STATE, "rtpp_knn_table_build has not been called");
    *d_idx = s->knn_table;
    if (k) *k = s->knn_k;
    return RTPP_OK;
}

int rtpp_knn_nodes_dev(rtpp_ctx *ctx, int k, int32_t *d_idx, double *d_d2) {
    if (!ctx || !d_idx) return RTPP_ERR_INVALID;
    if (k < 1 || k > 32) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "k must be in 1..32");
    int rc = rtpp_nodes_build_internal(ctx);
    if (rc) return rc;
    GridView g;
    rc = rtpp_grid_view(ctx, &g);
    if (rc) return rc;
    if (g.n == 0) return RTPP_OK;
    size_t blocks = ((size_t)g.n + KNN_NODES_WARPS * KNN_NODES_BATCH - 1) / (KNN_NODES_WARPS * KNN_NODES_BATCH); // one batch per warp, hardware-scheduled
    size_t cap = (size_t)ctx->sm_count * 128;
    knn_nodes_kernel<<<(int)(blocks < cap ? blocks : cap), 32 * KNN_NODES_WARPS, 0, ctx->stream>>>(g, k, d_idx, d_d2);
    RTPP_LAUNCHED(ctx);
    return RTPP_OK;
}

} // extern "C"
