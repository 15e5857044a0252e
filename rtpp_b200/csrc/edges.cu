// edges.cu — K1: batched edge collision checks by warp-cooperative dyadic ray marching.
//
// Replaces (bit-exactly) the reference's scalar routines
//   E1 RTRRTplanner::CrashDetection        RTRRT.cpp:202-222  (= RTFMT.cpp:242-262), divisor 5
//   E2 RTFMTplanner::CrashDetectionNDObs   RTFMT.cpp:264-283, static plane only
//   E3 AMRRTplanner::CrashDetection        AMRRT.cpp:266-286, divisor 2
//   E4 AMRRTplanner::CrashDetectionStart   AMRRT.cpp:287-308, linear march + last free point
//   E5b BImap::RTRT_CrashDetection         CDRT.cpp:584-601, binary64 points, divisor 10
//
// Mapping: a warp owns 32 consecutive edges.  Each lane loads its edge with one streaming 16-byte
// load, tests the two end pixels and marches its own interior probes while the edge is short
// (<= 15 interior probes: the k-NN / tree-edge case).  Edges with >= 31 interior probes are then
// marched by the whole warp, lane-strided over the probes and OR-reduced with __any_sync.  The 32
// verdicts leave as one ballot word (4 coalesced bytes per warp).  The verdict is a pure OR over
// the probe set, so probe order is free (the reference's early exit cannot change it).
//
// Roofline: algorithmic bytes per edge = 16 + 4*n_probes + 1/8; the edge list streams
// from HBM, the bit-plane (0.9-2 MB) is L2/L1 resident.
#include <stdlib.h>

#include "nodes.cuh"

static constexpr int kShortMaxP = 16; // P <= 16 -> thread marches alone; P >= 32 -> warp marches

// The warp-level body of K1: every lane brings one edge (e = ps.x, ps.y, pe.x, pe.y; valid = false
// for padding lanes) and gets the ballot word of the 32 verdicts.
__device__ __forceinline__ unsigned warp_edges_verdict(const PlaneView &pl, const float4 e, const bool valid, const MarchThresholds &mt) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    uint32_t acc = 0; // bit 0 = blocked
    int P = 1;
    if (valid) {
        P = rt_march_P(mt, rt_sqlen(e.x, e.y, e.z, e.w));
        if (rt_edge_safe(pl, e.x, e.y, e.z, e.w)) {
            // The reference's probe set is the dyadic levels t = 1/2 | 1/4, 3/4 | odd/8 | odd/16 ... up to P
            // (RTRRT.cpp:210-218); the verdict is their OR, so the levels are issued as straight-line groups
            // with immediate t and 1-t (both exact): no loop counter, no t accumulation, and the first level
            // (every edge has it: M >= 2) goes out together with the two end-point probes.
            auto at = [&](const float t, const float omt) -> uint32_t {
                const int x = rt_trunc_pos(__fadd_rn(__fmul_rn(e.x, t), __fmul_rn(omt, e.z)));
                const int y = rt_trunc_pos(__fadd_rn(__fmul_rn(e.y, t), __fmul_rn(omt, e.w)));
                return rt_probe_word(pl, y, x);
            };
            acc = rt_probe_word(pl, rt_trunc_pos(e.y), rt_trunc_pos(e.x)) | rt_probe_word(pl, rt_trunc_pos(e.w), rt_trunc_pos(e.z)) |
                  at(0.5f, 0.5f);
            if (P >= 4 && P <= kShortMaxP) {
                acc |= at(0.25f, 0.75f) | at(0.75f, 0.25f);
                if (P >= 8) {
                    acc |= at(0.125f, 0.875f) | at(0.375f, 0.625f) | at(0.625f, 0.375f) | at(0.875f, 0.125f);
                    if (P >= 16)
                        acc |= at(0.0625f, 0.9375f) | at(0.1875f, 0.8125f) | at(0.3125f, 0.6875f) | at(0.4375f, 0.5625f) |
                               at(0.5625f, 0.4375f) | at(0.6875f, 0.3125f) | at(0.8125f, 0.1875f) | at(0.9375f, 0.0625f);
                }
            }
            acc &= 1u;
        } else {
            acc = rt_probe(pl, __float2int_rz(e.y), __float2int_rz(e.x)) | rt_probe(pl, __float2int_rz(e.w), __float2int_rz(e.z));
            if (!acc && P <= kShortMaxP) {
                const float inv = rt_inv_pow2(P);
                for (int j = 1; j < P; j++) acc |= rt_probe_interior(pl, e.x, e.y, e.z, e.w, __fmul_rn((float)j, inv));
            }
        }
    }
    unsigned longmask = __ballot_sync(full, !acc && P > kShortMaxP);
    while (longmask) {
        const int src = __ffs(longmask) - 1;
        longmask &= longmask - 1;
        const float sx = __shfl_sync(full, e.x, src), sy = __shfl_sync(full, e.y, src);
        const float tx = __shfl_sync(full, e.z, src), ty = __shfl_sync(full, e.w, src);
        const int Ps = __shfl_sync(full, P, src);
        const float inv = rt_inv_pow2(Ps);
        bool hit = false;
        for (int j0 = 1; j0 < Ps; j0 += 32) { // warp-uniform trip count, early exit per chunk
            const int j = j0 + lane;
            if (j < Ps) hit = rt_probe_interior(pl, sx, sy, tx, ty, __fmul_rn((float)j, inv));
            if (__any_sync(full, hit)) {
                hit = true;
                break;
            }
        }
        if (lane == src) acc = hit;
    }
    return __ballot_sync(full, acc != 0);
}

// chunk_iters == 0: persistent grid-stride (one resident wave); > 0: block b owns the 8*chunk_iters
// consecutive 32-edge words starting at b*8*chunk_iters (warps interleaved, so each step of the block
// reads 4 KB contiguous) and the hardware block scheduler balances the tail.
__global__ void __launch_bounds__(256, 8)
edge_check_f32_kernel(PlaneView pl, const float4 *__restrict__ edges, size_t n, MarchThresholds mt,
                      uint32_t *__restrict__ out_bits, int chunk_iters) {
    const int lane = threadIdx.x & 31;
    const size_t nwords = (n + 31) >> 5;
    size_t w, stride, end;
    if (chunk_iters > 0) {
        const size_t base = (size_t)blockIdx.x * (blockDim.x >> 5) * chunk_iters;
        w = base + (threadIdx.x >> 5), stride = blockDim.x >> 5;
        end = min(nwords, base + stride * chunk_iters);
    } else {
        w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, stride = ((size_t)gridDim.x * blockDim.x) >> 5, end = nwords;
    }
    // software pipeline: the next 16-byte edge of this lane is in flight while the current one is marched
    float4 nxt = make_float4(0.f, 0.f, 0.f, 0.f);
    if (w < end && (w << 5) + lane < n) nxt = ld_stream_f4(edges + (w << 5) + lane);
    for (; w < end; w += stride) {
        const size_t i = (w << 5) + lane;
        const float4 e = nxt;
        const size_t in = ((w + stride) << 5) + lane;
        if (w + stride < end && in < n) nxt = ld_stream_f4(edges + in);
        const unsigned bits = warp_edges_verdict(pl, e, i < n, mt);
        if (lane == 0) out_bits[w] = bits;
    }
}

// E5b: binary64 end points, binary32 interpolation parameter (CDRT.cpp:586-600).
__device__ __forceinline__ bool probe_interior_f64(const PlaneView &p, double psx, double psy, double pex, double pey,
                                                   float t) {
    const float omt = __fsub_rn(1.0f, t);
    const double td = (double)t, od = (double)omt;
    int x = __double2int_rz(__dadd_rn(__dmul_rn(psx, td), __dmul_rn(od, pex)));
    int y = __double2int_rz(__dadd_rn(__dmul_rn(psy, td), __dmul_rn(od, pey)));
    return rt_probe(p, y, x);
}

__global__ void __launch_bounds__(256)
edge_check_f64_kernel(PlaneView pl, const double4 *__restrict__ edges, size_t n, float divisor,
                      uint32_t *__restrict__ out_bits) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const size_t warp0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    const size_t nwords = (n + 31) >> 5;
    for (size_t w = warp0; w < nwords; w += nwarps) {
        const size_t i = (w << 5) + lane;
        bool blocked = false;
        int P = 1;
        double4 e = make_double4(0., 0., 0., 0.);
        if (i < n) {
            e = edges[i];
            blocked = rt_probe(pl, __double2int_rz(e.y), __double2int_rz(e.x)) |
                      rt_probe(pl, __double2int_rz(e.w), __double2int_rz(e.z));
            const float len = __double2float_rn(rt_dist64(e.x, e.y, e.z, e.w)); // "float len = ps % pe"
            P = rt_pow2_floor(rt_march_M(len, divisor));
        }
        if (!blocked && P > 1 && P <= kShortMaxP) {
            const float inv = 1.0f / (float)P;
            for (int j = 1; j < P; j++) blocked |= probe_interior_f64(pl, e.x, e.y, e.z, e.w, __fmul_rn((float)j, inv));
        }
        unsigned longmask = __ballot_sync(full, !blocked && P > kShortMaxP);
        while (longmask) {
            const int src = __ffs(longmask) - 1;
            longmask &= longmask - 1;
            const double sx = __shfl_sync(full, e.x, src), sy = __shfl_sync(full, e.y, src);
            const double tx = __shfl_sync(full, e.z, src), ty = __shfl_sync(full, e.w, src);
            const int Ps = __shfl_sync(full, P, src);
            const float inv = 1.0f / (float)Ps;
            bool hit = false;
            for (int j0 = 1; j0 < Ps; j0 += 32) {
                const int j = j0 + lane;
                if (j < Ps) hit = probe_interior_f64(pl, sx, sy, tx, ty, __fmul_rn((float)j, inv));
                if (__any_sync(full, hit)) {
                    hit = true;
                    break;
                }
            }
            if (lane == src) blocked = hit;
        }
        const unsigned bits = __ballot_sync(full, blocked);
        if (lane == 0) out_bits[w] = bits;
    }
}

// E4: linear march i = 1..M-1, t = (float)i/M weighting pe; needs the FIRST blocked index
// (a min-reduction, not an OR).  One warp per edge: lanes take 32 consecutive i, the first chunk
// with a hit gives f = base + ffs(ballot) - 1; pstart = point(f-1) (or ps when f == 1).
__device__ __forceinline__ float2 first_hit_point(float psx, float psy, float pex, float pey, int i, float Mf) {
    const float t = __fdiv_rn((float)i, Mf);
    const float omt = __fsub_rn(1.0f, t);
    float2 r;
    r.x = __fadd_rn(__fmul_rn(pex, t), __fmul_rn(omt, psx));
    r.y = __fadd_rn(__fmul_rn(pey, t), __fmul_rn(omt, psy));
    return r;
}

__global__ void __launch_bounds__(256)
edge_first_hit_kernel(PlaneView pl, const float4 *__restrict__ edges, size_t n, uint32_t *__restrict__ out_bits,
                      float2 *__restrict__ pstart) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const size_t warp0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    const size_t nwords = (n + 31) >> 5;
    for (size_t w = warp0; w < nwords; w += nwarps) {
        unsigned bits = 0;
        const size_t base = w << 5;
        const int cnt = (int)min((size_t)32, n - base);
        // lane k preloads edge base+k; the warp then walks the 32 edges one at a time
        float4 mine = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lane < cnt) mine = ld_stream_f4(edges + base + lane);
        for (int k = 0; k < cnt; k++) {
            const float psx = __shfl_sync(full, mine.x, k), psy = __shfl_sync(full, mine.y, k);
            const float pex = __shfl_sync(full, mine.z, k), pey = __shfl_sync(full, mine.w, k);
            bool blocked = rt_probe(pl, __float2int_rz(psy), __float2int_rz(psx)) |
                           rt_probe(pl, __float2int_rz(pey), __float2int_rz(pex));
            float2 ps = make_float2(psx, psy);
            if (!blocked) {
                const int M = rt_march_M(rt_dist(psx, psy, pex, pey), 2.0f);
                const float Mf = (float)M;
                int f = 0; // first blocked index, 0 = none
                for (int i0 = 1; i0 < M; i0 += 32) {
                    const int i = i0 + lane;
                    bool hit = false;
                    if (i < M) {
                        const float2 q = first_hit_point(psx, psy, pex, pey, i, Mf);
                        hit = rt_probe(pl, __float2int_rz(q.y), __float2int_rz(q.x));
                    }
                    const unsigned hm = __ballot_sync(full, hit);
                    if (hm) {
                        f = i0 + __ffs(hm) - 1;
                        break;
                    }
                }
                if (f) {
                    blocked = true;
                    if (f > 1) ps = first_hit_point(psx, psy, pex, pey, f - 1, Mf);
                } else
                    ps = make_float2(pex, pey);
            }
            if (blocked) bits |= 1u << k;
            if (lane == 0) pstart[base + k] = ps;
        }
        if (lane == 0) out_bits[w] = bits;
    }
}

// Tick-level form of K1: the edges are implied by the resident tree — edge e < n-1 is the tree edge
// (parent(e+1) -> e+1) of UpdateTree (AMRRT.cpp:185-199), edge n-1 + i*k + j is the rewire edge
// (node i -> knn[i][j]) of RewireRoot (AMRRT.cpp:496-500) — so a tick moves no edge list at all:
// 4 index bytes + two 8-byte coordinate gathers (L2-resident node table) per edge.
__global__ void __launch_bounds__(256, 6)
edge_check_resident_kernel(PlaneView pl, const float2 *__restrict__ xy, const float2 *__restrict__ sxy, const cpword *__restrict__ cp,
                           const int *__restrict__ knn_pos, int n, int k, MarchThresholds mt, uint32_t *__restrict__ out_bits,
                           int chunk_iters) {
    const int lane = threadIdx.x & 31;
    const size_t n_tree = (size_t)(n - 1), n_edges = n_tree + (size_t)n * k;
    const size_t nwords = (n_edges + 31) >> 5;
    const bool small = n_edges < 0xffffffffull; // 32-bit row arithmetic (always, up to 1.6e8 nodes at k = 25)
    const size_t base = (size_t)blockIdx.x * (blockDim.x >> 5) * chunk_iters;
    const size_t stride = blockDim.x >> 5, end = min(nwords, base + stride * chunk_iters);
    size_t w = base + (threadIdx.x >> 5);
    // Two-stage software pipeline over the block's chunk: the node ids of word w+2 and the coordinates of
    // word w+1 are in flight while word w is marched (the id -> coordinate gather is a dependent pair of
    // L2 accesses; issued a full step apart, neither stalls the march).
    auto load_ids = [&](size_t ww) -> int2 { // (a, b): b >= 0 indexes xy (tree edge), b <= -2 encodes sorted position ~b... -1 = none
        const size_t i = (ww << 5) + lane;
        int2 r = make_int2(-1, -1);
        if (ww < end && i < n_edges) {
            if (i < n_tree) {
                const int b = (int)i + 1, a = cp_parent(cp[b]);
                if (a >= 0) r = make_int2(a, b);
            } else {
                const size_t q = i - n_tree;
                const int a = small ? (int)((unsigned)q / (unsigned)k) : (int)(q / (unsigned)k);
                const int bp = knn_pos[q];
                if (bp >= 0) r = make_int2(a, -2 - bp);
            }
        }
        return r;
    };
    auto load_xy = [&](int2 id) -> float4 {
        if (id.x < 0) return make_float4(0.f, 0.f, 0.f, 0.f);
        const float2 pa = xy[id.x], pb = id.y >= 0 ? xy[id.y] : sxy[-2 - id.y];
        return make_float4(pa.x, pa.y, pb.x, pb.y);
    };
    int2 id1 = load_ids(w), id2 = load_ids(w + stride);
    float4 e1 = load_xy(id1);
    bool v1 = id1.x >= 0;
    for (; w < end; w += stride) {
        const float4 e = e1;
        const bool valid = v1;
        e1 = load_xy(id2), v1 = id2.x >= 0;
        id2 = load_ids(w + 2 * stride);
        const unsigned bits = warp_edges_verdict(pl, e, valid, mt);
        if (lane == 0) out_bits[w] = bits;
    }
}

__global__ void __launch_bounds__(256) dist_kernel(const float4 *__restrict__ e, size_t n, float *__restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float4 v = e[i];
        out[i] = rt_dist(v.x, v.y, v.z, v.w);
    }
}

// March thresholds (see common.cuh): smallest squared length s with M(s) >= 4, 8, 16, 32, found by
// bisection over the float bit patterns with the same IEEE binary32 operations as the kernels
// (host code is built without FMA contraction or fast-math).
static int host_march_M(float s, int divisor) {
    volatile float len = sqrtf(s);
    volatile float q = len / (float)divisor;
    volatile float v = q + 2.0f;
    return (int)v;
}
void rtpp_march_thresholds(int divisor, MarchThresholds *out) {
    for (int m = 0; m < 4; m++) {
        const int target = 4 << m;
        const float s_hi = 1.0e12f; // len 1e6: M is far above every target and still fits an int
        uint32_t lo = 0, hi;        // invariant: M(lo) < target <= M(hi)
        memcpy(&hi, &s_hi, 4);
        while (hi - lo > 1) {
            const uint32_t mid = lo + (hi - lo) / 2;
            float s;
            memcpy(&s, &mid, 4);
            if (host_march_M(s, divisor) >= target) hi = mid;
            else lo = mid;
        }
        memcpy(&out->t[m], &hi, 4);
    }
    out->divisor = (float)divisor;
}

// grid sizing: a multiple of the SM count, enough resident warps to cover L2 latency
// Persistent grid-stride kernels: exactly one resident wave (blocks per SM from the occupancy
// calculator, so a register-count change can never leave a half-empty second wave).
template <class K> static int edge_grid(rtpp_ctx *ctx, K kernel, size_t n, int per_block) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, per_block, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
    size_t need = (n + per_block - 1) / per_block;
    size_t cap = (size_t)ctx->sm_count * per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

extern "C" {

int rtpp_edges_check_dev(rtpp_ctx *ctx, const float *d_xyxy, size_t n, int divisor, int plane_mask, uint32_t *d_bits) {
    if (!ctx || (n && (!d_xyxy || !d_bits))) return RTPP_ERR_INVALID;
    if (divisor <= 0) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "divisor must be positive");
    PlaneView pv;
    int rc = rtpp_select_plane(ctx, plane_mask, &pv);
    if (rc) return rc;
    if (n == 0) return RTPP_OK;
    MarchThresholds mt;
    rtpp_march_thresholds(divisor, &mt);
    // blocks of 8 warps x 8 steps (32 KB of the edge list each): small enough that the concurrently running blocks
    // do not fall into step on the same HBM channels (chunks >= 16 steps measurably do), and the hardware block
    // scheduler evens out the tail that a one-wave persistent grid leaves (0.186 -> 0.170 ms on the C4 step)
    static int chunk = getenv("RTPP_K1_CHUNK") ? atoi(getenv("RTPP_K1_CHUNK")) : 8;
    if (chunk > 0) {
        const size_t per_block = (size_t)8 * chunk * 32;
        edge_check_f32_kernel<<<(unsigned)((n + per_block - 1) / per_block), 256, 0, ctx->stream>>>(pv, (const float4 *)d_xyxy, n, mt, d_bits, chunk);
    } else
        edge_check_f32_kernel<<<edge_grid(ctx, edge_check_f32_kernel, n, 256), 256, 0, ctx->stream>>>(pv, (const float4 *)d_xyxy, n, mt, d_bits, 0);
    RTPP_LAUNCHED(ctx);
    return RTPP_OK;
}

int rtpp_tick_edges_check_dev(rtpp_ctx *ctx, int divisor, int plane_mask, uint32_t *d_bits, size_t *n_edges) {
    if (!ctx || !d_bits) return RTPP_ERR_INVALID;
    NodeStore *s = ctx->nodes;
    if (!s || !s->knn_table) RTPP_FAIL(ctx, RTPP_ERR_STATE, "rtpp_knn_table_build has not been called");
    if (s->knn_n != s->n || s->n_sorted != s->n) RTPP_FAIL(ctx, RTPP_ERR_STATE, "nodes were appended after rtpp_knn_table_build");
    if (divisor <= 0) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "divisor must be positive");
    PlaneView pv;
    int rc = rtpp_select_plane(ctx, plane_mask, &pv);
    if (rc) return rc;
    const size_t ne = (size_t)(s->n - 1) + (size_t)s->n * s->knn_k;
    if (n_edges) *n_edges = ne;
    MarchThresholds mt;
    rtpp_march_thresholds(divisor, &mt);
    const int chunk = 8;
    const size_t per_block = (size_t)8 * chunk * 32;
    edge_check_resident_kernel<<<(unsigned)((ne + per_block - 1) / per_block), 256, 0, ctx->stream>>>(pv, s->xy, s->sxy, s->cp, s->knn_pos, s->n, s->knn_k,
                                                                                                      mt, d_bits, chunk);
    RTPP_LAUNCHED(ctx);
    return RTPP_OK;
}

// One tick through host buffers: new dynamic plane in (NULL = unchanged), verdict words out.
int rtpp_tick_edges_check(rtpp_ctx *ctx, const uint8_t *dynamic_plane, int divisor, int plane_mask, uint32_t *blocked_bits,
                          size_t *n_edges) {
    if (!ctx || !blocked_bits) return RTPP_ERR_INVALID;
    NodeStore *s = ctx->nodes;
    if (!s || !s->knn_table) RTPP_FAIL(ctx, RTPP_ERR_STATE, "rtpp_knn_table_build has not been called");
    int rc;
    if (dynamic_plane) {
        void *d = nullptr;
        rc = rtpp_scratch(ctx, 0, (size_t)ctx->W * ctx->H, &d);
        if (rc) return rc;
        RTPP_CK(ctx, cudaMemcpyAsync(d, dynamic_plane, (size_t)ctx->W * ctx->H, cudaMemcpyHostToDevice, ctx->stream));
        rc = rtpp_map_set_plane_dev(ctx, RTPP_PLANE_DYNAMIC, (const uint8_t *)d, ctx->W, ctx->H);
        if (rc) return rc;
    }
    const size_t ne = (size_t)(s->n - 1) + (size_t)s->n * s->knn_k, nwords = (ne + 31) / 32;
    void *d_b = nullptr;
    rc = rtpp_scratch(ctx, 1, nwords * 4, &d_b);
    if (rc) return rc;
    rc = rtpp_tick_edges_check_dev(ctx, divisor, plane_mask, (uint32_t *)d_b, n_edges);
    if (rc) return rc;
    RTPP_CK(ctx, cudaMemcpyAsync(blocked_bits, d_b, nwords * 4, cudaMemcpyDeviceToHost, ctx->stream));
    RTPP_CK(ctx, cudaStreamSynchronize(ctx->stream));
    return RTPP_OK;
}

int rtpp_edges_check_f64_dev(rtpp_ctx *ctx, const double *d_xyxy, size_t n, int divisor, int plane_mask, uint32_t *d_bits) {
    if (!ctx || (n && (!d_xyxy || !d_bits))) return RTPP_ERR_INVALID;
    if (divisor <= 0) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "divisor must be positive");
    PlaneView pv;
    int rc = rtpp_select_plane(ctx, plane_mask, &pv);
    if (rc) return rc;
    if (n == 0) return RTPP_OK;
    edge_check_f64_kernel<<<edge_grid(ctx, edge_check_f64_kernel, n, 256), 256, 0, ctx->stream>>>(pv, (const double4 *)d_xyxy, n, (float)divisor, d_bits);
    RTPP_LAUNCHED(ctx);
    return RTPP_OK;
}

int rtpp_edges_first_hit_dev(rtpp_ctx *ctx, const float *d_xyxy, size_t n, uint32_t *d_bits, float *d_pstart) {
    if (!ctx || (n && (!d_xyxy || !d_bits || !d_pstart))) return RTPP_ERR_INVALID;
    PlaneView pv;
    int rc = rtpp_select_plane(ctx, 3, &pv);
    if (rc) return rc;
    if (n == 0) return RTPP_OK;
    edge_first_hit_kernel<<<edge_grid(ctx, edge_first_hit_kernel, n, 256), 256, 0, ctx->stream>>>(pv, (const float4 *)d_xyxy, n, d_bits, (float2 *)d_pstart);
    RTPP_LAUNCHED(ctx);
    return RTPP_OK;
}

// ---- host-buffer entry points: copy in, run, copy out, synchronise ----
static int edges_host(rtpp_ctx *ctx, const void *xyxy, size_t n, size_t elem, int divisor, int plane_mask,
                      uint32_t *bits, float *pstart, int kind) {
    if (!ctx || (n && (!xyxy || !bits))) return RTPP_ERR_INVALID;
    if (n == 0) return RTPP_OK;
    RTPP_CK(ctx, cudaSetDevice(ctx->device));
    void *d_e = nullptr, *d_b = nullptr, *d_p = nullptr;
    const size_t nwords = (n + 31) / 32;
    int rc = rtpp_scratch(ctx, 0, n * 4 * elem, &d_e);
    if (rc) return rc;
    rc = rtpp_scratch(ctx, 1, nwords * 4, &d_b);
    if (rc) return rc;
    RTPP_CK(ctx, cudaMemcpyAsync(d_e, xyxy, n * 4 * elem, cudaMemcpyHostToDevice, ctx->stream));
    if (kind == 0) rc = rtpp_edges_check_dev(ctx, (const float *)d_e, n, divisor, plane_mask, (uint32_t *)d_b);
    else if (kind == 1) rc = rtpp_edges_check_f64_dev(ctx, (const double *)d_e, n, divisor, plane_mask, (uint32_t *)d_b);
    else {
        rc = rtpp_scratch(ctx, 2, n * 8, &d_p);
        if (rc) return rc;
        rc = rtpp_edges_first_hit_dev(ctx, (const float *)d_e, n, (uint32_t *)d_b, (float *)d_p);
    }
    if (rc) return rc;
    RTPP_CK(ctx, cudaMemcpyAsync(bits, d_b, nwords * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (kind == 2) RTPP_CK(ctx, cudaMemcpyAsync(pstart, d_p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    RTPP_CK(ctx, cudaStreamSynchronize(ctx->stream));
    return RTPP_OK;
}

int rtpp_edges_check(rtpp_ctx *ctx, const float *xyxy, size_t n, int divisor, int plane_mask, uint32_t *bits) {
    return edges_host(ctx, xyxy, n, sizeof(float), divisor, plane_mask, bits, nullptr, 0);
}
int rtpp_edges_check_f64(rtpp_ctx *ctx, const double *xyxy, size_t n, int divisor, int plane_mask, uint32_t *bits) {
    return edges_host(ctx, xyxy, n, sizeof(double), divisor, plane_mask, bits, nullptr, 1);
}
int rtpp_edges_first_hit(rtpp_ctx *ctx, const float *xyxy, size_t n, uint32_t *bits, float *pstart) {
    if (n && !pstart) return RTPP_ERR_INVALID;
    return edges_host(ctx, xyxy, n, sizeof(float), 2, 3, bits, pstart, 2);
}

int rtpp_dist(rtpp_ctx *ctx, const float *xyxy, size_t n, float *out) {
    if (!ctx || (n && (!xyxy || !out))) return RTPP_ERR_INVALID;
    if (n == 0) return RTPP_OK;
    void *d_e = nullptr, *d_o = nullptr;
    int rc = rtpp_scratch(ctx, 0, n * 16, &d_e);
    if (rc) return rc;
    rc = rtpp_scratch(ctx, 1, n * 4, &d_o);
    if (rc) return rc;
    RTPP_CK(ctx, cudaMemcpyAsync(d_e, xyxy, n * 16, cudaMemcpyHostToDevice, ctx->stream));
    dist_kernel<<<edge_grid(ctx, dist_kernel, n, 256), 256, 0, ctx->stream>>>((const float4 *)d_e, n, (float *)d_o);
    RTPP_LAUNCHED(ctx);
    RTPP_CK(ctx, cudaMemcpyAsync(out, d_o, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    RTPP_CK(ctx, cudaStreamSynchronize(ctx->stream));
    return RTPP_OK;
}

} // extern "C"
