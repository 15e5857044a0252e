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

#include <algorithm>

#include "edge_march.cuh"
#include "nodes.cuh"

// Block b owns the 8*chunk_iters (<= 64) consecutive 32-edge words starting at b*8*chunk_iters; its 8 warps are
// interleaved over them, so every step of the block reads 4 KB contiguous, and the hardware block scheduler
// balances the tail (a one-wave persistent grid measured 0.186 ms on the C4 step, this 0.170 ms; chunks of >= 16
// steps put the concurrently running blocks in step on the same HBM channels and are measurably slower).
// The block stages its verdict words and writes them as full lines at the end (256 B per block instead of 64
// scattered 4-byte stores), which also makes a mapped host buffer a cheap destination.
// kSentinel: entries whose first coordinate is NaN stand for "no edge" (verdict 0): the materialised tick list.
// kPeers: the check fused with its all-gather (SURVEY 8(e) grain 2): the block's 256-byte burst of verdict words also
// goes, with plain stores over NVLink, to the same word range of every peer GPU's result buffer (CUDA IPC mappings),
// so the exchange rides under the march instead of following it as a collective.
// The ranks take the batch's 2048-edge blocks round-robin (block g belongs to rank g mod world): the cost of an edge
// varies by two orders of magnitude with its length (the C4 batch starts with 1e6 tree edges, 40 k of them thousands of
// pixels long), and contiguous ranges left one of two GPUs with 63 % of the work.
struct PeerBits {
    uint32_t *p[RTPP_MAX_PEERS - 1];
    int n;
    int blk_stride, blk_first; // this launch's block b is block b * blk_stride + blk_first of the batch
    uint32_t *defer_list;      // very long edges, marched by edge_deferred_kernel
    unsigned *defer_count;
};
template <bool kSentinel, bool kPeers = false>
__global__ void __launch_bounds__(256, 8)
edge_check_f32_kernel(PlaneView pl, const float4 *__restrict__ edges, size_t n, MarchThresholds mt,
                      uint32_t *__restrict__ out_bits, int chunk_iters, PeerBits peers = PeerBits()) {
    __shared__ uint32_t stage[64];
    const int lane = threadIdx.x & 31;
    const size_t blk = kPeers ? (size_t)blockIdx.x * peers.blk_stride + peers.blk_first : (size_t)blockIdx.x;
    const size_t base = blk * 8 * chunk_iters; // first word of the block
    const int n_rel = (int)min((size_t)(256 * chunk_iters), n - (base << 5)); // edges of this block
    const float4 *ep = edges + (base << 5) + threadIdx.x;
    // software pipeline through shared memory: the next 16-byte edge of this lane is fetched with cp.async (L2 ->
    // shared, no L1 allocation, no registers held across the march) while the current one is marched; each thread
    // only ever reads the slot it filled itself, so cp.async.wait_group is the only synchronisation
    __shared__ float4 ebuf[2][256];
    if ((int)threadIdx.x < n_rel)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(&ebuf[0][threadIdx.x])), "l"(ep) : "memory");
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    for (int rel = threadIdx.x; rel - lane < n_rel; rel += 256) { // rel = index of this lane's edge inside the block
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rel < n_rel) e = ebuf[(rel >> 8) & 1][threadIdx.x];
        ep += 256;
        if (rel + 256 < n_rel)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(&ebuf[((rel >> 8) & 1) ^ 1][threadIdx.x])), "l"(ep)
                         : "memory");
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        const unsigned bits = kPeers ? warp_edges_verdict(pl, e, rel < n_rel && (!kSentinel || e.x == e.x), mt, peers.defer_list, peers.defer_count,
                                                          (uint32_t)((base << 5) + rel))
                                     : warp_edges_verdict(pl, e, rel < n_rel && (!kSentinel || e.x == e.x), mt);
        if (lane == 0) stage[rel >> 5] = bits;
    }
    __syncthreads();
    if ((int)(threadIdx.x << 5) < n_rel) {
        const uint32_t v = stage[threadIdx.x];
        out_bits[base + threadIdx.x] = v;
        if (kPeers) {
#pragma unroll
            for (int k = 0; k < RTPP_MAX_PEERS - 1; k++)
                if (k < peers.n) peers.p[k][base + threadIdx.x] = v;
        }
    }
}

// One warp per deferred edge: lane-strided probes with an early exit per 32-probe chunk (the same probe set as the
// in-kernel march); a blocked edge ORs its bit into this rank's and every peer's result word.
__global__ void __launch_bounds__(256)
edge_deferred_kernel(PlaneView pl, const float4 *__restrict__ edges, MarchThresholds mt, uint32_t *__restrict__ out_bits, PeerBits peers) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned nwarps = (gridDim.x * blockDim.x) >> 5, count = *peers.defer_count;
    for (unsigned i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < count; i += nwarps) {
        const uint32_t idx = peers.defer_list[i];
        const float4 e = edges[idx];
        const int P = rt_march_P(mt, rt_sqlen(e.x, e.y, e.z, e.w));
        const float inv = rt_inv_pow2(P);
        bool hit = false;
        for (int j0 = 1; j0 < P; j0 += 32) {
            const int j = j0 + lane;
            if (j < P) hit = rt_probe_interior(pl, e.x, e.y, e.z, e.w, __fmul_rn((float)j, inv));
            if (__any_sync(full, hit)) {
                hit = true;
                break;
            }
        }
        if (hit && lane == 0) {
            const uint32_t bit = 1u << (idx & 31);
            atomicOr(out_bits + (idx >> 5), bit);
#pragma unroll
            for (int k = 0; k < RTPP_MAX_PEERS - 1; k++)
                if (k < peers.n) atomicOr(peers.p[k] + (idx >> 5), bit);
        }
    }
}

// After a rank's fused check has finished (stream order), tell every peer (its flag slot `rank` := epoch, system scope)
// and wait until every peer has told us: from then on this rank's result buffer holds the whole batch's verdict words.
// The wait is bounded (rtpp_peer_set_timeout_ms): when a peer never arrives the rank's result buffer is filled with 1-bits —
// every edge reads BLOCKED, fail closed — and the sticky timeout counter of its flag array is raised (rtpp_peer_status).
struct PeerFlags {
    int *p[RTPP_MAX_PEERS];
};
static constexpr int kPeerFlagTimeouts = RTPP_MAX_PEERS; // same slot as tick.cu's kFlagTimeouts
__global__ void __launch_bounds__(256) peer_signal_wait_kernel(PeerFlags flags_all, int rank, int world, int epoch, uint32_t *own_bits, size_t n_words,
                                                              long long spin_limit) {
    const int r = threadIdx.x;
    int *theirs = nullptr, *mine = nullptr, *own = nullptr;
#pragma unroll
    for (int k = 0; k < RTPP_MAX_PEERS; k++) {
        if (k == r) theirs = flags_all.p[k] + rank;
        if (k == rank) mine = flags_all.p[k] + r, own = flags_all.p[k];
    }
    const bool peer = r < world && r != rank;
    bool timed_out = false;
    if (peer) *(volatile int *)theirs = epoch;
    __threadfence_system();
    if (peer) {
        long long it = 0;
        while (*(volatile int *)mine < epoch && it < spin_limit) {
            __nanosleep(64);
            it++;
        }
        timed_out = *(volatile int *)mine < epoch;
    }
    const int any_timeout = __syncthreads_or(timed_out);
    __threadfence_system();
    if (any_timeout) {
        for (size_t i = threadIdx.x; i < n_words; i += blockDim.x) own_bits[i] = 0xffffffffu;
        if (threadIdx.x == 0) atomicAdd(own + kPeerFlagTimeouts, 1);
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

// E4: linear march i = 1..M-1, t = (float)i/M weighting pe; needs the FIRST blocked index (a min-reduction, not an
// OR) and the last free point before it.  A lane owns an edge and walks it in order until the first hit (Steer's
// edges are short: M = len/2 + 2 is a handful of probes, and the walk stops early); edges with more than kFirstHitLane
// interior probes are marched by the whole warp afterwards: lanes take 32 consecutive i, the first chunk with a hit
// gives f = base + ffs(ballot) - 1.  pstart = point(f-1) (ps when f == 1, pe when nothing is hit).
__device__ __forceinline__ float2 first_hit_point(float psx, float psy, float pex, float pey, int i, float Mf) {
    const float t = __fdiv_rn((float)i, Mf);
    const float omt = __fsub_rn(1.0f, t);
    float2 r;
    r.x = __fadd_rn(__fmul_rn(pex, t), __fmul_rn(omt, psx));
    r.y = __fadd_rn(__fmul_rn(pey, t), __fmul_rn(omt, psy));
    return r;
}

constexpr int kFirstHitLane = 48;

__global__ void __launch_bounds__(256)
edge_first_hit_kernel(PlaneView pl, const float4 *__restrict__ edges, size_t n, uint32_t *__restrict__ out_bits,
                      float2 *__restrict__ pstart) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const size_t warp0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    const size_t nwords = (n + 31) >> 5;
    for (size_t w = warp0; w < nwords; w += nwarps) {
        const size_t e = (w << 5) + lane;
        const bool have = e < n;
        float4 mine = make_float4(0.f, 0.f, 0.f, 0.f);
        if (have) mine = ld_stream_f4(edges + e);
        const float psx = mine.x, psy = mine.y, pex = mine.z, pey = mine.w;
        bool blocked = false, longer = false;
        float2 ps = make_float2(psx, psy);
        int M = 0;
        if (have) {
            blocked = rt_probe(pl, __float2int_rz(psy), __float2int_rz(psx)) | rt_probe(pl, __float2int_rz(pey), __float2int_rz(pex));
            if (!blocked) {
                M = rt_march_M(rt_dist(psx, psy, pex, pey), 2.0f);
                longer = M - 1 > kFirstHitLane;
                if (!longer) {
                    const float Mf = (float)M;
                    float2 prev = ps; // the last free point so far
                    for (int i = 1; i < M; i++) {
                        const float2 q = first_hit_point(psx, psy, pex, pey, i, Mf);
                        if (rt_probe(pl, __float2int_rz(q.y), __float2int_rz(q.x))) {
                            blocked = true;
                            break;
                        }
                        prev = q;
                    }
                    ps = blocked ? prev : make_float2(pex, pey);
                }
            }
        }
        unsigned todo = __ballot_sync(full, longer);
        while (todo) { // long edges: the whole warp marches one at a time
            const int k = __ffs(todo) - 1;
            todo &= todo - 1;
            const float ax = __shfl_sync(full, psx, k), ay = __shfl_sync(full, psy, k);
            const float bx = __shfl_sync(full, pex, k), by = __shfl_sync(full, pey, k);
            const int Mk = __shfl_sync(full, M, k);
            const float Mf = (float)Mk;
            int f = 0; // first blocked index, 0 = none
            for (int i0 = 1; i0 < Mk; i0 += 32) {
                const int i = i0 + lane;
                bool hit = false;
                if (i < Mk) {
                    const float2 q = first_hit_point(ax, ay, bx, by, i, Mf);
                    hit = rt_probe(pl, __float2int_rz(q.y), __float2int_rz(q.x));
                }
                const unsigned hm = __ballot_sync(full, hit);
                if (hm) {
                    f = i0 + __ffs(hm) - 1;
                    break;
                }
            }
            if (lane == k) {
                blocked = f != 0;
                if (f > 1) ps = first_hit_point(ax, ay, bx, by, f - 1, Mf);
                else if (f == 0) ps = make_float2(bx, by);
            }
        }
        const unsigned bits = __ballot_sync(full, blocked);
        if (have) pstart[e] = ps;
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

// The tick's edge list materialised in HBM (16 B/edge, streamed at the K1 rate instead of gathered edge by edge):
// the node -> k-NN part only changes when the table is rebuilt, the tree part (parent -> child) is refreshed per tick.
__global__ void __launch_bounds__(256) edge_list_fill_rewire_kernel(const float2 *__restrict__ xy, const float2 *__restrict__ sxy,
                                                                    const int *__restrict__ knn_pos, int n, int k, float4 *__restrict__ out) {
    const size_t m = (size_t)n * k;
    for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < m; r += (size_t)gridDim.x * blockDim.x) {
        const int a = (int)(r / (unsigned)k), bp = knn_pos[r];
        float4 e = make_float4(__int_as_float(0x7fc00000), 0.f, 0.f, 0.f);
        if (bp >= 0) {
            const float2 pa = xy[a], pb = sxy[bp];
            e = make_float4(pa.x, pa.y, pb.x, pb.y);
        }
        out[r] = e;
    }
}
__global__ void __launch_bounds__(256) edge_list_fill_tree_kernel(const float2 *__restrict__ xy, const cpword *__restrict__ cp, int n,
                                                                  float4 *__restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n - 1; i += gridDim.x * blockDim.x) {
        const int b = i + 1, a = cp_parent(cp[b]);
        float4 e = make_float4(__int_as_float(0x7fc00000), 0.f, 0.f, 0.f);
        if (a >= 0) {
            const float2 pa = xy[a], pb = xy[b];
            e = make_float4(pa.x, pa.y, pb.x, pb.y);
        }
        out[i] = e;
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
    if (((uintptr_t)d_xyxy & 15) != 0) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "the edge list must be 16-byte aligned (float4 loads)");
    PlaneView pv;
    int rc = rtpp_select_plane(ctx, plane_mask, &pv);
    if (rc) return rc;
    if (n == 0) return RTPP_OK;
    MarchThresholds mt;
    rtpp_march_thresholds(divisor, &mt);
    static int chunk = getenv("RTPP_K1_CHUNK") ? std::max(1, std::min(8, atoi(getenv("RTPP_K1_CHUNK")))) : 8;
    const size_t per_block = (size_t)8 * chunk * 32;
    edge_check_f32_kernel<false><<<(unsigned)((n + per_block - 1) / per_block), 256, 0, ctx->stream>>>(pv, (const float4 *)d_xyxy, n, mt, d_bits, chunk);
    RTPP_LAUNCHED(ctx);
    return RTPP_OK;
}

// Fused check + all-gather over peer memory.  bits_all[r] / flags_all[r]: rank r's result buffer (whole batch) and flag
// array (world ints) as mapped in THIS process (rtpp_peer_alloc / rtpp_peer_open).  Every rank holds the whole batch
// (the tree is replicated) and checks its share of the blocks.
int rtpp_edges_check_allgather_dev(rtpp_ctx *ctx, const float *d_xyxy, size_t n, int divisor, int plane_mask, uint32_t *const *bits_all,
                                   int32_t *const *flags_all, int rank, int world, int epoch) {
    if (!ctx || !bits_all || !flags_all || (n && !d_xyxy)) return RTPP_ERR_INVALID;
    if (world < 1 || world > RTPP_MAX_PEERS || rank < 0 || rank >= world) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "bad rank / world size");
    if (divisor <= 0) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "divisor must be positive");
    if (((uintptr_t)d_xyxy & 15) != 0) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "the edge list must be 16-byte aligned");
    if (n > 0xffffffffull) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "at most 2^32 - 1 edges per fused batch");
    PlaneView pv;
    int rc = rtpp_select_plane(ctx, plane_mask, &pv);
    if (rc) return rc;
    const int chunk = 8;
    const size_t per_block = (size_t)8 * chunk * 32, blocks = (n + per_block - 1) / per_block;
    const size_t mine = blocks > (size_t)rank ? (blocks - rank + world - 1) / world : 0;
    if (mine) {
        MarchThresholds mt;
        rtpp_march_thresholds(divisor, &mt);
        // deferral list (worst case: every edge of this rank's share) with its counter in front (scratch is shared with
        // other entry points, so nothing is assumed to survive between calls)
        void *d_list = nullptr;
        rc = rtpp_scratch(ctx, 3, (mine * per_block + 64) * 4 + 16, &d_list);
        if (rc) return rc;
        RTPP_CK(ctx, cudaMemsetAsync(d_list, 0, 16, ctx->stream));
        PeerBits peers;
        peers.n = 0, peers.blk_stride = world, peers.blk_first = rank;
        peers.defer_count = (unsigned *)d_list, peers.defer_list = (uint32_t *)d_list + 4;
        for (int r = 0; r < world; r++)
            if (r != rank) peers.p[peers.n++] = bits_all[r];
        edge_check_f32_kernel<false, true><<<(unsigned)mine, 256, 0, ctx->stream>>>(pv, (const float4 *)d_xyxy, n, mt, bits_all[rank], chunk, peers);
        RTPP_LAUNCHED(ctx);
        edge_deferred_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(pv, (const float4 *)d_xyxy, mt, bits_all[rank], peers);
        RTPP_LAUNCHED(ctx);
    }
    if (world > 1) {
        PeerFlags pf;
        for (int r = 0; r < RTPP_MAX_PEERS; r++) pf.p[r] = r < world ? flags_all[r] : nullptr;
        const long long spin_limit = (long long)(ctx->peer_timeout_ms * 1e6 / 100.0) + 1; // one poll is ~100 ns
        peer_signal_wait_kernel<<<1, 256, 0, ctx->stream>>>(pf, rank, world, epoch, bits_all[rank], (n + 31) / 32, spin_limit);
        RTPP_LAUNCHED(ctx);
    }
    return RTPP_OK;
}

// peer-visible device memory: cudaMalloc + an IPC handle other ranks on this node can open
int rtpp_peer_alloc(rtpp_ctx *ctx, size_t bytes, void **dptr, unsigned char handle[64]) {
    if (!ctx || !dptr || !handle) return RTPP_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    RTPP_CK(ctx, cudaSetDevice(ctx->device));
    RTPP_CK(ctx, cudaMalloc(dptr, bytes ? bytes : 1));
    RTPP_CK(ctx, cudaMemset(*dptr, 0, bytes ? bytes : 1));
    cudaIpcMemHandle_t h;
    RTPP_CK(ctx, cudaIpcGetMemHandle(&h, *dptr));
    memcpy(handle, &h, 64);
    return RTPP_OK;
}
int rtpp_peer_open(rtpp_ctx *ctx, const unsigned char handle[64], void **peer_ptr) {
    if (!ctx || !handle || !peer_ptr) return RTPP_ERR_INVALID;
    RTPP_CK(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    RTPP_CK(ctx, cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return RTPP_OK;
}
int rtpp_peer_close(rtpp_ctx *ctx, void *peer_ptr) {
    if (!ctx) return RTPP_ERR_INVALID;
    RTPP_CK(ctx, cudaIpcCloseMemHandle(peer_ptr));
    return RTPP_OK;
}

// The tick's kernels; with h_bits != NULL the verdict words are also copied to that (pageable) host buffer.
static int tick_edges_run(rtpp_ctx *ctx, int divisor, int plane_mask, uint32_t *d_bits, size_t *n_edges, uint32_t *h_bits) {
    NodeStore *s = ctx->nodes;
    if (!s || !s->knn_table) RTPP_FAIL(ctx, RTPP_ERR_STATE, "rtpp_knn_table_build has not been called");
    if (s->knn_n != s->n || s->n_sorted != s->n) RTPP_FAIL(ctx, RTPP_ERR_STATE, "nodes were appended after rtpp_knn_table_build");
    if (divisor <= 0) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "divisor must be positive");
    PlaneView pv;
    int rc = rtpp_select_plane(ctx, plane_mask, &pv);
    if (rc) return rc;
    const size_t ne = (size_t)(s->n - 1) + (size_t)s->n * s->knn_k, nwords = (ne + 31) / 32;
    if (n_edges) *n_edges = ne;
    // large ticks: the windowed kernel of tick.cu (map window in shared memory) + a repack into this call's edge order
    static const bool no_window = getenv("RTPP_TICK_NO_WINDOW") != nullptr;
    if (ne >= ((size_t)1 << 16) && !no_window) {
        bool ran = false;
        rc = rtpp_tick_window_run_packed(ctx, divisor, plane_mask, d_bits, &ran);
        if (rc) return rc;
        if (ran) {
            if (h_bits) RTPP_CK(ctx, cudaMemcpyAsync(h_bits, d_bits, nwords * 4, cudaMemcpyDeviceToHost, ctx->stream));
            return RTPP_OK;
        }
    }
    MarchThresholds mt;
    rtpp_march_thresholds(divisor, &mt);
    const int chunk = 8;
    const size_t per_block = (size_t)8 * chunk * 32;
    // Large ticks stream a materialised edge list (16 B/edge at the K1 rate beats one gather per edge); it costs
    // 16 B/edge of HBM, so small ticks — and any tick whose list cannot be allocated — use the gathering kernel.
    bool use_list = ne >= ((size_t)1 << 20);
    if (use_list && s->edge_list_cap < ne) {
        if (s->edge_list) {
            RTPP_CK(ctx, cudaStreamSynchronize(ctx->stream));
            cudaFree(s->edge_list);
        }
        s->edge_list = nullptr, s->edge_list_cap = 0, s->edge_list_n = -1;
        const size_t want = (size_t)(s->capacity - 1) + s->capacity * s->knn_k;
        // a tick list that does not fit is not an error — the gathering kernel checks the same edges from the node table — but it
        // is recorded in the ctx so that a caller who expects the streaming rate can see why it did not get it
        if (cudaMalloc(&s->edge_list, want * sizeof(float4)) == cudaSuccess) s->edge_list_cap = want;
        else {
            (void)cudaGetLastError();
            ctx->err = "tick edge list (" + std::to_string(want * sizeof(float4)) + " bytes) could not be allocated: using the gathering kernel";
            use_list = false;
        }
    }
    if (!use_list) {
        edge_check_resident_kernel<<<(unsigned)((ne + per_block - 1) / per_block), 256, 0, ctx->stream>>>(pv, s->xy, s->sxy, s->cp, s->knn_pos, s->n,
                                                                                                          s->knn_k, mt, d_bits, chunk);
        RTPP_LAUNCHED(ctx);
        if (h_bits) RTPP_CK(ctx, cudaMemcpyAsync(h_bits, d_bits, nwords * 4, cudaMemcpyDeviceToHost, ctx->stream));
        return RTPP_OK;
    }
    if (s->edge_list_n != s->n) {
        edge_list_fill_rewire_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(s->xy, s->sxy, s->knn_pos, s->n, s->knn_k, s->edge_list + (s->n - 1));
        RTPP_LAUNCHED(ctx);
        s->edge_list_n = s->n;
    }
    if (s->n > 1) {
        edge_list_fill_tree_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(s->xy, s->cp, s->n, s->edge_list);
        RTPP_LAUNCHED(ctx);
    }
    edge_check_f32_kernel<true><<<(unsigned)((ne + per_block - 1) / per_block), 256, 0, ctx->stream>>>(pv, s->edge_list, ne, mt, d_bits, chunk);
    RTPP_LAUNCHED(ctx);
    if (h_bits) RTPP_CK(ctx, cudaMemcpyAsync(h_bits, d_bits, nwords * 4, cudaMemcpyDeviceToHost, ctx->stream));
    return RTPP_OK;
}

int rtpp_tick_edges_check_dev(rtpp_ctx *ctx, int divisor, int plane_mask, uint32_t *d_bits, size_t *n_edges) {
    if (!ctx || !d_bits) return RTPP_ERR_INVALID;
    return tick_edges_run(ctx, divisor, plane_mask, d_bits, n_edges, nullptr);
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
    // A pinned (page-locked, mapped) result buffer is written by the kernel itself, one 256-byte burst per block as
    // the blocks retire, so the read-back rides under the march instead of following it (3 MB per 26 M-edge tick).
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, blocked_bits) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) {
        rc = tick_edges_run(ctx, divisor, plane_mask, (uint32_t *)at.devicePointer, n_edges, nullptr);
    } else {
        (void)cudaGetLastError();
        void *d_b = nullptr;
        rc = rtpp_scratch(ctx, 1, nwords * 4, &d_b);
        if (rc) return rc;
        rc = tick_edges_run(ctx, divisor, plane_mask, (uint32_t *)d_b, n_edges, blocked_bits);
    }
    if (rc) return rc;
    RTPP_CK(ctx, cudaStreamSynchronize(ctx->stream));
    return RTPP_OK;
}

int rtpp_edges_check_f64_dev(rtpp_ctx *ctx, const double *d_xyxy, size_t n, int divisor, int plane_mask, uint32_t *d_bits) {
    if (!ctx || (n && (!d_xyxy || !d_bits))) return RTPP_ERR_INVALID;
    if (divisor <= 0) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "divisor must be positive");
    if (((uintptr_t)d_xyxy & 31) != 0) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "the binary64 edge list must be 32-byte aligned (double4 loads)");
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
    if (((uintptr_t)d_xyxy & 15) != 0) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "the edge list must be 16-byte aligned (float4 loads)");
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
