// common.cuh — context, error plumbing and the exact device arithmetic shared by every kernel of
// librtpp_gpu.so (sm_100a).  No CPU fallback anywhere: if a call cannot run on the GPU it fails.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>

#include "../../include/rtpp_gpu.h"

// ------------------------------------------------------------------------------------------------
// Bit-packed occupancy planes
// ------------------------------------------------------------------------------------------------
// A plane is stored as 32x32-pixel tiles; a tile is 32 consecutive uint32 (one 128-byte line):
// word r of a tile holds row r of the tile, bit c = pixel column c.  A probe is therefore one
// 4-byte load; every probe of a short edge normally lands in the same 128-byte line.
// Pixels outside [0,W)x[0,H) inside the padded tiles are 0.
struct PlaneView {
    const uint32_t *__restrict__ w;
    int W, H;      // pixels
    int tilesX;    // tiles per tile-row
};

// ------------------------------------------------------------------------------------------------
// Context
// ------------------------------------------------------------------------------------------------
struct NodeStore; // grid.cu
struct rtpp_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr; // read-backs that overlap the next kernel slice
    cudaEvent_t slice_done[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    long long launches = 0;
    std::string err;

    // planes: [0] = S, [1] = D, [2] = S|D   (uint32 tiles)
    uint32_t *plane[3] = {nullptr, nullptr, nullptr};
    bool plane_set[2] = {false, false};
    int W = 0, H = 0, tilesX = 0, tilesY = 0;
    size_t plane_words = 0;

    // grow-only device scratch for the host-buffer entry points
    void *scratch[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t scratch_bytes[4] = {0, 0, 0, 0};

    NodeStore *nodes = nullptr;
    double peer_timeout_ms = 1000.0; // bound of the fused multi-GPU exchanges' flag wait (rtpp_peer_set_timeout_ms)
};

#define RTPP_CK(ctx, call)                                                                          \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                       \
            return e__ == cudaErrorMemoryAllocation ? RTPP_ERR_NOMEM : RTPP_ERR_CUDA;               \
        }                                                                                           \
    } while (0)

#define RTPP_FAIL(ctx, code, msg)                                                                   \
    do {                                                                                            \
        (ctx)->err = (msg);                                                                         \
        return (code);                                                                              \
    } while (0)

// count + check a kernel launch
#define RTPP_LAUNCHED(ctx)                                                                          \
    do {                                                                                            \
        (ctx)->launches++;                                                                          \
        RTPP_CK(ctx, cudaGetLastError());                                                           \
    } while (0)

int rtpp_scratch(rtpp_ctx *ctx, int slot, size_t bytes, void **out);
// raster.cu: the shapes a batch of dynamic planes was drawn from (kept by a caller that wants them erased, not memset)
#include <vector>
struct RasterShapes {
    std::vector<int32_t> circles4, poly_xy, poly_sizes, poly_inst;
};
int rtpp_raster_batch(rtpp_ctx *ctx, uint32_t *planes, size_t stride_words, int n_planes, int W, int H, const int32_t *circles4,
                      int n_circles, const int32_t *poly_xy, const int32_t *poly_sizes, const int32_t *poly_inst, int n_polys,
                      const RasterShapes *prev);
int rtpp_select_plane(rtpp_ctx *ctx, int plane_mask, PlaneView *pv);
// map.cu: byte plane -> bit tiles; word-wise OR of two tile planes (b == nullptr: copy)
int rtpp_plane_pack(rtpp_ctx *ctx, const uint8_t *d_src, int W, int H, uint32_t *d_dst);
int rtpp_plane_merge(rtpp_ctx *ctx, const uint32_t *a, const uint32_t *b, uint32_t *out, size_t words);
struct MarchThresholds;
void rtpp_march_thresholds(int divisor, MarchThresholds *out);

// ------------------------------------------------------------------------------------------------
// Exact arithmetic (bit-equal to the reference's scalar C++ built without FMA contraction)
// ------------------------------------------------------------------------------------------------
// RTpoint::operator% (reference RTRRT.h:47-52): sqrtf32(dx*dx + dy*dy), each op rounded once.
__device__ __forceinline__ float rt_dist(float ax, float ay, float bx, float by) {
    float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by);
    return __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
}
// BIpoint::operator% (reference CDRT.h:73-78), binary64.
__device__ __forceinline__ double rt_dist64(double ax, double ay, double bx, double by) {
    double dx = __dsub_rn(ax, bx), dy = __dsub_rn(ay, by);
    return __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}
// "int32_t MaximumDiscrete = len / D + 2" (reference RTRRT.cpp:209): binary32 divide, add, truncate.
__device__ __forceinline__ int rt_march_M(float len, float divisor) {
    return __float2int_rz(__fadd_rn(__fdiv_rn(len, divisor), 2.0f));
}
// Largest power of two <= M (M >= 2): the last value of the reference's "denominator" loop
// (RTRRT.cpp:210).  Interior probes are t = j/P, j = 1..P-1 (the dyadic fractions i/2^k are
// exact in binary32, so the reduced and unreduced forms give the same t).
__device__ __forceinline__ int rt_pow2_floor(int M) { return M < 2 ? 1 : (1 << (31 - __clz(M))); }

// One pixel probe = cv::Mat::at<uint8_t>(row, col) != 0 on a continuous Mat.  Coordinates that
// leave the row but stay inside the buffer wrap exactly as base[row*cols+col] does; reads outside
// the buffer (undefined in the reference) are DEFINED as blocked — same rule as the oracle.
// (W*H < 2^31 is enforced when a plane is set, so the wrapped index fits 32 bits.)
__device__ __forceinline__ bool rt_probe(const PlaneView &p, int row, int col) {
    if (!((unsigned)col < (unsigned)p.W && (unsigned)row < (unsigned)p.H)) { // cold path
        const long long lin = (long long)row * p.W + col;
        if (lin < 0 || lin >= (long long)p.W * p.H) return true;
        const unsigned l32 = (unsigned)lin;
        row = (int)(l32 / (unsigned)p.W);
        col = (int)(l32 - (unsigned)row * (unsigned)p.W);
    }
    const uint32_t word = __ldg(p.w + ((((unsigned)(row >> 5) * p.tilesX + (col >> 5)) << 5) + (row & 31)));
    return (word >> (col & 31)) & 1u;
}

// Truncation of a non-negative binary32 below 2^23 without the conversion pipe: adding 2^23 with
// round-toward-zero leaves floor(v) in the low mantissa bits (one FADD.RZ instead of F2I.TRUNC,
// which issues at a quarter of the FP32 rate).  Only used where 0 <= v < 2^23 is proven (safe edges).
__device__ __forceinline__ int rt_trunc_pos(float v) { return __float_as_int(__fadd_rz(v, 8388608.0f)) - 0x4B000000; }

// Probe with no bounds test: only for coordinates proven to be inside the plane.
__device__ __forceinline__ uint32_t rt_probe_word(const PlaneView &p, int row, int col) {
    return __ldg(p.w + ((((unsigned)(row >> 5) * p.tilesX + (col >> 5)) << 5) + (row & 31))) >> (col & 31);
}
// An edge is "safe" when both end points lie in [0, W-1) x [0, H-1): every interpolated point
// fl(fl(a*t) + fl((1-t)*b)) then stays below the plane size (W, H <= 2^20), so its probes need no
// bounds test.
__device__ __forceinline__ bool rt_edge_safe(const PlaneView &p, float ax, float ay, float bx, float by) {
    const float wx = (float)(p.W - 1), wy = (float)(p.H - 1);
    return ax >= 0.f && ay >= 0.f && bx >= 0.f && by >= 0.f && ax < wx && bx < wx && ay < wy && by < wy;
}

// The march resolution P = rt_pow2_floor(rt_march_M(sqrt(s), divisor)) is a monotone step function
// of the squared length s = dx*dx + dy*dy (sqrt, divide, add and truncate are all monotone), so for
// short edges it is read off four thresholds computed once on the host with the same IEEE
// arithmetic: P = 2 below t[0], 4 below t[1], 8 below t[2], 16 below t[3]; longer edges take the
// exact sqrt/divide path.  This removes the sqrt and divide sequences (and their slow paths for
// zero-length edges, e.g. a node against itself) from the common case.
struct MarchThresholds {
    float t[4];
    float divisor;
};
__device__ __forceinline__ int rt_march_P(const MarchThresholds &mt, float s) {
    if (s < mt.t[3]) return s < mt.t[1] ? (s < mt.t[0] ? 2 : 4) : (s < mt.t[2] ? 8 : 16);
    return rt_pow2_floor(rt_march_M(__fsqrt_rn(s), mt.divisor));
}
// 1/P exactly for P = 2^e: exponent field 127 - e
__device__ __forceinline__ float rt_inv_pow2(int P) { return __int_as_float((127 - (31 - __clz(P))) << 23); }
__device__ __forceinline__ float rt_sqlen(float ax, float ay, float bx, float by) {
    const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by);
    return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}

// Interior probe j of P for the dyadic march (reference RTRRT.cpp:214-217):
//   t = j/P;  x = (int)(ps.x*t + (1-t)*pe.x);  y likewise;  unfused, truncating.
__device__ __forceinline__ bool rt_probe_interior(const PlaneView &p, float psx, float psy, float pex, float pey,
                                                  float t) {
    float omt = __fsub_rn(1.0f, t);
    int x = __float2int_rz(__fadd_rn(__fmul_rn(psx, t), __fmul_rn(omt, pex)));
    int y = __float2int_rz(__fadd_rn(__fmul_rn(psy, t), __fmul_rn(omt, pey)));
    return rt_probe(p, y, x);
}

// Whole edge by one thread (used inside the planner engines and the relax kernels, where edges
// are short).  Returns the reference's CrashDetection verdict.
__device__ __forceinline__ bool rt_edge_blocked(const PlaneView &p, float psx, float psy, float pex, float pey,
                                                const MarchThresholds &mt) {
    bool hit = rt_probe(p, __float2int_rz(psy), __float2int_rz(psx)) | rt_probe(p, __float2int_rz(pey), __float2int_rz(pex));
    if (hit) return true;
    const int P = rt_march_P(mt, rt_sqlen(psx, psy, pex, pey));
    const float inv = rt_inv_pow2(P);
    if (rt_edge_safe(p, psx, psy, pex, pey)) {
        uint32_t acc = 0;
        float t = inv; // t_j = j/P accumulated exactly (dyadic)
        for (int j = 1; j < P; j++) {
            const float omt = __fsub_rn(1.0f, t);
            const int x = rt_trunc_pos(__fadd_rn(__fmul_rn(psx, t), __fmul_rn(omt, pex)));
            const int y = rt_trunc_pos(__fadd_rn(__fmul_rn(psy, t), __fmul_rn(omt, pey)));
            acc |= rt_probe_word(p, y, x);
            t = __fadd_rn(t, inv);
        }
        return acc & 1u;
    }
    for (int j = 1; j < P; j++) hit |= rt_probe_interior(p, psx, psy, pex, pey, __fmul_rn((float)j, inv));
    return hit;
}

// Whole edge by one warp (all lanes pass the same edge): lanes stride over the interior probes and
// OR-reduce; for long edges such as goal <-> root (reference RTRRT.cpp:362).
__device__ __forceinline__ bool rt_edge_blocked_warp(const PlaneView &p, float psx, float psy, float pex, float pey,
                                                     const MarchThresholds &mt) {
    const int lane = threadIdx.x & 31;
    if (rt_probe(p, __float2int_rz(psy), __float2int_rz(psx)) | rt_probe(p, __float2int_rz(pey), __float2int_rz(pex))) return true;
    const int P = rt_march_P(mt, rt_sqlen(psx, psy, pex, pey));
    const float inv = rt_inv_pow2(P);
    for (int j0 = 1; j0 < P; j0 += 32) {
        const int j = j0 + lane;
        bool hit = false;
        if (j < P) hit = rt_probe_interior(p, psx, psy, pex, pey, __fmul_rn((float)j, inv));
        if (__any_sync(0xffffffffu, hit)) return true;
    }
    return false;
}

// streaming 16-byte load that does not allocate in L1 (edge lists are read once; L1 is kept for
// the bit-plane lines)
__device__ __forceinline__ float4 ld_stream_f4(const float4 *ptr) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(ptr));
    return r;
}
