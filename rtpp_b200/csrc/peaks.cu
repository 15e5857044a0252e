// peaks.cu — the roofline denominators SURVEY §8(d) says are NOT in MEASURED_PEAKS.json and must be measured in the
// same run as the kernels they bound: L2 bandwidth for the map-probe side of K1 (random 32-byte sectors and whole
// 128-byte lines out of an L2-resident buffer of plane size) and the FP32 issue rate for the neighbour kernel K3;
// the HBM copy rate is measured next to them so that the three share one clock state.  Timed with CUDA events on
// the ctx stream, after warm-up, best of `reps`.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kGatherLoads = 256; // per thread, 8 independent in flight

// Random 4-byte probes, one per 32-byte sector, L1 bypassed (ld.global.cg): what a bit-plane probe costs the L2 when
// it misses L1.  `mask` = number of sectors - 1 (power of two).
__global__ void __launch_bounds__(kThreads) l2_sector_gather_kernel(const uint32_t *__restrict__ buf, uint32_t mask,
                                                                    uint32_t *__restrict__ sink) {
    uint32_t s = (blockIdx.x * kThreads + threadIdx.x) * 2654435761u + 12345u;
    uint32_t acc = 0;
#pragma unroll 1
    for (int i = 0; i < kGatherLoads; i += 8) {
        uint32_t v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            s = s * 1664525u + 1013904223u; // LCG: the addresses do not depend on the loaded data
            v[j] = __ldcg(buf + (size_t)((s >> 8) & mask) * 8u);
        }
#pragma unroll
        for (int j = 0; j < 8; j++) acc ^= v[j];
    }
    if (acc == 0x9E3779B9u) sink[0] = acc; // keeps the loads alive, practically never taken
}

// Whole lines, coalesced 16 B per lane, L1 bypassed: the streaming rate of the L2.  Every thread issues `loads`
// 16-byte loads in groups of four; the index wraps inside the buffer (n16 a power of two), so the rate does not depend
// on how the buffer size compares with the grid.
__global__ void __launch_bounds__(kThreads) l2_stream_kernel(const uint4 *__restrict__ buf, size_t n16, int loads,
                                                             uint32_t *__restrict__ sink) {
    uint32_t acc = 0;
    const size_t stride = (size_t)gridDim.x * kThreads, m = n16 - 1;
    size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
#pragma unroll 1
    for (int l = 0; l < loads; l += 4) {
        uint4 v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) v[j] = __ldcg(buf + ((i + j * stride) & m));
#pragma unroll
        for (int j = 0; j < 4; j++) acc ^= v[j].x ^ v[j].y ^ v[j].z ^ v[j].w;
        i += 4 * stride;
    }
    if (acc == 0x9E3779B9u) sink[0] = acc;
}

// 8 independent FFMA chains per thread: the FP32 pipe's issue peak (2 flops per lane per instruction).
__global__ void __launch_bounds__(kThreads) fp32_issue_kernel(int iters, float seed, float *__restrict__ sink) {
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; j++) a[j] = seed + (float)(threadIdx.x + j);
    const float m = 0.999f, c = 0.001f;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++)
#pragma unroll
            for (int j = 0; j < 8; j++) a[j] = __fmaf_rn(a[j], m, c);
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; j++) s += a[j];
    if (s == -1.0f) sink[0] = s;
}

__global__ void __launch_bounds__(kThreads) hbm_copy_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst,
                                                            size_t n16) {
    const size_t stride = (size_t)gridDim.x * kThreads;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n16; i += stride) dst[i] = src[i];
}

} // namespace

extern "C" int rtpp_measure_peaks(rtpp_ctx *ctx, size_t l2_bytes, int reps, double out[4]) {
    if (!ctx || !out || reps < 1) return RTPP_ERR_INVALID;
    if (l2_bytes < (1u << 20) || (l2_bytes & (l2_bytes - 1))) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "l2_bytes: power of two >= 1 MiB");
    RTPP_CK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t hbm_bytes = (size_t)1 << 30; // > L2 (126 MB)
    uint32_t *l2buf = nullptr, *sink = nullptr;
    uint4 *src = nullptr, *dst = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int rc = RTPP_OK;
    auto fail = [&](cudaError_t e, const char *what) {
        ctx->err = std::string(what) + ": " + cudaGetErrorString(e);
        rc = e == cudaErrorMemoryAllocation ? RTPP_ERR_NOMEM : RTPP_ERR_CUDA;
    };
#define PK(call)                                                                                    \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            fail(e__, #call);                                                                       \
            goto done;                                                                              \
        }                                                                                           \
    } while (0)
    {
        PK(cudaMalloc(&l2buf, l2_bytes));
        PK(cudaMalloc(&sink, 256));
        PK(cudaMalloc(&src, hbm_bytes));
        PK(cudaMalloc(&dst, hbm_bytes));
        PK(cudaMemsetAsync(l2buf, 0x5A, l2_bytes, st));
        PK(cudaMemsetAsync(src, 0x3C, hbm_bytes, st));
        PK(cudaMemsetAsync(dst, 0, hbm_bytes, st));
        PK(cudaEventCreate(&e0));
        PK(cudaEventCreate(&e1));
        const int grid = ctx->sm_count * 8;
        float ms;
        // [0] random 32-byte sectors out of L2
        {
            const uint32_t mask = (uint32_t)(l2_bytes / 32 - 1);
            const int g = ctx->sm_count * 64;
            double best = 0;
            for (int r = 0; r < reps + 2; r++) {
                PK(cudaEventRecord(e0, st));
                l2_sector_gather_kernel<<<g, kThreads, 0, st>>>(l2buf, mask, sink);
                ctx->launches++;
                PK(cudaGetLastError());
                PK(cudaEventRecord(e1, st));
                PK(cudaEventSynchronize(e1));
                PK(cudaEventElapsedTime(&ms, e0, e1));
                const double gbs = (double)g * kThreads * kGatherLoads * 32.0 / (ms * 1e6);
                if (r >= 2 && gbs > best) best = gbs;
            }
            out[0] = best;
        }
        // [1] coalesced lines out of L2
        {
            const size_t n16 = l2_bytes / 16;
            const int loads = 1024; // per thread: grid x 256 x 1024 x 16 B = 4.6 GiB per launch
            double best = 0;
            for (int r = 0; r < reps + 2; r++) {
                PK(cudaEventRecord(e0, st));
                l2_stream_kernel<<<grid, kThreads, 0, st>>>((const uint4 *)l2buf, n16, loads, sink);
                ctx->launches++;
                PK(cudaGetLastError());
                PK(cudaEventRecord(e1, st));
                PK(cudaEventSynchronize(e1));
                PK(cudaEventElapsedTime(&ms, e0, e1));
                const double gbs = (double)grid * kThreads * loads * 16.0 / (ms * 1e6);
                if (r >= 2 && gbs > best) best = gbs;
            }
            out[1] = best;
        }
        // [2] FP32 FFMA issue, TFLOP/s
        {
            const int iters = 1024;
            double best = 0;
            for (int r = 0; r < reps + 2; r++) {
                PK(cudaEventRecord(e0, st));
                fp32_issue_kernel<<<grid, kThreads, 0, st>>>(iters, 1.0f, (float *)sink);
                ctx->launches++;
                PK(cudaGetLastError());
                PK(cudaEventRecord(e1, st));
                PK(cudaEventSynchronize(e1));
                PK(cudaEventElapsedTime(&ms, e0, e1));
                const double tf = (double)grid * kThreads * iters * 128.0 * 2.0 / (ms * 1e9);
                if (r >= 2 && tf > best) best = tf;
            }
            out[2] = best;
        }
        // [3] HBM copy, read + write bytes
        {
            const size_t n16 = hbm_bytes / 16;
            double best = 0;
            for (int r = 0; r < reps + 2; r++) {
                PK(cudaEventRecord(e0, st));
                hbm_copy_kernel<<<ctx->sm_count * 16, kThreads, 0, st>>>(src, dst, n16);
                ctx->launches++;
                PK(cudaGetLastError());
                PK(cudaEventRecord(e1, st));
                PK(cudaEventSynchronize(e1));
                PK(cudaEventElapsedTime(&ms, e0, e1));
                const double gbs = 2.0 * (double)hbm_bytes / (ms * 1e6);
                if (r >= 2 && gbs > best) best = gbs;
            }
            out[3] = best;
        }
    }
done:
#undef PK
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(l2buf);
    cudaFree(sink);
    cudaFree(src);
    cudaFree(dst);
    return rc;
}
