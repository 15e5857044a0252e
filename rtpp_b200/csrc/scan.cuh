// scan.cuh — exclusive prefix sum over int counts (hash cell starts, children list starts).
// Three launches: per-block totals, scan of the totals (one block), per-block scan + offset.
// Each block owns kScanTile consecutive elements, so every load and store is coalesced.
#pragma once
#include "common.cuh"

static constexpr int kScanThreads = 256;
static constexpr int kScanPerThread = 8;
static constexpr int kScanTile = kScanThreads * kScanPerThread;

__device__ __forceinline__ int scan_block_reduce(int v, int *smem) {
    for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = v;
    __syncthreads();
    int t = 0;
    for (int w = 0; w < kScanThreads / 32; w++) t += smem[w];
    __syncthreads();
    return t;
}

static __global__ void __launch_bounds__(kScanThreads) scan_totals_kernel(const int *__restrict__ count, int m, int *__restrict__ totals) {
    __shared__ int smem[kScanThreads / 32];
    const int base = blockIdx.x * kScanTile;
    int v = 0;
    for (int j = 0; j < kScanPerThread; j++) {
        const int i = base + j * kScanThreads + threadIdx.x;
        if (i < m) v += count[i];
    }
    const int t = scan_block_reduce(v, smem);
    if (threadIdx.x == 0) totals[blockIdx.x] = t;
}

// in-place exclusive scan of up to 4096*... block totals by one block (sequential chunks of 1024)
static __global__ void __launch_bounds__(1024) scan_of_totals_kernel(int *totals, int nb, int *grand_total) {
    __shared__ int part[1024];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < nb ? totals[i] : 0;
        part[threadIdx.x] = v;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {
            const int a = threadIdx.x >= off ? part[threadIdx.x - off] : 0;
            __syncthreads();
            part[threadIdx.x] += a;
            __syncthreads();
        }
        if (i < nb) totals[i] = carry + part[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += part[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0 && grand_total) *grand_total = carry;
}

// start[i] = exclusive prefix of count; cursor[i] = start[i] (count and cursor may alias);
// start[m] = grand total is written by scan_of_totals_kernel.
static __global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const int *count, int m, const int *__restrict__ totals,
                                                                        int *start, int *cursor) {
    __shared__ int warp_sums[kScanThreads / 32];
    __shared__ int running;
    const int base = blockIdx.x * kScanTile;
    if (threadIdx.x == 0) running = totals[blockIdx.x];
    __syncthreads();
    for (int j = 0; j < kScanPerThread; j++) {
        const int i = base + j * kScanThreads + threadIdx.x;
        const int v = i < m ? count[i] : 0;
        int x = v;
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        for (int off = 1; off < 32; off <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, off);
            if (lane >= off) x += y;
        }
        if (lane == 31) warp_sums[w] = x;
        __syncthreads();
        int before = 0;
        for (int k = 0; k < w; k++) before += warp_sums[k];
        const int excl = running + before + x - v;
        if (i < m) {
            start[i] = excl;
            cursor[i] = excl;
        }
        __syncthreads();
        if (threadIdx.x == kScanThreads - 1) running = excl + v;
        __syncthreads();
    }
}

// host helper: start[0..m] and cursor[0..m) from count[0..m); totals = scratch of >= ceil(m/tile)+1 ints
static inline int rtpp_exclusive_scan(rtpp_ctx *ctx, const int *count, int m, int *start, int *cursor, int *totals) {
    const int nb = (m + kScanTile - 1) / kScanTile;
    scan_totals_kernel<<<nb, kScanThreads, 0, ctx->stream>>>(count, m, totals);
    RTPP_LAUNCHED(ctx);
    scan_of_totals_kernel<<<1, 1024, 0, ctx->stream>>>(totals, nb, start + m);
    RTPP_LAUNCHED(ctx);
    scan_apply_kernel<<<nb, kScanThreads, 0, ctx->stream>>>(count, m, totals, start, cursor);
    RTPP_LAUNCHED(ctx);
    return RTPP_OK;
}
