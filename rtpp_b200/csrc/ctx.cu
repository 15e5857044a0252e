// ctx.cu — lifecycle, streams, buffers of librtpp_gpu.so.
#include "common.cuh"

extern void rtpp_nodes_destroy(rtpp_ctx *ctx); // grid.cu

extern "C" {

int rtpp_abi_version(void) { return RTPP_ABI_VERSION; }

const char *rtpp_status_string(int s) {
    switch (s) {
    case RTPP_OK: return "ok";
    case RTPP_ERR_INVALID: return "invalid argument";
    case RTPP_ERR_CUDA: return "CUDA error";
    case RTPP_ERR_NOMEM: return "out of memory";
    case RTPP_ERR_STATE: return "invalid call order";
    case RTPP_ERR_CAPACITY: return "capacity exceeded";
    default: return "unknown status";
    }
}

int rtpp_create(int device, rtpp_ctx **out) {
    if (!out) return RTPP_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return RTPP_ERR_CUDA; // no GPU: fail loudly
    if (device < 0 || device >= count) return RTPP_ERR_INVALID;
    rtpp_ctx *ctx = new (std::nothrow) rtpp_ctx();
    if (!ctx) return RTPP_ERR_NOMEM;
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) {
        delete ctx;
        return RTPP_ERR_CUDA;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        delete ctx;
        return RTPP_ERR_CUDA;
    }
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return RTPP_ERR_CUDA;
    }
    ctx->stream = ctx->own_stream;
    if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) ctx->copy_stream = nullptr;
    for (int i = 0; i < 5 && ctx->copy_stream; i++)
        if (cudaEventCreateWithFlags(&ctx->slice_done[i], cudaEventDisableTiming) != cudaSuccess) {
            cudaStreamDestroy(ctx->copy_stream);
            ctx->copy_stream = nullptr;
        }
    *out = ctx;
    return RTPP_OK;
}

int rtpp_destroy(rtpp_ctx *ctx) {
    if (!ctx) return RTPP_ERR_INVALID;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    rtpp_nodes_destroy(ctx);
    for (int i = 0; i < 3; i++) cudaFree(ctx->plane[i]);
    for (int i = 0; i < 4; i++) cudaFree(ctx->scratch[i]);
    if (ctx->copy_stream) {
        cudaStreamSynchronize(ctx->copy_stream);
        for (int i = 0; i < 5; i++)
            if (ctx->slice_done[i]) cudaEventDestroy(ctx->slice_done[i]);
        cudaStreamDestroy(ctx->copy_stream);
    }
    cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return RTPP_OK;
}

const char *rtpp_last_error(rtpp_ctx *ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }

int rtpp_set_stream(rtpp_ctx *ctx, void *s) {
    if (!ctx) return RTPP_ERR_INVALID;
    ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
    return RTPP_OK;
}

int rtpp_sync(rtpp_ctx *ctx) {
    if (!ctx) return RTPP_ERR_INVALID;
    RTPP_CK(ctx, cudaStreamSynchronize(ctx->stream));
    return RTPP_OK;
}

long long rtpp_kernel_launches(rtpp_ctx *ctx) { return ctx ? ctx->launches : -1; }

int rtpp_device_sm_count(rtpp_ctx *ctx, int *n) {
    if (!ctx || !n) return RTPP_ERR_INVALID;
    *n = ctx->sm_count;
    return RTPP_OK;
}

int rtpp_dev_alloc(rtpp_ctx *ctx, size_t bytes, void **dptr) {
    if (!ctx || !dptr) return RTPP_ERR_INVALID;
    RTPP_CK(ctx, cudaSetDevice(ctx->device));
    RTPP_CK(ctx, cudaMalloc(dptr, bytes ? bytes : 1));
    return RTPP_OK;
}
int rtpp_dev_free(rtpp_ctx *ctx, void *dptr) {
    if (!ctx) return RTPP_ERR_INVALID;
    RTPP_CK(ctx, cudaFree(dptr));
    return RTPP_OK;
}
int rtpp_host_alloc(rtpp_ctx *ctx, size_t bytes, void **hptr) {
    if (!ctx || !hptr) return RTPP_ERR_INVALID;
    RTPP_CK(ctx, cudaHostAlloc(hptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return RTPP_OK;
}
int rtpp_host_free(rtpp_ctx *ctx, void *hptr) {
    if (!ctx) return RTPP_ERR_INVALID;
    RTPP_CK(ctx, cudaFreeHost(hptr));
    return RTPP_OK;
}
int rtpp_h2d(rtpp_ctx *ctx, void *dst, const void *src, size_t bytes) {
    if (!ctx || (!dst && bytes) || (!src && bytes)) return RTPP_ERR_INVALID;
    RTPP_CK(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return RTPP_OK;
}
int rtpp_d2h(rtpp_ctx *ctx, void *dst, const void *src, size_t bytes) {
    if (!ctx || (!dst && bytes) || (!src && bytes)) return RTPP_ERR_INVALID;
    RTPP_CK(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return RTPP_OK;
}

} // extern "C"

// grow-only scratch
int rtpp_scratch(rtpp_ctx *ctx, int slot, size_t bytes, void **out) {
    if (bytes > ctx->scratch_bytes[slot]) {
        RTPP_CK(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->scratch[slot]) RTPP_CK(ctx, cudaFree(ctx->scratch[slot]));
        ctx->scratch[slot] = nullptr;
        ctx->scratch_bytes[slot] = 0;
        size_t want = bytes + bytes / 4 + 256;
        RTPP_CK(ctx, cudaMalloc(&ctx->scratch[slot], want));
        ctx->scratch_bytes[slot] = want;
    }
    *out = ctx->scratch[slot];
    return RTPP_OK;
}
