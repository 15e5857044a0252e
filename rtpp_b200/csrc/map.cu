// map.cu — occupancy planes: byte plane -> bit-packed 32x32 tiles, S|D merge, unpack.
// Replaces the cv::Mat planes of the reference (Obsmap RTRRT.cpp:69-79, DObsmap RTRRT.cpp:138-169).
#include "common.cuh"

// One warp packs 32 consecutive pixels of one row into one word with a ballot.
// Grid-stride over all (tile, row-in-tile) words; lane = pixel column inside the tile row.
__global__ void __launch_bounds__(256) pack_plane_kernel(const uint8_t *__restrict__ src, int W, int H, int tilesX,
                                                        size_t n_words, uint32_t *__restrict__ dst) {
    const int lane = threadIdx.x & 31;
    size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t wi = warp; wi < n_words; wi += nwarps) {
        size_t tile = wi >> 5;
        int r = (int)(wi & 31);
        int ty = (int)(tile / tilesX), tx = (int)(tile - (size_t)ty * tilesX);
        int y = (ty << 5) + r, x = (tx << 5) + lane;
        bool on = (y < H && x < W) ? (src[(size_t)y * W + x] != 0) : false;
        uint32_t word = __ballot_sync(0xffffffffu, on);
        if (lane == 0) dst[wi] = word;
    }
}

__global__ void __launch_bounds__(256) merge_planes_kernel(const uint4 *__restrict__ a, const uint4 *__restrict__ b,
                                                          uint4 *__restrict__ out, size_t n4) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        uint4 u = a[i], v = b[i];
        out[i] = make_uint4(u.x | v.x, u.y | v.y, u.z | v.z, u.w | v.w);
    }
}

__global__ void __launch_bounds__(256) unpack_plane_kernel(PlaneView p, uint8_t *__restrict__ dst) {
    size_t n = (size_t)p.W * p.H;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int y = (int)(i / p.W), x = (int)(i - (size_t)y * p.W);
        dst[i] = rt_probe(p, y, x) ? 1 : 0;
    }
}

// shared with the planner engines (one pack kernel, one merge kernel in the library)
int rtpp_plane_pack(rtpp_ctx *ctx, const uint8_t *d_src, int W, int H, uint32_t *d_dst) {
    const int tilesX = (W + 31) >> 5, tilesY = (H + 31) >> 5;
    const size_t words = (size_t)tilesX * tilesY * 32;
    int grid = (int)((words * 32 + 255) / 256);
    if (grid > ctx->sm_count * 16) grid = ctx->sm_count * 16;
    pack_plane_kernel<<<grid, 256, 0, ctx->stream>>>(d_src, W, H, tilesX, words, d_dst);
    RTPP_LAUNCHED(ctx);
    return RTPP_OK;
}
// out = a | b word by word (b == nullptr: out = a); words is a multiple of 32 (whole tiles)
int rtpp_plane_merge(rtpp_ctx *ctx, const uint32_t *a, const uint32_t *b, uint32_t *out, size_t words) {
    if (!b) {
        RTPP_CK(ctx, cudaMemcpyAsync(out, a, words * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        return RTPP_OK;
    }
    const size_t n4 = words / 4;
    int grid = (int)((n4 + 255) / 256);
    if (grid > ctx->sm_count * 8) grid = ctx->sm_count * 8;
    merge_planes_kernel<<<grid, 256, 0, ctx->stream>>>((const uint4 *)a, (const uint4 *)b, (uint4 *)out, n4);
    RTPP_LAUNCHED(ctx);
    return RTPP_OK;
}

static int ensure_planes(rtpp_ctx *ctx, int W, int H) {
    if (W <= 0 || H <= 0 || W > (1 << 20) || H > (1 << 20)) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "bad plane shape");
    if ((long long)W * H >= (1LL << 31)) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "plane larger than 2^31 pixels");
    if (ctx->plane[0] && (ctx->W != W || ctx->H != H)) {
        if (ctx->plane_set[0] || ctx->plane_set[1]) {
            // a new map: drop everything
            RTPP_CK(ctx, cudaStreamSynchronize(ctx->stream));
            for (int i = 0; i < 3; i++) {
                cudaFree(ctx->plane[i]);
                ctx->plane[i] = nullptr;
            }
            ctx->plane_set[0] = ctx->plane_set[1] = false;
        }
    }
    if (!ctx->plane[0]) {
        ctx->W = W, ctx->H = H;
        ctx->tilesX = (W + 31) >> 5, ctx->tilesY = (H + 31) >> 5;
        ctx->plane_words = (size_t)ctx->tilesX * ctx->tilesY * 32;
        for (int i = 0; i < 3; i++) {
            RTPP_CK(ctx, cudaMalloc(&ctx->plane[i], ctx->plane_words * 4));
            RTPP_CK(ctx, cudaMemsetAsync(ctx->plane[i], 0, ctx->plane_words * 4, ctx->stream));
        }
    }
    return RTPP_OK;
}

static int remerge(rtpp_ctx *ctx) { return rtpp_plane_merge(ctx, ctx->plane[0], ctx->plane[1], ctx->plane[2], ctx->plane_words); }

int rtpp_select_plane(rtpp_ctx *ctx, int plane_mask, PlaneView *pv) {
    if (!ctx->plane[0]) RTPP_FAIL(ctx, RTPP_ERR_STATE, "no map plane has been set");
    if (plane_mask < 1 || plane_mask > 3) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "plane_mask must be 1, 2 or 3");
    pv->w = ctx->plane[plane_mask - 1];
    pv->W = ctx->W, pv->H = ctx->H, pv->tilesX = ctx->tilesX;
    return RTPP_OK;
}

extern "C" {

int rtpp_map_set_plane_dev(rtpp_ctx *ctx, int plane, const uint8_t *d_src, int W, int H) {
    if (!ctx || !d_src) return RTPP_ERR_INVALID;
    if (plane != RTPP_PLANE_STATIC && plane != RTPP_PLANE_DYNAMIC) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "bad plane id");
    RTPP_CK(ctx, cudaSetDevice(ctx->device));
    int rc = ensure_planes(ctx, W, H);
    if (rc) return rc;
    rc = rtpp_plane_pack(ctx, d_src, W, H, ctx->plane[plane - 1]);
    if (rc) return rc;
    ctx->plane_set[plane - 1] = true;
    return remerge(ctx);
}

int rtpp_map_set_plane(rtpp_ctx *ctx, int plane, const uint8_t *src, int W, int H) {
    if (!ctx || !src) return RTPP_ERR_INVALID;
    if (W <= 0 || H <= 0) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "bad plane shape");
    void *d = nullptr;
    int rc = rtpp_scratch(ctx, 0, (size_t)W * H, &d);
    if (rc) return rc;
    RTPP_CK(ctx, cudaMemcpyAsync(d, src, (size_t)W * H, cudaMemcpyHostToDevice, ctx->stream));
    rc = rtpp_map_set_plane_dev(ctx, plane, (const uint8_t *)d, W, H);
    if (rc) return rc;
    RTPP_CK(ctx, cudaStreamSynchronize(ctx->stream));
    return RTPP_OK;
}

int rtpp_map_clear_plane(rtpp_ctx *ctx, int plane) {
    if (!ctx) return RTPP_ERR_INVALID;
    if (plane != RTPP_PLANE_STATIC && plane != RTPP_PLANE_DYNAMIC) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "bad plane id");
    if (!ctx->plane[0]) RTPP_FAIL(ctx, RTPP_ERR_STATE, "no map plane has been set");
    RTPP_CK(ctx, cudaMemsetAsync(ctx->plane[plane - 1], 0, ctx->plane_words * 4, ctx->stream));
    return remerge(ctx);
}

int rtpp_map_get_plane(rtpp_ctx *ctx, int plane_mask, uint8_t *out) {
    if (!ctx || !out) return RTPP_ERR_INVALID;
    PlaneView pv;
    int rc = rtpp_select_plane(ctx, plane_mask, &pv);
    if (rc) return rc;
    void *d = nullptr;
    size_t n = (size_t)ctx->W * ctx->H;
    rc = rtpp_scratch(ctx, 0, n, &d);
    if (rc) return rc;
    int grid = (int)((n + 255) / 256);
    if (grid > ctx->sm_count * 16) grid = ctx->sm_count * 16;
    unpack_plane_kernel<<<grid, 256, 0, ctx->stream>>>(pv, (uint8_t *)d);
    RTPP_LAUNCHED(ctx);
    RTPP_CK(ctx, cudaMemcpyAsync(out, d, n, cudaMemcpyDeviceToHost, ctx->stream));
    RTPP_CK(ctx, cudaStreamSynchronize(ctx->stream));
    return RTPP_OK;
}

int rtpp_map_shape(rtpp_ctx *ctx, int *W, int *H) {
    if (!ctx || !W || !H) return RTPP_ERR_INVALID;
    *W = ctx->W, *H = ctx->H;
    return RTPP_OK;
}

} // extern "C"
