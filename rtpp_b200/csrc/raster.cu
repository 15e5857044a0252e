// raster.cu — D1: the dynamic-obstacle plane drawn on the GPU, straight into the bit-packed tiles.
//
// Replaces the OpenCV calls of DrawDObsmap (reference RTRRT.cpp:138-169 = RTFMT.cpp:178-209,
// AMRRT.cpp:202-233, CDRT.cpp:605-636): cv::Mat::zeros, cv::circle(FILLED) per outline point
// (type 0) and cv::fillPoly (type 1), value 127.  The rotation of the outline (sinf32/cosf32 and the
// truncating float->int of RTRRT.cpp:145-152) stays on the host — the kernels receive the integer
// pixel points — so glibc's results are used bit for bit.  Pixel-exact with the host rasteriser and
// with python cv2 4.13 for in-image shapes (tests/test_raster_gpu.py):
//   circle  = midpoint circle, four clipped spans per step (one warp per circle)
//   polygon = 8-connected Bresenham outline left-to-right (closed form of the error recurrence: the
//             minor-axis offset of point i is (2*minor*i + major - 1) / (2*major)) + even-odd scan
//             line fill on 16.16 fixed-point abscissae, left end rounds to nearest, right end
//             truncates (one thread per polygon row).
// This removes the 7-17 MB zero-fill + H2D plane upload per tick: a tick ships a few hundred bytes.
#include <limits.h>
#include <string.h>

#include <string>
#include <vector>

#include "common.cuh"

#define FULL 0xffffffffu

struct TileTarget {
    uint32_t *w;
    int W, H, tilesX;
    size_t stride; // words between the planes of a batch (primitive p draws into plane inst(p))
    int erase;     // 1: clear the pixels of the shapes instead of setting them (undo of the previous tick's drawing)
};
__device__ __forceinline__ TileTarget tile_plane(const TileTarget &t, int inst) {
    TileTarget r = t;
    r.w = t.w + (size_t)inst * t.stride;
    return r;
}

// set pixels [xa, xb] of row y (clipped) with atomicOr on the tile words
__device__ __forceinline__ void tile_span(const TileTarget &t, int y, int xa, int xb) {
    if ((unsigned)y >= (unsigned)t.H) return;
    xa = max(xa, 0), xb = min(xb, t.W - 1);
    if (xa > xb) return;
    const size_t row_base = ((size_t)(y >> 5) * t.tilesX << 5) + (y & 31);
    for (int wx = xa >> 5; wx <= (xb >> 5); wx++) {
        const int lo = max(xa, wx << 5) & 31, hi = min(xb, (wx << 5) + 31) & 31;
        const uint32_t mask = (hi == 31 ? 0xffffffffu : ((1u << (hi + 1)) - 1u)) & ~((1u << lo) - 1u);
        if (t.erase) atomicAnd(&t.w[row_base + ((size_t)wx << 5)], ~mask);
        else atomicOr(&t.w[row_base + ((size_t)wx << 5)], mask);
    }
}
__device__ __forceinline__ void tile_put(const TileTarget &t, int x, int y) {
    if ((unsigned)x >= (unsigned)t.W || (unsigned)y >= (unsigned)t.H) return;
    uint32_t *wd = &t.w[((((size_t)(y >> 5) * t.tilesX + (x >> 5)) << 5) + (y & 31))];
    if (t.erase) atomicAnd(wd, ~(1u << (x & 31)));
    else atomicOr(wd, 1u << (x & 31));
}

// Filled circle = cv::circle(FILLED): the midpoint recurrence (err += 2*dy+1; if err > 0: err -= 2*dx-1, dx--)
// keeps err = dx^2 + dy^2 - r^2, so the half-width drawn at step dy is the closed form
// dx(dy) = floor(sqrt(r^2 - dy^2)) while dx >= dy (one decrement per step suffices in the first octant).
// One thread per (circle, step): its four clipped spans.  step_of[j] = (circle, dy).
__device__ __forceinline__ int isqrt_floor(int v) {
    int s = (int)sqrtf((float)v);
    while (s * s > v) s--;
    while ((s + 1) * (s + 1) <= v) s++;
    return s;
}
__global__ void __launch_bounds__(128) raster_circles_kernel(TileTarget t0, const int4 *__restrict__ circles, int n_circles) {
    // one block per circle, threads over the steps dy = 0, 1, ... of the first octant (the step list used to be built on
    // the host: with one plane per planner instance that was hundreds of thousands of entries per tick)
    if ((int)blockIdx.x >= n_circles) return;
    const int4 c = circles[blockIdx.x]; // cx, cy, r, plane
    if (c.z < 0 || c.z > 23000) return; // r^2 must fit the exact integer arithmetic below
    const TileTarget t = tile_plane(t0, c.w);
    for (int dy = blockIdx.y * blockDim.x + threadIdx.x; dy <= c.z; dy += gridDim.y * blockDim.x) {
        const int dx = isqrt_floor(c.z * c.z - dy * dy);
        if (dx < dy) break; // past the octant's end (dx - dy only falls from here on)
        tile_span(t, c.y - dy, c.x - dx, c.x + dx);
        tile_span(t, c.y + dy, c.x - dx, c.x + dx);
        tile_span(t, c.y - dx, c.x - dy, c.x + dy);
        tile_span(t, c.y + dx, c.x - dy, c.x + dy);
    }
}

// polygon vertices: pts[first .. first+count) ; edge k joins vertex k-1 (cyclic) to vertex k
struct PolyDesc {
    int first, count, ymin, ymax, inst;
};

// outline: one block per polygon, its warps take the edges in turn, lanes stride over the Bresenham points
__global__ void __launch_bounds__(128) raster_poly_outline_kernel(TileTarget t0, const int2 *__restrict__ pts, const PolyDesc *__restrict__ polys,
                                                                 int n_polys) {
    if ((int)blockIdx.x >= n_polys) return;
    const int lane = threadIdx.x & 31;
    const PolyDesc pd = polys[blockIdx.x];
    const TileTarget t = tile_plane(t0, pd.inst);
    for (int k = blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5); k < pd.count; k += gridDim.y * (blockDim.x >> 5)) {
        int2 a = pts[pd.first + (k + pd.count - 1) % pd.count], b = pts[pd.first + k];
        if (b.x < a.x) {
            const int2 s = a;
            a = b, b = s;
        }
        const int ddx = b.x - a.x, sgn = b.y < a.y ? -1 : 1, ddy = (b.y - a.y) * sgn;
        const bool steep = ddy > ddx;
        const long long big = steep ? ddy : ddx, small = steep ? ddx : ddy;
        for (long long i = lane; i <= big; i += 32) {
            const long long off = big ? (2 * small * i + big - 1) / (2 * big) : 0; // minor-axis offset of point i
            const int x = steep ? a.x + (int)off : a.x + (int)i;
            const int y = steep ? a.y + sgn * (int)i : a.y + sgn * (int)off;
            tile_put(t, x, y);
        }
    }
}

// fill: one thread per (polygon, row); at most kMaxPolyEdges edges per polygon
static constexpr int kMaxPolyEdges = 64;
__device__ __forceinline__ void raster_poly_fill_row(const TileTarget &t, const int2 *__restrict__ pts, const PolyDesc &pd, int y) {
    long long xs[kMaxPolyEdges], sl[kMaxPolyEdges];
    int m = 0;
    for (int k = 0; k < pd.count && m < kMaxPolyEdges; k++) {
        const int2 a = pts[pd.first + (k + pd.count - 1) % pd.count], b = pts[pd.first + k];
        if (a.y == b.y) continue;
        const long long ax = (long long)a.x << 16, bx = (long long)b.x << 16;
        const long long slope = (bx - ax) / (b.y - a.y);
        const int ytop = min(a.y, b.y), ybot = max(a.y, b.y);
        if (y < ytop || y >= ybot) continue;
        const long long x = (a.y < b.y ? ax : bx) + (long long)(y - ytop) * slope;
        int j = m++; // insertion sort by (x, slope)
        while (j > 0 && (xs[j - 1] > x || (xs[j - 1] == x && sl[j - 1] > slope))) xs[j] = xs[j - 1], sl[j] = sl[j - 1], j--;
        xs[j] = x, sl[j] = slope;
    }
    for (int k = 0; k + 1 < m; k += 2) {
        const int xa = (int)((xs[k] + 32768) >> 16), xb = (int)(xs[k + 1] >> 16);
        if (xa < t.W && xb >= 0) tile_span(t, y, xa, xb);
    }
}
__global__ void __launch_bounds__(128) raster_poly_fill_kernel(TileTarget t0, const int2 *__restrict__ pts, const PolyDesc *__restrict__ polys,
                                                              int n_polys) {
    // one block per polygon, one thread per row of its clipped bounding box
    if ((int)blockIdx.x >= n_polys) return;
    const PolyDesc pd = polys[blockIdx.x];
    const TileTarget t = tile_plane(t0, pd.inst);
    for (int y = max(pd.ymin, 0) + (int)(blockIdx.y * blockDim.x + threadIdx.x); y < pd.ymax && y < t.H; y += gridDim.y * blockDim.x)
        raster_poly_fill_row(t, pts, pd, y);
}


// One pass over a set of shapes: circle c = (cx, cy, r, plane), polygon p into plane poly_inst[p]; erase = 1 clears their
// pixels.  The shapes go up in ONE copy — [circles | polygon vertices | polygon descriptors]; steps, edges and rows are
// enumerated by the kernels (one block per shape).
static int raster_shapes(rtpp_ctx *ctx, TileTarget t, const int32_t *circles4, int n_circles, const int32_t *poly_xy, const int32_t *poly_sizes,
                         const int32_t *poly_inst, int n_polys) {
    std::vector<PolyDesc> polys((size_t)(n_polys > 0 ? n_polys : 0));
    size_t npts = 0;
    int max_rows = 1, max_cnt = 1, max_r = 0; // the second grid dimension spreads a big shape over several blocks
    for (int p = 0; p < n_polys; p++) {
        const int cnt = poly_sizes[p];
        if (cnt > kMaxPolyEdges) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "polygon has too many vertices");
        int ymin = INT_MAX, ymax = INT_MIN;
        for (int k = 0; k < cnt; k++) {
            const int y = poly_xy[2 * (npts + k) + 1];
            ymin = y < ymin ? y : ymin, ymax = y > ymax ? y : ymax;
        }
        polys[p] = PolyDesc{(int)npts, cnt, ymin, ymax, poly_inst ? poly_inst[p] : 0};
        npts += cnt;
        const int rows = (ymax < t.H ? ymax : t.H) - (ymin > 0 ? ymin : 0);
        max_rows = rows > max_rows ? rows : max_rows, max_cnt = cnt > max_cnt ? cnt : max_cnt;
    }
    for (int c = 0; c < n_circles; c++) max_r = circles4[4 * c + 2] > max_r && circles4[4 * c + 2] <= 23000 ? circles4[4 * c + 2] : max_r;
    if (n_circles <= 0 && n_polys <= 0) return RTPP_OK;
    auto up16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
    const size_t o_cir = 0, o_pts = up16((size_t)(n_circles > 0 ? n_circles : 0) * 16), o_pol = up16(o_pts + npts * 8),
                 total = up16(o_pol + polys.size() * sizeof(PolyDesc)) + 16;
    std::vector<char> stage(total, 0);
    if (n_circles > 0) memcpy(stage.data() + o_cir, circles4, (size_t)n_circles * 16);
    if (npts) memcpy(stage.data() + o_pts, poly_xy, npts * 8);
    if (!polys.empty()) memcpy(stage.data() + o_pol, polys.data(), polys.size() * sizeof(PolyDesc));
    void *d = nullptr;
    int rc = rtpp_scratch(ctx, 2, total, &d);
    if (rc) return rc;
    // (a pageable source is staged by cudaMemcpyAsync before it returns, so `stage` may go out of scope; the copy is ordered
    // on the stream behind the kernels of an earlier pass that read the same scratch)
    RTPP_CK(ctx, cudaMemcpyAsync(d, stage.data(), total, cudaMemcpyHostToDevice, ctx->stream));
    const char *base = (const char *)d;
    if (n_circles > 0) {
        raster_circles_kernel<<<dim3(n_circles, (max_r * 3 / 4 + 2 + 127) / 128), 128, 0, ctx->stream>>>(t, (const int4 *)(base + o_cir), n_circles);
        RTPP_LAUNCHED(ctx);
    }
    if (n_polys > 0) {
        const int2 *dp = (const int2 *)(base + o_pts);
        const PolyDesc *dpo = (const PolyDesc *)(base + o_pol);
        raster_poly_outline_kernel<<<dim3(n_polys, (max_cnt + 3) / 4), 128, 0, ctx->stream>>>(t, dp, dpo, n_polys);
        RTPP_LAUNCHED(ctx);
        raster_poly_fill_kernel<<<dim3(n_polys, (max_rows + 127) / 128), 128, 0, ctx->stream>>>(t, dp, dpo, n_polys);
        RTPP_LAUNCHED(ctx);
    }
    return RTPP_OK;
}

// Draw a batch of planes (n_planes tile planes `stride_words` apart): every plane is cleared
// (DObsmap = cv::Mat::zeros, RTRRT.cpp:140), then circle c = (cx, cy, r, plane) and polygon p (poly_inst[p])
// are drawn into their planes with one launch per primitive kind.  Used by the ctx planes, the engines'
// map slots (one plane) and the CDRT engine (one dynamic plane per planner instance).
// prev_* (optional): the shapes these planes hold from the previous call and nothing else — they are then cleared by
// erasing exactly those shapes instead of a memset of every plane (1024 planners x 0.95 MB = 1 GB per tick on GAME0).
int rtpp_raster_batch(rtpp_ctx *ctx, uint32_t *planes, size_t stride_words, int n_planes, int W, int H, const int32_t *circles4,
                      int n_circles, const int32_t *poly_xy, const int32_t *poly_sizes, const int32_t *poly_inst, int n_polys,
                      const RasterShapes *prev) {
    const int tilesX = (W + 31) >> 5, tilesY = (H + 31) >> 5;
    const size_t words = (size_t)tilesX * tilesY * 32;
    TileTarget t{planes, W, H, tilesX, stride_words, 0};
    if (prev) {
        t.erase = 1;
        int rc = raster_shapes(ctx, t, prev->circles4.data(), (int)prev->circles4.size() / 4, prev->poly_xy.data(), prev->poly_sizes.data(),
                               prev->poly_inst.data(), (int)prev->poly_sizes.size());
        if (rc) return rc;
        t.erase = 0;
    } else if (n_planes > 1 && stride_words == words) {
        RTPP_CK(ctx, cudaMemsetAsync(planes, 0, words * 4 * (size_t)n_planes, ctx->stream));
    } else {
        for (int i = 0; i < n_planes; i++) RTPP_CK(ctx, cudaMemsetAsync(planes + (size_t)i * stride_words, 0, words * 4, ctx->stream));
    }
    return raster_shapes(ctx, t, circles4, n_circles, poly_xy, poly_sizes, poly_inst, n_polys);
}

// one plane; circles as (cx, cy, r) triples
int rtpp_raster_into(rtpp_ctx *ctx, uint32_t *plane, int W, int H, const int32_t *circles, int n_circles, const int32_t *poly_xy,
                     const int32_t *poly_sizes, int n_polys) {
    std::vector<int32_t> c4((size_t)n_circles * 4);
    for (int i = 0; i < n_circles; i++) c4[4 * i] = circles[3 * i], c4[4 * i + 1] = circles[3 * i + 1], c4[4 * i + 2] = circles[3 * i + 2], c4[4 * i + 3] = 0;
    return rtpp_raster_batch(ctx, plane, 0, 1, W, H, c4.data(), n_circles, poly_xy, poly_sizes, nullptr, n_polys, nullptr);
}

extern "C" {

// DrawDObsmap on the device: clears the dynamic plane, draws n_circles filled circles (cx, cy, r
// int triples: one per outline point of a type-0 obstacle) and n_polys filled polygons (poly_xy:
// concatenated (x, y) int pairs; poly_sizes[p] vertices each), then refreshes the merged S|D plane.
int rtpp_map_draw_dynamic(rtpp_ctx *ctx, const int32_t *circles, int n_circles, const int32_t *poly_xy, const int32_t *poly_sizes,
                          int n_polys) {
    if (!ctx || n_circles < 0 || n_polys < 0 || (n_circles && !circles) || (n_polys && (!poly_xy || !poly_sizes))) return RTPP_ERR_INVALID;
    if (!ctx->plane[0]) RTPP_FAIL(ctx, RTPP_ERR_STATE, "no map plane has been set");
    int rc = rtpp_raster_into(ctx, ctx->plane[1], ctx->W, ctx->H, circles, n_circles, poly_xy, poly_sizes, n_polys);
    if (rc) return rc;
    ctx->plane_set[1] = true;
    return rtpp_plane_merge(ctx, ctx->plane[0], ctx->plane[1], ctx->plane[2], ctx->plane_words);
}

} // extern "C"
