// raster.hpp — host rasteriser of the dynamic-obstacle plane for the planner classes (product code).
//
// Takes over the two OpenCV calls of DrawDObsmap (reference RTRRT.cpp:153,166): filled circles and
// filled polygons of value 127 on an 8-bit plane, LINE_8, no shift.  Circle: midpoint algorithm,
// four clipped spans per step.  Polygon: 8-connected Bresenham outline drawn left to right, then an
// even-odd scan-line fill on 16.16 fixed-point edge abscissae where the left end of a span rounds
// to nearest and the right end truncates.  Verified pixel-for-pixel against python cv2 4.13
// (tests/test_planners_cpu.py).  (The GPU rasteriser is the "next" row of SURVEY §8(f).)
#pragma once
#include <algorithm>
#include <climits>
#include <cstdint>
#include <vector>

namespace rtpp {

struct Plane8 {
    int rows = 0, cols = 0;
    std::vector<uint8_t> data;
    void create(int r, int c) { rows = r, cols = c, data.assign((size_t)r * c, 0); }
    void zero() { std::fill(data.begin(), data.end(), 0); }
    bool empty() const { return data.empty(); }
    // cv::Mat::at<uint8_t>(row, col) of a continuous matrix; outside the buffer reads as occupied
    template <class T = uint8_t> T at(int row, int col) const {
        long long lin = (long long)row * cols + col;
        if (lin < 0 || lin >= (long long)rows * cols) return (T)255;
        return (T)data[(size_t)lin];
    }
    void span(int y, int xa, int xb, uint8_t v) {
        if ((unsigned)y >= (unsigned)rows) return;
        xa = std::max(xa, 0), xb = std::min(xb, cols - 1);
        for (int x = xa; x <= xb; x++) data[(size_t)y * cols + x] = v;
    }
    void put(int x, int y, uint8_t v) {
        if ((unsigned)x < (unsigned)cols && (unsigned)y < (unsigned)rows) data[(size_t)y * cols + x] = v;
    }
};

// 16-bit id plane (BImap::BIamap / BIimap, CDRT.h:296-297): cv::Mat CV_16U stand-in
struct Plane16 {
    int rows = 0, cols = 0;
    std::vector<uint16_t> data;
    void create(int r, int c, uint16_t fill) { rows = r, cols = c, data.assign((size_t)r * c, fill); }
    bool empty() const { return data.empty(); }
    uint16_t at(int row, int col) const { // at<uint16_t>(row, col) of a continuous matrix; outside the buffer = 0xFFFF
        long long lin = (long long)row * cols + col;
        if (lin < 0 || lin >= (long long)rows * cols) return 0xFFFF;
        return data[(size_t)lin];
    }
    void span(int y, int xa, int xb, uint16_t v) {
        if ((unsigned)y >= (unsigned)rows) return;
        xa = std::max(xa, 0), xb = std::min(xb, cols - 1);
        for (int x = xa; x <= xb; x++) data[(size_t)y * cols + x] = v;
    }
    void put(int x, int y, uint16_t v) {
        if ((unsigned)x < (unsigned)cols && (unsigned)y < (unsigned)rows) data[(size_t)y * cols + x] = v;
    }
};

inline void fill_circle(Plane8 &im, int cx, int cy, int r, uint8_t v) {
    if (r < 0) return;
    int e = 0, dx = r, dy = 0, up = 1, down = 2 * r - 1;
    for (; dx >= dy; dy++) {
        im.span(cy - dy, cx - dx, cx + dx, v), im.span(cy + dy, cx - dx, cx + dx, v);
        im.span(cy - dx, cx - dy, cx + dy, v), im.span(cy + dx, cx - dy, cx + dy, v);
        e += up, up += 2;
        if (e > 0) e -= down, dx--, down -= 2;
    }
}

template <class Im, class T> inline void draw_line8(Im &im, int xa, int ya, int xb, int yb, T v) {
    if (xb < xa) std::swap(xa, xb), std::swap(ya, yb);
    const int ddx = xb - xa, sgn = yb < ya ? -1 : 1, ddy = (yb - ya) * sgn;
    const bool steep = ddy > ddx;
    const int big = steep ? ddy : ddx, small = steep ? ddx : ddy;
    int e = big - 2 * small, x = xa, y = ya;
    for (int k = 0; k <= big; k++) {
        im.put(x, y, v);
        const bool bump = e < 0;
        e += bump ? 2 * (big - small) : -2 * small;
        if (steep) y += sgn, x += bump;
        else x += 1, y += bump ? sgn : 0;
    }
}

inline void fill_polygon(Plane8 &im, const std::vector<std::pair<int, int>> &pts, uint8_t v) {
    struct Edge {
        int ytop, ybot;
        int64_t x, slope;
    };
    const int n = (int)pts.size();
    if (n == 0) return;
    std::vector<Edge> edges;
    for (int i = 0; i < n; i++) {
        const auto &a = pts[(i + n - 1) % n], &b = pts[i];
        draw_line8(im, a.first, a.second, b.first, b.second, v);
        if (a.second == b.second) continue;
        const int64_t ax = (int64_t)a.first << 16, bx = (int64_t)b.first << 16;
        const int64_t slope = (bx - ax) / (b.second - a.second);
        edges.push_back(a.second < b.second ? Edge{a.second, b.second, ax, slope} : Edge{b.second, a.second, bx, slope});
    }
    if (edges.size() < 2) return;
    std::sort(edges.begin(), edges.end(), [](const Edge &p, const Edge &q) {
        return p.ytop != q.ytop ? p.ytop < q.ytop : (p.x != q.x ? p.x < q.x : p.slope < q.slope);
    });
    int ylast = INT_MIN;
    for (const Edge &e : edges) ylast = std::max(ylast, e.ybot);
    ylast = std::min(ylast, im.rows);
    std::vector<int> active; // indices into edges, kept ordered by x like the active-edge list
    size_t next = 0;
    for (int y = edges[0].ytop; y < ylast; y++) {
        // merge: drop finished edges, insert the edges starting on this row at their x position
        std::vector<int> merged;
        size_t ai = 0;
        while (ai < active.size() || (next < edges.size() && edges[next].ytop == y)) {
            if (ai < active.size() && edges[active[ai]].ybot == y) {
                ai++;
                continue;
            }
            const bool take_active = ai < active.size() &&
                                     (!(next < edges.size() && edges[next].ytop == y) || edges[active[ai]].x < edges[next].x);
            merged.push_back(take_active ? active[ai++] : (int)next++);
        }
        for (size_t k = 0; k + 1 < merged.size(); k += 2) {
            Edge &l = edges[merged[k]], &r = edges[merged[k + 1]];
            if (y >= 0) {
                const int64_t lo = std::min(l.x, r.x), hi = std::max(l.x, r.x);
                const int xa = (int)((lo + 32768) >> 16), xb = (int)(hi >> 16);
                if (xa < im.cols && xb >= 0) im.span(y, xa, xb, v);
            }
            l.x += l.slope, r.x += r.slope;
        }
        std::stable_sort(merged.begin(), merged.end(), [&](int p, int q) { return edges[p].x < edges[q].x; });
        active.swap(merged);
    }
}

// cv::fillConvexPoly (LINE_8, no shift) for DrawBImap (CDRT.cpp:143-144): the outline as 8-connected lines, then
// one scan line at a time between a left-going and a right-going chain that both start at the top-most vertex.
// Chain abscissae are 16.16 fixed point; a chain entering an edge (x0,y) -> (x1,y1) advances per line by
// (2*(x1-x0) + (y1-y)) / (2*(y1-y)) (C integer division); both span ends round to nearest.
// Verified pixel-for-pixel against the planes python cv2 4.13 draws (tests/test_planners_cpu.py).
template <class Im, class T> inline void fill_convex_polygon(Im &im, const std::vector<std::pair<int, int>> &pts, T v) {
    const int n = (int)pts.size();
    if (n == 0) return;
    int top = 0, y_lo = pts[0].second, y_hi = pts[0].second, x_lo = pts[0].first, x_hi = pts[0].first;
    for (int i = 0; i < n; i++) {
        const std::pair<int, int> &a = pts[(i + n - 1) % n], &b = pts[i];
        draw_line8(im, a.first, a.second, b.first, b.second, v);
        if (b.second < y_lo) y_lo = b.second, top = i;
        y_hi = std::max(y_hi, b.second), x_lo = std::min(x_lo, b.first), x_hi = std::max(x_hi, b.first);
    }
    if (n < 3 || x_hi < 0 || y_hi < 0 || x_lo >= im.cols || y_lo >= im.rows) return;
    y_hi = std::min(y_hi, im.rows - 1);
    struct Chain {
        int vertex, dir, until;
        long long x, step;
    };
    Chain ch[2] = {{top, 1, y_lo, -65536, 0}, {top, n - 1, y_lo, -65536, 0}};
    int moves_left = n;
    for (int y = y_lo; y <= y_hi; y++) {
        for (Chain &c : ch) {
            if (y < c.until) continue;
            int a = c.vertex;
            bool entered = false;
            while (moves_left-- > 0) {
                const int b = (a + c.dir) % n;
                if (pts[b].second > y) {
                    const long long xa = (long long)pts[a].first * 65536, xb = (long long)pts[b].first * 65536;
                    const int dy = pts[b].second - y;
                    c.until = pts[b].second, c.step = (2 * (xb - xa) + dy) / (2 * dy), c.x = xa, c.vertex = b;
                    entered = true;
                    break;
                }
                a = b;
            }
            if (!entered) break;
        }
        if (moves_left < 0) return;
        if (y >= 0) {
            const long long xl = std::min(ch[0].x, ch[1].x), xr = std::max(ch[0].x, ch[1].x);
            const int a = (int)((xl + 32768) >> 16), b = (int)((xr + 32768) >> 16);
            if (b >= 0 && a < im.cols) im.span(y, a, b, v);
        }
        ch[0].x += ch[0].step, ch[1].x += ch[1].step;
    }
}

} // namespace rtpp
