// Stand-in for <opencv2/core/core.hpp> — TEST INFRASTRUCTURE ONLY (oracle/_ref build).
//
// The reference planners include OpenCV, which is not installed in this image.  To compile the
// reference's own planner sources *in place* (oracle/Makefile, outputs in oracle/_ref/), this header
// provides the handful of cv:: names those sources touch.  It is written from the documented
// behaviour of the OpenCV API, not from OpenCV sources:
//   * cv::Mat is a continuous row-major buffer; at<T>(row,col) indexes base[row*cols+col] with no
//     bounds check inside the buffer (release-mode OpenCV); reads outside the buffer — undefined in
//     the reference — return an "occupied" sentinel so that runs stay deterministic.
//   * imread() returns an image previously registered by the harness (no PNG codec here; the real
//     decode happens in Python cv2 when the fixtures are made, tools/make_map_fixtures.py).
//   * circle()/fillPoly() forward to a rasteriser hook (oracle/raster.hpp) pinned against Python
//     cv2 output.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

// AMRRT.cpp:371 calls an unqualified min(); with the real OpenCV/CGAL headers some transitive
// include makes it visible at global scope.
using std::max;
using std::min;

#define CV_8U 0
#define CV_8UC3 16
#define CV_16U 2

namespace cv {

enum { IMREAD_GRAYSCALE = 0, IMREAD_COLOR = 1, THRESH_BINARY = 0, FILLED = -1 };

struct Point {
    int x = 0, y = 0;
    Point() {}
    Point(int x_, int y_) : x(x_), y(y_) {}
};
struct Scalar {
    double v[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : v{a, b, c, d} {}
    double &operator[](int i) { return v[i]; }
    const double &operator[](int i) const { return v[i]; }
};

class Mat {
  public:
    int rows = 0, cols = 0;
    int type_ = CV_8U;
    std::shared_ptr<std::vector<uint8_t>> buf;

    Mat() {}
    Mat(int r, int c, int t) : rows(r), cols(c), type_(t), buf(std::make_shared<std::vector<uint8_t>>((size_t)r * c * esz(t), 0)) {}
    static int esz(int t) { return t == CV_16U ? 2 : (t == CV_8UC3 ? 3 : 1); }
    static Mat zeros(int r, int c, int t) { return Mat(r, c, t); }
    // CDRT.cpp:123-124 writes cv::Mat::ones(rows, cols, CV_16U) * 0xFFFF: every element = 0xFFFF
    static Mat ones(int r, int c, int t) {
        Mat m(r, c, t);
        if (t == CV_16U) {
            uint16_t *p = reinterpret_cast<uint16_t *>(m.data());
            for (size_t i = 0; i < (size_t)r * c; i++) p[i] = 1;
        } else
            std::memset(m.data(), 1, m.buf->size());
        return m;
    }
    Mat operator*(double s) const { // saturating scale of a 16U/8U plane
        Mat m = clone();
        if (type_ == CV_16U) {
            uint16_t *p = reinterpret_cast<uint16_t *>(m.data());
            for (size_t i = 0; i < (size_t)rows * cols; i++) p[i] = (uint16_t)std::min(65535.0, std::max(0.0, std::nearbyint(p[i] * s)));
        } else if (type_ == CV_8U) {
            for (size_t i = 0; i < (size_t)rows * cols; i++) m.data()[i] = (uint8_t)std::min(255.0, std::max(0.0, std::nearbyint(m.data()[i] * s)));
        }
        return m;
    }
    bool empty() const { return !buf || rows == 0 || cols == 0; }
    int type() const { return type_; }
    void convertTo(Mat &dst, int) const { dst = clone(); }
    Mat clone() const {
        Mat m;
        m.rows = rows, m.cols = cols, m.type_ = type_;
        if (buf) m.buf = std::make_shared<std::vector<uint8_t>>(*buf);
        return m;
    }
    uint8_t *data() { return buf ? buf->data() : nullptr; }
    const uint8_t *data() const { return buf ? buf->data() : nullptr; }
    template <class T> T &at(int row, int col) {
        static thread_local T oob;
        long long lin = (long long)row * cols + col;
        if (!buf || lin < 0 || lin >= (long long)rows * cols) {
            std::memset(&oob, 0xff, sizeof(T));
            return oob;
        }
        return reinterpret_cast<T *>(buf->data())[lin];
    }
};

// --- harness-side registry standing in for the file system + PNG decoder ---
inline std::map<std::string, Mat> &shim_images() {
    static std::map<std::string, Mat> reg;
    return reg;
}
inline void shim_register_image(const std::string &name, int rows, int cols, const uint8_t *gray) {
    Mat m(rows, cols, CV_8U);
    std::memcpy(m.data(), gray, (size_t)rows * cols);
    shim_images()[name] = m;
}
inline Mat imread(const std::string &name, int = IMREAD_COLOR) {
    auto it = shim_images().find(name);
    return it == shim_images().end() ? Mat() : it->second.clone();
}
inline double threshold(const Mat &src, Mat &dst, double th, double maxval, int) {
    Mat out(src.rows, src.cols, CV_8U);
    for (size_t i = 0; i < (size_t)src.rows * src.cols; i++) out.data()[i] = src.data()[i] > th ? (uint8_t)maxval : 0;
    dst = out;
    return th;
}
inline void bitwise_not(const Mat &src, Mat &dst) {
    Mat out(src.rows, src.cols, CV_8U);
    for (size_t i = 0; i < (size_t)src.rows * src.cols; i++) out.data()[i] = (uint8_t)~src.data()[i];
    dst = out;
}

// rasteriser hooks (set by the harness; default = no drawing)
typedef void (*shim_circle_fn)(uint8_t *img, int rows, int cols, int cx, int cy, int radius, uint8_t value);
typedef void (*shim_fillpoly_fn)(uint8_t *img, int rows, int cols, const int *xy, int npts, uint8_t value);
inline shim_circle_fn &shim_circle_hook() {
    static shim_circle_fn f = nullptr;
    return f;
}
inline shim_fillpoly_fn &shim_fillpoly_hook() {
    static shim_fillpoly_fn f = nullptr;
    return f;
}
inline void circle(Mat &img, Point c, int radius, const Scalar &color, int = 1) {
    if (shim_circle_hook() && img.type() == CV_8U) shim_circle_hook()(img.data(), img.rows, img.cols, c.x, c.y, radius, (uint8_t)color.v[0]);
}
inline void fillPoly(Mat &img, const std::vector<std::vector<Point>> &polys, const Scalar &color) {
    if (!shim_fillpoly_hook() || img.type() != CV_8U) return;
    for (const auto &p : polys) {
        std::vector<int> xy;
        for (const Point &q : p) xy.push_back(q.x), xy.push_back(q.y);
        shim_fillpoly_hook()(img.data(), img.rows, img.cols, xy.data(), (int)p.size(), (uint8_t)color.v[0]);
    }
}
// fillConvexPoly on the 16-bit id planes of BImap::DrawBImap (CDRT.cpp:143-144); other depths: no-op
typedef void (*shim_fillconvex16_fn)(uint16_t *img, int rows, int cols, const int *xy, int npts, uint16_t value);
inline shim_fillconvex16_fn &shim_fillconvex16_hook() {
    static shim_fillconvex16_fn f = nullptr;
    return f;
}
inline void fillConvexPoly(Mat &img, const Point *pts, int npts, const Scalar &color, int = 8, int = 0) {
    if (!shim_fillconvex16_hook() || img.type() != CV_16U) return;
    std::vector<int> xy;
    for (int i = 0; i < npts; i++) xy.push_back(pts[i].x), xy.push_back(pts[i].y);
    shim_fillconvex16_hook()(reinterpret_cast<uint16_t *>(img.data()), img.rows, img.cols, xy.data(), npts, (uint16_t)color.v[0]);
}
inline void line(Mat &, Point, Point, const Scalar &, int = 1) {}
inline void namedWindow(const std::string &, int = 0) {}
inline void imshow(const std::string &, const Mat &) {}
inline int waitKey(int = 0) { return -1; }

} // namespace cv
