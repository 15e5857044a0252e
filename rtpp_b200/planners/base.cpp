// base.cpp — shared host code of the drop-in planner classes (see rtpp_types.h).
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>

#include "rtpp_types.h"

rtppPlannerBase::rtppPlannerBase(int device) {
    int rc = rtpp_create(device, &ctx_);
    if (rc) {
        ctx_ = nullptr;
        fprintf(stderr, "rtpp planner: no CUDA device / rtpp_create failed (%s); there is no CPU fallback\n", rtpp_status_string(rc));
    }
}
rtppPlannerBase::~rtppPlannerBase() {
    if (ctx_) rtpp_destroy(ctx_);
}
bool rtppPlannerBase::Fail(const char *who, const char *what, int rc) {
    fprintf(stderr, "%s: %s failed: %s (%s)\n", who, what, rtpp_status_string(rc), ctx_ ? rtpp_last_error(ctx_) : "no ctx");
    return false;
}
int rtppPlannerBase::NextRand() {
    if (rand_stream_) return rand_pos_ < rand_n_ ? rand_stream_[rand_pos_++] : 0;
    return rand();
}
// reference RTRRT.cpp:57-80: grey image -> threshold(th, 255, BINARY) -> bitwise_not = Obsmap
void rtppPlannerBase::LoadGray(const uint8_t *gray, int rows, int cols, int th) {
    if (!gray || rows <= 0 || cols <= 0) {
        fprintf(stderr, "Failed to load input image.\n");
        return;
    }
    Freemap.create(rows, cols), Obsmap.create(rows, cols), DObsmap.create(rows, cols);
    for (size_t i = 0; i < (size_t)rows * cols; i++) {
        const uint8_t f = gray[i] > th ? 255 : 0;
        Freemap.data[i] = f;
        Obsmap.data[i] = (uint8_t)~f;
    }
    shapeX = cols, shapeY = rows;
    if (!ctx_) return;
    int rc = rtpp_map_set_plane(ctx_, RTPP_PLANE_STATIC, Obsmap.data.data(), cols, rows);
    if (rc) {
        Fail("InitializeInvironment", "rtpp_map_set_plane", rc);
        return;
    }
    OnStaticPlane();
}
bool rtppPlannerBase::LoadPGM(const char *path, int th) {
    std::ifstream f(path, std::ios::binary);
    std::string magic;
    int w = 0, h = 0, maxv = 0;
    if (f) {
        f >> magic;
        auto skip = [&] {
            f >> std::ws;
            while (f.peek() == '#') f.ignore(1 << 20, '\n'), f >> std::ws;
        };
        skip(), f >> w, skip(), f >> h, skip(), f >> maxv;
        f.get();
    }
    std::vector<uint8_t> g;
    if (f && magic == "P5" && w > 0 && h > 0 && maxv == 255) {
        g.resize((size_t)w * h);
        f.read((char *)g.data(), g.size());
    }
    if (!f || g.empty()) {
        fprintf(stderr, "Failed to load input image.\n");
        return false;
    }
    LoadGray(g.data(), h, w, th);
    return true;
}
void rtppPlannerBase::UploadDObsmap() {
    if (!ctx_ || DObsmap.empty()) return;
    rtpp_map_set_plane(ctx_, RTPP_PLANE_DYNAMIC, DObsmap.data.data(), shapeX, shapeY);
    OnDynamicPlane();
}
void rtppPlannerBase::SetDObsmap(const uint8_t *plane) {
    if (DObsmap.empty()) return;
    if (plane) memcpy(DObsmap.data.data(), plane, DObsmap.data.size());
    else DObsmap.zero();
    auto_draw_ = false;
    UploadDObsmap();
}
// reference RTRRT.cpp:138-169 (= RTFMT.cpp:178-209, AMRRT.cpp:202-233)
void rtppPlannerBase::DrawDObsmap(void) {
    if (DObsmap.empty()) return;
    DObsmap.zero();
    std::vector<int32_t> circles, poly_xy, poly_sizes;
    bool device_ok = true;
    for (const auto &kv : DObstacles) {
        const RTrrt_DynamicObstacle &obs = kv.second;
        const float sinangle = sinf(obs.angle), cosangle = cosf(obs.angle);
        std::vector<std::pair<int, int>> pts;
        for (const RTpoint &p : obs.OutLine) {
            const float a = cosangle * p.x, b = sinangle * p.y, c = sinangle * p.x, d = cosangle * p.y;
            pts.push_back({(int)((a - b) + obs.x), (int)((c + d) + obs.y)});
        }
        if (obs.DObsType == 0)
            for (const auto &q : pts) {
                rtpp::fill_circle(DObsmap, q.first, q.second, (int)obs.radius, 127);
                circles.push_back(q.first), circles.push_back(q.second), circles.push_back((int)obs.radius);
            }
        else if (obs.DObsType == 1) {
            rtpp::fill_polygon(DObsmap, pts, 127);
            for (const auto &q : pts) poly_xy.push_back(q.first), poly_xy.push_back(q.second);
            poly_sizes.push_back((int32_t)pts.size());
            device_ok &= pts.size() <= 64; // the device rasteriser's vertex limit; beyond it the byte plane is uploaded
        }
    }
    // The public DObsmap member stays what the reference's is; the device planes are drawn from the same integer shapes by
    // the library's pixel-identical rasteriser, so a tick ships poses instead of byte planes.
    if (!ctx_) return;
    if (device_ok) {
        const int nc = (int)circles.size() / 3, np = (int)poly_sizes.size();
        int rc = rtpp_map_draw_dynamic(ctx_, circles.data(), nc, poly_xy.data(), poly_sizes.data(), np);
        if (rc) Fail("DrawDObsmap", "rtpp_map_draw_dynamic", rc), device_ok = false;
        else if (!OnDynamicShapes(circles.data(), nc, poly_xy.data(), poly_sizes.data(), np)) OnDynamicPlane();
    }
    if (!device_ok) UploadDObsmap();
}
bool rtppPlannerBase::EdgeBlocked(const RTpoint &ps, const RTpoint &pe, int divisor, int mask) {
    if (!ctx_) return true;
    const float e[4] = {ps.x, ps.y, pe.x, pe.y};
    uint32_t bits = 1;
    int rc = rtpp_edges_check(ctx_, e, 1, divisor, mask, &bits);
    if (rc) Fail("CrashDetection", "rtpp_edges_check", rc);
    return bits & 1u;
}
