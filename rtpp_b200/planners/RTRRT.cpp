// RTRRT.cpp — see RTRRT.h.  Every public method names the reference lines whose behaviour it keeps.
#include "RTRRT.h"

#include <cfloat>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>

bool RTRRTplanner::Fail(const char *what, int rc) {
    fprintf(stderr, "RTRRTplanner: %s failed: %s (%s)\n", what, rtpp_status_string(rc), ctx_ ? rtpp_last_error(ctx_) : "no ctx");
    return false;
}

RTRRTplanner::RTRRTplanner(int device) {
    agent = {-1, -1};
    int rc = rtpp_create(device, &ctx_);
    if (rc) {
        ctx_ = nullptr;
        fprintf(stderr, "RTRRTplanner: no CUDA device / rtpp_create failed (%s); there is no CPU fallback\n", rtpp_status_string(rc));
    }
}

RTRRTplanner::~RTRRTplanner() {
    if (eng_) rtpp_engine_destroy(eng_);
    if (ctx_) rtpp_destroy(ctx_);
}

// reference RTRRT.cpp:57-80: grey image -> threshold(th, 255, BINARY) -> bitwise_not = Obsmap
void RTRRTplanner::InitializeInvironment(const uint8_t *gray, int rows, int cols, int th) {
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
        Fail("rtpp_map_set_plane", rc);
        return;
    }
    if (eng_) {
        rtpp_engine_destroy(eng_);
        eng_ = nullptr, started_ = false;
    }
}

void RTRRTplanner::InitializeInvironment(const char *IMGmap, int th) {
    std::ifstream f(IMGmap, std::ios::binary);
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
    if (!f || magic != "P5" || w <= 0 || h <= 0 || maxv != 255) {
        fprintf(stderr, "Failed to load input image.\n");
        return;
    }
    std::vector<uint8_t> g((size_t)w * h);
    f.read((char *)g.data(), g.size());
    if (!f) {
        fprintf(stderr, "Failed to load input image.\n");
        return;
    }
    InitializeInvironment(g.data(), h, w, th);
}

bool RTRRTplanner::EnsureEngine() {
    if (eng_) return true;
    if (!ctx_ || shapeX <= 0) return false;
    int rc = rtpp_engine_create(ctx_, 1, 1, node_capacity_, 5, shapeX, shapeY, max_samples_, &eng_);
    if (rc) return Fail("rtpp_engine_create", rc);
    rc = rtpp_engine_set_map(eng_, 0, 1, Obsmap.data.data(), shapeX, shapeY);
    if (rc) return Fail("rtpp_engine_set_map", rc);
    return true;
}

// reference RTRRT.cpp:11-36
void RTRRTplanner::SetAgent(RTpoint &point) {
    if (point.x >= shapeX || point.y >= shapeY) {
        printf("SetAgent error: Init point is too large!!\r\n");
        return;
    }
    if (Obsmap.at<uint8_t>((int)point.y, (int)point.x)) {
        printf("SetAgent error: robot can't be here!!\r\n");
        return;
    }
    if (!started_) {
        if (!EnsureEngine()) return;
        root = point, agent = root;
        const int32_t slot = 0;
        const float a[2] = {point.x, point.y}, g[2] = {goal.x, goal.y};
        int rc = rtpp_engine_init_instances(eng_, &slot, a, g);
        if (rc) {
            Fail("rtpp_engine_init_instances", rc);
            return;
        }
        started_ = true;
    } else
        agent = point;
}

// reference RTRRT.cpp:38-55
void RTRRTplanner::SetGoal(RTpoint &point) {
    if (point.x >= shapeX || point.y >= shapeY) {
        printf("SetGoal error: Init point is too large!!\r\n");
        return;
    }
    if (Obsmap.at<uint8_t>((int)point.y, (int)point.x)) {
        printf("SetAgent error: robot can't be here!!\r\n");
        return;
    }
    if (goal != point) goal = point;
}

void RTRRTplanner::AddDynamicObstacle(int num, RTrrt_DynamicObstacle &obs) { DObstacles[num] = obs; }
void RTRRTplanner::DelDynamicObstacle(int num) { DObstacles.erase(num); }

void RTRRTplanner::SetDObsmap(const uint8_t *plane) {
    if (DObsmap.empty()) return;
    if (plane) memcpy(DObsmap.data.data(), plane, DObsmap.data.size());
    else DObsmap.zero();
    if (!ctx_) return;
    rtpp_map_set_plane(ctx_, RTPP_PLANE_DYNAMIC, DObsmap.data.data(), shapeX, shapeY);
    if (EnsureEngine()) rtpp_engine_set_map(eng_, 0, 2, DObsmap.data.data(), shapeX, shapeY);
}

// reference RTRRT.cpp:138-169: fresh zero plane, circles / polygon of value 127 at the rotated outline
void RTRRTplanner::DrawDObsmap(void) {
    if (DObsmap.empty()) return;
    DObsmap.zero();
    std::vector<int32_t> circles, poly_xy, poly_sizes;
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
        }
    }
    // The public DObsmap member above stays what the reference's is (callers draw it); the device planes are drawn from
    // the same integer shapes by the library (pixel-identical rasteriser), so a tick uploads poses, not two byte planes.
    if (!ctx_) return;
    bool device_ok = true;
    for (int32_t v : poly_sizes) device_ok &= v <= 64; // the device rasteriser's vertex limit; beyond it the byte plane is uploaded
    if (device_ok) {
        const int nc = (int)circles.size() / 3, np = (int)poly_sizes.size();
        int rc = rtpp_map_draw_dynamic(ctx_, circles.data(), nc, poly_xy.data(), poly_sizes.data(), np);
        if (!rc && EnsureEngine()) rc = rtpp_engine_draw_dynamic(eng_, 0, circles.data(), nc, poly_xy.data(), poly_sizes.data(), np);
        if (rc) Fail("rtpp_*_draw_dynamic", rc), device_ok = false;
    }
    if (!device_ok) SetDObsmap(DObsmap.data.data());
}

void RTRRTplanner::UploadAgentGoal() {
    const float a[2] = {agent.x, agent.y}, g[2] = {goal.x, goal.y};
    rtpp_engine_set_inputs(eng_, a, g, nullptr, nullptr);
}

// reference RTRRT.cpp:92-136
void RTRRTplanner::UpdateTree(void) {
    if (!started_) return;
    UploadAgentGoal();
    int rc = rtpp_engine_tick(eng_, 1, 0, dt_ns_, 0);
    if (rc) Fail("rtpp_engine_tick", rc);
    long long st[16];
    if (rtpp_engine_read_stats(eng_, st) == 0) {
        float xy[2];
        int n = 0;
        // the root id may have moved: fetch its coordinates
        std::vector<float> all((size_t)st[0] * 2);
        if (rtpp_engine_read_tree(eng_, 0, (int)st[0], all.data(), nullptr, nullptr, nullptr, &n) == 0 && st[1] < n) {
            xy[0] = all[2 * st[1]], xy[1] = all[2 * st[1] + 1];
            root = {xy[0], xy[1]};
        }
    }
}

// reference RTRRT.cpp:226-355.  The samples are drawn ON THE DEVICE (SamplePoint, RTRRT.cpp:234-262: rtpp_engine_set_rand; glibc's
// sinf / cosf / atan2f bits, csrc/glibc_math.cuh) from the rand() values handed over here: the injected stream when there is one,
// else this process's rand().  Values the tick did not consume stay in `pending_` for the next one.
void RTRRTplanner::ExpandRewiring(long RunTime) {
    if (!started_) return;
    int32_t n;
    const int32_t *src;
    if (rand_stream_) {
        n = (int32_t)std::min<long long>(max_samples_, rand_n_ - rand_pos_);
        src = rand_stream_ + rand_pos_;
    } else {
        while ((int)pending_.size() < max_samples_) pending_.push_back(rand());
        n = max_samples_, src = pending_.data();
    }
    std::vector<int32_t> r((size_t)max_samples_, 0);
    memcpy(r.data(), src, (size_t)n * 4);
    int rc = rtpp_engine_set_rand(eng_, r.data(), &n);
    if (!rc) rc = rtpp_engine_tick(eng_, 0, RunTime, dt_ns_, 0);
    long long st[16] = {0};
    if (!rc) rc = rtpp_engine_read_stats(eng_, st);
    if (rc) {
        Fail("ExpandRewiring", rc);
        return;
    }
    if (st[9]) fprintf(stderr, "RTRRTplanner: engine capacity exceeded (status %lld)\n", st[9]);
    if (rand_stream_) rand_pos_ += st[14];
    else pending_.erase(pending_.begin(), pending_.begin() + (size_t)std::min<long long>(st[14], (long long)pending_.size()));
}

void RTRRTplanner::Start(void) {}

// reference RTRRT.cpp:483-493
void RTRRTplanner::MainTask(void) {
    DrawDObsmap();
    UpdateTree();
    ExpandRewiring(50);
}

// reference RTRRT.cpp:202-222 — a single edge through the batched kernel (callers with many edges
// should use rtpp_edges_check directly).
bool RTRRTplanner::CrashDetection(const RTpoint &ps, const RTpoint &pe) {
    if (!ctx_) return true;
    const float e[4] = {ps.x, ps.y, pe.x, pe.y};
    uint32_t bits = 1;
    int rc = rtpp_edges_check(ctx_, e, 1, 5, RTPP_PLANE_STATIC | RTPP_PLANE_DYNAMIC, &bits);
    if (rc) Fail("rtpp_edges_check", rc);
    return bits & 1u;
}

int RTRRTplanner::NodeCount() {
    long long st[16];
    if (!started_ || rtpp_engine_read_stats(eng_, st)) return 0;
    return (int)st[0];
}
int RTRRTplanner::ReadTree(float *xy, int32_t *parent, float *cost, uint8_t *col, int max_nodes) {
    int n = 0;
    if (!started_ || rtpp_engine_read_tree(eng_, 0, max_nodes, xy, parent, cost, col, &n)) return 0;
    return n;
}
void RTRRTplanner::QueueSizes(int *qr, int *qs) {
    long long st[16] = {0};
    if (started_) rtpp_engine_read_stats(eng_, st);
    *qr = (int)st[7], *qs = (int)st[8];
}
long long RTRRTplanner::ClockReads() {
    long long st[16] = {0};
    if (started_) rtpp_engine_read_stats(eng_, st);
    return st[3];
}

void RTRRTplanner::SyncNodeMap() {
    Node_Map.clear();
    const int n = NodeCount();
    if (n <= 0) return;
    std::vector<float> xy((size_t)n * 2), cost(n);
    std::vector<int32_t> par(n);
    std::vector<uint8_t> col(n);
    ReadTree(xy.data(), par.data(), cost.data(), col.data(), n);
    for (int i = 0; i < n; i++) {
        RTrrt_NodeAttribute &a = Node_Map[{xy[2 * i], xy[2 * i + 1]}];
        a.cost = {col[i] != 0, cost[i]};
        a.par = par[i] >= 0 ? RTpoint{xy[2 * par[i]], xy[2 * par[i] + 1]} : RTpoint{-1, -1};
    }
    for (int i = 0; i < n; i++)
        if (par[i] >= 0) Node_Map[{xy[2 * par[i]], xy[2 * par[i] + 1]}].subs.insert({xy[2 * i], xy[2 * i + 1]});
}

// reference RTRRT.cpp:356-478, all branches on the device (the depth-<=5 fallback of :401-437 is csrc/path_fallback.cuh):
// only the path's node ids and the coordinates they index come back.
RTrrt_Cost RTRRTplanner::GetPath(std::list<RTpoint> &path) {
    c_base = -1;
    path.clear();
    if (goal.x <= 0 || !started_) return {1, 3.4e38f};
    UploadAgentGoal();
    int rc = rtpp_engine_tick(eng_, 0, 0, dt_ns_, 1);
    long long st[16] = {0};
    if (!rc) rc = rtpp_engine_read_stats(eng_, st);
    if (rc) {
        Fail("GetPath", rc);
        return {1, 3.4e38f};
    }
    float pcost, cb;
    const uint32_t pc = (uint32_t)st[12], cbb = (uint32_t)st[13];
    memcpy(&pcost, &pc, 4), memcpy(&cb, &cbb, 4);
    c_base = cb;
    if (st[10] > 0) { // st[11] == 0: collision-free path to the goal; 1: the fallback chain root..nearest reachable node
        std::vector<int32_t> ids(st[10]);
        int len = 0;
        rtpp_engine_read_path(eng_, 0, (int)ids.size(), ids.data(), &len);
        const int n = NodeCount();
        std::vector<float> xy((size_t)n * 2);
        ReadTree(xy.data(), nullptr, nullptr, nullptr, n);
        for (int id : ids) path.push_back(id == -2 ? goal : RTpoint{xy[2 * id], xy[2 * id + 1]});
    }
    if (st[11] == 0 && st[10] > 0) return {0, pcost};
    if (pcost == 2.4e38f) return {1, 2.4e38f}; // >2000 hops (reference :392-396, :428-432)
    return {1, 3.4e38f};
}
