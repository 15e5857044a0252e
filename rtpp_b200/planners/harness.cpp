// harness.cpp — extern "C" test harness over the drop-in planner classes, shaped like the
// oracle/_ref harness so the Python parity tests drive both the same way.
#include <cstring>

#include "RTRRT.h"

#define PL_EXPORT extern "C" __attribute__((visibility("default")))

PL_EXPORT void *pl_rtrrt_create(const uint8_t *obs, int rows, int cols, int device) {
    RTRRTplanner *p = new RTRRTplanner(device);
    if (!p->ok()) {
        delete p;
        return nullptr;
    }
    std::vector<uint8_t> gray((size_t)rows * cols);
    for (size_t i = 0; i < gray.size(); i++) gray[i] = obs[i] ? 0 : 255;
    p->InitializeInvironment(gray.data(), rows, cols, 128);
    return p;
}
PL_EXPORT void pl_rtrrt_destroy(void *h) { delete (RTRRTplanner *)h; }
PL_EXPORT void pl_rtrrt_set_clock(void *h, long long dt) { ((RTRRTplanner *)h)->SetVirtualClock(dt); }
PL_EXPORT void pl_rtrrt_set_rand(void *h, const int32_t *r, long long n) { ((RTRRTplanner *)h)->SetRandStream(r, n); }
PL_EXPORT long long pl_rtrrt_rand_pos(void *h) { return ((RTRRTplanner *)h)->RandPos(); }
PL_EXPORT void pl_rtrrt_set_agent(void *h, float x, float y) {
    RTpoint p{x, y};
    ((RTRRTplanner *)h)->SetAgent(p);
}
PL_EXPORT void pl_rtrrt_set_goal(void *h, float x, float y) {
    RTpoint p{x, y};
    ((RTRRTplanner *)h)->SetGoal(p);
}
PL_EXPORT void pl_rtrrt_set_dobs(void *h, const uint8_t *plane) { ((RTRRTplanner *)h)->SetDObsmap(plane); }
PL_EXPORT void pl_rtrrt_add_obstacle(void *h, int num, int type, float x, float y, float angle, float radius,
                                     const float *outline_xy, int npts) {
    RTrrt_DynamicObstacle o;
    o.DObsType = type, o.x = x, o.y = y, o.angle = angle, o.radius = radius, o.sampleRadius = 0;
    for (int i = 0; i < npts; i++) o.OutLine.push_back({outline_xy[2 * i], outline_xy[2 * i + 1]});
    ((RTRRTplanner *)h)->AddDynamicObstacle(num, o);
}
PL_EXPORT void pl_rtrrt_draw_dobs(void *h) { ((RTRRTplanner *)h)->DrawDObsmap(); }
PL_EXPORT void pl_rtrrt_get_dobs(void *h, uint8_t *plane) {
    RTRRTplanner *p = (RTRRTplanner *)h;
    memcpy(plane, p->DObsmap.data.data(), p->DObsmap.data.size());
}
PL_EXPORT void pl_rtrrt_update_tree(void *h) { ((RTRRTplanner *)h)->UpdateTree(); }
PL_EXPORT void pl_rtrrt_expand_rewiring(void *h, long ms) { ((RTRRTplanner *)h)->ExpandRewiring(ms); }
PL_EXPORT void pl_rtrrt_main_task(void *h) { ((RTRRTplanner *)h)->MainTask(); }
PL_EXPORT int pl_rtrrt_crash(void *h, float ax, float ay, float bx, float by) {
    return ((RTRRTplanner *)h)->CrashDetection({ax, ay}, {bx, by});
}
PL_EXPORT int pl_rtrrt_num_nodes(void *h) { return ((RTRRTplanner *)h)->NodeCount(); }
PL_EXPORT int pl_rtrrt_dump(void *h, float *xy, int32_t *parent, float *cost, uint8_t *col) {
    RTRRTplanner *p = (RTRRTplanner *)h;
    return p->ReadTree(xy, parent, cost, col, p->NodeCount());
}
PL_EXPORT void pl_rtrrt_root(void *h, float *xy) {
    RTpoint r = ((RTRRTplanner *)h)->Root();
    xy[0] = r.x, xy[1] = r.y;
}
PL_EXPORT void pl_rtrrt_queue_sizes(void *h, int *qr, int *qs) { ((RTRRTplanner *)h)->QueueSizes(qr, qs); }
PL_EXPORT long long pl_rtrrt_clock_reads(void *h) { return ((RTRRTplanner *)h)->ClockReads(); }
PL_EXPORT int pl_rtrrt_get_path(void *h, float *out_xy, int cap, int *collision, float *cost) {
    std::list<RTpoint> path;
    RTrrt_Cost c = ((RTRRTplanner *)h)->GetPath(path);
    *collision = c.collision, *cost = c.cost;
    int n = 0;
    for (const RTpoint &q : path) {
        if (n < cap) out_xy[2 * n] = q.x, out_xy[2 * n + 1] = q.y;
        n++;
    }
    return n;
}
PL_EXPORT int pl_rtrrt_node_map_size(void *h) {
    RTRRTplanner *p = (RTRRTplanner *)h;
    p->SyncNodeMap();
    size_t subs = 0;
    for (auto &kv : p->Node_Map) subs += kv.second.subs.size();
    return (int)(p->Node_Map.size() * 1000003 + subs) & 0x7fffffff;
}

// host rasteriser alone (no GPU needed): used by the CPU tests against cv2
PL_EXPORT void pl_raster_circle(uint8_t *img, int rows, int cols, int cx, int cy, int r, int v) {
    rtpp::Plane8 p;
    p.rows = rows, p.cols = cols, p.data.assign(img, img + (size_t)rows * cols);
    rtpp::fill_circle(p, cx, cy, r, (uint8_t)v);
    memcpy(img, p.data.data(), p.data.size());
}
PL_EXPORT void pl_raster_fillpoly(uint8_t *img, int rows, int cols, const int *xy, int npts, int v) {
    rtpp::Plane8 p;
    p.rows = rows, p.cols = cols, p.data.assign(img, img + (size_t)rows * cols);
    std::vector<std::pair<int, int>> pts;
    for (int i = 0; i < npts; i++) pts.push_back({xy[2 * i], xy[2 * i + 1]});
    rtpp::fill_polygon(p, pts, (uint8_t)v);
    memcpy(img, p.data.data(), p.data.size());
}
