// ref_rtrrt.cpp — harness over the reference's OWN RT-RRT* source, compiled in place
// (TEST INFRASTRUCTURE ONLY; output oracle/_ref/libref_rtrrt.so; no reference source is copied).
// Build: see oracle/Makefile (needs /root/reference, -I oracle/shim for the OpenCV/CGAL stand-ins).
#include "ref_hooks.h"
#define private public
#include "RTrrt/RTRRT.cpp" // = /root/reference/RTRRT/lib/RTrrt/RTRRT.cpp via -I
#undef private
#include "ref_common.hpp"
#include "raster.hpp"

struct RefRTRRT {
    RTRRTplanner p;
};

REF_EXPORT void *ref_rtrrt_create(const uint8_t *obs, int rows, int cols) {
    // rebuild a grayscale image whose InitializeInvironment(th=128) result is exactly `obs`
    std::vector<uint8_t> gray((size_t)rows * cols);
    for (size_t i = 0; i < gray.size(); i++) gray[i] = obs[i] ? 0 : 255;
    cv::shim_register_image("ref_map", rows, cols, gray.data());
    cv::shim_circle_hook() = orc_raster::circle_filled;
    cv::shim_fillpoly_hook() = orc_raster::fill_poly;
    RefRTRRT *h = new RefRTRRT();
    h->p.InitializeInvironment("ref_map", 128);
    h->p.DObsmap = cv::Mat::zeros(rows, cols, CV_8U);
    return h;
}
REF_EXPORT void ref_rtrrt_destroy(void *h) { delete (RefRTRRT *)h; }
REF_EXPORT void ref_rtrrt_set_agent(void *h, float x, float y) {
    RTpoint p{x, y};
    ((RefRTRRT *)h)->p.SetAgent(p);
}
REF_EXPORT void ref_rtrrt_set_goal(void *h, float x, float y) {
    RTpoint p{x, y};
    ((RefRTRRT *)h)->p.SetGoal(p);
}
REF_EXPORT void ref_rtrrt_set_dobs(void *h, const uint8_t *plane) {
    RTRRTplanner &p = ((RefRTRRT *)h)->p;
    p.DObsmap = cv::Mat::zeros(p.shapeY, p.shapeX, CV_8U);
    if (plane) memcpy(p.DObsmap.data(), plane, (size_t)p.shapeX * p.shapeY);
}
REF_EXPORT void ref_rtrrt_get_dobs(void *h, uint8_t *plane) {
    RTRRTplanner &p = ((RefRTRRT *)h)->p;
    memcpy(plane, p.DObsmap.data(), (size_t)p.shapeX * p.shapeY);
}
// type 0: circles at outline points; type 1: polygon (RTRRT.h:92-106)
REF_EXPORT void ref_rtrrt_add_obstacle(void *h, int num, int type, float x, float y, float angle, float radius,
                                       const float *outline_xy, int npts) {
    RTrrt_DynamicObstacle o;
    o.DObsType = type, o.x = x, o.y = y, o.angle = angle, o.radius = radius, o.sampleRadius = 0;
    for (int i = 0; i < npts; i++) o.OutLine.push_back({outline_xy[2 * i], outline_xy[2 * i + 1]});
    ((RefRTRRT *)h)->p.AddDynamicObstacle(num, o);
}
REF_EXPORT void ref_rtrrt_draw_dobs(void *h) { ((RefRTRRT *)h)->p.DrawDObsmap(); }
REF_EXPORT void ref_rtrrt_update_tree(void *h) { ((RefRTRRT *)h)->p.UpdateTree(); }
REF_EXPORT void ref_rtrrt_expand_rewiring(void *h, long ms) { ((RefRTRRT *)h)->p.ExpandRewiring(ms); }
REF_EXPORT void ref_rtrrt_crash(void *h, const float *xyxy, long long n, uint8_t *out) {
    RTRRTplanner &p = ((RefRTRRT *)h)->p;
    for (long long i = 0; i < n; i++) {
        RTpoint a{xyxy[4 * i], xyxy[4 * i + 1]}, b{xyxy[4 * i + 2], xyxy[4 * i + 3]};
        out[i] = p.CrashDetection(a, b);
    }
}
REF_EXPORT void ref_rtrrt_dist(const float *xyxy, long long n, float *out) {
    for (long long i = 0; i < n; i++) {
        RTpoint a{xyxy[4 * i], xyxy[4 * i + 1]}, b{xyxy[4 * i + 2], xyxy[4 * i + 3]};
        out[i] = a % b;
    }
}
REF_EXPORT int ref_rtrrt_num_nodes(void *h) { return (int)((RefRTRRT *)h)->p.Node_Map.size(); }
// nodes in insertion order (the stand-in Kd_tree keeps its points in insertion order)
REF_EXPORT int ref_rtrrt_dump(void *h, float *xy, int32_t *parent, float *cost, uint8_t *col) {
    RTRRTplanner &p = ((RefRTRRT *)h)->p;
    std::map<RTpoint, int> id;
    int n = (int)p.NeighborSearchTree.pts.size();
    for (int i = 0; i < n; i++) id[{(float)p.NeighborSearchTree.pts[i].x(), (float)p.NeighborSearchTree.pts[i].y()}] = i;
    for (int i = 0; i < n; i++) {
        RTpoint q{(float)p.NeighborSearchTree.pts[i].x(), (float)p.NeighborSearchTree.pts[i].y()};
        const RTrrt_NodeAttribute &a = p.Node_Map[q];
        xy[2 * i] = q.x, xy[2 * i + 1] = q.y;
        auto it = id.find(a.par);
        parent[i] = (a.par.x < 0 || it == id.end()) ? -1 : it->second;
        cost[i] = a.cost.cost;
        col[i] = a.cost.collision;
    }
    return n;
}
REF_EXPORT void ref_rtrrt_root(void *h, float *xy) {
    xy[0] = ((RefRTRRT *)h)->p.root.x, xy[1] = ((RefRTRRT *)h)->p.root.y;
}
REF_EXPORT int ref_rtrrt_queue_sizes(void *h, int *qr, int *qs) {
    *qr = (int)((RefRTRRT *)h)->p.Qr.size(), *qs = (int)((RefRTRRT *)h)->p.Qs.size();
    return 0;
}
REF_EXPORT int ref_rtrrt_get_path(void *h, float *out_xy, int cap, int *collision, float *cost, float *c_base) {
    RTRRTplanner &p = ((RefRTRRT *)h)->p;
    std::list<RTpoint> path;
    RTrrt_Cost c = p.GetPath(path);
    *collision = c.collision, *cost = c.cost, *c_base = p.c_base;
    int n = 0;
    for (const RTpoint &q : path) {
        if (n < cap) out_xy[2 * n] = q.x, out_xy[2 * n + 1] = q.y;
        n++;
    }
    return n;
}
