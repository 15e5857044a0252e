// ref_rtfmt.cpp — harness over the reference's OWN RT-FMT source, compiled in place
// (TEST INFRASTRUCTURE ONLY; output oracle/_ref/libref_rtfmt.so; no reference source is copied).
#include "ref_hooks.h"
#define private public
#include "RTFMT/RTFMT.cpp" // = /root/reference/RTFMT/lib/RTFMT/RTFMT.cpp via -I
#undef private
#include "ref_common.hpp"
#include "raster.hpp"

struct RefRTFMT {
    RTFMTplanner p;
};
static std::map<RTpoint, int> ids_of(RTFMTplanner &p) {
    std::map<RTpoint, int> id;
    for (int i = 0; i < (int)p.NeighborSearchTree.pts.size(); i++)
        id[{(float)p.NeighborSearchTree.pts[i].x(), (float)p.NeighborSearchTree.pts[i].y()}] = i;
    return id;
}
REF_EXPORT void *ref_rtfmt_create(const uint8_t *obs, int rows, int cols) {
    std::vector<uint8_t> gray((size_t)rows * cols);
    for (size_t i = 0; i < gray.size(); i++) gray[i] = obs[i] ? 0 : 255;
    cv::shim_register_image("ref_map", rows, cols, gray.data());
    cv::shim_circle_hook() = orc_raster::circle_filled;
    cv::shim_fillpoly_hook() = orc_raster::fill_poly;
    RefRTFMT *h = new RefRTFMT();
    h->p.InitializeInvironment("ref_map", 128);
    h->p.DObsmap = cv::Mat::zeros(rows, cols, CV_8U);
    return h;
}
REF_EXPORT void ref_rtfmt_destroy(void *h) { delete (RefRTFMT *)h; }
REF_EXPORT void ref_rtfmt_set_agent(void *h, float x, float y) {
    RTpoint p{x, y};
    ((RefRTFMT *)h)->p.SetAgent(p);
}
REF_EXPORT void ref_rtfmt_set_goal(void *h, float x, float y) {
    RTpoint p{x, y};
    ((RefRTFMT *)h)->p.SetGoal(p);
}
REF_EXPORT void ref_rtfmt_set_dobs(void *h, const uint8_t *plane) {
    RTFMTplanner &p = ((RefRTFMT *)h)->p;
    p.DObsmap = cv::Mat::zeros(p.shapeY, p.shapeX, CV_8U);
    if (plane) memcpy(p.DObsmap.data(), plane, (size_t)p.shapeX * p.shapeY);
}
// The reference hard-codes SampleN 2000 (RTFMT.cpp:392); Start() is called once per 2000 samples
// wanted.  Start() also resets V_open/z, which is idempotent before any expansion.
REF_EXPORT void ref_rtfmt_start(void *h) { ((RefRTFMT *)h)->p.Start(); }
REF_EXPORT void ref_rtfmt_update_tree(void *h) { ((RefRTFMT *)h)->p.UpdateTree(); }
REF_EXPORT void ref_rtfmt_expand(void *h, int iters, int with_rewire) {
    RTFMTplanner &p = ((RefRTFMT *)h)->p;
    for (int i = 0; i < iters; i++) {
        p.ExpandFMT();
        if (with_rewire) {
            p.RewireFromObstacles();
            p.RewireFromRoot();
        }
    }
}
REF_EXPORT void ref_rtfmt_crash(void *h, const float *xyxy, long long n, int ndobs, uint8_t *out) {
    RTFMTplanner &p = ((RefRTFMT *)h)->p;
    for (long long i = 0; i < n; i++) {
        RTpoint a{xyxy[4 * i], xyxy[4 * i + 1]}, b{xyxy[4 * i + 2], xyxy[4 * i + 3]};
        out[i] = ndobs ? p.CrashDetectionNDObs(a, b) : p.CrashDetection(a, b);
    }
}
REF_EXPORT int ref_rtfmt_num_nodes(void *h) { return (int)((RefRTFMT *)h)->p.NeighborSearchTree.pts.size(); }
REF_EXPORT int ref_rtfmt_dump(void *h, float *xy, int32_t *parent, float *cost, uint8_t *col, uint8_t *state) {
    RTFMTplanner &p = ((RefRTFMT *)h)->p;
    std::map<RTpoint, int> id = ids_of(p);
    int n = (int)p.NeighborSearchTree.pts.size();
    for (int i = 0; i < n; i++) {
        RTpoint q{(float)p.NeighborSearchTree.pts[i].x(), (float)p.NeighborSearchTree.pts[i].y()};
        const RTFMT_NodeAttribute &a = p.Node_Map[q];
        xy[2 * i] = q.x, xy[2 * i + 1] = q.y;
        auto it = id.find(a.par);
        parent[i] = (a.par.x < 0 || it == id.end()) ? -1 : it->second;
        cost[i] = a.cost.cost;
        col[i] = a.cost.collision;
        uint8_t s = 0;
        if (p.V_open.count(q)) s = 1;
        else if (p.V_opennew.count(q)) s = 2;
        else if (p.V_closed.count(q)) s = 3;
        if (p.V_toopen.count(q)) s |= 4;
        state[i] = s;
    }
    return n;
}
REF_EXPORT void ref_rtfmt_z(void *h, float *xy) { xy[0] = ((RefRTFMT *)h)->p.z.x, xy[1] = ((RefRTFMT *)h)->p.z.y; }
REF_EXPORT int ref_rtfmt_get_path(void *h, float *out_xy, int cap) {
    std::list<RTpoint> path;
    ((RefRTFMT *)h)->p.GetPath(path);
    int n = 0;
    for (const RTpoint &q : path) {
        if (n < cap) out_xy[2 * n] = q.x, out_xy[2 * n + 1] = q.y;
        n++;
    }
    return n;
}
