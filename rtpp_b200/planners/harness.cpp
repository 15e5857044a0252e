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

// ------------------------------------------------------------------------------------- RT-FMT
#include "RTFMT.h"
PL_EXPORT void *pl_rtfmt_create(const uint8_t *obs, int rows, int cols, int device) {
    RTFMTplanner *p = new RTFMTplanner(device);
    if (!p->ok()) {
        delete p;
        return nullptr;
    }
    std::vector<uint8_t> gray((size_t)rows * cols);
    for (size_t i = 0; i < gray.size(); i++) gray[i] = obs[i] ? 0 : 255;
    p->InitializeInvironment(gray.data(), rows, cols, 128);
    return p;
}
PL_EXPORT void pl_rtfmt_destroy(void *h) { delete (RTFMTplanner *)h; }
PL_EXPORT void pl_rtfmt_set_rand(void *h, const int32_t *r, long long n) { ((RTFMTplanner *)h)->SetRandStream(r, n); }
PL_EXPORT long long pl_rtfmt_rand_pos(void *h) { return ((RTFMTplanner *)h)->RandPos(); }
PL_EXPORT void pl_rtfmt_set_agent(void *h, float x, float y) {
    RTpoint p{x, y};
    ((RTFMTplanner *)h)->SetAgent(p);
}
PL_EXPORT void pl_rtfmt_set_goal(void *h, float x, float y) {
    RTpoint p{x, y};
    ((RTFMTplanner *)h)->SetGoal(p);
}
PL_EXPORT void pl_rtfmt_set_dobs(void *h, const uint8_t *plane) { ((RTFMTplanner *)h)->SetDObsmap(plane); }
PL_EXPORT void pl_rtfmt_start(void *h, int sample_n) {
    ((RTFMTplanner *)h)->SetSampleN(sample_n);
    ((RTFMTplanner *)h)->Start();
}
PL_EXPORT void pl_rtfmt_update_tree(void *h) { ((RTFMTplanner *)h)->UpdateTree(); }
PL_EXPORT void pl_rtfmt_expand(void *h, int iters, int with_rewire) {
    RTFMTplanner *p = (RTFMTplanner *)h;
    if (iters <= 3) { // the reference's own call sequence, one method at a time
        for (int i = 0; i < iters; i++) {
            p->ExpandFMT();
            if (with_rewire) p->RewireFromObstacles(), p->RewireFromRoot();
        }
    } else
        p->RunRounds(iters, with_rewire != 0);
}
PL_EXPORT int pl_rtfmt_num_nodes(void *h) { return ((RTFMTplanner *)h)->NodeCount(); }
PL_EXPORT int pl_rtfmt_dump(void *h, float *xy, int32_t *parent, float *cost, uint8_t *col, uint8_t *state) {
    RTFMTplanner *p = (RTFMTplanner *)h;
    memcpy(xy, p->NodesXY().data(), p->NodesXY().size() * 4);
    return p->ReadTree(parent, cost, col, state);
}
PL_EXPORT void pl_rtfmt_z(void *h, float *xy) {
    RTpoint z = ((RTFMTplanner *)h)->Z();
    xy[0] = z.x, xy[1] = z.y;
}
PL_EXPORT int pl_rtfmt_get_path(void *h, float *out_xy, int cap) {
    std::list<RTpoint> path;
    ((RTFMTplanner *)h)->GetPath(path);
    int n = 0;
    for (const RTpoint &q : path) {
        if (n < cap) out_xy[2 * n] = q.x, out_xy[2 * n + 1] = q.y;
        n++;
    }
    return n;
}
PL_EXPORT int pl_rtfmt_set_sizes(void *h, int *sizes6) {
    RTFMTplanner *p = (RTFMTplanner *)h;
    p->SyncSets();
    sizes6[0] = (int)p->V_unv.size(), sizes6[1] = (int)p->V_open.size(), sizes6[2] = (int)p->V_opennew.size();
    sizes6[3] = (int)p->V_closed.size(), sizes6[4] = (int)p->V_toopen.size(), sizes6[5] = (int)p->Node_Map.size();
    return 0;
}

// ------------------------------------------------------------------------------------- AM-RRT*
#include "AMRRT.h"
PL_EXPORT void *pl_amrrt_create(const uint8_t *obs, int rows, int cols, int device) {
    AMRRTplanner *p = new AMRRTplanner(device);
    if (!p->ok()) {
        delete p;
        return nullptr;
    }
    std::vector<uint8_t> gray((size_t)rows * cols);
    for (size_t i = 0; i < gray.size(); i++) gray[i] = obs[i] ? 0 : 255;
    p->InitializeInvironment(gray.data(), rows, cols, 128);
    return p;
}
PL_EXPORT void pl_amrrt_destroy(void *h) { delete (AMRRTplanner *)h; }
PL_EXPORT void pl_amrrt_set_clock(void *h, long long dt) { ((AMRRTplanner *)h)->SetVirtualClock(dt); }
PL_EXPORT void pl_amrrt_set_rand(void *h, const int32_t *r, long long n) { ((AMRRTplanner *)h)->SetRandStream(r, n); }
PL_EXPORT long long pl_amrrt_rand_pos(void *h) { return ((AMRRTplanner *)h)->RandPos(); }
PL_EXPORT void pl_amrrt_set_agent(void *h, float x, float y) {
    RTpoint p{x, y};
    ((AMRRTplanner *)h)->SetAgent(p);
}
PL_EXPORT void pl_amrrt_set_goal(void *h, float x, float y) {
    RTpoint p{x, y};
    ((AMRRTplanner *)h)->SetGoal(p);
}
PL_EXPORT void pl_amrrt_set_dobs(void *h, const uint8_t *plane) { ((AMRRTplanner *)h)->SetDObsmap(plane); }
PL_EXPORT void pl_amrrt_set_diffusion(void *h, const int32_t *cells, int n, const float *m, int dim, float ratio) {
    ((AMRRTplanner *)h)->InitializeDiffusion(cells, n, m, dim, ratio);
}
PL_EXPORT float pl_amrrt_diffusivity(void *h, float ax, float ay, float bx, float by) {
    RTpoint a{ax, ay}, b{bx, by};
    return ((AMRRTplanner *)h)->DiffusivityMetric(a, b);
}
PL_EXPORT void pl_amrrt_set_outpath(void *h, int nonempty) {
    AMRRTplanner *p = (AMRRTplanner *)h;
    p->OutPath.clear();
    if (nonempty) p->OutPath.push_back({0, 0});
}
PL_EXPORT void pl_amrrt_main_task(void *h) { ((AMRRTplanner *)h)->MainTask(); }
PL_EXPORT int pl_amrrt_num_nodes(void *h) { return ((AMRRTplanner *)h)->NodeCount(); }
PL_EXPORT int pl_amrrt_dump(void *h, float *xy, int32_t *parent, float *cost, uint8_t *col) {
    AMRRTplanner *p = (AMRRTplanner *)h;
    return p->ReadTree(xy, parent, cost, col, p->NodeCount());
}
PL_EXPORT int pl_amrrt_get_path(void *h, float *out_xy, int cap, int *collision, float *cost) {
    std::list<RTpoint> path;
    RTrrt_Cost c = ((AMRRTplanner *)h)->GetPath(path);
    *collision = c.collision, *cost = c.cost;
    int n = 0;
    for (const RTpoint &q : path) {
        if (n < cap) out_xy[2 * n] = q.x, out_xy[2 * n + 1] = q.y;
        n++;
    }
    return n;
}

// ------------------------------------------------------------------------------------- CDRT-RRT*
#include "CDRT.h"
PL_EXPORT void *pl_cdrt_create(int W, int H, int device) {
    BImap *p = new BImap(device);
    if (!p->ok()) {
        delete p;
        return nullptr;
    }
    p->SetShape((uint32_t)W, (uint32_t)H);
    return p;
}
PL_EXPORT void pl_cdrt_destroy(void *h) { delete (BImap *)h; }
PL_EXPORT int pl_cdrt_set_graph(void *h, int ncut, const double *cut_se, const double *cut_core, const int *cp_off, const int *cp, const int *cc_off,
                                const int *cc, int npoly, const int *pv_off, const double *pv, const double *poly_core, const int *pc_off,
                                const int *pc, const int *pp_off, const int *pp) {
    return ((BImap *)h)->SetGraph(ncut, cut_se, cut_core, cp_off, cp, cc_off, cc, npoly, pv_off, pv, poly_core, pc_off, pc, pp_off, pp);
}
PL_EXPORT void pl_cdrt_draw_bimap(void *h, uint16_t *amap, uint16_t *imap) {
    BImap *p = (BImap *)h;
    p->DrawBImap(false);
    if (amap) memcpy(amap, p->BIamap.data.data(), p->BIamap.data.size() * 2);
    if (imap) memcpy(imap, p->BIimap.data.data(), p->BIimap.data.size() * 2);
}
PL_EXPORT void pl_cdrt_set_params(void *h, double r_th, double p_th, int nearK, int RewireN) {
    BImap *p = (BImap *)h;
    p->r_th = r_th, p->p_th = p_th, p->nearK = nearK, p->RewireN = RewireN;
}
PL_EXPORT void pl_cdrt_set_clock(void *h, long long dt) { ((BImap *)h)->SetVirtualClock(dt); }
PL_EXPORT void pl_cdrt_set_rand(void *h, const int32_t *r, long long n) { ((BImap *)h)->SetRandStream(r, n); }
PL_EXPORT void pl_cdrt_set_agent(void *h, double x, double y) { ((BImap *)h)->RTRT_SetAgent({x, y}); }
PL_EXPORT void pl_cdrt_set_goal(void *h, double x, double y) { ((BImap *)h)->RTRT_SetGoal({x, y}); }
PL_EXPORT void pl_cdrt_add_obstacle(void *h, int num, int type, double x, double y, double angle, double radius, const double *outline_xy, int npts,
                                    double check_len) {
    CTrtrt_DynamicObstacle o;
    o.DObsType = type, o.x = x, o.y = y, o.angle = angle, o.radius = radius, o.sampleRadius = 0;
    for (int i = 0; i < npts; i++) o.OutLine.push_back({outline_xy[2 * i], outline_xy[2 * i + 1]});
    ((BImap *)h)->RTRT_AddDynamicObstacle(num, o, check_len);
}
PL_EXPORT void pl_cdrt_move_obstacle(void *h, int num, double x, double y, double angle) {
    CTrtrt_DynamicObstacle &o = ((BImap *)h)->DObstacles[num]; // what the demos do (CDRT2/test/testD.cpp:189-190)
    o.x = x, o.y = y, o.angle = angle;
}
PL_EXPORT void pl_cdrt_start(void *h, double resolution) { ((BImap *)h)->RTRT_start(resolution); }
PL_EXPORT void pl_cdrt_task(void *h) { ((BImap *)h)->RTRT_task(); }
PL_EXPORT int pl_cdrt_get_path(void *h, double *out_xy, int cap, int *collision, double *cost) {
    BImap *p = (BImap *)h;
    CTrtrt_Cost c = p->GetPath();
    *collision = c.collision, *cost = c.cost;
    int n = 0;
    for (const BIpoint &q : p->agent_path) {
        if (n < cap) out_xy[2 * n] = q.x, out_xy[2 * n + 1] = q.y;
        n++;
    }
    return n;
}
PL_EXPORT int pl_cdrt_stats(void *h, long long *out24) { return ((BImap *)h)->ReadStats(out24); }
PL_EXPORT int pl_cdrt_num_nodes(void *h) {
    BImap *p = (BImap *)h;
    p->SyncNodeMap();
    return (int)p->Node_Map.size();
}
// Node_Map order (after pl_cdrt_num_nodes): xy, parent xy, cost, collision, invM, #children
PL_EXPORT int pl_cdrt_dump(void *h, double *xy, double *par_xy, double *cost, uint8_t *col, int32_t *invM, int32_t *nsubs) {
    BImap *p = (BImap *)h;
    int i = 0;
    for (const auto &kv : p->Node_Map) {
        xy[2 * i] = kv.first.x, xy[2 * i + 1] = kv.first.y;
        par_xy[2 * i] = kv.second.par.x, par_xy[2 * i + 1] = kv.second.par.y;
        cost[i] = kv.second.cost.cost, col[i] = kv.second.cost.collision, invM[i] = kv.second.invM;
        nsubs[i] = (int)kv.second.subs.size();
        i++;
    }
    return i;
}
// host fillConvexPoly alone (no GPU needed)
PL_EXPORT void pl_fill_convex16(uint16_t *img, int rows, int cols, const int *xy, int npts, int v) {
    rtpp::Plane16 p;
    p.rows = rows, p.cols = cols, p.data.assign(img, img + (size_t)rows * cols);
    std::vector<std::pair<int, int>> pts;
    for (int i = 0; i < npts; i++) pts.push_back({xy[2 * i], xy[2 * i + 1]});
    rtpp::fill_convex_polygon(p, pts, (uint16_t)v);
    memcpy(img, p.data.data(), p.data.size() * 2);
}
