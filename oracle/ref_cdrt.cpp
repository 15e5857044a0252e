// ref_cdrt.cpp — harness over the reference's OWN CDRT-RRT* sources, compiled in place
// (TEST INFRASTRUCTURE ONLY; output oracle/_ref/libref_cdrt.so; no reference source is copied).
// Build: oracle/Makefile (-DREF_HOOK_STD_RAND so that std::random_shuffle, CDRT.cpp:330, draws from
// the injected stream; -I oracle/shim for the OpenCV/CGAL stand-ins; nlohmann/json is vendored in
// the reference tree and used as is).
//
// What is the reference's own compiled code here: BImap::StartCut and everything under it
// (CDRTmapcd.cpp, CDRTcommon.cpp), DrawBImap, RTRT_start, RTRT_Neighbor, RTRT_CrashDetection,
// RTRT_DrawDObsmap, RTRT_CrashDObsIsland, RTRT_UpdateTreeCost, RTRT_AddPotentialLinks,
// RTRT_Sampling, RTRT_task, GetPath, RTRT_AddDynamicObstacle (CDRT.cpp).  MaptoBInavi's UDP round
// trip to approx_work.py is bypassed: the polygon groups it would receive are handed in directly
// (tools/make_cdrt_fixtures.py runs approx_work's cv2 calls in Python).
#include "ref_hooks.h"
#include "CDRT/CDRT.cpp" // = /root/reference/CDRT2/lib/CDRT/CDRT.cpp via -I
#include "CDRT/CDRTcommon.cpp"
#include "CDRT/CDRTmapcd.cpp"
#include "ref_common.hpp"
#include "raster.hpp"

static void fill_convex16(uint16_t *img, int rows, int cols, const int *xy, int npts, uint16_t v) {
    orc_raster::fill_convex_poly<uint16_t>(img, rows, cols, xy, npts, v);
}

struct RefCDRT {
    BImap m;
};

REF_EXPORT void *ref_cdrt_create(int W, int H, double ratio, double rsize) {
    cv::shim_circle_hook() = orc_raster::circle_filled;
    cv::shim_fillpoly_hook() = orc_raster::fill_poly;
    cv::shim_fillconvex16_hook() = fill_convex16;
    RefCDRT *h = new RefCDRT();
    h->m.shapeX = W, h->m.shapeY = H, h->m.mapratio = ratio, h->m.robotsize = rsize;
    h->m.DObsmap = cv::Mat::zeros(H, W, CV_8U);
    return h;
}
REF_EXPORT void ref_cdrt_destroy(void *h) { delete (RefCDRT *)h; }

// one connected free region: polygon 0 = outer boundary, 1.. = holes (approx_work.py:get_approx_list)
REF_EXPORT int ref_cdrt_add_group(void *h, const double *pts, const int *poly_sizes, int npoly) {
    BImap &m = ((RefCDRT *)h)->m;
    BIpolygons polys(npoly);
    size_t o = 0;
    for (int i = 0; i < npoly; i++)
        for (int j = 0; j < poly_sizes[i]; j++, o++) polys[i].push_back({pts[2 * o], pts[2 * o + 1]});
    m.BIgraphList.emplace_back();
    m.StartCut(polys, m.BIgraphList.back());
    return (int)m.BIgraphList.size() - 1;
}
// install a ready-made graph (tests hand the fixture graph to reference and oracle alike)
REF_EXPORT int ref_cdrt_set_graph(void *h, int ncut, const double *cut_se, const double *cut_core, const int *cp_off, const int *cp,
                                  const int *cc_off, const int *cc, int npoly, const int *pv_off, const double *pv,
                                  const double *poly_core, const int *pc_off, const int *pc, const int *pp_off, const int *pp) {
    BImap &m = ((RefCDRT *)h)->m;
    m.BIgraphList.emplace_back();
    BIgraph &g = m.BIgraphList.back();
    for (int i = 0; i < ncut; i++) {
        BIcutline c{i, BIline({cut_se[4 * i], cut_se[4 * i + 1]}, {cut_se[4 * i + 2], cut_se[4 * i + 3]}), {cut_core[2 * i], cut_core[2 * i + 1]}, {}, {}};
        c.line.S = {cut_se[4 * i], cut_se[4 * i + 1]}, c.line.E = {cut_se[4 * i + 2], cut_se[4 * i + 3]};
        for (int j = cp_off[i]; j < cp_off[i + 1]; j++) c.polygonlink.insert(cp[j]);
        for (int j = cc_off[i]; j < cc_off[i + 1]; j++) c.cutlinelink.insert(cc[j]);
        g.cutlineList.push_back(c);
    }
    for (int i = 0; i < npoly; i++) {
        BIfreepolygon p;
        p.index = i;
        for (int j = pv_off[i]; j < pv_off[i + 1]; j++) p.polygon.push_back({pv[2 * j], pv[2 * j + 1]});
        p.core = {poly_core[2 * i], poly_core[2 * i + 1]};
        for (int j = pc_off[i]; j < pc_off[i + 1]; j++) p.cutlinelink.insert(pc[j]);
        for (int j = pp_off[i]; j < pp_off[i + 1]; j++) p.polygonlink.insert(pp[j]);
        g.freepolygonList.push_back(p);
    }
    g.cutlineBaseNum = ncut, g.freepolygonBaseNum = npoly;
    return (int)m.BIgraphList.size() - 1;
}
// sizes: [ncut, npoly, sum cut->poly, sum cut->cut, sum poly vertices, sum poly->cut, sum poly->poly]
REF_EXPORT void ref_cdrt_graph_sizes(void *h, int g, int *out) {
    const BIgraph &G = ((RefCDRT *)h)->m.BIgraphList[g];
    int s[7] = {(int)G.cutlineList.size(), (int)G.freepolygonList.size(), 0, 0, 0, 0, 0};
    for (const BIcutline &c : G.cutlineList) s[2] += c.polygonlink.size(), s[3] += c.cutlinelink.size();
    for (const BIfreepolygon &p : G.freepolygonList) s[4] += p.polygon.size(), s[5] += p.cutlinelink.size(), s[6] += p.polygonlink.size();
    for (int i = 0; i < 7; i++) out[i] = s[i];
}
REF_EXPORT void ref_cdrt_graph_dump(void *h, int g, double *cut_se, double *cut_core, int *cp_off, int *cp, int *cc_off, int *cc,
                                    int *pv_off, double *pv, double *poly_core, int *pc_off, int *pc, int *pp_off, int *pp) {
    const BIgraph &G = ((RefCDRT *)h)->m.BIgraphList[g];
    int a = 0, b = 0;
    for (size_t i = 0; i < G.cutlineList.size(); i++) {
        const BIcutline &c = G.cutlineList[i];
        cut_se[4 * i] = c.line.S.x, cut_se[4 * i + 1] = c.line.S.y, cut_se[4 * i + 2] = c.line.E.x, cut_se[4 * i + 3] = c.line.E.y;
        cut_core[2 * i] = c.core.x, cut_core[2 * i + 1] = c.core.y;
        cp_off[i] = a, cc_off[i] = b;
        for (int v : c.polygonlink) cp[a++] = v;
        for (int v : c.cutlinelink) cc[b++] = v;
    }
    cp_off[G.cutlineList.size()] = a, cc_off[G.cutlineList.size()] = b;
    int v0 = 0, c0 = 0, p0 = 0;
    for (size_t i = 0; i < G.freepolygonList.size(); i++) {
        const BIfreepolygon &p = G.freepolygonList[i];
        pv_off[i] = v0, pc_off[i] = c0, pp_off[i] = p0;
        for (const BIpoint &q : p.polygon) pv[2 * v0] = q.x, pv[2 * v0 + 1] = q.y, v0++;
        poly_core[2 * i] = p.core.x, poly_core[2 * i + 1] = p.core.y;
        for (int v : p.cutlinelink) pc[c0++] = v;
        for (int v : p.polygonlink) pp[p0++] = v;
    }
    pv_off[G.freepolygonList.size()] = v0, pc_off[G.freepolygonList.size()] = c0, pp_off[G.freepolygonList.size()] = p0;
}
// DrawBImap(false): the two 16-bit id planes (CDRT.cpp:121-175)
REF_EXPORT void ref_cdrt_draw_bimap(void *h, uint16_t *amap, uint16_t *imap) {
    BImap &m = ((RefCDRT *)h)->m;
    m.DrawBImap(false);
    if (amap) memcpy(amap, m.BIamap.data(), (size_t)m.shapeX * m.shapeY * 2);
    if (imap) memcpy(imap, m.BIimap.data(), (size_t)m.shapeX * m.shapeY * 2);
}
REF_EXPORT void ref_cdrt_set_planes(void *h, const uint16_t *amap, const uint16_t *imap) {
    BImap &m = ((RefCDRT *)h)->m;
    m.BIamap = cv::Mat(m.shapeY, m.shapeX, CV_16U), m.BIimap = cv::Mat(m.shapeY, m.shapeX, CV_16U);
    memcpy(m.BIamap.data(), amap, (size_t)m.shapeX * m.shapeY * 2);
    memcpy(m.BIimap.data(), imap, (size_t)m.shapeX * m.shapeY * 2);
}
REF_EXPORT void ref_cdrt_set_params(void *h, double r_th, double p_th, int nearK, int RewireN) {
    BImap &m = ((RefCDRT *)h)->m;
    m.r_th = r_th, m.p_th = p_th, m.nearK = nearK, m.RewireN = RewireN;
}
REF_EXPORT void ref_cdrt_set_agent(void *h, double x, double y) { ((RefCDRT *)h)->m.RTRT_SetAgent({x, y}); }
REF_EXPORT void ref_cdrt_set_goal(void *h, double x, double y) { ((RefCDRT *)h)->m.RTRT_SetGoal({x, y}); }
REF_EXPORT void ref_cdrt_add_obstacle(void *h, int num, int type, double x, double y, double angle, double radius,
                                      const double *outline_xy, int npts, double check_len) {
    CTrtrt_DynamicObstacle o;
    o.DObsType = type, o.x = x, o.y = y, o.angle = angle, o.radius = radius, o.sampleRadius = 0;
    for (int i = 0; i < npts; i++) o.OutLine.push_back({outline_xy[2 * i], outline_xy[2 * i + 1]});
    ((RefCDRT *)h)->m.RTRT_AddDynamicObstacle(num, o, check_len);
}
REF_EXPORT void ref_cdrt_move_obstacle(void *h, int num, double x, double y, double angle) {
    CTrtrt_DynamicObstacle &o = ((RefCDRT *)h)->m.DObstacles[num]; // demos mutate .x/.y in place (CDRT2/test/testD.cpp:189-190)
    o.x = x, o.y = y, o.angle = angle;
}
REF_EXPORT int ref_cdrt_check_points(void *h, int num, double *out, int cap) {
    const std::vector<BIpoint> &c = ((RefCDRT *)h)->m.DObstacles[num].CheckPoints;
    for (size_t i = 0; i < c.size() && (int)i < cap; i++) out[2 * i] = c[i].x, out[2 * i + 1] = c[i].y;
    return (int)c.size();
}
REF_EXPORT void ref_cdrt_start(void *h, double resolution) { ((RefCDRT *)h)->m.RTRT_start(resolution); }
REF_EXPORT void ref_cdrt_task(void *h) { ((RefCDRT *)h)->m.RTRT_task(); }
REF_EXPORT void ref_cdrt_draw_dobs(void *h, uint8_t *plane) {
    BImap &m = ((RefCDRT *)h)->m;
    m.RTRT_DrawDObsmap();
    if (plane) memcpy(plane, m.DObsmap.data(), (size_t)m.shapeX * m.shapeY);
}
REF_EXPORT void ref_cdrt_set_dobs(void *h, const uint8_t *plane) {
    BImap &m = ((RefCDRT *)h)->m;
    m.DObsmap = cv::Mat::zeros(m.shapeY, m.shapeX, CV_8U);
    if (plane) memcpy(m.DObsmap.data(), plane, (size_t)m.shapeX * m.shapeY);
}
REF_EXPORT void ref_cdrt_get_dobs(void *h, uint8_t *plane) {
    BImap &m = ((RefCDRT *)h)->m;
    memcpy(plane, m.DObsmap.data(), (size_t)m.shapeX * m.shapeY);
}
REF_EXPORT void ref_cdrt_crash_dobs_island(void *h) { ((RefCDRT *)h)->m.RTRT_CrashDObsIsland(); }
REF_EXPORT void ref_cdrt_update_tree_cost(void *h) { ((RefCDRT *)h)->m.RTRT_UpdateTreeCost(); }
// E5: full RTRT_CrashDetection (topology switch + fp64 ray march), CDRT.cpp:507-603
REF_EXPORT void ref_cdrt_crash(void *h, const double *xyxy, const int32_t *inv_s, const int32_t *inv_e, long long n, uint8_t *out) {
    BImap &m = ((RefCDRT *)h)->m;
    for (long long i = 0; i < n; i++)
        out[i] = m.RTRT_CrashDetection({xyxy[4 * i], xyxy[4 * i + 1]}, inv_s[i], {xyxy[4 * i + 2], xyxy[4 * i + 3]}, inv_e[i]);
}
REF_EXPORT int ref_cdrt_neighbor(void *h, double x, double y, int32_t invM, double *out, int cap) {
    std::list<BIpoint> near;
    ((RefCDRT *)h)->m.RTRT_Neighbor({x, y}, near, invM);
    int n = 0;
    for (const BIpoint &p : near) {
        if (n < cap) out[2 * n] = p.x, out[2 * n + 1] = p.y;
        n++;
    }
    return n;
}
REF_EXPORT int ref_cdrt_num_nodes(void *h) { return (int)((RefCDRT *)h)->m.Node_Map.size(); }
// nodes in Node_Map order (sorted by (x, y))
REF_EXPORT int ref_cdrt_dump(void *h, double *xy, double *par_xy, double *cost, uint8_t *col, int32_t *invM, int32_t *nsubs) {
    BImap &m = ((RefCDRT *)h)->m;
    int i = 0;
    for (const auto &kv : m.Node_Map) {
        xy[2 * i] = kv.first.x, xy[2 * i + 1] = kv.first.y;
        par_xy[2 * i] = kv.second.par.x, par_xy[2 * i + 1] = kv.second.par.y;
        cost[i] = kv.second.cost.cost, col[i] = kv.second.cost.collision, invM[i] = kv.second.invM;
        nsubs[i] = (int)kv.second.subs.size();
        i++;
    }
    return i;
}
// [agent.x, agent.y, agent_tag.x, agent_tag.y], [agent_apl, agent_island, eQr, larger, smaller, samples left, islands with trees, goalIsSet]
REF_EXPORT void ref_cdrt_state(void *h, double *d, int *s) {
    BImap &m = ((RefCDRT *)h)->m;
    d[0] = m.agent.x, d[1] = m.agent.y, d[2] = m.agent_tag.x, d[3] = m.agent_tag.y;
    s[0] = m.agent_apl, s[1] = m.agent_island, s[2] = (int)m.eQr.size(), s[3] = (int)m.largerSet.size(), s[4] = (int)m.smallerSet.size();
    s[5] = (int)m.NodeBSample_List.size(), s[6] = (int)m.obsKdtreeMap.size(), s[7] = m.goalIsSet;
}
REF_EXPORT int ref_cdrt_tree_islands(void *h, int32_t *out, int32_t *npts, int cap) {
    BImap &m = ((RefCDRT *)h)->m;
    int n = 0;
    for (const auto &kv : m.obsKdtreeMap) {
        if (n < cap) out[n] = kv.first, npts[n] = (int)kv.second.size();
        n++;
    }
    return n;
}
REF_EXPORT int ref_cdrt_samples(void *h, double *xy, int32_t *cut, int cap) {
    BImap &m = ((RefCDRT *)h)->m;
    int n = 0;
    for (const auto &s : m.NodeBSample_List) {
        if (n < cap) xy[2 * n] = s.first.x, xy[2 * n + 1] = s.first.y, cut[n] = s.second;
        n++;
    }
    return n;
}
REF_EXPORT int ref_cdrt_get_path(void *h, double *out_xy, int cap, int *collision, double *cost) {
    BImap &m = ((RefCDRT *)h)->m;
    CTrtrt_Cost c = m.GetPath();
    *collision = c.collision, *cost = c.cost;
    int n = 0;
    for (const BIpoint &q : m.agent_path) {
        if (n < cap) out_xy[2 * n] = q.x, out_xy[2 * n + 1] = q.y;
        n++;
    }
    return n;
}
