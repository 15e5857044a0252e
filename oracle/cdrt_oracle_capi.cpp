// cdrt_oracle_capi.cpp — C entry points of the CDRT-RRT* CPU oracle (TEST INFRASTRUCTURE ONLY).
// Same names and argument lists as the reference harness oracle/ref_cdrt.cpp (orc_cdrt_* vs
// ref_cdrt_*), so that one Python wrapper (oracle/loader.py:CdrtCpu) drives both side by side.
#include <cstring>

#include "cdrt_oracle.hpp"

using namespace orc;
#define ORC_EXPORT extern "C" __attribute__((visibility("default")))

struct HCDRT {
    CDRT p;
    std::vector<uint16_t> amap, imap;
    std::vector<int32_t> rand;
};
static CDRT &P(void *h) { return ((HCDRT *)h)->p; }

ORC_EXPORT void *orc_cdrt_create(int W, int H, double, double) {
    HCDRT *h = new HCDRT();
    h->p.W = W, h->p.H = H;
    h->p.dobs.assign((size_t)W * H, 0);
    return h;
}
ORC_EXPORT void orc_cdrt_destroy(void *h) { delete (HCDRT *)h; }
ORC_EXPORT int orc_cdrt_set_graph(void *h, int ncut, const double *cut_se, const double *cut_core, const int *cp_off, const int *cp,
                                  const int *cc_off, const int *cc, int npoly, const int *, const double *, const double *,
                                  const int *pc_off, const int *pc, const int *, const int *) {
    CdrtGraph g;
    g.cut.resize(ncut), g.poly.resize(npoly);
    for (int i = 0; i < ncut; i++) {
        g.cut[i].S = {cut_se[4 * i], cut_se[4 * i + 1]}, g.cut[i].E = {cut_se[4 * i + 2], cut_se[4 * i + 3]};
        g.cut[i].core = {cut_core[2 * i], cut_core[2 * i + 1]};
        g.cut[i].polys.assign(cp + cp_off[i], cp + cp_off[i + 1]);
        g.cut[i].cuts.assign(cc + cc_off[i], cc + cc_off[i + 1]);
    }
    for (int i = 0; i < npoly; i++) g.poly[i].cuts.assign(pc + pc_off[i], pc + pc_off[i + 1]);
    P(h).graphs.push_back(g);
    return (int)P(h).graphs.size() - 1;
}
ORC_EXPORT void orc_cdrt_set_planes(void *h, const uint16_t *amap, const uint16_t *imap) {
    HCDRT *H = (HCDRT *)h;
    size_t n = (size_t)H->p.W * H->p.H;
    H->amap.assign(amap, amap + n), H->imap.assign(imap, imap + n);
    H->p.amap = H->amap.data(), H->p.imap = H->imap.data();
}
// fill_convex_poly of one polygon into a 16-bit plane (the DrawBImap primitive, CDRT.cpp:143-144)
ORC_EXPORT void orc_fill_convex16(uint16_t *img, int rows, int cols, const int *xy, int npts, int value) {
    orc_raster::fill_convex_poly<uint16_t>(img, rows, cols, xy, npts, (uint16_t)value);
}
ORC_EXPORT void orc_cdrt_set_params(void *h, double r_th, double p_th, int nearK, int RewireN) {
    P(h).r_th = r_th, P(h).p_th = p_th, P(h).nearK = nearK, P(h).RewireN = RewireN;
}
ORC_EXPORT void orc_cdrt_set_clock(void *h, long long dt_ns, long long n0) { P(h).clk.dt_ns = dt_ns, P(h).clk.n = n0; }
ORC_EXPORT long long orc_cdrt_clock_reads(void *h) { return P(h).clk.n; }
ORC_EXPORT void orc_cdrt_set_rand(void *h, const int32_t *r, long long n) {
    HCDRT *H = (HCDRT *)h;
    H->rand.assign(r, r + n);
    H->p.rnd.r = H->rand.data(), H->p.rnd.n = n, H->p.rnd.pos = 0, H->p.rnd.exhausted = false;
}
ORC_EXPORT long long orc_cdrt_rand_pos(void *h) { return P(h).rnd.pos; }
ORC_EXPORT int orc_cdrt_rand_exhausted(void *h) { return P(h).rnd.exhausted; }
ORC_EXPORT void orc_cdrt_set_agent(void *h, double x, double y) { P(h).set_agent({x, y}); }
ORC_EXPORT void orc_cdrt_set_goal(void *h, double x, double y) { P(h).set_goal({x, y}); }
ORC_EXPORT void orc_cdrt_add_obstacle(void *h, int num, int type, double x, double y, double angle, double radius, const double *outline_xy,
                                      int npts, double check_len) {
    CdrtObstacle o;
    o.type = type, o.x = x, o.y = y, o.angle = angle, o.radius = radius;
    for (int i = 0; i < npts; i++) o.outline.push_back({outline_xy[2 * i], outline_xy[2 * i + 1]});
    P(h).add_obstacle(num, o, check_len);
}
ORC_EXPORT void orc_cdrt_move_obstacle(void *h, int num, double x, double y, double angle) {
    CdrtObstacle &o = P(h).obstacles[num];
    o.x = x, o.y = y, o.angle = angle;
}
ORC_EXPORT int orc_cdrt_check_points(void *h, int num, double *out, int cap) {
    const std::vector<P64> &c = P(h).obstacles[num].check;
    for (size_t i = 0; i < c.size() && (int)i < cap; i++) out[2 * i] = c[i].x, out[2 * i + 1] = c[i].y;
    return (int)c.size();
}
ORC_EXPORT void orc_cdrt_start(void *h, double resolution) { P(h).start(resolution); }
ORC_EXPORT void orc_cdrt_task(void *h) { P(h).task(); }
ORC_EXPORT void orc_cdrt_draw_dobs(void *h, uint8_t *plane) {
    P(h).draw_dobs();
    if (plane) memcpy(plane, P(h).dobs.data(), P(h).dobs.size());
}
ORC_EXPORT void orc_cdrt_set_dobs(void *h, const uint8_t *plane) {
    if (plane) memcpy(P(h).dobs.data(), plane, P(h).dobs.size());
    else std::fill(P(h).dobs.begin(), P(h).dobs.end(), 0);
}
ORC_EXPORT void orc_cdrt_get_dobs(void *h, uint8_t *plane) { memcpy(plane, P(h).dobs.data(), P(h).dobs.size()); }
ORC_EXPORT void orc_cdrt_crash_dobs_island(void *h) { P(h).crash_dobs_island(); }
ORC_EXPORT void orc_cdrt_update_tree_cost(void *h) { P(h).update_tree_cost(); }
ORC_EXPORT void orc_cdrt_crash(void *h, const double *xyxy, const int32_t *inv_s, const int32_t *inv_e, long long n, uint8_t *out) {
    for (long long i = 0; i < n; i++)
        out[i] = P(h).crash({xyxy[4 * i], xyxy[4 * i + 1]}, inv_s[i], {xyxy[4 * i + 2], xyxy[4 * i + 3]}, inv_e[i]);
}
ORC_EXPORT int orc_cdrt_neighbor(void *h, double x, double y, int32_t invM, double *out, int cap) {
    std::vector<P64> near;
    P(h).neighbor({x, y}, near, invM);
    for (size_t i = 0; i < near.size() && (int)i < cap; i++) out[2 * i] = near[i].x, out[2 * i + 1] = near[i].y;
    return (int)near.size();
}
ORC_EXPORT int orc_cdrt_num_nodes(void *h) { return (int)P(h).nodes.size(); }
ORC_EXPORT int orc_cdrt_dump(void *h, double *xy, double *par_xy, double *cost, uint8_t *col, int32_t *invM, int32_t *nsubs) {
    int i = 0;
    for (const auto &kv : P(h).nodes) {
        xy[2 * i] = kv.first.x, xy[2 * i + 1] = kv.first.y;
        par_xy[2 * i] = kv.second.par.x, par_xy[2 * i + 1] = kv.second.par.y;
        cost[i] = kv.second.cost.c, col[i] = kv.second.cost.col, invM[i] = kv.second.invM;
        nsubs[i] = (int)kv.second.subs.size();
        i++;
    }
    return i;
}
ORC_EXPORT void orc_cdrt_state(void *h, double *d, int *s) {
    CDRT &m = P(h);
    d[0] = m.agent.x, d[1] = m.agent.y, d[2] = m.agent_tag.x, d[3] = m.agent_tag.y;
    s[0] = m.agent_apl, s[1] = m.agent_island, s[2] = (int)m.eQr.size(), s[3] = (int)m.largerSet.size(), s[4] = (int)m.smallerSet.size();
    s[5] = (int)m.samples.size(), s[6] = (int)m.trees.size(), s[7] = m.goalIsSet;
}
ORC_EXPORT int orc_cdrt_tree_islands(void *h, int32_t *out, int32_t *npts, int cap) {
    int n = 0;
    for (const auto &kv : P(h).trees) {
        if (n < cap) out[n] = kv.first, npts[n] = (int)kv.second.size();
        n++;
    }
    return n;
}
ORC_EXPORT int orc_cdrt_samples(void *h, double *xy, int32_t *cut, int cap) {
    int n = 0;
    for (const auto &s : P(h).samples) {
        if (n < cap) xy[2 * n] = s.first.x, xy[2 * n + 1] = s.first.y, cut[n] = s.second;
        n++;
    }
    return n;
}
ORC_EXPORT int orc_cdrt_get_path(void *h, double *out_xy, int cap, int *collision, double *cost) {
    Cost64 c = P(h).get_path();
    *collision = c.col, *cost = c.c;
    int n = 0;
    for (const P64 &q : P(h).agent_path) {
        if (n < cap) out_xy[2 * n] = q.x, out_xy[2 * n + 1] = q.y;
        n++;
    }
    return n;
}
ORC_EXPORT int orc_cdrt_error(void *h) { return P(h).error; }
// [pops, rewired, sampled, edge_checks, knn, pushes]
ORC_EXPORT void orc_cdrt_counters(void *h, long long *out) {
    CDRT &m = P(h);
    out[0] = m.n_pops, out[1] = m.n_rewired, out[2] = m.n_sampled, out[3] = m.n_edge_checks, out[4] = m.n_knn, out[5] = m.n_pushes;
}
