// cdrt_oracle.hpp — CPU ORACLE of the CDRT-RRT* replanning loop (test infrastructure, NOT product).
//
// Restates BImap's RTRT_* member functions of the reference (CDRT2/lib/CDRT/CDRT.cpp) over the
// same containers (point-keyed std::map / std::set / priority queue), so that every quirk of
// operator[] insertion and set ordering comes out the same.  Out of scope and taken as input:
// the convex decomposition (CDRTmapcd.cpp StartCut) and MaptoBInavi/approx_work.py — the graph and
// the two id planes arrive as fixtures made by the reference's own code (tools/make_cdrt_fixtures.py).
//
// PINNED: tests/test_oracle_cdrt.py runs this against oracle/_ref/libref_cdrt.so (the reference's
// CDRT.cpp compiled in place) tick by tick under the same virtual clock and injected rand()
// stream, and tests/golden/cdrt_*.npz holds the resulting vectors for the GPU box.
// UNPINNED library boundaries, as everywhere: CGAL k-NN order among equal distances (here:
// insertion order), cv::circle / cv::fillPoly (oracle/raster.hpp, pinned to Python cv2 4.13).
// Where BIpoint::operator==/!= use a 1e-16 tolerance (CDRT.h:39-48) this file compares exactly:
// identical for every |coordinate| >= 0.5, and free space never comes that close to the border.
#pragma once
#include <cmath>
#include <cstdint>
#include <list>
#include <map>
#include <queue>
#include <set>
#include <vector>

#include "raster.hpp"
#include "rtpp_oracle.hpp"

namespace orc {

struct P64 { // BIpoint, CDRT.h:34-90
    double x, y;
    bool operator==(const P64 &o) const { return x == o.x && y == o.y; }
    bool operator!=(const P64 &o) const { return !(*this == o); }
    bool operator<(const P64 &o) const { return x != o.x ? x < o.x : y < o.y; }
};
inline double dist64(const P64 &a, const P64 &b) { return dist64(a.x, a.y, b.x, b.y); }

struct Cost64 { // CTrtrt_Cost, CDRT.h:92-131
    bool col;
    double c;
    Cost64 operator+(const Cost64 &o) const { return {col || o.col, c + o.c}; }
    bool operator<(const Cost64 &o) const { return col == o.col ? c < o.c : col < o.col; }
    bool operator>(const Cost64 &o) const { return col == o.col ? c > o.c : col > o.col; }
};
constexpr double kDoubleMax = 1.79769e+307; // CDRT.h:359
constexpr int32_t kOrigin = 0x40000000;     // originIndexFun, CDRT.h:226
inline int32_t obs_inv(int32_t island) { return (int32_t)(0x80000000u | (uint32_t)island); }

struct CdrtGraph { // the parts of BIgraph (CDRT.h:240-248) the planner loop reads; links ascending like std::set
    struct Cut {
        P64 S, E, core;
        std::vector<int> polys, cuts;
    };
    struct Poly {
        std::vector<int> cuts;
    };
    std::vector<Cut> cut;
    std::vector<Poly> poly;
};
inline bool has(const std::vector<int> &v, int x) { return std::binary_search(v.begin(), v.end(), x); }

struct CdrtObstacle { // CTrtrt_DynamicObstacle, CDRT.h:250-261
    int type;
    double x, y, angle, radius;
    std::vector<P64> outline, check;
};

struct CDRT {
    struct Node { // CDRT.h:230-238
        Cost64 cost{false, 0};
        P64 par{0, 0};
        std::set<P64> subs;
        int32_t invM = 0;
    };
    struct QE {
        Cost64 c;
        P64 S, E;
        // std::greater<VEpair>: min-heap on (cost, edge.S, edge.E), CDRT.h:263-281,328
        bool operator<(const QE &o) const { // "lower priority than o"
            if (c < o.c) return false;
            if (o.c < c) return true;
            if (S != o.S) return o.S < S;
            return o.E < E;
        }
    };

    int W = 0, H = 0;
    const uint16_t *amap = nullptr, *imap = nullptr;
    std::vector<uint8_t> dobs;
    std::vector<CdrtGraph> graphs;
    double r_th = 10, p_th = 0.85;
    int nearK = 20, RewireN = 100;
    P64 agent{0, 0}, agent_tag{0, 0}, goal{0, 0};
    int agent_apl = 0, agent_island = 0;
    bool goalIsSet = false;
    std::list<P64> agent_path;
    std::map<P64, Node> nodes;
    std::list<std::pair<P64, int>> samples;     // NodeBSample_List
    std::map<int, std::vector<P64>> trees;      // obsKdtreeMap: points in insertion order
    std::map<int, CdrtObstacle> obstacles;
    std::map<int, std::set<P64>> Mnode;
    std::priority_queue<QE> eQr;
    std::set<P64> largerSet, smallerSet;
    VClock clk;
    RandStream rnd;
    bool timers_init = false;
    long long updateTimeOld = 0, smallerTimeOld = 0;
    int error = 0; // 1: the reference would have hit undefined behaviour here (empty path, ...)
    long long n_pops = 0, n_rewired = 0, n_sampled = 0, n_edge_checks = 0, n_knn = 0, n_pushes = 0;

    uint16_t id16(const uint16_t *P, int row, int col) const { // cv::Mat::at<uint16_t> (row wrap; outside = 0xFFFF)
        long long lin = (long long)row * W + col;
        if (lin < 0 || lin >= (long long)W * H) return 0xFFFF;
        return P[lin];
    }
    bool dprobe(int row, int col) const { return probe(dobs.data(), W, H, row, col); }
    const CdrtGraph &G() const { return graphs[agent_apl]; }

    // CGAL::Orthogonal_k_neighbor_search over one island's points (CDRT.cpp:205,231,261,890)
    void knn(const std::vector<P64> &pts, const P64 &q, int k, std::vector<std::pair<P64, double>> &out) {
        n_knn++;
        std::vector<Nbr> c(pts.size());
        for (size_t i = 0; i < pts.size(); i++) c[i] = {sqdist64(q.x, q.y, pts[i].x, pts[i].y), (int)i};
        size_t kk = std::min<size_t>((size_t)k, c.size());
        std::partial_sort(c.begin(), c.begin() + kk, c.end());
        out.clear();
        for (size_t j = 0; j < kk; j++) out.push_back({pts[c[j].id], c[j].d2});
    }

    // RTRT_Neighbor, CDRT.cpp:177-274
    void neighbor(const P64 &x, std::vector<P64> &Xnear, int32_t invM) {
        Xnear.clear();
        const CdrtGraph &g = G();
        std::vector<std::pair<P64, double>> nn;
        auto cut_nodes = [&](int c) {
            auto it = Mnode.find(c);
            if (it != Mnode.end()) Xnear.insert(Xnear.end(), it->second.begin(), it->second.end());
        };
        if (invM == kOrigin) {
            for (int c : g.poly[agent_island].cuts) cut_nodes(c);
            auto t = trees.find(agent_island);
            if (t != trees.end()) {
                knn(t->second, agent, nearK, nn);
                for (auto &p : nn) Xnear.push_back(p.first);
            }
        } else if ((uint32_t)invM & 0xC0000000u) {
            const int island = invM & 0x3FFFFFFF;
            for (int c : g.poly[island].cuts) cut_nodes(c);
            auto t = trees.find(island);
            if (t != trees.end()) {
                knn(t->second, x, nearK, nn);
                for (auto &p : nn)
                    if (!(p.second < 1e-16)) Xnear.push_back(p.first);
            }
            if (island == agent_island) Xnear.push_back(agent);
        } else {
            const int ci = invM & 0x3FFFFFFF;
            for (int c : g.cut[ci].cuts) cut_nodes(c);
            for (int island : g.cut[ci].polys) {
                auto t = trees.find(island);
                if (t != trees.end()) {
                    knn(t->second, x, nearK, nn);
                    for (auto &p : nn)
                        if (!(p.second < 1e-16)) Xnear.push_back(p.first);
                }
                if (island == agent_island) Xnear.push_back(agent);
            }
        }
    }
    void neighbor(const P64 &x, std::vector<P64> &Xnear) { // CDRT.cpp:177-182
        Xnear.clear();
        auto it = nodes.find(x);
        if (it == nodes.end()) return;
        neighbor(x, Xnear, it->second.invM);
    }

    // RTRT_CrashDetection, CDRT.cpp:507-603: topology switch (E5a) then the dynamic-plane march (E5b)
    bool crash(const P64 &ps, int32_t invMps, const P64 &pe, int32_t invMpe) {
        n_edge_checks++;
        const CdrtGraph &g = G();
        bool march = false;
        const int is = invMps & 0x3FFFFFFF, ie = invMpe & 0x3FFFFFFF;
        const uint32_t ks = (uint32_t)invMps & 0xC0000000u, ke = (uint32_t)invMpe & 0xC0000000u;
        if (ks == 0) {
            if (ke == 0) {
                if (!has(g.cut[is].cuts, ie)) return true;
                int island = g.cut[is].polys.front();
                if (!has(g.cut[ie].polys, island)) island = g.cut[is].polys.back();
                if (trees.count(island)) march = true;
            } else if (ke == 0x40000000u) {
                if (!has(g.poly[agent_island].cuts, is)) return true;
                if (trees.count(agent_island)) march = true;
            } else if (ke == 0x80000000u) {
                if (!has(g.cut[is].polys, ie)) return true;
                if (trees.count(ie)) march = true;
            }
        } else if (ks == 0x40000000u) {
            if (ke == 0) {
                if (!has(g.poly[agent_island].cuts, ie)) return true;
            } else if (ke == 0x80000000u) {
                if (agent_island != ie) return true;
            }
            if (trees.count(agent_island)) march = true;
        } else if (ks == 0x80000000u) {
            if (ke == 0) {
                if (!has(g.poly[is].cuts, ie)) return true;
            } else if (ke == 0x40000000u) {
                if (agent_island != is) return true;
            } else if (ke == 0x80000000u) {
                if (is != ie) return true;
            }
            if (trees.count(is)) march = true;
        }
        if (!march) return false;
        Map m;
        m.D = dobs.data(), m.W = W, m.H = H;
        return edge_blocked_f64(m, ps.x, ps.y, pe.x, pe.y, 10, PLANE_D);
    }

    // RTRT_AddDynamicObstacle, CDRT.cpp:425-500 (binary32 intermediates as written there)
    void add_obstacle(int num, const CdrtObstacle &o_in, double CheckLen) {
        CdrtObstacle &o = obstacles[num] = o_in;
        if (!(CheckLen > 0)) return;
        o.check.clear();
        if (o.type == 0) {
            float radius = (float)o.radius;
            float two_pi = 2 * 3.14159265f;
            int n = (int)((double)(two_pi * radius) / CheckLen);
            float dangle = two_pi / (float)n;
            for (const P64 &p : o.outline)
                for (int i = 0; i < n; i++) {
                    float a = dangle * (float)i;
                    float x = (float)((double)(sinf(a) * radius) + p.x);
                    float y = (float)((double)(cosf(a) * radius) + p.y);
                    o.check.push_back({(double)x, (double)y});
                }
        } else if (o.type == 1) {
            const size_t nv = o.outline.size();
            for (size_t e = 0; e < nv; e++) { // edges (0,1), (1,2), ..., (nv-1, 0)
                const P64 &a = o.outline[e], &b = o.outline[(e + 1) % nv];
                float dx = (float)(b.x - a.x), dy = (float)(b.y - a.y);
                float length = std::sqrt(dx * dx + dy * dy);
                uint32_t n = (uint32_t)((double)length / CheckLen) + 1;
                for (uint32_t i = 0; i < n; i++) {
                    float t = (float)i / (float)n;
                    o.check.push_back({a.x + (double)(t * dx), a.y + (double)(t * dy)});
                }
            }
        }
    }
    static void obstacle_pixel(const CdrtObstacle &o, float s, float c, const P64 &p, int *ix, int *iy) { // CDRT.cpp:617-618
        *ix = (int)((double)c * p.x - (double)s * p.y + o.x);
        *iy = (int)((double)s * p.x + (double)c * p.y + o.y);
    }
    // RTRT_DrawDObsmap, CDRT.cpp:605-636
    void draw_dobs() {
        dobs.assign((size_t)W * H, 0);
        for (const auto &kv : obstacles) {
            const CdrtObstacle &o = kv.second;
            float s = sinf((float)o.angle), c = cosf((float)o.angle);
            if (o.type == 0) {
                for (const P64 &p : o.outline) {
                    int ix, iy;
                    obstacle_pixel(o, s, c, p, &ix, &iy);
                    orc_raster::circle_filled(dobs.data(), H, W, ix, iy, (int)o.radius, (uint8_t)127);
                }
            } else if (o.type == 1) {
                std::vector<int> xy;
                for (const P64 &p : o.outline) {
                    int ix, iy;
                    obstacle_pixel(o, s, c, p, &ix, &iy);
                    xy.push_back(ix), xy.push_back(iy);
                }
                orc_raster::fill_poly(dobs.data(), H, W, xy.data(), (int)o.outline.size(), (uint8_t)127);
            }
        }
    }
    // islands touched by the obstacles' check points, CDRT.cpp:644-668
    void covered_islands(std::set<int> &out) {
        for (const auto &kv : obstacles) {
            const CdrtObstacle &o = kv.second;
            float s = sinf((float)o.angle), c = cosf((float)o.angle);
            for (const P64 &p : o.check) {
                int ix, iy;
                obstacle_pixel(o, s, c, p, &ix, &iy);
                if (ix >= W || iy >= H || ix < 0 || iy < 0) continue;
                if ((int)id16(amap, iy, ix) != agent_apl) continue;
                int inum = id16(imap, iy, ix);
                if (!(inum & 0xC000)) out.insert(inum);
            }
        }
    }
    void prune_subs(Node &n) {
        for (auto it = n.subs.begin(); it != n.subs.end();)
            if (!nodes.count(*it)) it = n.subs.erase(it);
            else ++it;
    }
    // RTRT_CrashDObsIsland, CDRT.cpp:638-762
    void crash_dobs_island() {
        const CdrtGraph &g = G();
        std::set<int> now, gone;
        covered_islands(now);
        for (const auto &kv : trees)
            if (now.count(kv.first)) now.erase(kv.first);
            else gone.insert(kv.first);
        std::set<int> seen;
        for (int island : gone)
            for (int c : g.poly[island].cuts) {
                if (seen.count(c)) continue;
                for (const P64 &x : Mnode[c]) {
                    Node &xn = nodes[x];
                    Node &pn = nodes[xn.par];
                    if (((uint32_t)pn.invM & 0xC0000000u) == 0x80000000u && gone.count(pn.invM & 0x3FFFFFFF)) {
                        P64 a = xn.par; // first ancestor that is not an island node
                        while (((uint32_t)nodes[a].invM & 0xC0000000u) == 0x80000000u) a = nodes[a].par;
                        xn.par = a;
                        nodes[a].subs.insert(x);
                    }
                }
                seen.insert(c);
            }
        for (int island : gone) {
            for (const P64 &p : trees[island]) nodes.erase(p);
            trees.erase(island);
        }
        seen.clear();
        for (int island : gone) {
            if (island == agent_island) prune_subs(nodes[agent]);
            for (int c : g.poly[island].cuts) {
                if (seen.count(c)) continue;
                for (const P64 &x : Mnode[c]) prune_subs(nodes[x]);
                seen.insert(c);
            }
        }
        for (int k : now) trees[k];
    }
    // RTRT_UpdateTreeCost, CDRT.cpp:764-796
    void update_tree_cost() {
        std::queue<P64> q;
        q.push(agent);
        while (!q.empty()) {
            P64 p = q.front();
            Node &pn = nodes[p];
            for (const P64 &s : pn.subs) {
                Node &sn = nodes[s];
                Cost64 c = pn.cost + Cost64{crash(p, pn.invM, s, sn.invM), dist64(p, s)};
                if (sn.cost < c) {
                    largerSet.insert(s);
                    smallerSet.erase(s);
                } else if (c < sn.cost)
                    smallerSet.insert(s);
                sn.cost = c;
                q.push(s);
            }
            q.pop();
        }
    }
    void push(const Cost64 &c, const P64 &S, const P64 &E) {
        n_pushes++;
        eQr.push({c, S, E});
    }
    // RTRT_AddPotentialLinks, CDRT.cpp:797-830
    void add_potential_links(const P64 &p, bool larger) {
        std::vector<P64> Xnear;
        neighbor(p, Xnear);
        Node &pn = nodes[p];
        for (const P64 &q : Xnear) {
            Node &qn = nodes[q];
            if (larger) {
                Cost64 h = qn.cost + Cost64{false, dist64(p, q)};
                if (pn.cost > h) {
                    if (goalIsSet) h.c += dist64(p, goal);
                    push(h, q, p);
                }
            } else {
                Cost64 h = pn.cost + Cost64{false, dist64(p, q)};
                if (qn.cost > h) {
                    if (goalIsSet) h.c += dist64(q, goal);
                    push(h, p, q);
                }
            }
        }
    }
    // choose-parent shared by RTRT_Sampling (CDRT.cpp:848-864, 903-919) and GetPath (:1094-1108)
    void best_parent(const P64 &x, int32_t invM, const std::vector<P64> &Xnear, Cost64 *cmin, P64 *pmin) {
        *cmin = {true, kDoubleMax};
        *pmin = {-1, -1};
        for (const P64 &q : Xnear) {
            Node &qn = nodes[q];
            Cost64 c = qn.cost + Cost64{crash(x, invM, q, qn.invM), dist64(x, q)};
            if (*cmin > c) *cmin = c, *pmin = q;
        }
    }
    // RTRT_Sampling, CDRT.cpp:832-925
    bool sampling(P64 *out, std::vector<P64> &Xnear) {
        float p1 = rnd.u();
        if ((double)p1 < p_th || trees.empty()) {
            if (samples.empty()) return false;
            P64 x = samples.front().first;
            int ci = samples.front().second;
            samples.pop_front();
            nodes[x].invM = ci;
            neighbor(x, Xnear);
            Cost64 c;
            P64 par;
            best_parent(x, ci, Xnear, &c, &par);
            Node &xn = nodes[x];
            xn.cost = c, xn.par = par;
            nodes[par].subs.insert(x);
            Mnode[ci].insert(x);
            *out = x;
            return true;
        }
        for (int err = 0; err < 20; err++) {
            bool ok = false;
            double x = 0, y = 0;
            int island = 0;
            for (int rej = 0; rej < 100; rej++) {
                x = (double)(rnd.u() * (float)(uint32_t)W);
                y = (double)(rnd.u() * (float)(uint32_t)H);
                int ix = (int)x, iy = (int)y;
                if (dprobe(iy, ix) || (int)id16(amap, iy, ix) != agent_apl) continue;
                island = id16(imap, iy, ix);
                if (trees.count(island)) {
                    ok = true;
                    break;
                }
            }
            if (!ok) continue;
            std::vector<std::pair<P64, double>> nn;
            knn(trees[island], {x, y}, 1, nn);
            if (!nn.empty() && nn[0].second < 10 * r_th * r_th) continue;
            P64 p{x, y};
            const int32_t inv = obs_inv(island);
            nodes[p].invM = inv;
            neighbor(p, Xnear);
            Cost64 c;
            P64 par;
            best_parent(p, inv, Xnear, &c, &par);
            Node &pn = nodes[p];
            pn.cost = c, pn.par = par;
            nodes[par].subs.insert(p);
            trees[island].push_back(p);
            *out = p;
            return true;
        }
        return false;
    }
    bool placeable(const P64 &p) const { // the checks shared by SetAgent/SetGoal/GetPath, CDRT.cpp:386-398
        if (p.x >= (double)(uint32_t)W || p.y >= (double)(uint32_t)H) return false;
        uint32_t a = id16(amap, (int)p.y, (int)p.x), i = id16(imap, (int)p.y, (int)p.x);
        return !(a == 0xffff || (i & 0xC000));
    }
    void set_agent(const P64 &p) {
        if (placeable(p)) agent_tag = p;
    }
    void set_goal(const P64 &p) {
        if (placeable(p)) goal = p, goalIsSet = true;
    }
    // GetPath(point, path), CDRT.cpp:1076-1123
    Cost64 get_path(const P64 &point, std::list<P64> &path) {
        path.clear();
        if (!placeable(point)) return {true, kDoubleMax};
        const int32_t inv = obs_inv(id16(imap, (int)point.y, (int)point.x));
        std::vector<P64> Xnear;
        neighbor(point, Xnear, inv);
        Cost64 c;
        P64 par;
        best_parent(point, inv, Xnear, &c, &par);
        path.push_front(point);
        P64 xi = par;
        while (xi.x >= 0) {
            path.push_front(xi);
            xi = nodes[xi].par;
            if (path.size() > 2000) {
                path.clear();
                return {true, (double)3.4e38f};
            }
        }
        return c;
    }
    Cost64 get_path() { return goalIsSet ? get_path(goal, agent_path) : Cost64{true, kDoubleMax}; }

    // RTRT_start, CDRT.cpp:276-383
    bool start(double Resolution) {
        agent = agent_tag;
        if (!placeable(agent)) return false;
        agent_apl = id16(amap, (int)agent.y, (int)agent.x);
        agent_island = id16(imap, (int)agent.y, (int)agent.x);
        const CdrtGraph &g = G();
        trees.clear();
        samples.clear();
        std::vector<std::pair<P64, int>> tmp;
        for (size_t j = 0; j < g.cut.size(); j++) {
            const P64 &S = g.cut[j].S, &E = g.cut[j].E;
            double dlen = dist64(S, E);
            int dnum = (int)(dlen / Resolution + 4);
            if (20 < dlen) {
                double t = 8.0 / dlen, u = 1 - t;
                tmp.push_back({{t * S.x + u * E.x, t * S.y + u * E.y}, (int)j});
                tmp.push_back({{u * S.x + t * E.x, u * S.y + t * E.y}, (int)j});
            }
            for (int k = 1; k < dnum; k++) {
                if (2 * k == dnum) continue;
                double t = (double)k / dnum, t1 = 1.0 - t;
                tmp.push_back({{t1 * S.x + t * E.x, t1 * S.y + t * E.y}, (int)j});
            }
        }
        // std::random_shuffle (libstdc++): swap(v[i], v[rand() % (i + 1)]) for i = 1..n-1
        for (size_t i = 1; i < tmp.size(); i++) {
            size_t j = (size_t)(rnd.next() % (int32_t)(i + 1));
            if (i != j) std::swap(tmp[i], tmp[j]);
        }
        samples.assign(tmp.begin(), tmp.end());
        Mnode.clear();
        nodes.clear();
        Node &a = nodes[agent];
        a.invM = kOrigin, a.cost = {false, 0}, a.par = {-1, -1};
        for (size_t j = 0; j < g.cut.size(); j++) {
            Node &n = nodes[g.cut[j].core];
            n.cost = {true, kDoubleMax}, n.par = {-1, -1}, n.subs.clear(), n.invM = (int)j;
            Mnode[(int)j].insert(g.cut[j].core);
        }
        std::priority_queue<QE> pq;
        std::set<P64> exist;
        exist.insert(agent);
        std::vector<P64> near;
        neighbor(agent, near);
        for (const P64 &q : near) pq.push({Cost64{false, dist64(agent, q)}, agent, q});
        while (!pq.empty()) {
            QE e = pq.top();
            pq.pop();
            if (exist.count(e.E)) continue;
            Node &ne = nodes[e.E];
            if (ne.cost > e.c) {
                ne.par = e.S, ne.cost = e.c;
                nodes[e.S].subs.insert(e.E);
                exist.insert(e.E);
                neighbor(e.E, near);
                for (const P64 &q : near)
                    if (!exist.count(q)) pq.push({e.c + Cost64{false, dist64(e.E, q)}, e.E, q});
            }
        }
        return true;
    }

    // the agent hand-over inside RTRT_task, CDRT.cpp:946-996
    void move_agent() {
        const int tag_island = id16(imap, (int)agent_tag.y, (int)agent_tag.x);
        if (agent_island == tag_island) {
            Node &t = nodes[agent_tag];
            t.invM = kOrigin, t.cost = {false, 0}, t.par = {-1, -1};
            t.subs = nodes[agent].subs;
            nodes.erase(agent);
            for (const P64 &s : nodes[agent_tag].subs) nodes[s].par = agent_tag;
        } else {
            std::list<P64> path;
            get_path(agent_tag, path);
            if (path.size() < 2) { // reference: pop_front()/front() on an empty list
                error = 1;
                return;
            }
            Node &an = nodes[agent];
            path.pop_front();
            P64 x1 = path.front();
            Node &x1n = nodes[x1];
            for (const P64 &s : an.subs) {
                if (s == x1) continue;
                nodes[s].par = x1;
                x1n.subs.insert(s);
            }
            nodes.erase(agent);
            Node &t = nodes[agent_tag];
            t.invM = kOrigin, t.cost = {false, 0}, t.par = {-1, -1};
            auto it = path.begin(), old = it;
            for (++it; it != path.end(); ++it) {
                nodes[*it].subs.insert(*old);
                nodes[*old].subs.erase(*it);
                nodes[*old].par = *it;
                old = it;
            }
        }
        agent_island = tag_island;
        agent = agent_tag;
    }

    // RTRT_task, CDRT.cpp:927-1068
    void task() {
        if (!timers_init) { // function-local statics, CDRT.cpp:929-930
            updateTimeOld = clk.now();
            smallerTimeOld = clk.now();
            timers_init = true;
        }
        clk.now(); // Time0
        draw_dobs();
        crash_dobs_island();
        if (dist64(agent, agent_tag) > r_th || (double)(clk.now() - updateTimeOld) > 50e6) {
            updateTimeOld = clk.now();
            if (agent != agent_tag) move_agent();
            if (error) return;
            update_tree_cost();
        }
        for (const P64 &p : largerSet)
            if (nodes.count(p)) add_potential_links(p, true);
        largerSet.clear();
        if ((double)(clk.now() - smallerTimeOld) > 100e6) {
            smallerTimeOld = clk.now();
            for (const P64 &p : smallerSet)
                if (nodes.count(p)) add_potential_links(p, false);
            smallerSet.clear();
        }
        for (int s = 0; s < RewireN; s++) {
            P64 x;
            std::vector<P64> Xnear;
            if (!sampling(&x, Xnear)) continue;
            n_sampled++;
            Node &xn = nodes[x];
            for (const P64 &q : Xnear) {
                Node &qn = nodes[q];
                Cost64 h = xn.cost + Cost64{false, dist64(q, x)};
                if (qn.cost > h) {
                    if (goalIsSet) h.c += dist64(q, goal);
                    push(h, x, q);
                }
            }
        }
        const long long t0 = clk.now();
        while (!eQr.empty() && (double)(clk.now() - t0) < 20e6) {
            QE e = eQr.top();
            eQr.pop();
            n_pops++;
            if (!nodes.count(e.S) || !nodes.count(e.E)) continue;
            Node &sn = nodes[e.S];
            Node &en = nodes[e.E];
            Cost64 hat = sn.cost + Cost64{false, dist64(e.S, e.E)};
            if (en.cost > hat) {
                Cost64 tc = hat + Cost64{crash(e.S, sn.invM, e.E, en.invM), 0};
                if (en.cost > tc) {
                    en.cost = tc;
                    nodes[en.par].subs.erase(e.E);
                    nodes[e.S].subs.insert(e.E);
                    en.par = e.S;
                    n_rewired++;
                    add_potential_links(e.E, false);
                }
            }
        }
    }
};

} // namespace orc
