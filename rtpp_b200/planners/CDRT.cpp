// CDRT.cpp — BImap drop-in over the rtpp_cdrt_* engine; see CDRT.h.  Reference: CDRT2/lib/CDRT/CDRT.cpp.
#include "CDRT.h"

#include <cstdlib>
#include <cstring>

bool BImap::Fail(const char *what, int rc) {
    fprintf(stderr, "BImap::%s failed: %s (%s)\n", what, rtpp_status_string(rc), ctx_ ? rtpp_last_error(ctx_) : "no ctx");
    return false;
}

BImap::BImap(int device) {
    int rc = rtpp_create(device, &ctx_);
    if (rc) {
        fprintf(stderr, "BImap: no CUDA context (%s); there is no CPU fallback\n", rtpp_status_string(rc));
        ctx_ = nullptr;
    }
}
BImap::~BImap() {
    if (eng_) rtpp_cdrt_destroy(eng_);
    if (ctx_) rtpp_destroy(ctx_);
}

int BImap::SetGraph(int ncut, const double *cut_se, const double *cut_core, const int32_t *cp_off, const int32_t *cp, const int32_t *cc_off,
                    const int32_t *cc, int npoly, const int32_t *pv_off, const double *pv, const double *poly_core, const int32_t *pc_off,
                    const int32_t *pc, const int32_t *pp_off, const int32_t *pp) {
    BIgraphList.emplace_back();
    BIgraph &g = BIgraphList.back();
    for (int i = 0; i < ncut; i++) {
        BIcutline c{i, BIline({cut_se[4 * i], cut_se[4 * i + 1]}, {cut_se[4 * i + 2], cut_se[4 * i + 3]}), {cut_core[2 * i], cut_core[2 * i + 1]}, {}, {}};
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
        if (pp_off)
            for (int j = pp_off[i]; j < pp_off[i + 1]; j++) p.polygonlink.insert(pp[j]);
        g.freepolygonList.push_back(p);
    }
    g.cutlineBaseNum = ncut, g.freepolygonBaseNum = npoly;
    return (int)BIgraphList.size() - 1;
}

// CDRT.cpp:121-175: BIamap = region index, BIimap = free-polygon index, 0xFFFF elsewhere (vertices truncate to int)
void BImap::DrawBImap(bool) {
    BIamap.create(shapeY, shapeX, 0xFFFF);
    BIimap.create(shapeY, shapeX, 0xFFFF);
    for (size_t gi = 0; gi < BIgraphList.size(); gi++)
        for (size_t pi = 0; pi < BIgraphList[gi].freepolygonList.size(); pi++) {
            std::vector<std::pair<int, int>> pts;
            for (const BIpoint &q : BIgraphList[gi].freepolygonList[pi].polygon) pts.push_back({(int)q.x, (int)q.y});
            rtpp::fill_convex_polygon(BIamap, pts, (uint16_t)gi);
            rtpp::fill_convex_polygon(BIimap, pts, (uint16_t)pi);
        }
}

bool BImap::Placeable(const BIpoint &p, const char *who) { // CDRT.cpp:386-398
    if (p.x >= shapeX || p.y >= shapeY) {
        printf("%s error: Init point is too large!!\r\n", who);
        return false;
    }
    uint32_t a = BIamap.at((int)p.y, (int)p.x), i = BIimap.at((int)p.y, (int)p.x);
    if (a == 0xffff || (i & 0xC000)) {
        printf("%s error: robot can't be here!!\r\n", who);
        return false;
    }
    return true;
}
void BImap::RTRT_SetAgent(const BIpoint &point) {
    if (!Placeable(point, "SetAgent")) return;
    agent_tag = point;
    if (eng_) rtpp_cdrt_set_agent(eng_, 0, point.x, point.y);
}
void BImap::RTRT_SetGoal(const BIpoint &point) {
    if (!Placeable(point, "SetGoal")) return;
    goal = point, goalIsSet = true;
    if (eng_) rtpp_cdrt_set_goal(eng_, 0, point.x, point.y);
}
void BImap::RTRT_AddDynamicObstacle(int num, CTrtrt_DynamicObstacle &obs, double CheckLen) {
    DObstacles[num] = obs;
    check_len_[num] = CheckLen;
    sent_.erase(num);
}
void BImap::RTRT_DelDynamicObstacle(int num) {
    DObstacles.erase(num);
    check_len_.erase(num);
    if (eng_ && sent_.count(num)) rtpp_cdrt_del_obstacle(eng_, 0, num);
    sent_.erase(num);
}
// the demos mutate DObstacles[num].x/.y in place: ship whatever the map holds now
void BImap::PushObstacles() {
    for (auto &kv : DObstacles) {
        const CTrtrt_DynamicObstacle &o = kv.second;
        auto it = sent_.find(kv.first);
        if (it == sent_.end() || it->second != std::make_pair(o.DObsType, o.OutLine.size())) {
            std::vector<double> xy;
            for (const BIpoint &p : o.OutLine) xy.push_back(p.x), xy.push_back(p.y);
            rtpp_cdrt_add_obstacle(eng_, 0, kv.first, o.DObsType, o.x, o.y, o.angle, o.radius, xy.data(), (int)o.OutLine.size(), check_len_[kv.first]);
            sent_[kv.first] = {o.DObsType, o.OutLine.size()};
        } else
            rtpp_cdrt_move_obstacle(eng_, 0, kv.first, o.x, o.y, o.angle);
    }
    for (auto it = sent_.begin(); it != sent_.end();)
        if (!DObstacles.count(it->first)) {
            rtpp_cdrt_del_obstacle(eng_, 0, it->first);
            it = sent_.erase(it);
        } else
            ++it;
}
void BImap::RefillRand() { // un-injected runs draw from glibc rand() exactly as the reference does
    rand_.resize(rand_capacity_);
    for (int &v : rand_) v = rand();
}

// RTRT_start, CDRT.cpp:276-383
void BImap::RTRT_start(double Resolution) {
    if (!ctx_) return;
    if (!Placeable(agent_tag, "RTRT_task")) return;
    agent = agent_tag;
    agent_apl = BIamap.at((int)agent.y, (int)agent.x);
    agent_island = BIimap.at((int)agent.y, (int)agent.x);
    const BIgraph &g = BIgraphList[agent_apl];
    const int ncut = (int)g.cutlineList.size(), npoly = (int)g.freepolygonList.size();
    std::vector<double> cut_se, cut_core, pv;
    std::vector<int32_t> cp_off{0}, cp, cc_off{0}, cc, pv_off{0}, pc_off{0}, pc;
    for (const BIcutline &c : g.cutlineList) {
        cut_se.insert(cut_se.end(), {c.line.S.x, c.line.S.y, c.line.E.x, c.line.E.y});
        cut_core.insert(cut_core.end(), {c.core.x, c.core.y});
        cp.insert(cp.end(), c.polygonlink.begin(), c.polygonlink.end()), cp_off.push_back((int)cp.size());
        cc.insert(cc.end(), c.cutlinelink.begin(), c.cutlinelink.end()), cc_off.push_back((int)cc.size());
    }
    for (const BIfreepolygon &p : g.freepolygonList) {
        for (const BIpoint &q : p.polygon) pv.push_back(q.x), pv.push_back(q.y);
        pv_off.push_back((int)pv.size() / 2);
        pc.insert(pc.end(), p.cutlinelink.begin(), p.cutlinelink.end()), pc_off.push_back((int)pc.size());
    }
    std::vector<uint16_t> isl((size_t)shapeX * shapeY);
    for (size_t i = 0; i < isl.size(); i++) isl[i] = BIamap.data[i] == (uint16_t)agent_apl ? BIimap.data[i] : 0xFFFF;
    if (eng_) rtpp_cdrt_destroy(eng_), eng_ = nullptr;
    sent_.clear();
    if (!rand_injected_) RefillRand();
    int rc = rtpp_cdrt_create(ctx_, 1, node_capacity_, queue_capacity_, (int)shapeX, (int)shapeY, ncut, cut_se.data(), cut_core.data(), cp_off.data(),
                              cp.data(), cc_off.data(), cc.data(), npoly, pv_off.data(), pv.data(), pc_off.data(), pc.data(), isl.data(), Resolution,
                              r_th, p_th, nearK, RewireN, (int)std::max<size_t>(rand_.size(), 1), &eng_);
    if (rc) {
        Fail("RTRT_start", rc);
        eng_ = nullptr;
        return;
    }
    PushObstacles();
    rtpp_cdrt_set_agent(eng_, 0, agent_tag.x, agent_tag.y);
    if (goalIsSet) rtpp_cdrt_set_goal(eng_, 0, goal.x, goal.y);
    rc = rtpp_cdrt_start(eng_, 0, rand_.data(), (int)rand_.size());
    if (rc) Fail("RTRT_start", rc);
}

// RTRT_task, CDRT.cpp:927-1068 (+ the GetPath the demos call right after it)
void BImap::RTRT_task(void) {
    if (!eng_) return;
    PushObstacles();
    int rc = rtpp_cdrt_tick(eng_, dt_ns_, 1);
    if (rc) {
        Fail("RTRT_task", rc);
        return;
    }
    long long st[24];
    if (rtpp_cdrt_read_stats(eng_, st) == 0) {
        agent_island = (int32_t)st[2];
        if (st[11]) fprintf(stderr, "BImap::RTRT_task: engine status %lld\n", st[11]);
        if (!rand_injected_ && st[4] > rand_capacity_ / 2) { // keep the un-injected stream from running dry
            RefillRand();
            rtpp_cdrt_set_rand(eng_, 0, rand_.data(), (int)rand_.size());
        }
    }
    agent = agent_tag;
}

CTrtrt_Cost BImap::GetPath(void) {
    agent_path.clear();
    if (!goalIsSet || !eng_) return {true, doubleMax};
    long long st[24];
    if (rtpp_cdrt_read_stats(eng_, st)) return {true, doubleMax};
    const int n = (int)st[17];
    std::vector<double> xy(2 * (size_t)std::max(n, 1));
    int len = 0;
    rtpp_cdrt_read_path(eng_, 0, n, xy.data(), &len);
    for (int i = 0; i < n; i++) agent_path.push_back({xy[2 * i], xy[2 * i + 1]});
    double c;
    memcpy(&c, &st[19], 8);
    return {st[18] != 0, c};
}

int BImap::ReadStats(long long *out24) { return eng_ ? rtpp_cdrt_read_stats(eng_, out24) : RTPP_ERR_STATE; }
int BImap::ReadTree(double *xy, int32_t *parent, double *cost, uint8_t *col, int32_t *invM, uint8_t *alive, int max_nodes) {
    int n = 0;
    if (!eng_ || rtpp_cdrt_read_tree(eng_, 0, max_nodes, xy, parent, cost, col, invM, alive, &n)) return -1;
    return n;
}
void BImap::SyncNodeMap() {
    Node_Map.clear();
    int n = ReadTree(nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0);
    if (n <= 0) return;
    std::vector<double> xy(2 * (size_t)n), cost(n);
    std::vector<int32_t> par(n), inv(n);
    std::vector<uint8_t> col(n), alive(n);
    ReadTree(xy.data(), par.data(), cost.data(), col.data(), inv.data(), alive.data(), n);
    for (int i = 0; i < n; i++) {
        if (!alive[i]) continue;
        Node &nd = Node_Map[{xy[2 * i], xy[2 * i + 1]}];
        nd.cost = {col[i] != 0, cost[i]}, nd.invM = inv[i];
        nd.par = par[i] >= 0 ? BIpoint{xy[2 * par[i]], xy[2 * par[i] + 1]} : BIpoint{-1, -1};
    }
    for (int i = 0; i < n; i++)
        if (alive[i] && par[i] >= 0 && alive[par[i]]) Node_Map[{xy[2 * par[i]], xy[2 * par[i] + 1]}].subs.insert({xy[2 * i], xy[2 * i + 1]});
}
