// RTFMT.cpp — see RTFMT.h.
#include "RTFMT.h"

#include <cfloat>
#include <cstring>

RTFMTplanner::~RTFMTplanner() {
    if (eng_) rtpp_fmt_destroy(eng_);
}
void RTFMTplanner::OnStaticPlane() {
    if (eng_) rtpp_fmt_destroy(eng_), eng_ = nullptr;
    started_ = have_agent_ = false;
    nodes_.clear();
}
void RTFMTplanner::OnDynamicPlane() {
    if (eng_) rtpp_fmt_set_map(eng_, 0, 2, DObsmap.data.data(), shapeX, shapeY);
}

// reference RTFMT.cpp:11-36
void RTFMTplanner::SetAgent(RTpoint &point) {
    if (point.x >= shapeX || point.y >= shapeY) {
        printf("SetAgent error: Init point is too large!!\r\n");
        return;
    }
    if (Obsmap.at<uint8_t>((int)point.y, (int)point.x)) {
        printf("SetAgent error: robot can't be here!!\r\n");
        return;
    }
    if (!have_agent_) {
        nodes_.assign({point.x, point.y});
        root = point, agent = root, have_agent_ = true;
    } else {
        agent = point;
        if (started_) {
            const float a[2] = {agent.x, agent.y};
            rtpp_fmt_set_inputs(eng_, a, nullptr);
        }
    }
}
// reference RTFMT.cpp:38-55
void RTFMTplanner::SetGoal(RTpoint &point) {
    if (point.x >= shapeX || point.y >= shapeY) {
        printf("SetGoal error: Init point is too large!!\r\n");
        return;
    }
    if (Obsmap.at<uint8_t>((int)point.y, (int)point.x)) {
        printf("SetAgent error: robot can't be here!!\r\n");
        return;
    }
    if (goal != point) {
        goal = point;
        if (started_) {
            const float g[2] = {goal.x, goal.y};
            rtpp_fmt_set_inputs(eng_, nullptr, g);
        }
    }
}

bool RTFMTplanner::Restart() {
    if (!ctx_ || !have_agent_) return false;
    if (eng_) rtpp_fmt_destroy(eng_), eng_ = nullptr;
    const int n = NodeCount();
    int rc = rtpp_fmt_create(ctx_, 1, n + 32, 5, shapeX, shapeY, &eng_);
    if (rc) return Fail("RTFMTplanner", "rtpp_fmt_create", rc);
    rc = rtpp_fmt_set_map(eng_, 0, 1, Obsmap.data.data(), shapeX, shapeY);
    if (!rc) rc = rtpp_fmt_set_map(eng_, 0, 2, DObsmap.data.data(), shapeX, shapeY);
    const int32_t slot = 0, nn = n;
    std::vector<float> xy((size_t)(n + 32) * 2, 0.f);
    memcpy(xy.data(), nodes_.data(), nodes_.size() * 4);
    const float g[2] = {goal.x, goal.y};
    if (!rc) rc = rtpp_fmt_start(eng_, &slot, &nn, xy.data(), g);
    if (rc) return Fail("RTFMTplanner", "rtpp_fmt_start", rc);
    started_ = true;
    return true;
}

// reference RTFMT.cpp:393-422: SampleN free-space samples (static plane only) join the node set; the
// V_* sets restart from {root}.  (Calling Start() again adds another SampleN samples, as in the reference.)
void RTFMTplanner::Start(void) {
    if (!have_agent_) return;
    root = agent;
    for (int i = 0; i < sample_n_;) {
        const float rand_x = NextRand() / (float(RAND_MAX)) * shapeX;
        const float rand_y = NextRand() / (float(RAND_MAX)) * shapeY;
        if (Obsmap.at<uint8_t>((int)rand_y, (int)rand_x) == 0) {
            nodes_.push_back(rand_x), nodes_.push_back(rand_y);
            i++;
        }
        if (rand_stream_ && rand_pos_ >= rand_n_) break;
    }
    Restart();
}

void RTFMTplanner::Tick(int update, int iters, int mask, int path) {
    if (!started_) return;
    int rc = rtpp_fmt_tick(eng_, update, iters, mask, path);
    if (rc) Fail("RTFMTplanner", "rtpp_fmt_tick", rc);
}

// reference RTFMT.cpp:704-728
void RTFMTplanner::MainTask(void) {
    if (auto_draw_) DrawDObsmap();
    Tick(1, 500, 7, 0);
}

int RTFMTplanner::ReadTree(int32_t *parent, float *cost, uint8_t *col, uint8_t *state) {
    int n = 0;
    if (!started_ || rtpp_fmt_read_tree(eng_, 0, NodeCount(), parent, cost, col, state, &n)) return 0;
    return n;
}
RTpoint RTFMTplanner::Z() {
    long long st[16] = {0};
    if (started_) rtpp_fmt_read_stats(eng_, st);
    z = st[2] >= 0 && started_ ? RTpoint{nodes_[2 * st[2]], nodes_[2 * st[2] + 1]} : RTpoint{-1, -1};
    return z;
}

void RTFMTplanner::SyncSets() {
    Node_Map.clear(), V_unv.clear(), V_open.clear(), V_opennew.clear(), V_toopen.clear(), V_closed.clear();
    const int n = NodeCount();
    std::vector<int32_t> par(n);
    std::vector<float> cost(n);
    std::vector<uint8_t> col(n), st(n);
    if (ReadTree(par.data(), cost.data(), col.data(), st.data()) != n) return;
    auto pt = [&](int i) { return RTpoint{nodes_[2 * i], nodes_[2 * i + 1]}; };
    for (int i = 0; i < n; i++) {
        RTFMT_NodeAttribute &a = Node_Map[pt(i)];
        a.cost = {col[i] != 0, cost[i]};
        a.par = par[i] >= 0 ? pt(par[i]) : RTpoint{-1, -1};
        switch (st[i] & 3) {
        case 0: if (i) V_unv.insert(pt(i)); break;
        case 1: V_open.insert(pt(i)); break;
        case 2: V_opennew.insert(pt(i)); break;
        default: V_closed.insert(pt(i));
        }
        if (st[i] & 4) V_toopen.insert(pt(i));
    }
    for (int i = 0; i < n; i++)
        if (par[i] >= 0) Node_Map[pt(par[i])].subs.insert(pt(i));
}

// reference RTFMT.cpp:285-391, both branches on the device (the depth-<=5 fallback is csrc/path_fallback.cuh)
void RTFMTplanner::GetPath(std::list<RTpoint> &path) {
    path.clear();
    if (goal.x <= 0 || !started_) return;
    Tick(0, 0, 0, 1);
    long long st[16] = {0};
    if (rtpp_fmt_read_stats(eng_, st)) return;
    auto pt = [&](int i) { return RTpoint{nodes_[2 * i], nodes_[2 * i + 1]}; };
    // st[14] = 1: an entry with c_min < {0, 20000} exists and the path ends in the goal; 0: the fallback chain root..nearest node
    if (st[13] > 0) {
        std::vector<int32_t> ids(st[13]);
        int len = 0;
        rtpp_fmt_read_path(eng_, 0, (int)ids.size(), ids.data(), &len);
        for (int k = 0; k < len && k < (int)ids.size(); k++) path.push_back(ids[k] == -2 ? goal : pt(ids[k]));
    }
}
