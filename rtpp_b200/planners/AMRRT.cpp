// AMRRT.cpp — see AMRRT.h.
#include "AMRRT.h"

#include <cfloat>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

AMRRTplanner::~AMRRTplanner() {
    if (eng_) rtpp_engine_destroy(eng_);
}
void AMRRTplanner::OnStaticPlane() {
    if (eng_) rtpp_engine_destroy(eng_), eng_ = nullptr;
    started_ = false;
}
bool AMRRTplanner::OnDynamicShapes(const int32_t *circles, int n_circles, const int32_t *poly_xy, const int32_t *poly_sizes, int n_polys) {
    if (!EnsureEngine()) return true; // no engine yet: nothing to refresh
    const int rc = rtpp_engine_draw_dynamic(eng_, 0, circles, n_circles, poly_xy, poly_sizes, n_polys);
    if (rc) Fail("AMRRTplanner", "rtpp_engine_draw_dynamic", rc);
    return rc == 0;
}
void AMRRTplanner::OnDynamicPlane() {
    if (EnsureEngine()) rtpp_engine_set_map(eng_, 0, 2, DObsmap.data.data(), shapeX, shapeY);
}
bool AMRRTplanner::EnsureEngine() {
    if (eng_) return true;
    if (!ctx_ || shapeX <= 0) return false;
    int rc = rtpp_engine_create(ctx_, 2, 1, node_capacity_, 5, shapeX, shapeY, rand_capacity_, &eng_);
    if (rc) return Fail("AMRRTplanner", "rtpp_engine_create", rc);
    rc = rtpp_engine_set_map(eng_, 0, 1, Obsmap.data.data(), shapeX, shapeY);
    if (rc) return Fail("AMRRTplanner", "rtpp_engine_set_map", rc);
    UploadDiffusion();
    return true;
}
void AMRRTplanner::UploadDiffusion() {
    if (!eng_ || diffusion_matrix.empty()) return;
    const int dim = (int)diffusion_matrix[0].size();
    std::vector<int32_t> cells;
    std::vector<float> rows;
    // rows in table order; cells that map to each row
    std::vector<DIFFpoint> cell_of(diffusion_matrix.size(), DIFFpoint{-1, -1});
    for (const auto &kv : index_to_pos)
        if (kv.second >= 0 && kv.second < (int)cell_of.size()) cell_of[kv.second] = kv.first;
    for (size_t i = 0; i < diffusion_matrix.size(); i++) {
        if (cell_of[i].x < 0 || cell_of[i].y < 0) continue;
        cells.push_back(cell_of[i].x), cells.push_back(cell_of[i].y);
        for (int d = 0; d < dim; d++) rows.push_back(d < (int)diffusion_matrix[i].size() ? diffusion_matrix[i][d] : 0.f);
    }
    if (cells.empty()) return;
    int rc = rtpp_engine_set_diffusion(eng_, cells.data(), (int)cells.size() / 2, rows.data(), dim, diffusion_ratios);
    if (rc) Fail("AMRRTplanner", "rtpp_engine_set_diffusion", rc);
}

// reference AMRRT.cpp:58-93 with the tables handed over directly
void AMRRTplanner::InitializeDiffusion(const int32_t *cells_ij, int ncells, const float *matrix, int dim, float ratios) {
    diffusion_ratios = ratios;
    diffusion_matrix.assign(ncells, std::vector<float>());
    index_to_pos.clear();
    for (int i = 0; i < ncells; i++) {
        diffusion_matrix[i].assign(matrix + (size_t)i * dim, matrix + (size_t)(i + 1) * dim);
        index_to_pos[{cells_ij[2 * i], cells_ij[2 * i + 1]}] = i;
    }
    UploadDiffusion();
}

// the reference's JSON layout: DiffFile = [[f,...] x rows], IndexFile = [[i,j] x rows] (AMRRT.cpp:58-93)
static bool read_numbers(const char *path, std::vector<std::vector<double>> &rows) {
    std::ifstream f(path);
    if (!f.is_open()) return false;
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string s = ss.str();
    int depth = 0;
    std::vector<double> cur;
    for (size_t i = 0; i < s.size();) {
        const char c = s[i];
        if (c == '[') depth++, i++;
        else if (c == ']') {
            if (depth == 2) rows.push_back(cur), cur.clear();
            depth--, i++;
        } else if (c == '-' || c == '+' || (c >= '0' && c <= '9') || c == '.') {
            char *end = nullptr;
            const double v = strtod(s.c_str() + i, &end);
            if (depth == 2) cur.push_back(v);
            i = end - s.c_str();
        } else
            i++;
    }
    return true;
}
void AMRRTplanner::InitializeDiffusion(const char *DiffFile, const char *IndexFile, float ratios) {
    std::vector<std::vector<double>> diff, index;
    if (!read_numbers(DiffFile, diff) || !read_numbers(IndexFile, index)) {
        fprintf(stderr, "cannot open the JSON file.\n");
        return;
    }
    const int n = (int)std::min(diff.size(), index.size());
    if (n == 0) return;
    const int dim = (int)diff[0].size();
    std::vector<int32_t> cells;
    std::vector<float> m;
    for (int i = 0; i < n; i++) {
        cells.push_back((int32_t)index[i][0]), cells.push_back((int32_t)index[i][1]);
        for (int d = 0; d < dim; d++) m.push_back(d < (int)diff[i].size() ? (float)diff[i][d] : 0.f);
    }
    InitializeDiffusion(cells.data(), n, m.data(), dim, ratios);
}

// reference AMRRT.cpp:94-116 (host copy, for callers; the device has its own table)
float AMRRTplanner::DiffusivityMetric(RTpoint &x_s, RTpoint &x_e) {
    const int i_s = (int)(x_s.x * diffusion_ratios), j_s = (int)(x_s.y * diffusion_ratios);
    auto a = index_to_pos.find({i_s, j_s});
    if (a == index_to_pos.end()) return 3.4e38f;
    const int i_e = (int)(x_e.x * diffusion_ratios), j_e = (int)(x_e.y * diffusion_ratios);
    auto b = index_to_pos.find({i_e, j_e});
    if (b == index_to_pos.end()) return 3.4e38f;
    const std::vector<float> &ps = diffusion_matrix[a->second], &pe = diffusion_matrix[b->second];
    float len = 0;
    for (size_t i = 0; i < ps.size(); i++) {
        float dlen = ps[i] - pe[i];
        float sq = dlen * dlen;
        len = len + sq;
    }
    return sqrtf(len) / diffusion_ratios;
}

// reference AMRRT.cpp:13-38
void AMRRTplanner::SetAgent(RTpoint &point) {
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
            Fail("AMRRTplanner", "rtpp_engine_init_instances", rc);
            return;
        }
        started_ = true;
    } else
        agent = point;
}
// reference AMRRT.cpp:40-57
void AMRRTplanner::SetGoal(RTpoint &point) {
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

// reference AMRRT.cpp:156-200
void AMRRTplanner::UpdateTree(void) {
    if (!started_) return;
    const float a[2] = {agent.x, agent.y}, g[2] = {goal.x, goal.y};
    rtpp_engine_set_inputs(eng_, a, g, nullptr, nullptr);
    int rc = rtpp_engine_tick(eng_, 1, 0, dt_ns_, 0);
    long long st[16] = {0};
    if (!rc) rc = rtpp_engine_read_stats(eng_, st);
    if (rc) {
        Fail("AMRRTplanner", "UpdateTree", rc);
        return;
    }
    std::vector<float> xy((size_t)st[0] * 2);
    int n = 0;
    if (!rtpp_engine_read_tree(eng_, 0, (int)st[0], xy.data(), nullptr, nullptr, nullptr, &n) && st[1] < n)
        root = {xy[2 * st[1]], xy[2 * st[1] + 1]};
}

// reference AMRRT.cpp:590-617
void AMRRTplanner::MainTask(void) {
    if (!started_) return;
    if (auto_draw_) DrawDObsmap();
    UpdateTree();
    // the informed sampler's constants (AMRRT.cpp:418-432) are computed on the device from c_base, the new root and the goal
    std::vector<int32_t> r(rand_capacity_);
    int32_t n = rand_capacity_;
    if (rand_stream_) {
        n = (int32_t)std::min<long long>(rand_capacity_, rand_n_ - rand_pos_);
        memcpy(r.data(), rand_stream_ + rand_pos_, (size_t)n * 4);
    } else
        for (int i = 0; i < n; i++) r[i] = rand();
    const int32_t outpath = OutPath.size() != 0;
    int rc = rtpp_engine_set_am_inputs(eng_, r.data(), &n, nullptr, &outpath);
    if (!rc) rc = rtpp_engine_tick(eng_, 0, 80, dt_ns_, 0);
    long long st[16] = {0};
    if (!rc) rc = rtpp_engine_read_stats(eng_, st);
    if (rc) {
        Fail("AMRRTplanner", "MainTask", rc);
        return;
    }
    if (rand_stream_) rand_pos_ += st[14];
    if (st[9]) fprintf(stderr, "AMRRTplanner: engine status %lld\n", st[9]);
}

// reference AMRRT.cpp:287-308
bool AMRRTplanner::CrashDetectionStart(const RTpoint &ps, const RTpoint &pe, RTpoint &pstart) {
    pstart = ps;
    if (!ctx_) return true;
    const float e[4] = {ps.x, ps.y, pe.x, pe.y};
    uint32_t bits = 1;
    float out[2] = {ps.x, ps.y};
    int rc = rtpp_edges_first_hit(ctx_, e, 1, &bits, out);
    if (rc) Fail("AMRRTplanner", "rtpp_edges_first_hit", rc);
    pstart = {out[0], out[1]};
    return bits & 1u;
}

int AMRRTplanner::NodeCount() {
    long long st[16];
    if (!started_ || rtpp_engine_read_stats(eng_, st)) return 0;
    return (int)st[0];
}
int AMRRTplanner::ReadTree(float *xy, int32_t *parent, float *cost, uint8_t *col, int max_nodes) {
    int n = 0;
    if (!started_ || rtpp_engine_read_tree(eng_, 0, max_nodes, xy, parent, cost, col, &n)) return 0;
    return n;
}
void AMRRTplanner::SyncNodeMap() {
    Node_Map.clear();
    const int n = NodeCount();
    if (n <= 0) return;
    std::vector<float> xy((size_t)n * 2), cost(n);
    std::vector<int32_t> par(n);
    std::vector<uint8_t> col(n);
    ReadTree(xy.data(), par.data(), cost.data(), col.data(), n);
    auto pt = [&](int i) { return RTpoint{xy[2 * i], xy[2 * i + 1]}; };
    for (int i = 0; i < n; i++) {
        RTrrt_NodeAttribute &a = Node_Map[pt(i)];
        a.cost = {col[i] != 0, cost[i]};
        a.par = par[i] >= 0 ? pt(par[i]) : RTpoint{-1, -1};
    }
    for (int i = 0; i < n; i++)
        if (par[i] >= 0) Node_Map[pt(par[i])].subs.insert(pt(i));
}

// reference AMRRT.cpp:310-351
RTrrt_Cost AMRRTplanner::GetPath(std::list<RTpoint> &path) {
    c_base = -1;
    path.clear();
    if (goal.x <= 0 || !started_) return {1, 3.4e38f};
    const float a[2] = {agent.x, agent.y}, g[2] = {goal.x, goal.y};
    rtpp_engine_set_inputs(eng_, a, g, nullptr, nullptr);
    int rc = rtpp_engine_tick(eng_, 0, 0, dt_ns_, 1);
    long long st[16] = {0};
    if (!rc) rc = rtpp_engine_read_stats(eng_, st);
    if (rc) {
        Fail("AMRRTplanner", "GetPath", rc);
        return {1, 3.4e38f};
    }
    float pcost, cb;
    const uint32_t pc = (uint32_t)st[12], cbb = (uint32_t)st[13];
    memcpy(&pcost, &pc, 4), memcpy(&cb, &cbb, 4);
    c_base = cb;
    if (st[11] == 0 && st[10] > 0) {
        std::vector<int32_t> ids(st[10]);
        int len = 0;
        rtpp_engine_read_path(eng_, 0, (int)ids.size(), ids.data(), &len);
        const int n = NodeCount();
        std::vector<float> xy((size_t)n * 2);
        ReadTree(xy.data(), nullptr, nullptr, nullptr, n);
        for (int id : ids) path.push_back(id == -2 ? goal : RTpoint{xy[2 * id], xy[2 * id + 1]});
        return {0, pcost};
    }
    return {1, pcost == 2.4e38f ? 2.4e38f : 3.4e38f};
}
