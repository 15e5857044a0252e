// rtpp_oracle.hpp — CPU ORACLE (test infrastructure, NOT product code).
//
// A from-scratch CPU restatement of the hot path shared by the four planners of the reference
// project TZY-H/RTPP.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may link or load this; the product path (rtpp_b200/) never does.
//
// Pinning status (see DESIGN.md "Oracle"): the in-tree arithmetic (edge checks, cost algebra,
// planner loops) is PINNED against the reference's own sources compiled in place with stand-in
// OpenCV/CGAL headers (oracle/_ref, built by oracle/Makefile; golden vectors generated from it are
// committed under tests/golden/).  The k-NN semantics (CGAL 5.6, vendored in the reference but not
// compilable here: needs Boost) and cv::circle / cv::fillPoly (OpenCV, not vendored) are restated
// from their published behaviour: parity UNPINNED at those two library boundaries.
//
// Build: g++ -O3 -std=c++17 -ffp-contract=off (never -ffast-math): the author's build is
// `g++ -O3 -DNDEBUG -frounding-math -std=gnu++14` without -march, i.e. no FMA contraction.
//
// Each function cites the reference file:line it follows (paths relative to the reference root).
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <deque>
#include <list>
#include <map>
#include <queue>
#include <set>
#include <vector>

namespace orc {

// ------------------------------------------------------------------------------------------------
// Points, distance, cost algebra
// ------------------------------------------------------------------------------------------------

// RTpoint — RTRRT/lib/RTrrt/RTRRT.h:28-53 (identical in RTFMT.h:28-53, AMRRT.h:28-53).
struct Pt {
    float x, y;
    bool operator!=(const Pt &o) const { return x != o.x || y != o.y; }
    bool operator==(const Pt &o) const { return !(*this != o); }
    bool operator<(const Pt &o) const { return (x != o.x) ? (x < o.x) : (y < o.y); }
};

// RTpoint::operator% — RTRRT.h:47-52: sqrtf32(dx*dx + dy*dy), every op a separate binary32 rounding.
inline float dist(const Pt &a, const Pt &b) {
    float dx = a.x - b.x;
    float dy = a.y - b.y;
    float xx = dx * dx;
    float yy = dy * dy;
    return sqrtf(xx + yy);
}

// BIpoint::operator% — CDRT2/lib/CDRT/CDRT.h:73-78 (binary64).
inline double dist64(double ax, double ay, double bx, double by) {
    double dx = ax - bx, dy = ay - by;
    return std::sqrt(dx * dx + dy * dy);
}

// RTrrt_Cost — RTRRT.h:55-91 (= RTFMT_Cost RTFMT.h:55-91).  Lexicographic (collision, cost).
struct Cost {
    bool col;
    float c;
    Cost operator+(const Cost &o) const { return {col || o.col, c + o.c}; }
    bool operator==(const Cost &o) const { return col == o.col && c == o.c; }
    bool operator<(const Cost &o) const { return (col == o.col) ? (c < o.c) : (col < o.col); }
    bool operator>(const Cost &o) const { return (col == o.col) ? (c > o.c) : (col > o.col); }
};

// ------------------------------------------------------------------------------------------------
// Maps
// ------------------------------------------------------------------------------------------------

// Two row-major uint8 planes, as the reference's cv::Mat Obsmap / DObsmap (RTRRT.h:138-139).
// Non-zero = obstacle.  Either pointer may be null (plane absent = all free).
struct Map {
    const uint8_t *S = nullptr; // static  (Obsmap: 0/255, RTRRT.cpp:69-76)
    const uint8_t *D = nullptr; // dynamic (DObsmap: 0/127, RTRRT.cpp:140-168)
    int W = 0, H = 0;           // shapeX = cols, shapeY = rows (RTRRT.cpp:78-79)
};

enum { PLANE_S = 1, PLANE_D = 2 };

// cv::Mat::at<uint8_t>(row, col) on a continuous Mat is base[row*cols + col] with no bounds check
// in release builds.  In-buffer "overflows" (col >= cols wrapping into the next row) are therefore
// reproduced; reads outside the buffer are undefined in the reference and DEFINED here as blocked.
inline bool probe(const uint8_t *P, int W, int H, int row, int col) {
    if (!P) return false;
    long long lin = (long long)row * W + col;
    if (lin < 0 || lin >= (long long)W * H) return true;
    return P[lin] != 0;
}
inline bool probe_mask(const Map &m, int mask, int row, int col) {
    bool hit = false;
    if (mask & PLANE_S) hit = hit || probe(m.S, m.W, m.H, row, col);
    if (mask & PLANE_D) hit = hit || probe(m.D, m.W, m.H, row, col);
    return hit;
}

// ------------------------------------------------------------------------------------------------
// Edge checks
// ------------------------------------------------------------------------------------------------

// Number of pixel probes the dyadic march makes when nothing is hit: 2 endpoints + (2^m - 1)
// interior, 2^m = largest power of two <= M (SURVEY §8(d) formula).
inline int march_M(float len, int divisor) {
    float q = len / (float)divisor; // "len / 5": int promotes to float
    float s = q + 2.0f;             // "+ 2"
    return (int32_t)s;              // assignment to int32_t truncates
}
inline int nprobes_for_len(float len, int divisor) {
    int M = march_M(len, divisor);
    int interior = 0;
    for (float den = 2; den <= (float)M; den *= 2) interior = (int)den - 1;
    return 2 + interior;
}

// E1 RTRRTplanner::CrashDetection RTRRT.cpp:202-222 (= RTFMT.cpp:242-262, divisor 5),
// E2 RTFMTplanner::CrashDetectionNDObs RTFMT.cpp:264-283 (mask = PLANE_S),
// E3 AMRRTplanner::CrashDetection AMRRT.cpp:266-286 (divisor 2).
// Probe positions are the dyadic fractions i/2^j of the segment, NOT a DDA; thin walls can be
// stepped over — reproduced, not fixed.
inline bool edge_blocked(const Map &m, Pt ps, Pt pe, int divisor, int mask) {
    // endpoint tests: at<uint8_t>(ps.y, ps.x) converts float -> int by truncation
    if (probe_mask(m, mask, (int)ps.y, (int)ps.x) || probe_mask(m, mask, (int)pe.y, (int)pe.x)) return true;
    float len = dist(ps, pe);
    int32_t M = march_M(len, divisor);
    for (float den = 2; den <= (float)M; den *= 2) {
        for (int32_t i = 1; (float)i < den; i += 2) {
            float t = (float)i / den;
            float omt = 1.0f - t;
            float ax = ps.x * t, bx = omt * pe.x;
            float ay = ps.y * t, by = omt * pe.y;
            int32_t x = (int32_t)(ax + bx);
            int32_t y = (int32_t)(ay + by);
            if (probe_mask(m, mask, y, x)) return true;
        }
    }
    return false;
}

// E4 AMRRTplanner::CrashDetectionStart AMRRT.cpp:287-308: linear march i=1..M-1, t weights pe;
// returns the last free float point in *pstart.
inline bool edge_first_hit(const Map &m, Pt ps, Pt pe, Pt *pstart) {
    *pstart = ps;
    const int mask = PLANE_S | PLANE_D;
    if (probe_mask(m, mask, (int)ps.y, (int)ps.x) || probe_mask(m, mask, (int)pe.y, (int)pe.x)) return true;
    float len = dist(ps, pe);
    int32_t M = march_M(len, 2);
    for (int32_t i = 1; i < M; i++) {
        float t = (float)i / (float)M;
        float omt = 1.0f - t;
        float ax = pe.x * t, bx = omt * ps.x;
        float ay = pe.y * t, by = omt * ps.y;
        float x = ax + bx;
        float y = ay + by;
        if (probe_mask(m, mask, (int32_t)y, (int32_t)x)) return true;
        *pstart = {x, y};
    }
    *pstart = pe;
    return false;
}

// E5b BImap::RTRT_CrashDetection, the ray-march part only, CDRT.cpp:584-601: binary64 points,
// binary32 t, dynamic plane only, divisor 10.
inline bool edge_blocked_f64(const Map &m, double psx, double psy, double pex, double pey, int divisor, int mask) {
    if (probe_mask(m, mask, (int)psy, (int)psx) || probe_mask(m, mask, (int)pey, (int)pex)) return true;
    float len = (float)dist64(psx, psy, pex, pey);
    int32_t M = march_M(len, divisor);
    for (float den = 2; den <= (float)M; den *= 2) {
        for (int32_t i = 1; (float)i < den; i += 2) {
            float t = (float)i / den;
            float omt = 1.0f - t;
            double ax = psx * (double)t, bx = (double)omt * pex;
            double ay = psy * (double)t, by = (double)omt * pey;
            int32_t x = (int32_t)(ax + bx);
            int32_t y = (int32_t)(ay + by);
            if (probe_mask(m, mask, y, x)) return true;
        }
    }
    return false;
}

// ------------------------------------------------------------------------------------------------
// k-nearest neighbours (N1/N2): CGAL::Orthogonal_k_neighbor_search semantics
// ------------------------------------------------------------------------------------------------
// CGAL/Orthogonal_k_neighbor_search.h:51-79,121-150 + CGAL/Euclidean_distance.h:72-83:
// squared distance (qx-px)^2 + (qy-py)^2 in binary64 on float coordinates widened exactly;
// result = min(k, N) points of smallest d2, ascending.  CGAL's order among equal d2 is an artefact
// of kd-tree layout + std::sort (unspecified); the oracle rule is (d2, insertion index) ascending.
struct Nbr {
    double d2;
    int id;
    bool operator<(const Nbr &o) const { return (d2 != o.d2) ? (d2 < o.d2) : (id < o.id); }
};

inline double sqdist64(double qx, double qy, double px, double py) {
    double dx = qx - px, dy = qy - py;
    return dx * dx + dy * dy;
}

// Brute force: the definition.
inline void knn_brute(const double *pxy, int n, double qx, double qy, int k, std::vector<Nbr> &out) {
    out.clear();
    out.reserve(n);
    for (int i = 0; i < n; i++) out.push_back({sqdist64(qx, qy, pxy[2 * i], pxy[2 * i + 1]), i});
    int kk = std::min(k, n);
    std::partial_sort(out.begin(), out.begin() + kk, out.end());
    out.resize(kk);
}

// Incremental exact index (uniform grid + ring expansion) — only an accelerator for the oracle's
// planner loops; tests check it against knn_brute.  Mirrors Kd_tree::insert (CGAL/Kd_tree.h:461-466)
// as "append; every later query sees all points inserted so far".
class KnnIndex {
  public:
    void init(double w, double h, double cell) {
        cell_ = cell;
        gw_ = std::max(1, (int)std::ceil(w / cell));
        gh_ = std::max(1, (int)std::ceil(h / cell));
        cells_.assign((size_t)gw_ * gh_, {});
        xy_.clear();
    }
    int size() const { return (int)(xy_.size() / 2); }
    int insert(double x, double y) {
        int id = size();
        xy_.push_back(x);
        xy_.push_back(y);
        cells_[(size_t)cy(y) * gw_ + cx(x)].push_back(id);
        kd_dirty_ = true;
        return id;
    }
    double px(int i) const { return xy_[2 * i]; }
    double py(int i) const { return xy_[2 * i + 1]; }
    const double *data() const { return xy_.data(); }

    // kd_rebuild(true): answer queries from a kd-tree (sliding-midpoint splits, bucket size 10: CGAL/Splitters.h:35,159-212)
    // that is REBUILT on the first query after every insert — what CGAL's Kd_tree does (insert -> invalidate_build,
    // Kd_tree.h:433-466; const_build on the next search, :418-425).  Same answers as the grid and as knn_brute (the answer is
    // defined as the k smallest (d2, id)); only the cost differs.  Used for the "reference-shaped" CPU baseline.
    void kd_rebuild(bool on) { kd_on_ = on, kd_dirty_ = true; }
    long long kd_builds() const { return kd_builds_; }

    void query(double qx, double qy, int k, std::vector<Nbr> &out) const {
        out.clear();
        int n = size();
        int kk = std::min(k, n);
        if (kk <= 0) return;
        if (kd_on_) {
            if (kd_dirty_) kd_build();
            std::vector<Nbr> &heap = out;
            kd_search(0, qx, qy, kk, heap);
            std::sort_heap(heap.begin(), heap.end());
            return;
        }
        int c0x = cx(qx), c0y = cy(qy);
        std::vector<Nbr> &heap = out; // max-heap on (d2,id)
        int maxr = std::max(gw_, gh_);
        for (int r = 0; r <= maxr; r++) {
            int x0 = c0x - r, x1 = c0x + r, y0 = c0y - r, y1 = c0y + r;
            for (int yy = std::max(y0, 0); yy <= std::min(y1, gh_ - 1); yy++) {
                bool edge_row = (yy == y0 || yy == y1);
                int step = edge_row ? 1 : std::max(1, x1 - x0);
                for (int xx = x0; xx <= x1; xx += step) {
                    if (xx < 0 || xx >= gw_) continue;
                    for (int id : cells_[(size_t)yy * gw_ + xx]) {
                        Nbr nb{sqdist64(qx, qy, xy_[2 * id], xy_[2 * id + 1]), id};
                        if ((int)heap.size() < kk) {
                            heap.push_back(nb);
                            std::push_heap(heap.begin(), heap.end());
                        } else if (nb < heap.front()) {
                            std::pop_heap(heap.begin(), heap.end());
                            heap.back() = nb;
                            std::push_heap(heap.begin(), heap.end());
                        }
                    }
                }
            }
            if ((int)heap.size() == kk) {
                // every unscanned point lies outside the scanned square (clipped to the grid);
                // its distance to q is >= the gap to the nearest still-open side.
                double gap = HUGE_VAL;
                if (x0 > 0) gap = std::min(gap, qx - x0 * cell_);
                if (x1 < gw_ - 1) gap = std::min(gap, (x1 + 1) * cell_ - qx);
                if (y0 > 0) gap = std::min(gap, qy - y0 * cell_);
                if (y1 < gh_ - 1) gap = std::min(gap, (y1 + 1) * cell_ - qy);
                if (gap == HUGE_VAL) break;
                if (gap > 0 && heap.front().d2 < gap * gap) break;
            }
        }
        std::sort_heap(heap.begin(), heap.end());
    }

  private:
    struct KdNode {
        int lo, hi;       // leaf: ids kd_ids_[lo, hi)
        int dim;          // -1 = leaf
        double split;
        int left, right;
    };
    void kd_build() const {
        kd_ids_.resize((size_t)size());
        for (int i = 0; i < size(); i++) kd_ids_[i] = i;
        kd_nodes_.clear();
        kd_nodes_.reserve((size_t)size() / 4 + 8);
        kd_split(0, size());
        kd_dirty_ = false, kd_builds_++;
    }
    int kd_split(int lo, int hi) const {
        const int me = (int)kd_nodes_.size();
        kd_nodes_.push_back({lo, hi, -1, 0.0, -1, -1});
        if (hi - lo <= 10) return me;
        double mn[2] = {HUGE_VAL, HUGE_VAL}, mx[2] = {-HUGE_VAL, -HUGE_VAL};
        for (int i = lo; i < hi; i++)
            for (int d = 0; d < 2; d++) {
                const double v = xy_[2 * kd_ids_[i] + d];
                mn[d] = std::min(mn[d], v), mx[d] = std::max(mx[d], v);
            }
        const int dim = (mx[1] - mn[1] > mx[0] - mn[0]) ? 1 : 0;
        if (mx[dim] == mn[dim]) return me; // all points coincide: a (large) leaf
        // midpoint of the tight box; both sides are non-empty because min < mid <= max, so no sliding is needed
        const double mid = mn[dim] + (mx[dim] - mn[dim]) * 0.5;
        const double cut = mid > mn[dim] ? mid : mx[dim];
        int *b = kd_ids_.data() + lo, *e = kd_ids_.data() + hi;
        int *m = std::partition(b, e, [&](int id) { return xy_[2 * id + dim] < cut; });
        const int midx = lo + (int)(m - b);
        kd_nodes_[me].dim = dim, kd_nodes_[me].split = cut;
        const int l = kd_split(lo, midx);
        const int r = kd_split(midx, hi);
        kd_nodes_[me].left = l, kd_nodes_[me].right = r;
        return me;
    }
    void kd_search(int node, double qx, double qy, int kk, std::vector<Nbr> &heap) const {
        const KdNode &nd = kd_nodes_[node];
        if (nd.dim < 0) {
            for (int i = nd.lo; i < nd.hi; i++) {
                const int id = kd_ids_[i];
                Nbr nb{sqdist64(qx, qy, xy_[2 * id], xy_[2 * id + 1]), id};
                if ((int)heap.size() < kk) {
                    heap.push_back(nb);
                    std::push_heap(heap.begin(), heap.end());
                } else if (nb < heap.front()) {
                    std::pop_heap(heap.begin(), heap.end());
                    heap.back() = nb;
                    std::push_heap(heap.begin(), heap.end());
                }
            }
            return;
        }
        const double q = nd.dim ? qy : qx, diff = q - nd.split;
        const int near = diff < 0 ? nd.left : nd.right, far = diff < 0 ? nd.right : nd.left;
        kd_search(near, qx, qy, kk, heap);
        // a point beyond the plane is at least |diff| away in this coordinate; rounding is monotone, so its computed d2 is
        // >= diff*diff as computed here: pruning on strict > never drops a candidate that ties with the k-th
        if ((int)heap.size() < kk || !(diff * diff > heap.front().d2)) kd_search(far, qx, qy, kk, heap);
    }
    bool kd_on_ = false;
    mutable bool kd_dirty_ = true;
    mutable long long kd_builds_ = 0;
    mutable std::vector<int> kd_ids_;
    mutable std::vector<KdNode> kd_nodes_;

    int cx(double x) const { return std::min(gw_ - 1, std::max(0, (int)std::floor(x / cell_))); }
    int cy(double y) const { return std::min(gh_ - 1, std::max(0, (int)std::floor(y / cell_))); }
    double cell_ = 16;
    int gw_ = 1, gh_ = 1;
    std::vector<std::vector<int>> cells_;
    std::vector<double> xy_;
};

// ------------------------------------------------------------------------------------------------
// Deterministic stand-ins for the reference's two sources of non-determinism
// ------------------------------------------------------------------------------------------------

// Virtual clock.  The reference bounds its loops by wall-clock reads (utime_ns(), RTRRT.cpp:3-9,
// 228-230,300,329).  For parity every implementation (oracle, oracle/_ref with a redirected
// clock_gettime, the CUDA engines) uses the same rule: the k-th clock read returns k*dt_ns.
struct VClock {
    long long n = 0;
    long long dt_ns = 1000000;
    long long now() { return dt_ns * (n++); }
};

// Injected rand() stream (the reference calls un-seeded glibc rand(), RTRRT.cpp:234).  Values are
// raw rand() results in [0, RAND_MAX]; u = rand() / float(RAND_MAX) is computed in binary32 exactly
// as the reference does (float(RAND_MAX) == 2^31).
struct RandStream {
    const int32_t *r = nullptr;
    long long n = 0, pos = 0;
    bool exhausted = false;
    int32_t next() {
        if (pos >= n) {
            exhausted = true;
            return 0;
        }
        return r[pos++];
    }
    float u() { return (float)next() / 2147483648.0f; }
};

// Counter-based generator used by tests/bench to fill rand streams identically everywhere
// (splitmix64 finaliser; clamped so that float(r) < 2^31, i.e. u < 1: SURVEY §0 "u can be 1.0").
inline int32_t rand31(uint64_t seed, uint64_t i) {
    uint64_t z = seed * 0x9E3779B97F4A7C15ull + (i + 1) * 0xBF58476D1CE4E5B9ull;
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    uint32_t r = (uint32_t)(z >> 33);
    return (int32_t)std::min<uint32_t>(r, 2147483520u);
}

// ------------------------------------------------------------------------------------------------
// RT-RRT*  (RTRRT/lib/RTrrt/RTRRT.cpp)
// ------------------------------------------------------------------------------------------------
struct RTRRT {
    static constexpr int KnearN = 25;      // RTRRT.cpp:224
    static constexpr float RnearR = 8.0f;  // RTRRT.cpp:225
    int divisor = 5;                       // RTRRT.cpp:209

    struct Node {
        Cost cost;
        Pt par;
        std::set<Pt> subs;
        int id;
    };
    Map map;
    std::map<Pt, Node> nodes; // Node_Map, RTRRT.h:130
    KnnIndex index;           // NeighborSearchTree, RTRRT.h:129
    std::vector<Pt> by_id;
    std::queue<Pt> Qr, Qs;
    std::set<Pt> Qs_set;
    Pt root{-1, -1}, goal{-1, -1}, agent{-1, -1};
    float c_base = -1;
    VClock clk;
    // counters (for iterations/s)
    long long n_iters = 0, n_pops_r = 0, n_pops_s = 0, n_edge_checks = 0, n_knn = 0;

    void init(const Map &m, double cell = 32) {
        map = m;
        index.init(m.W, m.H, cell);
    }
    bool crash(const Pt &a, const Pt &b) {
        n_edge_checks++;
        return edge_blocked(map, a, b, divisor, PLANE_S | PLANE_D);
    }
    void knn(const Pt &q, int k, std::vector<Nbr> &out) {
        n_knn++;
        index.query((double)q.x, (double)q.y, k, out);
    }
    Pt pt_of(const Nbr &n) const { return by_id[n.id]; }
    int add_node(const Pt &p) {
        int id = index.insert((double)p.x, (double)p.y);
        by_id.push_back(p);
        return id;
    }

    // RTRRT.cpp:11-36
    bool SetAgent(const Pt &p) {
        if (p.x >= map.W || p.y >= map.H) return false;
        if (probe(map.S, map.W, map.H, (int)p.y, (int)p.x)) return false;
        if (nodes.empty()) {
            Node &n = nodes[p];
            n.cost = {false, 0};
            n.par = {-1, -1};
            root = p;
            n.id = add_node(p);
            agent = root;
        } else
            agent = p;
        return true;
    }
    // RTRRT.cpp:38-55
    bool SetGoal(const Pt &p) {
        if (p.x >= map.W || p.y >= map.H) return false;
        if (probe(map.S, map.W, map.H, (int)p.y, (int)p.x)) return false;
        if (goal != p) goal = p;
        return true;
    }

    // RTRRT.cpp:92-136
    void UpdateTree() {
        if (agent != root) {
            std::vector<Nbr> nb;
            knn(agent, 1, nb);
            Pt x_near = pt_of(nb[0]);
            if (x_near != root) {
                Pt x_temp = x_near, x_old{-1, -1};
                while (x_temp != root) {
                    Node &nt = nodes[x_temp];
                    Pt x_par = nt.par;
                    nt.par = x_old;
                    nt.subs.insert(x_par);
                    nt.subs.erase(x_old);
                    x_old = x_temp;
                    x_temp = x_par;
                }
                nodes[x_temp].par = x_old;
                nodes[x_temp].subs.erase(x_old);
                root = x_near;
                nodes[root].cost = {false, 0};
            }
        }
        std::queue<Pt> Qup;
        Qup.push(root);
        while (!Qup.empty()) {
            Pt now = Qup.front();
            Qup.pop();
            Node &nn = nodes[now];
            for (const Pt &sub : nn.subs) {
                Node &ns = nodes[sub];
                ns.cost = nn.cost + Cost{crash(now, sub), dist(now, sub)};
                Qup.push(sub);
            }
        }
    }

    // Sampler, RTRRT.cpp:234-262.  Host-only transcendental code (glibc sinf/cosf/atan2).
    Pt sample(RandStream &rs) {
        float p0 = rs.u();
        float rx, ry;
        // Guard: when c_base < |goal - root| (e.g. the agent moved since the last
        // GetPath) the reference's ellipse has b = sqrt(negative) = NaN and its rejection loop never
        // terminates.  Defined behaviour here and in the product's host sampler: sample uniformly.
        const bool degenerate = c_base >= 0 && c_base < dist(goal, root);
        if (p0 < 0.1f || c_base < 0 || degenerate) {
            rx = rs.u() * map.W;
            ry = rs.u() * map.H;
        } else {
            float c_min = dist(goal, root);
            float a = c_base / 2.0f;
            // unqualified sqrt()/atan2() on float arguments: the planner header pulls in <math.h>
            // (RTRRT.h:17 -> CGAL/Exact_predicates_inexact_constructions_kernel.h:24 ->
            // CGAL/double.h:26), whose libstdc++ wrapper makes the float overloads visible at global
            // scope, so these are sqrtf / atan2f (confirmed by the oracle/_ref build).
            float b = sqrtf(c_base * c_base - c_min * c_min) / 2.0f;
            float angle = atan2f(goal.y - root.y, goal.x - root.x);
            int guard = 0;
            while (true) {
                float th = rs.u() * 2 * 3.141592954f;
                float rr = sqrtf(rs.u());
                float x = a * cosf(th) * rr;
                float y = b * sinf(th) * rr;
                rx = cosf(angle) * x - sinf(angle) * y;
                ry = sinf(angle) * x + cosf(angle) * y;
                rx = rx + (goal.x + root.x) / 2;
                ry = ry + (goal.y + root.y) / 2;
                if (rx >= 0 && rx < map.W && ry >= 0 && ry < map.H) break;
                if (rs.exhausted || ++guard > 100000) break; // NaN ellipse loops forever in the reference
            }
        }
        return {rx, ry};
    }

    // One accepted-or-rejected sample: RTRRT.cpp:263-298
    void expand_one(const Pt &rp) {
        n_iters++;
        std::vector<Nbr> nb;
        knn(rp, KnearN, nb);
        Pt closest = pt_of(nb[0]);
        if (!crash(rp, closest)) {
            if (sqrtf((float)nb[0].d2) > RnearR) {
                Pt x_min = closest;
                Cost c_min = nodes[closest].cost + Cost{false, dist(rp, closest)};
                for (const Nbr &n : nb) {
                    Pt xn = pt_of(n);
                    Cost c_new = nodes[xn].cost + Cost{crash(rp, xn), dist(rp, xn)};
                    if (c_new < c_min) {
                        c_min = c_new;
                        x_min = xn;
                    }
                }
                Node &nn = nodes[rp];
                nn.par = x_min;
                nn.cost = c_min;
                nodes[x_min].subs.insert(rp);
                Qr.push(rp);
                nn.id = add_node(rp);
            } else
                Qr.push(closest);
        }
    }
    // RTRRT.cpp:302-319
    void rewire_pop_r() {
        n_pops_r++;
        Pt xr = Qr.front();
        Qr.pop();
        std::vector<Nbr> nb;
        knn(xr, KnearN, nb);
        for (const Nbr &n : nb) {
            Pt xn = pt_of(n);
            Cost c_old = nodes[xn].cost;
            Cost c_new = nodes[xr].cost + Cost{crash(xr, xn), dist(xr, xn)};
            if (c_new < c_old) {
                nodes[nodes[xn].par].subs.erase(xn);
                nodes[xr].subs.insert(xn);
                nodes[xn].par = xr;
                nodes[xn].cost = c_new;
                Qr.push(xn);
            }
        }
    }
    // RTRRT.cpp:331-352
    void rewire_pop_s() {
        n_pops_s++;
        Pt xs = Qs.front();
        Qs.pop();
        std::vector<Nbr> nb;
        knn(xs, KnearN, nb);
        for (const Nbr &n : nb) {
            Pt xn = pt_of(n);
            Cost c_old = nodes[xn].cost;
            Cost c_new = nodes[xs].cost + Cost{crash(xs, xn), dist(xs, xn)};
            if (c_new < c_old) {
                nodes[nodes[xn].par].subs.erase(xn);
                nodes[xs].subs.insert(xn);
                nodes[xn].par = xs;
                nodes[xn].cost = c_new;
            }
            if (Qs_set.count(xn) == 0) {
                Qs.push(xn);
                Qs_set.insert(xn);
            }
        }
    }
    // RTRRT.cpp:226-355 under the virtual clock.  `samples` supplies one point per outer iteration
    // (pre-drawn by sample(): the sampler depends only on c_base/root/goal/shape, which are fixed
    // for the duration of the call).  Returns the number of samples consumed.
    int ExpandRewiring(long run_ms, const Pt *samples, int n_samples) {
        long long t0 = clk.now();
        long long run = run_ms * 1000000LL;
        int used = 0;
        while (clk.now() - t0 < run) {
            if (used >= n_samples) break;
            expand_one(samples[used++]);
            while (Qr.size() && (clk.now() - t0 < run)) rewire_pop_r();
            if (Qs.size() == 0) {
                Qs_set.clear();
                Qs.push(root);
                Qs_set.insert(root);
            }
            while (Qs.size() && (clk.now() - t0 < run)) rewire_pop_s();
        }
        return used;
    }

    // RTRRT.cpp:356-478
    Cost GetPath(std::list<Pt> &path) {
        c_base = -1;
        path.clear();
        if (goal.x <= 0) return {true, 3.4e38f};
        if (!crash(goal, root)) {
            path.push_back(root);
            path.push_back(goal);
            return {false, dist(root, goal)};
        }
        std::vector<Nbr> nb;
        knn(goal, KnearN, nb);
        Pt x_min{-1, -1};
        Cost c_min{true, FLT_MAX};
        for (const Nbr &n : nb) {
            Pt xn = pt_of(n);
            Cost c_new = nodes[xn].cost + Cost{crash(goal, xn), dist(goal, xn)};
            if (c_new < c_min) {
                c_min = c_new;
                x_min = xn;
            }
        }
        if (c_min < Cost{true, 0.0f}) {
            c_base = 0;
            path.push_front(goal);
            Pt xi = x_min;
            while (xi.x >= 0) {
                c_base = c_base + dist(xi, path.front());
                path.push_front(xi);
                xi = nodes[xi].par;
                if (path.size() > 2000) {
                    path.clear();
                    return {true, 2.4e38f};
                }
            }
            return {false, c_base};
        } else {
            Pt xm{-1, -1};
            Cost cm{true, 3.3e38f};
            std::queue<std::pair<int, Pt>> Q;
            Q.push({0, root});
            while (!Q.empty()) {
                auto cur = Q.front();
                Q.pop();
                Node &nc = nodes[cur.second];
                if (cur.first < 5 && nc.subs.size())
                    for (const Pt &s : nc.subs)
                        if (!nodes[s].cost.col) Q.push({cur.first + 1, s});
                Cost c_now{nc.cost.col, dist(cur.second, goal)};
                if (c_now < cm) {
                    cm = c_now;
                    xm = cur.second;
                }
            }
            Pt xi = xm;
            while (xi.x >= 0) {
                path.push_front(xi);
                xi = nodes[xi].par;
                if (path.size() > 2000) {
                    path.clear();
                    return {true, 2.4e38f};
                }
            }
            return {true, 3.4e38f};
        }
    }
};

// ------------------------------------------------------------------------------------------------
// RT-FMT  (RTFMT/lib/RTFMT/RTFMT.cpp)
// ------------------------------------------------------------------------------------------------
struct RTFMT {
    int divisor = 5; // RTFMT.cpp:249
    int K_expand = 25, K_root = 20; // RTFMT.cpp:427,440,477,549 / 677
    struct Node {
        Cost cost;
        Pt par;
        std::set<Pt> subs;
        int id;
    };
    Map map;
    std::map<Pt, Node> nodes;
    KnnIndex index;
    std::vector<Pt> by_id;
    std::set<Pt> N_b, Qr_set, V_unv, V_open, V_opennew, V_toopen, V_closed;
    std::queue<Pt> Qo, Qr;
    std::list<Pt> X_near;
    Pt root{-1, -1}, goal{-1, -1}, agent{-1, -1}, z{-1, -1};
    long long n_edge_checks = 0, n_knn = 0, n_expand = 0;

    void init(const Map &m, double cell = 32) {
        map = m;
        index.init(m.W, m.H, cell);
    }
    bool crash(const Pt &a, const Pt &b) {
        n_edge_checks++;
        return edge_blocked(map, a, b, divisor, PLANE_S | PLANE_D);
    }
    bool crashND(const Pt &a, const Pt &b) {
        n_edge_checks++;
        return edge_blocked(map, a, b, divisor, PLANE_S);
    }
    void knn(const Pt &q, int k, std::vector<Nbr> &out) {
        n_knn++;
        index.query((double)q.x, (double)q.y, k, out);
    }
    Pt pt_of(const Nbr &n) const { return by_id[n.id]; }
    int add_point(const Pt &p) {
        by_id.push_back(p);
        return index.insert((double)p.x, (double)p.y);
    }
    // RTFMT.cpp:11-36
    bool SetAgent(const Pt &p) {
        if (p.x >= map.W || p.y >= map.H) return false;
        if (probe(map.S, map.W, map.H, (int)p.y, (int)p.x)) return false;
        if (nodes.empty()) {
            Node &n = nodes[p];
            n.cost = {false, 0};
            n.par = {-1, -1};
            root = p;
            n.id = add_point(p);
            agent = root;
        } else
            agent = p;
        return true;
    }
    bool SetGoal(const Pt &p) {
        if (p.x >= map.W || p.y >= map.H) return false;
        if (probe(map.S, map.W, map.H, (int)p.y, (int)p.x)) return false;
        if (goal != p) goal = p;
        return true;
    }
    // RTFMT.cpp:393-422: SampleN free-space samples (static plane only).  Returns rand draws used.
    void Start(RandStream &rs, int SampleN) {
        root = agent;
        V_open.clear();
        V_open.insert(root);
        V_opennew.clear();
        V_closed.clear();
        z = root;
        std::queue<Pt>().swap(Qr);
        std::queue<Pt>().swap(Qo);
        for (int i = 0; i < SampleN && !rs.exhausted;) {
            float rx = rs.u() * map.W;
            float ry = rs.u() * map.H;
            if (!probe(map.S, map.W, map.H, (int)ry, (int)rx)) {
                Pt p{rx, ry};
                bool fresh = nodes.find(p) == nodes.end();
                Node &n = nodes[p];
                n.cost = {true, 3.4e38f};
                n.par = {-1, -1};
                n.subs.clear();
                // the reference inserts a duplicate coordinate into the kd-tree too; the injected
                // streams used for parity never produce duplicates (asserted by the tests).
                n.id = add_point(p);
                (void)fresh;
                V_unv.insert(p);
                i++;
            }
        }
    }
    // RTFMT.cpp:92-176
    void UpdateTree() {
        if (agent != root) {
            std::vector<Nbr> nb;
            knn(agent, 1, nb);
            Pt x_near = pt_of(nb[0]);
            if (x_near != root) {
                Pt x_temp = x_near, x_old{-1, -1};
                while (x_temp != root) {
                    Node &nt = nodes[x_temp];
                    Pt x_par = nt.par;
                    nt.par = x_old;
                    nt.subs.insert(x_par);
                    nt.subs.erase(x_old);
                    x_old = x_temp;
                    x_temp = x_par;
                }
                nodes[x_temp].par = x_old;
                nodes[x_temp].subs.erase(x_old);
                root = x_near;
                nodes[root].cost = {false, 0};
            }
        }
        std::queue<Pt> Qup;
        Qup.push(root);
        while (!Qup.empty()) {
            Pt now = Qup.front();
            Qup.pop();
            Node &nn = nodes[now];
            for (const Pt &sub : nn.subs) {
                Node &ns = nodes[sub];
                bool col_old = ns.cost.col;
                ns.cost = nn.cost + Cost{crash(now, sub), dist(now, sub)};
                if (ns.cost.col && !col_old) N_b.insert(sub);
                Qup.push(sub);
            }
        }
    }
    Pt argmin_open() {
        Pt zm{-1, -1};
        Cost cm{true, 3.4e38f};
        for (const Pt &zo : V_open) {
            const Cost &c = nodes[zo].cost;
            if (c < cm) {
                cm = c;
                zm = zo;
            }
        }
        return zm;
    }
    // RTFMT.cpp:423-529
    void ExpandFMT() {
        n_expand++;
        std::vector<Nbr> nb;
        if (V_unv.size() && X_near.size() == 0 && z.x >= 0) {
            knn(z, K_expand, nb);
            for (const Nbr &n : nb) {
                Pt xn = pt_of(n);
                if (V_unv.count(xn)) X_near.push_back(xn);
            }
        }
        if (X_near.size()) {
            Pt x = X_near.front();
            X_near.pop_front();
            knn(x, K_expand, nb);
            float cost_min = 3.4e38f;
            Pt y_min{-1, -1};
            for (const Nbr &n : nb) {
                Pt yn = pt_of(n);
                if (V_open.count(yn)) {
                    float cost_now = nodes[yn].cost.c + dist(yn, x);
                    if (cost_min > cost_now) {
                        cost_min = cost_now;
                        y_min = yn;
                    }
                }
            }
            if (y_min.x >= 0 && !crash(x, y_min) && nodes[y_min].cost.col == 0) {
                V_opennew.insert(x);
                V_unv.erase(x);
                // NB reference quirk (RTFMT.cpp:448,460): cost_min already contains cost(y_min).c
                nodes[x].cost = nodes[y_min].cost + Cost{false, cost_min};
                nodes[x].par = y_min;
                nodes[y_min].subs.insert(x);
            }
        }
        if (X_near.size() == 0 && z.x >= 0) {
            V_open.insert(V_opennew.begin(), V_opennew.end());
            V_opennew.clear();
            V_open.erase(z);
            V_closed.insert(z);
            knn(z, K_expand, nb);
            for (const Nbr &n : nb) {
                Pt xn = pt_of(n);
                if (V_unv.count(xn) && !crashND(z, xn)) {
                    V_toopen.insert(z);
                    break;
                }
            }
            z = argmin_open();
        }
        if (z.x < 0) {
            V_open.insert(V_toopen.begin(), V_toopen.end());
            std::set<Pt> result;
            std::set_difference(V_closed.begin(), V_closed.end(), V_toopen.begin(), V_toopen.end(),
                                std::inserter(result, result.begin()));
            V_closed.swap(result);
            V_toopen.clear();
            z = argmin_open();
        }
    }
    // RTFMT.cpp:531-595
    void RewireFromObstacles() {
        if (Qo.size() == 0) {
            for (const Pt &xb : N_b) Qo.push(xb);
            N_b.clear();
        } else {
            Pt xb = Qo.front();
            Qo.pop();
            if (!probe(map.D, map.W, map.H, (int)xb.y, (int)xb.x)) {
                float cost_min = 3.4e38f;
                Pt y_min{-1, -1};
                std::vector<Nbr> nb;
                knn(xb, K_expand, nb);
                for (const Nbr &n : nb) {
                    Pt yn = pt_of(n);
                    if (yn != xb && (V_open.count(yn) || V_closed.count(yn))) {
                        float cost_now = nodes[yn].cost.c + dist(yn, xb);
                        if (cost_now < cost_min && nodes[yn].cost.col == 0) {
                            cost_min = cost_now;
                            y_min = yn;
                        }
                    }
                }
                // y_min == {-1,-1}: the reference probes pixel (-1,-1) — outside the buffer — defined as blocked
                if (!crash(y_min, xb) && cost_min < nodes[xb].cost.c) {
                    if (nodes[xb].par.x >= 0) nodes[nodes[xb].par].subs.erase(xb);
                    nodes[y_min].subs.insert(xb);
                    nodes[xb].par = y_min;
                    nodes[xb].cost = nodes[y_min].cost + Cost{false, dist(xb, y_min)};
                    Cost pc = nodes[xb].cost;
                    for (const Pt &ps : nodes[xb].subs) {
                        nodes[ps].cost = pc + Cost{crash(xb, ps), dist(xb, ps)};
                        Qo.push(ps);
                    }
                }
            }
        }
    }
    // RTFMT.cpp:665-703
    void RewireFromRoot() {
        if (Qr.size() == 0) {
            Qr_set.clear();
            Qr.push(root);
            Qr_set.insert(root);
        } else {
            Pt xb = Qr.front();
            Qr.pop();
            std::vector<Nbr> nb;
            knn(xb, K_root, nb);
            for (const Nbr &n : nb) {
                Pt xn = pt_of(n);
                if (!(xb != xn)) continue;
                if (V_open.count(xn) || V_closed.count(xn)) {
                    Cost c_old = nodes[xn].cost;
                    Cost c_new = nodes[xb].cost + Cost{crash(xb, xn), dist(xb, xn)};
                    if (c_new < c_old) {
                        nodes[nodes[xn].par].subs.erase(xn);
                        nodes[xb].subs.insert(xn);
                        nodes[xn].par = xb;
                        nodes[xn].cost = c_new;
                    }
                    if (Qr_set.count(xn) == 0) {
                        Qr.push(xn);
                        Qr_set.insert(xn);
                    }
                }
            }
        }
    }
    // RTFMT.cpp:704-728 minus DrawDObsmap (the D plane is an input here)
    void MainTask(int iters = 500) {
        UpdateTree();
        for (int i = 0; i < iters; i++) {
            ExpandFMT();
            RewireFromObstacles();
            RewireFromRoot();
        }
    }
    // RTFMT.cpp:285-391 (first branch; the depth-5 fallback mirrors RTRRT::GetPath)
    bool GetPath(std::list<Pt> &path) {
        path.clear();
        if (goal.x <= 0) return false;
        std::vector<Nbr> nb;
        knn(goal, 25, nb);
        Pt x_min{-1, -1};
        Cost c_min{true, FLT_MAX};
        for (const Nbr &n : nb) {
            Pt xn = pt_of(n);
            Cost c_new = nodes[xn].cost + Cost{crash(goal, xn), dist(goal, xn)};
            if (c_new < c_min) {
                c_min = c_new;
                x_min = xn;
            }
        }
        if (c_min < Cost{false, 20000.0f}) {
            path.push_front(goal);
            Pt xi = x_min;
            while (xi.x >= 0) {
                path.push_front(xi);
                xi = nodes[xi].par;
                if (path.size() > 2000) {
                    path.clear();
                    return false;
                }
            }
            return true;
        }
        Pt xm{-1, -1};
        Cost cm{true, 3.3e38f};
        std::queue<std::pair<int, Pt>> Q;
        Q.push({0, root});
        while (!Q.empty()) {
            auto cur = Q.front();
            Q.pop();
            Node &nc = nodes[cur.second];
            if (cur.first < 5 && nc.subs.size())
                for (const Pt &s : nc.subs)
                    if (!nodes[s].cost.col) Q.push({cur.first + 1, s});
            Cost c_now{nc.cost.col, dist(cur.second, goal)};
            if (c_now < cm) {
                cm = c_now;
                xm = cur.second;
            }
        }
        Pt xi = xm;
        while (xi.x >= 0) {
            path.push_front(xi);
            xi = nodes[xi].par;
            if (path.size() > 2000) {
                path.clear();
                return false;
            }
        }
        return false;
    }
};

// ------------------------------------------------------------------------------------------------
// AM-RRT*  (AMRRT/lib/AMrrt/AMRRT.cpp)
// ------------------------------------------------------------------------------------------------
struct AMRRT {
    static constexpr int KnearN = 25;           // AMRRT.cpp:153
    static constexpr float SteerLenMAX = 100.f; // AMRRT.cpp:353
    static constexpr float alpha = 0.9f, beta = 4.0f, k_max = 4.0f; // AMRRT.cpp:392-394
    struct Node {
        Cost cost;
        Pt par;
        std::set<Pt> subs;
        int id;
    };
    Map map;
    std::map<Pt, Node> nodes;
    KnnIndex index;
    std::vector<Pt> by_id;
    std::queue<Pt> Qroot;
    std::set<Pt> rewired_root;
    std::list<Pt> Sgoal;
    std::queue<Pt> Qgoal;
    std::set<Pt> seen_goal;
    std::list<Pt> OutPath;
    Pt root{-1, -1}, goal{-1, -1}, agent{-1, -1};
    float c_base = 0; // uninitialised in the reference (AMRRT.h:133); 0 here
    VClock clk;
    // diffusion table (AMRRT.cpp:58-93): cell (i,j) -> row; rows x dim floats
    std::map<std::pair<int, int>, int> index_to_pos;
    std::vector<float> diffusion;
    int diff_dim = 0;
    float diffusion_ratios = 1;
    long long n_edge_checks = 0, n_knn = 0, n_expand = 0, n_pops_root = 0;

    void init(const Map &m, double cell = 32) {
        map = m;
        index.init(m.W, m.H, cell);
    }
    bool crash(const Pt &a, const Pt &b) {
        n_edge_checks++;
        return edge_blocked(map, a, b, 2, PLANE_S | PLANE_D);
    }
    void knn(const Pt &q, int k, std::vector<Nbr> &out) {
        n_knn++;
        index.query((double)q.x, (double)q.y, k, out);
    }
    Pt pt_of(const Nbr &n) const { return by_id[n.id]; }
    int add_point(const Pt &p) {
        by_id.push_back(p);
        return index.insert((double)p.x, (double)p.y);
    }
    bool SetAgent(const Pt &p) {
        if (p.x >= map.W || p.y >= map.H) return false;
        if (probe(map.S, map.W, map.H, (int)p.y, (int)p.x)) return false;
        if (nodes.empty()) {
            Node &n = nodes[p];
            n.cost = {false, 0};
            n.par = {-1, -1};
            root = p;
            n.id = add_point(p);
            agent = root;
        } else
            agent = p;
        return true;
    }
    bool SetGoal(const Pt &p) {
        if (p.x >= map.W || p.y >= map.H) return false;
        if (probe(map.S, map.W, map.H, (int)p.y, (int)p.x)) return false;
        if (goal != p) goal = p;
        return true;
    }
    // AMRRT.cpp:94-116
    float DiffusivityMetric(const Pt &xs, const Pt &xe) {
        int is = (int)(xs.x * diffusion_ratios), js = (int)(xs.y * diffusion_ratios);
        auto a = index_to_pos.find({is, js});
        if (a == index_to_pos.end()) return 3.4e38f;
        int ie = (int)(xe.x * diffusion_ratios), je = (int)(xe.y * diffusion_ratios);
        auto b = index_to_pos.find({ie, je});
        if (b == index_to_pos.end()) return 3.4e38f;
        const float *ps = &diffusion[(size_t)a->second * diff_dim];
        const float *pe = &diffusion[(size_t)b->second * diff_dim];
        float len = 0;
        for (int i = 0; i < diff_dim; i++) {
            float d = ps[i] - pe[i];
            float dd = d * d;
            len = len + dd;
        }
        return sqrtf(len) / diffusion_ratios;
    }
    // AMRRT.cpp:156-200
    void UpdateTree() {
        if (agent != root) {
            std::vector<Nbr> nb;
            knn(agent, 1, nb);
            Pt x_near = pt_of(nb[0]);
            if (x_near != root) {
                Pt x_temp = x_near, x_old{-1, -1};
                while (x_temp != root) {
                    Node &nt = nodes[x_temp];
                    Pt x_par = nt.par;
                    nt.par = x_old;
                    nt.subs.insert(x_par);
                    nt.subs.erase(x_old);
                    x_old = x_temp;
                    x_temp = x_par;
                }
                nodes[x_temp].par = x_old;
                nodes[x_temp].subs.erase(x_old);
                root = x_near;
                nodes[root].cost = {false, 0};
            }
        }
        std::queue<Pt> Qup;
        Qup.push(root);
        while (!Qup.empty()) {
            Pt now = Qup.front();
            Qup.pop();
            Node &nn = nodes[now];
            for (const Pt &sub : nn.subs) {
                Node &ns = nodes[sub];
                ns.cost = nn.cost + Cost{crash(now, sub), dist(now, sub)};
                Qup.push(sub);
            }
        }
    }
    // AMRRT.cpp:354-390
    Pt Steer(Pt x_init, Pt x_end, RandStream &rs) {
        Pt x_start;
        n_edge_checks++;
        if (!edge_first_hit(map, x_init, x_end, &x_start)) {
            float len = dist(x_init, x_end);
            float p = SteerLenMAX / len;
            if (p < 1.0f) {
                float x = (1 - p) * x_init.x + p * x_end.x;
                float y = (1 - p) * x_init.y + p * x_end.y;
                return {x, y};
            }
            return x_end;
        }
        Pt x_min = x_start;
        float c_min = DiffusivityMetric(x_start, x_end);
        float r_sam = std::min(SteerLenMAX, dist(x_init, x_end));
        long long t0 = clk.now();
        while (clk.now() - t0 < 5000000LL) {
            float th = rs.u() * 2 * 3.141592954f;
            float rr = rs.u();
            rr = sqrtf(rr) * r_sam;
            float x = cosf(th) * rr + x_init.x;
            float y = sinf(th) * rr + x_init.y;
            Pt x_new{x, y};
            float c_new = DiffusivityMetric(x_new, x_end);
            if (c_new < c_min && !crash(x_init, x_new)) {
                x_min = x_new;
                c_min = c_new;
            }
        }
        return x_min;
    }
    // AMRRT.cpp:395-481
    void Expand(RandStream &rs) {
        n_expand++;
        Pt x_rand{0, 0};
        while (!rs.exhausted) {
            float p = rs.u();
            if (p > alpha && OutPath.size() == 0 && goal.x >= 0)
                x_rand = goal; // no break in the reference (AMRRT.cpp:402-403): re-draws
            else if (p < alpha / beta || OutPath.size() == 0) {
                float rx = rs.u() * map.W;
                float ry = rs.u() * map.H;
                if (!probe_mask(map, PLANE_S | PLANE_D, (int)ry, (int)rx)) {
                    x_rand = {rx, ry};
                    break;
                }
            } else if (goal.x >= 0) {
                float c_min = dist(goal, root);
                float a = c_base / 2.0f;
                float b = sqrtf(c_base * c_base - c_min * c_min) / 2.0f; // float overloads, see RTRRT::sample
                float angle = atan2f(goal.y - root.y, goal.x - root.x);
                float th = rs.u() * 2 * 3.141592954f;
                float rr = sqrtf(rs.u());
                float x = a * cosf(th) * rr;
                float y = b * sinf(th) * rr;
                float rx = cosf(angle) * x - sinf(angle) * y;
                float ry = sinf(angle) * x + cosf(angle) * y;
                rx = rx + (goal.x + root.x) / 2;
                ry = ry + (goal.y + root.y) / 2;
                if (rx >= 0 && rx < map.W && ry >= 0 && ry < map.H)
                    if (!probe_mask(map, PLANE_S | PLANE_D, (int)ry, (int)rx)) {
                        x_rand = {rx, ry};
                        break;
                    }
            }
        }
        if (rs.exhausted) return;
        std::vector<Nbr> nb;
        knn(x_rand, KnearN, nb);
        Pt x_nrst = pt_of(nb[0]);
        Pt x_new = Steer(x_nrst, x_rand, rs);
        knn(x_new, KnearN, nb);
        Pt nearest = pt_of(nb[0]);
        if (dist(nearest, x_new) > k_max) {
            Pt x_min = x_nrst;
            Cost c_min = nodes[x_nrst].cost + Cost{false, dist(x_nrst, x_new)};
            for (const Nbr &n : nb) {
                Pt xn = pt_of(n);
                Cost c_new = nodes[xn].cost + Cost{crash(xn, x_new), dist(xn, x_new)};
                if (c_new < c_min && !c_new.col) {
                    x_min = xn;
                    c_min = c_new;
                }
            }
            if (!c_min.col) {
                nodes[x_min].subs.insert(x_new);
                Node &nn = nodes[x_new];
                nn.cost = c_min;
                nn.par = x_min;
                nn.subs.clear();
                nn.id = add_point(x_new);
            }
        }
    }
    // AMRRT.cpp:483-519
    void RewireRoot() {
        if (Qroot.size() == 0) {
            Qroot.push(root);
            rewired_root.clear();
            rewired_root.insert(root);
        }
        long long t0 = clk.now();
        while (clk.now() - t0 < 20000000LL && Qroot.size()) {
            n_pops_root++;
            Pt xr = Qroot.front();
            Qroot.pop();
            std::vector<Nbr> nb;
            knn(xr, KnearN, nb);
            for (const Nbr &n : nb) {
                Pt xn = pt_of(n);
                if (!crash(xn, xr)) {
                    Cost c_old = nodes[xn].cost;
                    Cost c_new = nodes[xr].cost + Cost{false, dist(xn, xr)};
                    if (c_new < c_old && !c_new.col) {
                        nodes[nodes[xn].par].subs.erase(xn);
                        nodes[xn].par = xr;
                        nodes[xn].cost = c_new;
                        nodes[xr].subs.insert(xn);
                    }
                    if (rewired_root.count(xn) == 0) {
                        rewired_root.insert(xn);
                        Qroot.push(xn);
                    }
                }
            }
        }
    }
    // AMRRT.cpp:521-585
    void RewireGoal() {
        if (Qgoal.size() == 0 && Sgoal.size() == 0) {
            Sgoal.push_front(root);
            seen_goal.clear();
        }
        long long t0 = clk.now();
        while (clk.now() - t0 < 100000000LL && (Sgoal.size() || Qgoal.size())) {
            Pt xr;
            if (Sgoal.size()) {
                xr = Sgoal.front();
                Sgoal.pop_front();
            } else {
                xr = Qgoal.front();
                Qgoal.pop();
            }
            if (dist(root, xr) + dist(goal, xr) <= c_base) {
                std::list<Pt> X_next;
                typedef std::pair<float, Pt> NB;
                std::priority_queue<NB, std::vector<NB>, std::greater<NB>> sorted;
                std::vector<Nbr> nb;
                knn(xr, KnearN, nb);
                for (const Nbr &n : nb) {
                    Pt xn = pt_of(n);
                    sorted.push({DiffusivityMetric(goal, xn), xn});
                }
                while (sorted.size()) {
                    Pt xn = sorted.top().second;
                    sorted.pop();
                    if (!crash(xn, xr)) {
                        Cost c_new = nodes[xr].cost + Cost{false, dist(xn, xr)};
                        if (c_new < nodes[xn].cost && !c_new.col) {
                            nodes[nodes[xn].par].subs.erase(xn);
                            nodes[xn].par = xr;
                            nodes[xn].cost = c_new;
                            nodes[xr].subs.insert(xn);
                        }
                        if (seen_goal.count(xn) == 0) {
                            seen_goal.insert(xn);
                            X_next.push_back(xn);
                        }
                    }
                }
                for (const Pt &xn : X_next) {
                    Sgoal.push_front(xn);
                    Qgoal.push(xn);
                }
            }
            if (Sgoal.size())
                if (DiffusivityMetric(Sgoal.front(), goal) > DiffusivityMetric(xr, goal)) Sgoal.clear();
        }
    }
    // AMRRT.cpp:590-617 minus DrawDObsmap
    void MainTask(RandStream &rs) {
        UpdateTree();
        long long t0 = clk.now();
        while (clk.now() - t0 < 80000000LL && !rs.exhausted) {
            Expand(rs);
            RewireRoot();
            if (goal.x >= 0) {
                auto it = nodes.find(goal);
                if (it != nodes.end() && !it->second.cost.col) RewireGoal();
            }
        }
    }
    // AMRRT.cpp:310-351
    Cost GetPath(std::list<Pt> &path) {
        c_base = -1;
        path.clear();
        if (goal.x <= 0) return {true, 3.4e38f};
        std::vector<Nbr> nb;
        knn(goal, KnearN, nb);
        Pt x_min{-1, -1};
        Cost c_min{true, FLT_MAX};
        for (const Nbr &n : nb) {
            Pt xn = pt_of(n);
            Cost c_new = nodes[xn].cost + Cost{crash(goal, xn), dist(goal, xn)};
            if (c_new < c_min) {
                c_min = c_new;
                x_min = xn;
            }
        }
        if (c_min < Cost{true, 0.0f}) {
            c_base = 0;
            path.push_front(goal);
            Pt xi = x_min;
            while (xi.x >= 0) {
                c_base = c_base + dist(xi, path.front());
                path.push_front(xi);
                xi = nodes[xi].par;
                if (path.size() > 2000) {
                    c_base = -1;
                    path.clear();
                    return {true, 2.4e38f};
                }
            }
            return {false, c_base};
        }
        return {true, 3.4e38f};
    }
};

} // namespace orc
