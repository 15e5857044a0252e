// CDRT.h — drop-in CDRT-RRT* planner class (BImap) over the B200 engine (host side, C++).
//
// Mirrors the public surface of the reference's BImap (CDRT2/lib/CDRT/CDRT.h:33-357): same value types
// (BIpoint, CTrtrt_Cost, BIline, BIcutline, BIfreepolygon, BIgraph, Node, CTrtrt_DynamicObstacle), the same
// public members demos read and write (agent, agent_path, goal, Node_Map, DObstacles — mutated in place per
// frame, CDRT2/test/testD.cpp:189-190 — BIamap/BIimap, r_th, p_th, nearK, RewireN) and the same methods with
// the same error behaviour (printf + silent return).  What changed underneath:
//   * the tree, the per-island k-d trees, Mnode, eQr and the larger/smaller sets live on the GPU inside an
//     rtpp_cdrt_engine (include/rtpp_gpu.h); RTRT_task() is one engine tick (+ GetPath), Node_Map and
//     agent_path are materialised from it (SyncNodeMap());
//   * MaptoBInavi's UDP round trip to approx_work.py and the convex decomposition (StartCut) are out of scope:
//     the decomposition arrives through SetGraph() (the arrays of a BIgraph, e.g. from a fixture);
//   * cv::Mat planes are rtpp::Plane16; DrawBImap uses the host fillConvexPoly of raster.hpp;
//   * wall-clock gates become the virtual clock (SetVirtualClock), rand() draws can be injected (SetRandStream).
// No CPU fallback: construction fails loudly (ok() == false) without a CUDA device.
#ifndef RTPP_B200_CDRT_H
#define RTPP_B200_CDRT_H
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <list>
#include <map>
#include <set>
#include <vector>

#include "../../include/rtpp_gpu.h"
#include "raster.hpp"

struct BIpoint { // CDRT.h:34-90
    double x, y;
    bool operator==(const BIpoint &o) const { return std::abs(x - o.x) <= 1e-16 && std::abs(y - o.y) <= 1e-16; }
    bool operator!=(const BIpoint &o) const { return std::abs(x - o.x) > 1e-16 || std::abs(y - o.y) > 1e-16; }
    bool operator<(const BIpoint &o) const { return x != o.x ? x < o.x : y < o.y; }
    BIpoint operator+(const BIpoint &o) const { return {x + o.x, y + o.y}; }
    BIpoint operator-(const BIpoint &o) const { return {x - o.x, y - o.y}; }
    BIpoint operator*(const double &r) const { return {r * x, r * y}; }
    BIpoint operator/(const double &r) const { return {x / r, y / r}; }
    double operator%(const BIpoint &o) const {
        double dx = x - o.x, dy = y - o.y;
        return sqrt(dx * dx + dy * dy);
    }
    double cross(const BIpoint &o) const { return x * o.y - y * o.x; }
};
struct CTrtrt_Cost { // CDRT.h:92-131
    bool collision;
    double cost;
    CTrtrt_Cost operator+(const CTrtrt_Cost &o) const { return {collision || o.collision, cost + o.cost}; }
    bool operator==(const CTrtrt_Cost &o) const { return collision == o.collision && cost == o.cost; }
    bool operator<(const CTrtrt_Cost &o) const { return collision == o.collision ? cost < o.cost : collision < o.collision; }
    bool operator>(const CTrtrt_Cost &o) const { return collision == o.collision ? cost > o.cost : collision > o.collision; }
    bool operator<=(const CTrtrt_Cost &o) const { return collision == o.collision ? cost <= o.cost : collision <= o.collision; }
    bool operator>=(const CTrtrt_Cost &o) const { return collision == o.collision ? cost >= o.cost : collision >= o.collision; }
};
typedef std::vector<BIpoint> BIpolygon;
typedef std::vector<BIpolygon> BIpolygons;
struct BIline { // end points ordered, CDRT.h:155-182
    BIpoint S, E;
    BIline(BIpoint p1, BIpoint p2) {
        if (p2 < p1) S = p2, E = p1;
        else S = p1, E = p2;
    }
};
struct BIcutline {
    int32_t index;
    BIline line;
    BIpoint core;
    std::set<int32_t> polygonlink, cutlinelink;
};
struct BIfreepolygon {
    int32_t index;
    BIpolygon polygon;
    BIpoint core;
    std::set<int32_t> polygonlink, cutlinelink;
};
#define originIndexFun() ((int32_t)(0x40000000))
#define obsNodeIndexFun(x) ((int32_t)(0x80000000 | x))
#define cutNodeIndexFun(x) ((int32_t)x)
typedef struct {
    CTrtrt_Cost cost;
    BIpoint par;
    std::set<BIpoint> subs;
    int32_t invM;
} Node;
struct BIgraph {
    int32_t cutlineBaseNum = 0, freepolygonBaseNum = 0;
    std::vector<BIcutline> cutlineList;
    std::vector<BIfreepolygon> freepolygonList;
};
typedef struct {
    int DObsType; // -1 off, 0 circles at the outline points, 1 polygon
    double x, y, angle, radius, sampleRadius;
    std::vector<BIpoint> OutLine;
    std::vector<BIpoint> CheckPoints;
    std::set<int> ObsIslands;
} CTrtrt_DynamicObstacle;
#define doubleMax (1.79769e+307)

#ifndef RTPP_PLANNER_API
#define RTPP_PLANNER_API __attribute__((visibility("default")))
#endif
class RTPP_PLANNER_API BImap {
  private:
    rtpp_ctx *ctx_ = nullptr;
    rtpp_cdrt_engine *eng_ = nullptr;
    long long dt_ns_ = 20000;
    std::vector<int32_t> rand_;
    bool rand_injected_ = false;
    std::map<int, double> check_len_;      // CheckLen of each obstacle (RTRT_AddDynamicObstacle's third argument)
    std::map<int, std::pair<int, size_t>> sent_; // what the engine knows: (type, outline size)
    int node_capacity_ = 16384, queue_capacity_ = 1 << 21, rand_capacity_ = 1 << 20;
    bool Fail(const char *what, int rc);
    bool Placeable(const BIpoint &p, const char *who);
    void PushObstacles();
    void RefillRand();

  public:
    uint32_t shapeX = 0, shapeY = 0;
    double mapratio = 0, robotsize = 0;
    rtpp::Plane16 BIamap, BIimap;
    std::vector<BIgraph> BIgraphList;
    double r_th = 10, p_th = 0.85;
    int32_t nearK = 20, RewireN = 100;
    BIpoint agent{0, 0}, agent_tag{0, 0}, goal{0, 0};
    int32_t agent_apl = 0, agent_island = 0, goalIsSet = false;
    std::list<BIpoint> agent_path;
    std::map<BIpoint, Node> Node_Map; // filled by SyncNodeMap()
    std::map<int32_t, CTrtrt_DynamicObstacle> DObstacles;

    explicit BImap(int device = 0);
    ~BImap();
    BImap(const BImap &) = delete;
    BImap &operator=(const BImap &) = delete;
    bool ok() const { return ctx_ != nullptr; }

    // the decomposition of one connected free region (what StartCut leaves in a BIgraph); CSR link lists
    void SetShape(uint32_t cols, uint32_t rows) { shapeX = cols, shapeY = rows; }
    int SetGraph(int ncut, const double *cut_se, const double *cut_core, const int32_t *cp_off, const int32_t *cp, const int32_t *cc_off,
                 const int32_t *cc, int npoly, const int32_t *pv_off, const double *pv, const double *poly_core, const int32_t *pc_off,
                 const int32_t *pc, const int32_t *pp_off, const int32_t *pp);
    void DrawBImap(bool debugmap);
    void RTRT_task(void);
    void RTRT_start(double Resolution);
    void RTRT_SetAgent(const BIpoint &point);
    void RTRT_SetGoal(const BIpoint &point);
    CTrtrt_Cost GetPath(void);
    void RTRT_AddDynamicObstacle(int num, CTrtrt_DynamicObstacle &obs, double Check = 0);
    void RTRT_DelDynamicObstacle(int num);

    // --- additions (not in the reference) ---
    void SetVirtualClock(long long dt_ns) { dt_ns_ = dt_ns > 0 ? dt_ns : 1; }
    void SetRandStream(const int32_t *r, long long n) { rand_.assign(r, r + n), rand_injected_ = true; }
    void SyncNodeMap(); // device tree -> Node_Map (with subs)
    int ReadStats(long long *out24);
    int ReadTree(double *xy, int32_t *parent, double *cost, uint8_t *col, int32_t *invM, uint8_t *alive, int max_nodes);
};
#endif
