// rtpp_types.h — value types shared by the drop-in planner classes.  Same names, fields and operators
// as the reference declares (identically) in RTRRT.h:28-113, RTFMT.h:28-113 and AMRRT.h:28-113.
#ifndef RTPP_B200_TYPES_H
#define RTPP_B200_TYPES_H
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <list>
#include <map>
#include <queue>
#include <set>
#include <vector>

#include "../../include/rtpp_gpu.h"
#include "raster.hpp"

struct RTpoint {
    float x;
    float y;
    bool operator!=(const RTpoint &o) const { return x != o.x || y != o.y; }
    bool operator<(const RTpoint &o) const { return x != o.x ? x < o.x : y < o.y; }
    RTpoint operator+(const RTpoint &o) const { return {x + o.x, y + o.y}; }
    float operator%(const RTpoint &o) const { // Euclidean distance, binary32, unfused
        float dx = x - o.x, dy = y - o.y;
        float xx = dx * dx, yy = dy * dy;
        return sqrtf(xx + yy);
    }
};

struct RTrrt_Cost {
    bool collision;
    float cost;
    RTrrt_Cost operator+(const RTrrt_Cost &o) const { return {collision || o.collision, cost + o.cost}; }
    bool operator==(const RTrrt_Cost &o) const { return collision == o.collision && cost == o.cost; }
    bool operator<(const RTrrt_Cost &o) const { return collision == o.collision ? cost < o.cost : collision < o.collision; }
    bool operator>(const RTrrt_Cost &o) const { return o < *this; }
    bool operator<=(const RTrrt_Cost &o) const { return !(o < *this); }
    bool operator>=(const RTrrt_Cost &o) const { return !(*this < o); }
};

typedef struct {
    int DObsType; // -1 off, 0 circles at the outline points, 1 polygon
    float x, y, angle, radius, sampleRadius;
    std::vector<RTpoint> OutLine;
    std::vector<RTpoint> CheckPoints;
    std::set<int> ObsIslands;
} RTrrt_DynamicObstacle;

struct RTrrt_NodeAttribute {
    RTrrt_Cost cost;
    RTpoint par;
    std::set<RTpoint> subs;
};

typedef RTrrt_Cost RTFMT_Cost;
typedef RTrrt_DynamicObstacle RTFMT_DynamicObstacle;
typedef RTrrt_NodeAttribute RTFMT_NodeAttribute;

// What the three image-based planners share verbatim in the reference: planes, obstacle set,
// InitializeInvironment (RTRRT.cpp:57-80), DrawDObsmap (RTRRT.cpp:138-169), injected rand().
// the planner classes are the library's C++ surface: exported although the library is built with hidden visibility
#define RTPP_PLANNER_API __attribute__((visibility("default")))

class RTPP_PLANNER_API rtppPlannerBase {
  protected:
    rtpp_ctx *ctx_ = nullptr;
    const int32_t *rand_stream_ = nullptr;
    long long rand_n_ = 0, rand_pos_ = 0;
    int NextRand();
    bool Fail(const char *who, const char *what, int rc);
    virtual void OnStaticPlane() {}               // a new Obsmap was loaded
    virtual void OnDynamicPlane() {}              // DObsmap changed (upload to the engine)
    // the same change as integer shapes (DrawDObsmap): return true when the engine drew them itself on the device
    virtual bool OnDynamicShapes(const int32_t *, int, const int32_t *, const int32_t *, int) { return false; }
    void LoadGray(const uint8_t *gray, int rows, int cols, int th);
    bool LoadPGM(const char *path, int th);
    bool EdgeBlocked(const RTpoint &ps, const RTpoint &pe, int divisor, int mask);
    void UploadDObsmap();
    bool auto_draw_ = true; // MainTask() rasterises DObstacles itself unless a plane was injected

  public:
    int32_t shapeX = 0, shapeY = 0;
    std::map<int, RTrrt_DynamicObstacle> DObstacles;
    rtpp::Plane8 Freemap, Obsmap, DObsmap;
    explicit rtppPlannerBase(int device);
    virtual ~rtppPlannerBase();
    bool ok() const { return ctx_ != nullptr; }
    void AddDynamicObstacle(int num, RTrrt_DynamicObstacle &obs) { DObstacles[num] = obs; }
    void DelDynamicObstacle(int num) { DObstacles.erase(num); }
    void DrawDObsmap(void);
    void SetDObsmap(const uint8_t *plane); // inject a dynamic plane; MainTask() then stops redrawing it
    void SetRandStream(const int32_t *r, long long n) { rand_stream_ = r, rand_n_ = n, rand_pos_ = 0; }
    long long RandPos() const { return rand_pos_; }
};
#endif
