// AMRRT.h — drop-in AM-RRT* planner class over the B200 engine (planner_kind 2), host side, C++.
// Mirrors AMRRTplanner (reference AMRRT/lib/AMrrt/AMRRT.h:126-183).  MainTask runs UpdateTree on the
// device; the informed-sampler ellipse and the cosf / sinf of the drawn angles are evaluated on the device too (csrc/glibc_math.cuh),
// then runs the Expand / RewireRoot loop on the device with rand() drawn from a host-filled stream.
// Expand / RewireRoot / RewireGoal / Steer are not individually callable: they live inside MainTask.
#ifndef RTPP_B200_AMRRT_H
#define RTPP_B200_AMRRT_H
#include "rtpp_types.h"

struct DIFFpoint {
    int x;
    int y;
    bool operator<(const DIFFpoint &o) const { return x != o.x ? x < o.x : y < o.y; }
};

class RTPP_PLANNER_API AMRRTplanner : public rtppPlannerBase {
  private:
    RTpoint agent = {-1, -1};
    RTpoint root = {-1, -1};
    RTpoint goal = {-1, -1};
    float c_base = 0; // uninitialised in the reference (AMRRT.h:133)
    std::vector<std::vector<float>> diffusion_matrix;
    std::map<DIFFpoint, int> index_to_pos;
    float diffusion_ratios = 1;
    rtpp_engine *eng_ = nullptr;
    bool started_ = false;
    long long dt_ns_ = 500000;
    int node_capacity_ = 65536, rand_capacity_ = 262144;
    bool EnsureEngine();
    void UploadDiffusion();
    void OnStaticPlane() override;
    void OnDynamicPlane() override;
    bool OnDynamicShapes(const int32_t *circles, int n_circles, const int32_t *poly_xy, const int32_t *poly_sizes, int n_polys) override;

  public:
    std::map<RTpoint, RTrrt_NodeAttribute> Node_Map; // filled by SyncNodeMap()
    std::list<RTpoint> OutPath;

    explicit AMRRTplanner(int device = 0) : rtppPlannerBase(device) {}
    ~AMRRTplanner();
    void SetAgent(RTpoint &point);
    void SetGoal(RTpoint &point);
    float DiffusivityMetric(RTpoint &x_s, RTpoint &x_e);
    void InitializeInvironment(const char *IMGmap, int th) { LoadPGM(IMGmap, th); }
    void InitializeInvironment(const uint8_t *gray, int rows, int cols, int th) { LoadGray(gray, rows, cols, th); }
    void InitializeDiffusion(const char *DiffFile, const char *IndexFile, float ratios); // the reference's two JSON files
    void InitializeDiffusion(const int32_t *cells_ij, int ncells, const float *matrix, int dim, float ratios);
    void UpdateTree(void);
    void Start(void) {}
    void MainTask(void);
    bool CrashDetection(const RTpoint &ps, const RTpoint &pe) { return EdgeBlocked(ps, pe, 2, 3); }
    bool CrashDetectionStart(const RTpoint &ps, const RTpoint &pe, RTpoint &pstart);
    RTrrt_Cost GetPath(std::list<RTpoint> &path);

    // --- additions ---
    void SetVirtualClock(long long dt_ns) { dt_ns_ = dt_ns > 0 ? dt_ns : 1; }
    int NodeCount();
    int ReadTree(float *xy, int32_t *parent, float *cost, uint8_t *col, int max_nodes);
    void SyncNodeMap();
};
#endif
