// RTRRT.h — drop-in RT-RRT* planner class over the B200 hot path (host side, C++).
//
// Mirrors the public surface of the reference's RTRRTplanner (RTRRT/lib/RTrrt/RTRRT.h:28-158): same
// type names, method names, argument meaning and error behaviour (printf + silent return, no
// exceptions), so that callers such as RTRRT/test/testD.cpp keep compiling.  What changed underneath:
//   * the tree (Node_Map keys, costs, parents), the k-d tree and both queues live on the GPU inside
//     an rtpp_engine instance (include/rtpp_gpu.h); UpdateTree / ExpandRewiring / GetPath are engine
//     calls; Node_Map is materialised on demand (SyncNodeMap) for callers that iterate it;
//   * cv::Mat planes are rtpp::Plane8 (this image has no OpenCV); InitializeInvironment reads a binary
//     PGM (P5) or takes a raw grey buffer; DrawDObsmap uses the host rasteriser in raster.hpp;
//   * wall-clock budgets become the virtual clock (SetVirtualClock): ExpandRewiring(50) performs
//     50e6/dt_ns clock reads of work, reproducibly;
//   * rand() draws can be injected (SetRandStream) for parity runs; by default glibc rand() is used
//     exactly as the reference does.
// No CPU fallback: construction fails loudly (ok() == false, message on stderr) without a CUDA device.
#ifndef RTPP_B200_RTRRT_H
#define RTPP_B200_RTRRT_H
#include "rtpp_types.h"

class RTPP_PLANNER_API RTRRTplanner {
  private:
    RTpoint root = {-1, -1};
    RTpoint goal = {-1, -1};
    float c_base = -1;
    rtpp_ctx *ctx_ = nullptr;
    rtpp_engine *eng_ = nullptr;
    bool started_ = false;
    long long dt_ns_ = 5000;
    const int32_t *rand_stream_ = nullptr;
    long long rand_n_ = 0, rand_pos_ = 0;
    int node_capacity_ = 65536, max_samples_ = 16384;
    std::vector<int32_t> pending_; // rand() values drawn for the engine and not consumed yet
    bool EnsureEngine();
    void UploadAgentGoal();
    bool Fail(const char *what, int rc);

  public:
    RTpoint agent;
    int32_t shapeX = 0, shapeY = 0;
    std::map<int, RTrrt_DynamicObstacle> DObstacles;
    std::map<RTpoint, RTrrt_NodeAttribute> Node_Map; // filled by SyncNodeMap()
    std::list<RTpoint> OutPath;
    rtpp::Plane8 Freemap, Obsmap, DObsmap;

    explicit RTRRTplanner(int device = 0);
    ~RTRRTplanner();
    RTRRTplanner(const RTRRTplanner &) = delete;
    RTRRTplanner &operator=(const RTRRTplanner &) = delete;
    bool ok() const { return ctx_ != nullptr; }

    void SetAgent(RTpoint &point);
    void SetGoal(RTpoint &point);
    void InitializeInvironment(const char *IMGmap, int th);                       // binary PGM (P5)
    void InitializeInvironment(const uint8_t *gray, int rows, int cols, int th);  // raw grey plane
    void AddDynamicObstacle(int num, RTrrt_DynamicObstacle &obs);
    void DelDynamicObstacle(int num);
    void UpdateTree(void);
    void Start(void);
    void MainTask(void);
    void ExpandRewiring(long RunTime); // ms of virtual time
    void DrawDObsmap(void);
    bool CrashDetection(const RTpoint &ps, const RTpoint &pe);
    RTrrt_Cost GetPath(std::list<RTpoint> &path);

    // --- additions (not in the reference) ---
    void SetVirtualClock(long long dt_ns) { dt_ns_ = dt_ns > 0 ? dt_ns : 1; }
    void SetRandStream(const int32_t *r, long long n) { rand_stream_ = r, rand_n_ = n, rand_pos_ = 0; }
    long long RandPos() const { return rand_pos_; }
    void SetDObsmap(const uint8_t *plane);   // inject a dynamic plane instead of DrawDObsmap()
    void SyncNodeMap();                      // device tree -> Node_Map (with subs)
    int NodeCount();
    int ReadTree(float *xy, int32_t *parent, float *cost, uint8_t *col, int max_nodes);
    RTpoint Root() const { return root; }
    void QueueSizes(int *qr, int *qs);
    long long ClockReads();
};
#endif
