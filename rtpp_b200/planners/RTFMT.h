// RTFMT.h — drop-in RT-FMT planner class over the B200 RT-FMT engine (host side, C++).
// Mirrors RTFMTplanner (reference RTFMT/lib/RTFMT/RTFMT.h:115-173): same method names and meaning.
// The tree, the V_* sets and both queues live on the device (rtpp_fmt_engine); the std::set members
// the reference exposes (V_unv, V_open, ...) are materialised on demand by SyncSets().
#ifndef RTPP_B200_RTFMT_H
#define RTPP_B200_RTFMT_H
#include "rtpp_types.h"

class RTPP_PLANNER_API RTFMTplanner : public rtppPlannerBase {
  private:
    RTpoint agent = {-1, -1};
    RTpoint root = {-1, -1};
    RTpoint goal = {-1, -1};
    RTpoint z = {-1, -1};
    rtpp_fmt_engine *eng_ = nullptr;
    bool have_agent_ = false, started_ = false;
    int sample_n_ = 2000; // SampleN, RTFMT.cpp:392
    std::vector<float> nodes_; // node 0 = agent, then the samples of every Start() call
    bool Restart();
    void OnStaticPlane() override;
    void OnDynamicPlane() override;
    void Tick(int update, int iters, int mask, int path);

  public:
    std::map<RTpoint, RTFMT_NodeAttribute> Node_Map;
    std::set<RTpoint> N_b, V_unv, V_open, V_opennew, V_toopen, V_closed;
    std::list<RTpoint> OutPath;

    explicit RTFMTplanner(int device = 0) : rtppPlannerBase(device) {}
    ~RTFMTplanner();
    void SetAgent(RTpoint &point);
    void SetGoal(RTpoint &point);
    void InitializeInvironment(const char *IMGmap, int th) { LoadPGM(IMGmap, th); }
    void InitializeInvironment(const uint8_t *gray, int rows, int cols, int th) { LoadGray(gray, rows, cols, th); }
    void UpdateTree(void) { Tick(1, 0, 0, 0); }
    void Start(void);
    void MainTask(void);
    void ExpandFMT(void) { Tick(0, 1, 1, 0); }
    void RewireFromObstacles(void) { Tick(0, 1, 2, 0); }
    void RewireFromRoot(void) { Tick(0, 1, 4, 0); }
    bool CrashDetection(const RTpoint &ps, const RTpoint &pe) { return EdgeBlocked(ps, pe, 5, 3); }
    bool CrashDetectionNDObs(const RTpoint &ps, const RTpoint &pe) { return EdgeBlocked(ps, pe, 5, 1); }
    void GetPath(std::list<RTpoint> &path);

    // --- additions (not in the reference) ---
    void SetSampleN(int n) { sample_n_ = n; }
    void RunRounds(int iters, bool with_rewire) { Tick(0, iters, with_rewire ? 7 : 1, 0); }
    int NodeCount() const { return (int)nodes_.size() / 2; }
    const std::vector<float> &NodesXY() const { return nodes_; }
    int ReadTree(int32_t *parent, float *cost, uint8_t *col, uint8_t *state);
    RTpoint Z();
    void SyncSets(); // device state -> Node_Map and the V_* sets
};
#endif
