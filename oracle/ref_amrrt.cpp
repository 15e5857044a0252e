// ref_amrrt.cpp — harness over the reference's OWN AM-RRT* source, compiled in place
// (TEST INFRASTRUCTURE ONLY; output oracle/_ref/libref_amrrt.so; no reference source is copied).
#include "ref_hooks.h"
#include <nlohmann/json.hpp> // the reference's vendored copy (via -I), parsed before the access override
#define private public
#include "AMrrt/AMRRT.cpp" // = /root/reference/AMRRT/lib/AMrrt/AMRRT.cpp via -I
#undef private
#include "ref_common.hpp"
#include "raster.hpp"
#include "pool.hpp"

struct RefAMRRT {
    AMRRTplanner p;
};
REF_EXPORT void *ref_amrrt_create(const uint8_t *obs, int rows, int cols) {
    std::vector<uint8_t> gray((size_t)rows * cols);
    for (size_t i = 0; i < gray.size(); i++) gray[i] = obs[i] ? 0 : 255;
    cv::shim_register_image("ref_map", rows, cols, gray.data());
    cv::shim_circle_hook() = orc_raster::circle_filled;
    cv::shim_fillpoly_hook() = orc_raster::fill_poly;
    RefAMRRT *h = new RefAMRRT();
    h->p.InitializeInvironment("ref_map", 128);
    h->p.DObsmap = cv::Mat::zeros(rows, cols, CV_8U);
    h->p.c_base = 0; // uninitialised in the reference (AMRRT.h:133)
    h->p.diffusion_ratios = 1;
    return h;
}
REF_EXPORT void ref_amrrt_destroy(void *h) { delete (RefAMRRT *)h; }
REF_EXPORT void ref_amrrt_set_agent(void *h, float x, float y) {
    RTpoint p{x, y};
    ((RefAMRRT *)h)->p.SetAgent(p);
}
REF_EXPORT void ref_amrrt_set_goal(void *h, float x, float y) {
    RTpoint p{x, y};
    ((RefAMRRT *)h)->p.SetGoal(p);
}
REF_EXPORT void ref_amrrt_set_dobs(void *h, const uint8_t *plane) {
    AMRRTplanner &p = ((RefAMRRT *)h)->p;
    p.DObsmap = cv::Mat::zeros(p.shapeY, p.shapeX, CV_8U);
    if (plane) memcpy(p.DObsmap.data(), plane, (size_t)p.shapeX * p.shapeY);
}
// fills the private tables InitializeDiffusion (AMRRT.cpp:58-93) would read from JSON
REF_EXPORT void ref_amrrt_set_diffusion(void *h, const int32_t *cells_ij, int ncells, const float *matrix, int dim, float ratio) {
    AMRRTplanner &p = ((RefAMRRT *)h)->p;
    p.diffusion_ratios = ratio;
    p.diffusion_matrix.assign(ncells, std::vector<float>());
    p.index_to_pos.clear();
    for (int i = 0; i < ncells; i++) {
        p.diffusion_matrix[i].assign(matrix + (size_t)i * dim, matrix + (size_t)(i + 1) * dim);
        p.index_to_pos[{cells_ij[2 * i], cells_ij[2 * i + 1]}] = i;
    }
}
REF_EXPORT float ref_amrrt_diffusivity(void *h, float ax, float ay, float bx, float by) {
    RTpoint a{ax, ay}, b{bx, by};
    return ((RefAMRRT *)h)->p.DiffusivityMetric(a, b);
}
REF_EXPORT void ref_amrrt_set_outpath(void *h, int nonempty) {
    AMRRTplanner &p = ((RefAMRRT *)h)->p;
    p.OutPath.clear();
    if (nonempty) p.OutPath.push_back({0, 0});
}
REF_EXPORT float ref_amrrt_c_base(void *h) { return ((RefAMRRT *)h)->p.c_base; }
REF_EXPORT void ref_amrrt_root(void *h, float *xy) { xy[0] = ((RefAMRRT *)h)->p.root.x, xy[1] = ((RefAMRRT *)h)->p.root.y; }
REF_EXPORT void ref_amrrt_update_tree(void *h) { ((RefAMRRT *)h)->p.UpdateTree(); }
REF_EXPORT void ref_amrrt_expand(void *h) { ((RefAMRRT *)h)->p.Expand(); }
REF_EXPORT void ref_amrrt_rewire_root(void *h) { ((RefAMRRT *)h)->p.RewireRoot(); }
// MainTask minus DrawDObsmap (AMRRT.cpp:590-617): the D plane is injected
REF_EXPORT void ref_amrrt_main_task(void *h) {
    AMRRTplanner &p = ((RefAMRRT *)h)->p;
    p.UpdateTree();
    long long t0 = utime_ns();
    while (utime_ns() - t0 < 80e6 && !g_rand_exhausted) {
        p.Expand();
        p.RewireRoot();
        if (p.goal.x >= 0) {
            auto it = p.Node_Map.find(p.goal);
            if (it != p.Node_Map.end() && !(it->second.cost.collision)) p.RewireGoal();
        }
    }
}
REF_EXPORT void ref_amrrt_crash(void *h, const float *xyxy, long long n, uint8_t *out) {
    AMRRTplanner &p = ((RefAMRRT *)h)->p;
    for (long long i = 0; i < n; i++) {
        RTpoint a{xyxy[4 * i], xyxy[4 * i + 1]}, b{xyxy[4 * i + 2], xyxy[4 * i + 3]};
        out[i] = p.CrashDetection(a, b);
    }
}
// the same routine fanned out over host threads (CrashDetection only reads the two planes): persistent pool, chunks of
// 8192 edges dealt through an atomic cursor (pool.hpp)
REF_EXPORT void ref_amrrt_crash_mt(void *h, const float *xyxy, long long n, uint8_t *out, int nthreads) {
    AMRRTplanner *p = &((RefAMRRT *)h)->p;
    orc_pool(nthreads).run(n, 8192, [=](long long a, long long b) {
        for (long long i = a; i < b; i++) {
            RTpoint s{xyxy[4 * i], xyxy[4 * i + 1]}, e{xyxy[4 * i + 2], xyxy[4 * i + 3]};
            out[i] = p->CrashDetection(s, e);
        }
    });
}
REF_EXPORT void ref_amrrt_crash_start(void *h, const float *xyxy, long long n, uint8_t *out, float *pstart) {
    AMRRTplanner &p = ((RefAMRRT *)h)->p;
    for (long long i = 0; i < n; i++) {
        RTpoint a{xyxy[4 * i], xyxy[4 * i + 1]}, b{xyxy[4 * i + 2], xyxy[4 * i + 3]}, s;
        out[i] = p.CrashDetectionStart(a, b, s);
        pstart[2 * i] = s.x, pstart[2 * i + 1] = s.y;
    }
}
REF_EXPORT int ref_amrrt_num_nodes(void *h) { return (int)((RefAMRRT *)h)->p.NeighborSearchTree.pts.size(); }
REF_EXPORT int ref_amrrt_dump(void *h, float *xy, int32_t *parent, float *cost, uint8_t *col) {
    AMRRTplanner &p = ((RefAMRRT *)h)->p;
    std::map<RTpoint, int> id;
    int n = (int)p.NeighborSearchTree.pts.size();
    for (int i = 0; i < n; i++) id[{(float)p.NeighborSearchTree.pts[i].x(), (float)p.NeighborSearchTree.pts[i].y()}] = i;
    for (int i = 0; i < n; i++) {
        RTpoint q{(float)p.NeighborSearchTree.pts[i].x(), (float)p.NeighborSearchTree.pts[i].y()};
        const RTrrt_NodeAttribute &a = p.Node_Map[q];
        xy[2 * i] = q.x, xy[2 * i + 1] = q.y;
        auto it = id.find(a.par);
        parent[i] = (a.par.x < 0 || it == id.end()) ? -1 : it->second;
        cost[i] = a.cost.cost;
        col[i] = a.cost.collision;
    }
    return n;
}
REF_EXPORT int ref_amrrt_get_path(void *h, float *out_xy, int cap, int *collision, float *cost) {
    std::list<RTpoint> path;
    RTrrt_Cost c = ((RefAMRRT *)h)->p.GetPath(path);
    *collision = c.collision, *cost = c.cost;
    int n = 0;
    for (const RTpoint &q : path) {
        if (n < cap) out_xy[2 * n] = q.x, out_xy[2 * n + 1] = q.y;
        n++;
    }
    return n;
}
