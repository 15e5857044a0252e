// rtpp_oracle_capi.cpp — extern "C" surface of the CPU ORACLE (TEST INFRASTRUCTURE ONLY; see
// rtpp_oracle.hpp).  Loaded with ctypes by tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs; never by the product path.
#include "rtpp_oracle.hpp"
#include "raster.hpp"
#include "pool.hpp"
#include <cstring>
#include <thread>
#include <atomic>
#include <mutex>

using namespace orc;
#define ORC_EXPORT extern "C" __attribute__((visibility("default")))

// ---------------------------------------------------------------- primitives (bulk) ------------
static Map mk_map(const uint8_t *S, const uint8_t *D, int W, int H) {
    Map m;
    m.S = S, m.D = D, m.W = W, m.H = H;
    return m;
}

// batches are dealt in chunks to a persistent pool (pool.hpp): no thread start-up inside a timed call
template <class F> static void par_for(long long n, int nthreads, F f) {
    if (nthreads <= 1 || n < 4096) {
        f(0, n);
        return;
    }
    orc_pool(nthreads).run(n, 8192, f);
}

ORC_EXPORT void orc_edges_check_f32(const uint8_t *S, const uint8_t *D, int W, int H, const float *xyxy, long long n,
                                    int divisor, int mask, uint8_t *out, int nthreads) {
    Map m = mk_map(S, D, W, H);
    par_for(n, nthreads, [=](long long a, long long b) {
        for (long long i = a; i < b; i++)
            out[i] = edge_blocked(m, {xyxy[4 * i], xyxy[4 * i + 1]}, {xyxy[4 * i + 2], xyxy[4 * i + 3]}, divisor, mask);
    });
}
ORC_EXPORT void orc_edges_first_hit_f32(const uint8_t *S, const uint8_t *D, int W, int H, const float *xyxy,
                                        long long n, uint8_t *out, float *pstart, int nthreads) {
    Map m = mk_map(S, D, W, H);
    par_for(n, nthreads, [=](long long a, long long b) {
        for (long long i = a; i < b; i++) {
            Pt ps;
            out[i] = edge_first_hit(m, {xyxy[4 * i], xyxy[4 * i + 1]}, {xyxy[4 * i + 2], xyxy[4 * i + 3]}, &ps);
            pstart[2 * i] = ps.x, pstart[2 * i + 1] = ps.y;
        }
    });
}
ORC_EXPORT void orc_edges_check_f64(const uint8_t *S, const uint8_t *D, int W, int H, const double *xyxy, long long n,
                                    int divisor, int mask, uint8_t *out, int nthreads) {
    Map m = mk_map(S, D, W, H);
    par_for(n, nthreads, [=](long long a, long long b) {
        for (long long i = a; i < b; i++)
            out[i] = edge_blocked_f64(m, xyxy[4 * i], xyxy[4 * i + 1], xyxy[4 * i + 2], xyxy[4 * i + 3], divisor, mask);
    });
}
ORC_EXPORT void orc_edges_nprobes(const float *xyxy, long long n, int divisor, int32_t *out) {
    for (long long i = 0; i < n; i++)
        out[i] = nprobes_for_len(dist({xyxy[4 * i], xyxy[4 * i + 1]}, {xyxy[4 * i + 2], xyxy[4 * i + 3]}), divisor);
}
ORC_EXPORT long long orc_edges_nprobes_sum(const float *xyxy, long long n, int divisor) {
    long long s = 0;
    for (long long i = 0; i < n; i++)
        s += nprobes_for_len(dist({xyxy[4 * i], xyxy[4 * i + 1]}, {xyxy[4 * i + 2], xyxy[4 * i + 3]}), divisor);
    return s;
}
ORC_EXPORT void orc_dist_f32(const float *xyxy, long long n, float *out) {
    for (long long i = 0; i < n; i++) out[i] = dist({xyxy[4 * i], xyxy[4 * i + 1]}, {xyxy[4 * i + 2], xyxy[4 * i + 3]});
}
ORC_EXPORT void orc_rand31_fill(uint64_t seed, long long i0, long long n, int32_t *out) {
    for (long long i = 0; i < n; i++) out[i] = rand31(seed, (uint64_t)(i0 + i));
}

// k-NN.  pts: n float (x,y) pairs (widened exactly to double, as the reference inserts float
// coordinates into a double kd-tree: RTRRT.cpp:29,291).  q: nq double pairs.  Output rows of k:
// idx (-1 padded), d2 (+inf padded), count = min(k, n).
ORC_EXPORT void orc_knn_brute(const float *pts, int n, const double *q, long long nq, int k, int32_t *idx, double *d2,
                              int32_t *count, int nthreads) {
    std::vector<double> p(2 * (size_t)n);
    for (size_t i = 0; i < 2 * (size_t)n; i++) p[i] = (double)pts[i];
    const double *pp = p.data();
    par_for(nq, nthreads, [=](long long a, long long b) {
        std::vector<Nbr> out;
        for (long long i = a; i < b; i++) {
            knn_brute(pp, n, q[2 * i], q[2 * i + 1], k, out);
            for (int j = 0; j < k; j++) {
                idx[i * k + j] = j < (int)out.size() ? out[j].id : -1;
                d2[i * k + j] = j < (int)out.size() ? out[j].d2 : HUGE_VAL;
            }
            if (count) count[i] = (int)out.size();
        }
    });
}
ORC_EXPORT void orc_knn_grid(const float *pts, int n, double W, double H, double cell, const double *q, long long nq,
                             int k, int32_t *idx, double *d2, int32_t *count, int nthreads) {
    KnnIndex ix;
    ix.init(W, H, cell);
    for (int i = 0; i < n; i++) ix.insert((double)pts[2 * i], (double)pts[2 * i + 1]);
    const KnnIndex *pix = &ix;
    par_for(nq, nthreads, [=](long long a, long long b) {
        std::vector<Nbr> out;
        for (long long i = a; i < b; i++) {
            pix->query(q[2 * i], q[2 * i + 1], k, out);
            for (int j = 0; j < k; j++) {
                idx[i * k + j] = j < (int)out.size() ? out[j].id : -1;
                d2[i * k + j] = j < (int)out.size() ? out[j].d2 : HUGE_VAL;
            }
            if (count) count[i] = (int)out.size();
        }
    });
}

// The same queries through the kd-tree form of the index (sliding-midpoint, bucket 10, built once here); for the equality test
// against orc_knn_brute.  Single-threaded on purpose: the tree is built lazily on the first query.
ORC_EXPORT void orc_knn_kdtree(const float *pts, int n, double W, double H, const double *q, long long nq, int k, int32_t *idx,
                               double *d2, int32_t *count) {
    KnnIndex ix;
    ix.init(W, H, 64);
    for (int i = 0; i < n; i++) ix.insert((double)pts[2 * i], (double)pts[2 * i + 1]);
    ix.kd_rebuild(true);
    std::vector<Nbr> out;
    for (long long i = 0; i < nq; i++) {
        ix.query(q[2 * i], q[2 * i + 1], k, out);
        for (int j = 0; j < k; j++) {
            idx[i * k + j] = j < (int)out.size() ? out[j].id : -1;
            d2[i * k + j] = j < (int)out.size() ? out[j].d2 : HUGE_VAL;
        }
        if (count) count[i] = (int)out.size();
    }
}

// R3 on flat arrays: BFS from root over the children implied by parent[]; every tree edge gets
// cost[c] = cost[p] + {E(p,c), |p-c|} (RTRRT.cpp:121-135).  Unreached nodes keep their input
// cost/col.  Returns the number of nodes reached.
ORC_EXPORT int orc_tree_propagate(const uint8_t *S, const uint8_t *D, int W, int H, const float *xy,
                                  const int32_t *parent, int n, int root, int divisor, int mask, float *cost,
                                  uint8_t *col) {
    Map m = mk_map(S, D, W, H);
    std::vector<std::vector<int>> kids(n);
    for (int i = 0; i < n; i++)
        if (i != root && parent[i] >= 0 && parent[i] < n) kids[parent[i]].push_back(i);
    std::queue<int> Q;
    Q.push(root);
    int reached = 0;
    while (!Q.empty()) {
        int p = Q.front();
        Q.pop();
        reached++;
        Pt pp{xy[2 * p], xy[2 * p + 1]};
        for (int c : kids[p]) {
            Pt pc{xy[2 * c], xy[2 * c + 1]};
            Cost cc = Cost{col[p] != 0, cost[p]} + Cost{edge_blocked(m, pp, pc, divisor, mask), dist(pp, pc)};
            cost[c] = cc.c, col[c] = cc.col;
            Q.push(c);
        }
    }
    return reached;
}

// R2 fixpoint on flat arrays: repeat { for x in order: for s in knn[x]: relax(x -> s) } until no
// change — the least fixpoint the wave-synchronous atomicMin kernel must also reach.  The relax
// is RTRRT.cpp:305-319 (strict <, lexicographic).  Ties on cost keep the incumbent parent, so
// parents are compared only where costs are unique (tests do that).
// free_only = 1: AM-RRT* RewireRoot's rule (AMRRT.cpp:500-504): only over free edges, never to a colliding cost.
ORC_EXPORT int orc_relax_fixpoint(const uint8_t *S, const uint8_t *D, int W, int H, const float *xy, int n,
                                  const int32_t *knn_idx, int k, int divisor, int mask, float *cost, uint8_t *col,
                                  int32_t *parent, int max_sweeps, int free_only) {
    Map m = mk_map(S, D, W, H);
    int sweeps = 0;
    for (; sweeps < max_sweeps; sweeps++) {
        bool changed = false;
        for (int x = 0; x < n; x++) {
            Pt px{xy[2 * x], xy[2 * x + 1]};
            for (int j = 0; j < k; j++) {
                int s = knn_idx[(size_t)x * k + j];
                if (s < 0) continue;
                Pt ps{xy[2 * s], xy[2 * s + 1]};
                if (s == x) continue;
                const bool blocked = edge_blocked(m, px, ps, divisor, mask);
                if (free_only && blocked) continue;
                Cost c_new = Cost{col[x] != 0, cost[x]} + Cost{blocked, dist(px, ps)};
                Cost c_old{col[s] != 0, cost[s]};
                if (free_only && c_new.col) continue;
                if (c_new < c_old) {
                    cost[s] = c_new.c, col[s] = c_new.col, parent[s] = x;
                    changed = true;
                }
            }
        }
        if (!changed) break;
    }
    return sweeps;
}

// D1 rasteriser
ORC_EXPORT void orc_raster_circle(uint8_t *img, int rows, int cols, int cx, int cy, int radius, int value) {
    orc_raster::circle_filled(img, rows, cols, cx, cy, radius, (uint8_t)value);
}
ORC_EXPORT void orc_raster_fillpoly(uint8_t *img, int rows, int cols, const int *xy, int npts, int value) {
    orc_raster::fill_poly(img, rows, cols, xy, npts, (uint8_t)value);
}

// ---------------------------------------------------------------- RT-RRT* ----------------------
struct HRTRRT {
    RTRRT p;
    std::vector<uint8_t> S, D;
};
ORC_EXPORT void *orc_rtrrt_create(const uint8_t *S, int rows, int cols, int divisor) {
    HRTRRT *h = new HRTRRT();
    h->S.assign(S, S + (size_t)rows * cols);
    h->D.assign((size_t)rows * cols, 0);
    h->p.init(mk_map(h->S.data(), h->D.data(), cols, rows));
    h->p.divisor = divisor;
    return h;
}
ORC_EXPORT void orc_rtrrt_destroy(void *h) { delete (HRTRRT *)h; }
// mode 1: neighbour queries from a kd-tree rebuilt on the first query after every insert, as CGAL's Kd_tree does (same answers as
// the default grid; the cost profile of the reference).  Returns the number of rebuilds so far.
ORC_EXPORT long long orc_rtrrt_knn_mode(void *h, int mode) {
    KnnIndex &ix = ((HRTRRT *)h)->p.index;
    if (mode >= 0) ix.kd_rebuild(mode == 1);
    return ix.kd_builds();
}
ORC_EXPORT void orc_rtrrt_set_clock(void *h, long long dt_ns, long long n0) {
    ((HRTRRT *)h)->p.clk.dt_ns = dt_ns, ((HRTRRT *)h)->p.clk.n = n0;
}
ORC_EXPORT long long orc_rtrrt_clock_reads(void *h) { return ((HRTRRT *)h)->p.clk.n; }
ORC_EXPORT int orc_rtrrt_set_agent(void *h, float x, float y) { return ((HRTRRT *)h)->p.SetAgent({x, y}); }
ORC_EXPORT int orc_rtrrt_set_goal(void *h, float x, float y) { return ((HRTRRT *)h)->p.SetGoal({x, y}); }
ORC_EXPORT void orc_rtrrt_set_dobs(void *h, const uint8_t *plane) {
    HRTRRT *H = (HRTRRT *)h;
    if (plane) memcpy(H->D.data(), plane, H->D.size());
    else std::fill(H->D.begin(), H->D.end(), 0);
}
ORC_EXPORT void orc_rtrrt_update_tree(void *h) { ((HRTRRT *)h)->p.UpdateTree(); }
// draws up to n_max sample points from the rand stream exactly as RTRRT.cpp:234-262 would;
// offsets[i] = rand draws consumed before sample i (offsets has n_max+1 entries).
ORC_EXPORT int orc_rtrrt_presample(void *h, const int32_t *r, long long nr, int n_max, float *pts, long long *offsets) {
    RTRRT &p = ((HRTRRT *)h)->p;
    RandStream rs;
    rs.r = r, rs.n = nr;
    int i = 0;
    for (; i < n_max; i++) {
        offsets[i] = rs.pos;
        Pt s = p.sample(rs);
        if (rs.exhausted) break;
        pts[2 * i] = s.x, pts[2 * i + 1] = s.y;
    }
    offsets[i] = rs.pos;
    return i;
}
ORC_EXPORT int orc_rtrrt_expand_rewiring(void *h, long ms, const float *pts, int n) {
    std::vector<Pt> s(n);
    for (int i = 0; i < n; i++) s[i] = {pts[2 * i], pts[2 * i + 1]};
    return ((HRTRRT *)h)->p.ExpandRewiring(ms, s.data(), n);
}
ORC_EXPORT int orc_rtrrt_num_nodes(void *h) { return (int)((HRTRRT *)h)->p.by_id.size(); }
ORC_EXPORT int orc_rtrrt_dump(void *h, float *xy, int32_t *parent, float *cost, uint8_t *col) {
    RTRRT &p = ((HRTRRT *)h)->p;
    int n = (int)p.by_id.size();
    for (int i = 0; i < n; i++) {
        const RTRRT::Node &nd = p.nodes[p.by_id[i]];
        xy[2 * i] = p.by_id[i].x, xy[2 * i + 1] = p.by_id[i].y;
        auto it = p.nodes.find(nd.par);
        parent[i] = (nd.par.x < 0 || it == p.nodes.end()) ? -1 : it->second.id;
        cost[i] = nd.cost.c, col[i] = nd.cost.col;
    }
    return n;
}
ORC_EXPORT void orc_rtrrt_root(void *h, float *xy) { xy[0] = ((HRTRRT *)h)->p.root.x, xy[1] = ((HRTRRT *)h)->p.root.y; }
ORC_EXPORT void orc_rtrrt_queue_sizes(void *h, int *qr, int *qs) {
    *qr = (int)((HRTRRT *)h)->p.Qr.size(), *qs = (int)((HRTRRT *)h)->p.Qs.size();
}
ORC_EXPORT void orc_rtrrt_counters(void *h, long long *out5) {
    RTRRT &p = ((HRTRRT *)h)->p;
    out5[0] = p.n_iters, out5[1] = p.n_pops_r, out5[2] = p.n_pops_s, out5[3] = p.n_edge_checks, out5[4] = p.n_knn;
}
ORC_EXPORT int orc_rtrrt_get_path(void *h, float *out_xy, int cap, int *collision, float *cost, float *c_base) {
    RTRRT &p = ((HRTRRT *)h)->p;
    std::list<Pt> path;
    Cost c = p.GetPath(path);
    *collision = c.col, *cost = c.c, *c_base = p.c_base;
    int n = 0;
    for (const Pt &q : path) {
        if (n < cap) out_xy[2 * n] = q.x, out_xy[2 * n + 1] = q.y;
        n++;
    }
    return n;
}

// ---------------------------------------------------------------- RT-FMT -----------------------
struct HRTFMT {
    RTFMT p;
    std::vector<uint8_t> S, D;
};
ORC_EXPORT void *orc_rtfmt_create(const uint8_t *S, int rows, int cols) {
    HRTFMT *h = new HRTFMT();
    h->S.assign(S, S + (size_t)rows * cols);
    h->D.assign((size_t)rows * cols, 0);
    h->p.init(mk_map(h->S.data(), h->D.data(), cols, rows));
    return h;
}
ORC_EXPORT void orc_rtfmt_destroy(void *h) { delete (HRTFMT *)h; }
ORC_EXPORT int orc_rtfmt_set_agent(void *h, float x, float y) { return ((HRTFMT *)h)->p.SetAgent({x, y}); }
ORC_EXPORT int orc_rtfmt_set_goal(void *h, float x, float y) { return ((HRTFMT *)h)->p.SetGoal({x, y}); }
ORC_EXPORT void orc_rtfmt_set_dobs(void *h, const uint8_t *plane) {
    HRTFMT *H = (HRTFMT *)h;
    if (plane) memcpy(H->D.data(), plane, H->D.size());
    else std::fill(H->D.begin(), H->D.end(), 0);
}
ORC_EXPORT long long orc_rtfmt_start(void *h, const int32_t *r, long long nr, int sample_n) {
    RandStream rs;
    rs.r = r, rs.n = nr;
    ((HRTFMT *)h)->p.Start(rs, sample_n);
    return rs.pos;
}
ORC_EXPORT void orc_rtfmt_update_tree(void *h) { ((HRTFMT *)h)->p.UpdateTree(); }
ORC_EXPORT void orc_rtfmt_expand(void *h, int iters, int with_rewire) {
    RTFMT &p = ((HRTFMT *)h)->p;
    for (int i = 0; i < iters; i++) {
        p.ExpandFMT();
        if (with_rewire) {
            p.RewireFromObstacles();
            p.RewireFromRoot();
        }
    }
}
ORC_EXPORT int orc_rtfmt_num_nodes(void *h) { return (int)((HRTFMT *)h)->p.by_id.size(); }
// state: 0 unvisited, 1 open, 2 opennew, 3 closed (+4 if also in V_toopen)
ORC_EXPORT int orc_rtfmt_dump(void *h, float *xy, int32_t *parent, float *cost, uint8_t *col, uint8_t *state) {
    RTFMT &p = ((HRTFMT *)h)->p;
    int n = (int)p.by_id.size();
    for (int i = 0; i < n; i++) {
        const Pt &q = p.by_id[i];
        const RTFMT::Node &nd = p.nodes[q];
        xy[2 * i] = q.x, xy[2 * i + 1] = q.y;
        auto it = p.nodes.find(nd.par);
        parent[i] = (nd.par.x < 0 || it == p.nodes.end()) ? -1 : it->second.id;
        cost[i] = nd.cost.c, col[i] = nd.cost.col;
        uint8_t s = 0;
        if (p.V_open.count(q)) s = 1;
        else if (p.V_opennew.count(q)) s = 2;
        else if (p.V_closed.count(q)) s = 3;
        if (p.V_toopen.count(q)) s |= 4;
        state[i] = s;
    }
    return n;
}
ORC_EXPORT void orc_rtfmt_z(void *h, float *xy) { xy[0] = ((HRTFMT *)h)->p.z.x, xy[1] = ((HRTFMT *)h)->p.z.y; }
ORC_EXPORT void orc_rtfmt_counters(void *h, long long *out3) {
    RTFMT &p = ((HRTFMT *)h)->p;
    out3[0] = p.n_expand, out3[1] = p.n_edge_checks, out3[2] = p.n_knn;
}
ORC_EXPORT int orc_rtfmt_get_path(void *h, float *out_xy, int cap) {
    std::list<Pt> path;
    ((HRTFMT *)h)->p.GetPath(path);
    int n = 0;
    for (const Pt &q : path) {
        if (n < cap) out_xy[2 * n] = q.x, out_xy[2 * n + 1] = q.y;
        n++;
    }
    return n;
}

// ---------------------------------------------------------------- AM-RRT* ----------------------
struct HAMRRT {
    AMRRT p;
    std::vector<uint8_t> S, D;
    std::vector<int32_t> rbuf;
    RandStream rs;
};
ORC_EXPORT void *orc_amrrt_create(const uint8_t *S, int rows, int cols) {
    HAMRRT *h = new HAMRRT();
    h->S.assign(S, S + (size_t)rows * cols);
    h->D.assign((size_t)rows * cols, 0);
    h->p.init(mk_map(h->S.data(), h->D.data(), cols, rows));
    return h;
}
ORC_EXPORT void orc_amrrt_destroy(void *h) { delete (HAMRRT *)h; }
ORC_EXPORT void orc_amrrt_set_clock(void *h, long long dt_ns, long long n0) {
    ((HAMRRT *)h)->p.clk.dt_ns = dt_ns, ((HAMRRT *)h)->p.clk.n = n0;
}
ORC_EXPORT void orc_amrrt_set_rand(void *h, const int32_t *r, long long n) {
    HAMRRT *H = (HAMRRT *)h;
    H->rbuf.assign(r, r + n);
    H->rs = RandStream();
    H->rs.r = H->rbuf.data(), H->rs.n = n;
}
ORC_EXPORT long long orc_amrrt_rand_pos(void *h) { return ((HAMRRT *)h)->rs.pos; }
ORC_EXPORT int orc_amrrt_set_agent(void *h, float x, float y) { return ((HAMRRT *)h)->p.SetAgent({x, y}); }
ORC_EXPORT int orc_amrrt_set_goal(void *h, float x, float y) { return ((HAMRRT *)h)->p.SetGoal({x, y}); }
ORC_EXPORT void orc_amrrt_set_dobs(void *h, const uint8_t *plane) {
    HAMRRT *H = (HAMRRT *)h;
    if (plane) memcpy(H->D.data(), plane, H->D.size());
    else std::fill(H->D.begin(), H->D.end(), 0);
}
// diffusion table: cells (i,j) -> row index; rows x dim floats (AMRRT.cpp:58-93)
ORC_EXPORT void orc_amrrt_set_diffusion(void *h, const int32_t *cells_ij, int ncells, const float *matrix, int dim,
                                        float ratio) {
    AMRRT &p = ((HAMRRT *)h)->p;
    p.index_to_pos.clear();
    for (int i = 0; i < ncells; i++) p.index_to_pos[{cells_ij[2 * i], cells_ij[2 * i + 1]}] = i;
    p.diffusion.assign(matrix, matrix + (size_t)ncells * dim);
    p.diff_dim = dim;
    p.diffusion_ratios = ratio;
}
ORC_EXPORT float orc_amrrt_diffusivity(void *h, float ax, float ay, float bx, float by) {
    return ((HAMRRT *)h)->p.DiffusivityMetric({ax, ay}, {bx, by});
}
// OutPath is only ever filled by the caller (AMRRT/test/testD.cpp passes planner.OutPath to GetPath);
// non-empty switches the sampler to the informed ellipse (AMRRT.cpp:402-404)
ORC_EXPORT void orc_amrrt_set_outpath(void *h, int nonempty) {
    AMRRT &p = ((HAMRRT *)h)->p;
    p.OutPath.clear();
    if (nonempty) p.OutPath.push_back({0, 0});
}
ORC_EXPORT float orc_amrrt_c_base(void *h) { return ((HAMRRT *)h)->p.c_base; }
ORC_EXPORT void orc_amrrt_root(void *h, float *xy) { xy[0] = ((HAMRRT *)h)->p.root.x, xy[1] = ((HAMRRT *)h)->p.root.y; }
ORC_EXPORT void orc_amrrt_update_tree(void *h) { ((HAMRRT *)h)->p.UpdateTree(); }
ORC_EXPORT void orc_amrrt_expand(void *h) { ((HAMRRT *)h)->p.Expand(((HAMRRT *)h)->rs); }
ORC_EXPORT void orc_amrrt_rewire_root(void *h) { ((HAMRRT *)h)->p.RewireRoot(); }
ORC_EXPORT void orc_amrrt_main_task(void *h) { ((HAMRRT *)h)->p.MainTask(((HAMRRT *)h)->rs); }
ORC_EXPORT int orc_amrrt_num_nodes(void *h) { return (int)((HAMRRT *)h)->p.by_id.size(); }
ORC_EXPORT int orc_amrrt_dump(void *h, float *xy, int32_t *parent, float *cost, uint8_t *col) {
    AMRRT &p = ((HAMRRT *)h)->p;
    int n = (int)p.by_id.size();
    for (int i = 0; i < n; i++) {
        const AMRRT::Node &nd = p.nodes[p.by_id[i]];
        xy[2 * i] = p.by_id[i].x, xy[2 * i + 1] = p.by_id[i].y;
        auto it = p.nodes.find(nd.par);
        parent[i] = (nd.par.x < 0 || it == p.nodes.end()) ? -1 : it->second.id;
        cost[i] = nd.cost.c, col[i] = nd.cost.col;
    }
    return n;
}
ORC_EXPORT int orc_amrrt_get_path(void *h, float *out_xy, int cap, int *collision, float *cost) {
    std::list<Pt> path;
    Cost c = ((HAMRRT *)h)->p.GetPath(path);
    *collision = c.col, *cost = c.c;
    int n = 0;
    for (const Pt &q : path) {
        if (n < cap) out_xy[2 * n] = q.x, out_xy[2 * n + 1] = q.y;
        n++;
    }
    return n;
}

// ---------------------------------------------------------------- glibc transcendentals ---------
// Checker for rtpp_b200/csrc/glibc_math.cuh (the restatement of glibc's sinf / cosf / atanf / atan2f that the device samplers
// use): the same checksum computed with the host's own libm and with the header's host instantiation.
#include <math.h>
#include "../rtpp_b200/csrc/glibc_math.cuh"
static inline uint32_t orc_canon(float r) {
    uint32_t b;
    memcpy(&b, &r, 4);
    return r != r ? 0x7fc00000u : b;
}
template <class F> static unsigned long long orc_checksum(uint32_t first, unsigned long long count, int nthreads, F f) {
    std::vector<unsigned long long> part((size_t)std::max(nthreads, 1) * 64 + 64, 0);
    std::atomic<int> slot{0};
    par_for((long long)count, nthreads, [&](long long a, long long b) {
        unsigned long long acc = 0;
        for (long long i = a; i < b; i++) {
            const uint32_t bits = first + (uint32_t)i;
            float x;
            memcpy(&x, &bits, 4);
            acc += (unsigned long long)orc_canon(f(x)) * (2ull * bits + 1ull);
        }
        static std::mutex mu;
        std::lock_guard<std::mutex> lk(mu);
        part[0] += acc;
    });
    (void)slot;
    return part[0];
}
ORC_EXPORT int orc_glibc_host_fma(void) {
    volatile float x;
    const uint32_t xb = 0x4255b0a9u;
    memcpy((void *)&x, &xb, 4);
    const float r = sinf(x);
    uint32_t rb;
    memcpy(&rb, &r, 4);
    return rb == 0xbc7d08a9u ? 1 : 0;
}
// fn: 0 sinf, 1 cosf, 2 atanf.  impl: 0 = the host's libm, 1 = glibc_math.cuh without fused operations, 2 = with
ORC_EXPORT unsigned long long orc_glibc_checksum(int impl, int fn, uint32_t first, unsigned long long count, int nthreads) {
    namespace gm = glibc_math;
    if (impl == 0) {
        if (fn == 0) return orc_checksum(first, count, nthreads, [](float x) { return sinf(x); });
        if (fn == 1) return orc_checksum(first, count, nthreads, [](float x) { return cosf(x); });
        return orc_checksum(first, count, nthreads, [](float x) { return atanf(x); });
    }
    if (fn == 2) return orc_checksum(first, count, nthreads, [](float x) { return gm::atanf_(x); });
    if (impl == 2) {
        if (fn == 0) return orc_checksum(first, count, nthreads, [](float x) { return gm::sinf_<true>(x); });
        return orc_checksum(first, count, nthreads, [](float x) { return gm::cosf_<true>(x); });
    }
    if (fn == 0) return orc_checksum(first, count, nthreads, [](float x) { return gm::sinf_<false>(x); });
    return orc_checksum(first, count, nthreads, [](float x) { return gm::cosf_<false>(x); });
}
// out[i] = atan2f(yx[2i], yx[2i+1]): impl 0 = libm, 1 = glibc_math.cuh
ORC_EXPORT void orc_glibc_atan2f(int impl, const float *yx, long long n, float *out) {
    for (long long i = 0; i < n; i++) out[i] = impl ? glibc_math::atan2f_(yx[2 * i], yx[2 * i + 1]) : atan2f(yx[2 * i], yx[2 * i + 1]);
}
