// engine_cdrt.cu — sequential-exact CDRT-RRT* replanning engine: one warp runs one BImap planner.
//
// Replaces, behind rtpp_cdrt_* (include/rtpp_gpu.h), the RTRT_* members of the reference's BImap
// (CDRT2/lib/CDRT/CDRT.cpp):
//   RTRT_start            :276-383  -> rtpp_cdrt_start      (host: one-off set-up, then upload)
//   RTRT_AddDynamicObstacle :425-500, RTRT_SetAgent/SetGoal :384-423 -> host members of the engine
//   RTRT_DrawDObsmap      :605-636  -> per-instance bit-tile plane drawn by raster.cu
//   RTRT_task             :927-1068 -> cdrt_tick_kernel:
//       RTRT_CrashDObsIsland :638-762, the agent hand-over :946-996, RTRT_UpdateTreeCost :764-796,
//       RTRT_AddPotentialLinks :797-830, RTRT_Sampling :832-925, the eQr rewiring loop :1036-1067
//   RTRT_Neighbor         :177-274  -> cd_neighbors   (N2 per island, k = nearK)
//   RTRT_CrashDetection   :507-603  -> cd_crash       (E5a topology switch + E5b fp64 march)
//   GetPath               :1076-1123 -> cd_get_path
//
// Data layout (per instance, node ids instead of the reference's point keys):
//   xy double2, cost word (collision<<63 | binary64 bits: unsigned order == CTrtrt_Cost order,
//   CDRT.h:92-131), parent id, invM, flags.  Children sets are not stored: a node's children are
//   the live nodes whose parent it is (every place the reference edits `subs` keeps that invariant),
//   and RTRT_UpdateTreeCost rebuilds a CSR of them.  Mnode[cutline] = a slice kept sorted by (x, y)
//   (std::set order: it is the tie-break order of choose-parent); obsKdtreeMap[island] = a slice of
//   ids in insertion order scanned by the warp with an exact top-k (binary64 d2, ties by insertion
//   order, the oracle's rule).  eQr = a 32-ary min-heap on (cost word, S point, E point): one
//   coalesced 32-child load + warp arg-min per level (4 levels at 10^6 entries); its order is total,
//   so any heap pops the same sequence as std::priority_queue.
//   largerSet / smallerSet = flag bit + list.
// Wall-clock gates use the virtual clock (k-th read = k*dt); rand() is the injected stream.
// A re-sampled point bit-identical to an erased node would alias it in the reference's point-keyed
// map; ids here are never reused (probability < 1e-12 per sample; documented in DESIGN.md).
#include <math.h>
#include <math_constants.h>

#include <algorithm>
#include <map>
#include <queue>
#include <set>
#include <vector>

#include "common.cuh"

#define FULL 0xffffffffu
typedef unsigned long long u64;

static constexpr int kOrigin = 0x40000000;        // originIndexFun, CDRT.h:226
static constexpr double kDoubleMax = 1.79769e+307; // CDRT.h:359
enum { F_ALIVE = 1, F_LARGER = 2, F_SMALLER = 4 };
enum { ST_OK = 0, ST_NODES_FULL = 1, ST_QUEUE_FULL = 2, ST_ISLAND_FULL = 3, ST_REF_UB = 4, ST_CUT_FULL = 5, ST_SET_FULL = 6 };

struct CdInst {
    int n, agent_id, agent_island, goal_set;
    double2 agent_tag, goal;
    long long clk_n, update_old, smaller_old;
    int timers_init, samp_pos, samp_n, rand_pos, rand_n, rand_exhausted;
    int q_n, n_larger, n_smaller, n_trees, status, epoch;
    long long pops, rewired, sampled, pushes, checks;
    int path_len, path_col;
    double path_cost;
    int n_obs, n_chk;
    long long t_phase[4]; // SM cycles of the last tick: environment + tree update | larger/smaller sets | sampling | eQr loop
    long long t_fine[8];  // eQr loop of the last tick: cycles in pop | entry filter | edge check | neighbour list | pushes; edge checks, neighbours listed, rewires
};

struct CdHdr { // per-tick inputs of one instance
    double2 tag, goal;
    int goal_set, n_obs, n_chk, started;
};

struct CdView {
    int n_inst, cap, qcap, ncut, npoly, W, H, nearK, RewireN, max_obs, max_chk, path_cap, rand_cap, set_cap;
    double r_th, p_th, gate2;
    const int *cp_off, *cp, *cc_off, *cc, *pc_off, *pc; // cut->polys, cut->cuts, poly->cuts (ascending)
    const int *cut_cap_off, *isl_cap_off;              // slice offsets [ncut+1], [npoly+1]
    int cut_cap_total, isl_cap_total;
    const uint16_t *isl; // island id where BIamap == agent_apl, else 0xFFFF
    CdInst *inst;
    double2 *xy;
    u64 *cw;
    int *par, *inv;
    unsigned char *flag;
    int *cut_cnt, *cut_nodes, *isl_cnt, *isl_nodes, *isl_tree, *isl_mark;
    u64 *qkey;
    int2 *qse;
    int *larger, *smaller, *xnear, *child_start, *children, *frontier, *path;
    const int *rand;
    double4 *pose; // x, y, sinf(angle), cosf(angle) (binary32 values widened)
    double2 *chk;
    int *chk_obs;
    uint32_t *dplane;
    size_t plane_words;
    int tilesX;
    const CdHdr *hdr;
};

struct CdI { // per-instance pointers
    CdInst *g;
    double2 *xy;
    u64 *cw;
    int *par, *inv;
    unsigned char *flag;
    int *cut_cnt, *cut_nodes, *isl_cnt, *isl_nodes, *isl_tree, *isl_mark;
    u64 *qkey;
    int2 *qse;
    int *larger, *smaller, *xnear, *child_start, *children, *frontier, *path;
    const int *rand;
    PlaneView dpl;
    int agent_id, agent_island, n, q_n, status;
    long long clk_n, dt;
    int rand_pos, rand_n, samp_pos, samp_n, n_trees;
};

// ---------------------------------------------------------------------------------------------
// cost words
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 cw_make(bool col, double c) { return ((u64)col << 63) | ((u64)__double_as_longlong(c) & 0x7fffffffffffffffull); }
__device__ __forceinline__ double cw_cost(u64 w) { return __longlong_as_double((long long)(w & 0x7fffffffffffffffull)); }
// CTrtrt_Cost::operator+ (CDRT.h:97-100)
__device__ __forceinline__ u64 cw_add(u64 a, bool col, double c) { return ((a >> 63) | (u64)col) << 63 | ((u64)__double_as_longlong(__dadd_rn(cw_cost(a), c)) & 0x7fffffffffffffffull); }
__device__ __forceinline__ u64 warp_min_u64(u64 x) {
    const unsigned hi = (unsigned)(x >> 32);
    const unsigned mh = __reduce_min_sync(FULL, hi);
    const unsigned lo = hi == mh ? (unsigned)x : 0xffffffffu;
    const unsigned ml = __reduce_min_sync(FULL, lo);
    return ((u64)mh << 32) | ml;
}
__device__ __forceinline__ long long cd_now(CdI &in) { return in.dt * (in.clk_n++); }

// cv::Mat::at<uint16_t>(row, col) on the id plane (row wrap inside the buffer; outside = 0xFFFF)
__device__ __forceinline__ int cd_island_at(const CdView &v, int row, int col) {
    const long long lin = (long long)row * v.W + col;
    if (lin < 0 || lin >= (long long)v.W * v.H) return 0xFFFF;
    return v.isl[lin];
}
__device__ __forceinline__ bool csr_has(const int *off, const int *val, int row, int x) {
    for (int j = off[row]; j < off[row + 1]; j++)
        if (val[j] == x) return true;
    return false;
}

// E5b, CDRT.cpp:584-601
__device__ __forceinline__ bool cd_probe_interior(const PlaneView &p, double psx, double psy, double pex, double pey, float t) {
    const float omt = __fsub_rn(1.0f, t);
    const double td = (double)t, od = (double)omt;
    const int x = __double2int_rz(__dadd_rn(__dmul_rn(psx, td), __dmul_rn(od, pex)));
    const int y = __double2int_rz(__dadd_rn(__dmul_rn(psy, td), __dmul_rn(od, pey)));
    return rt_probe(p, y, x);
}
__device__ bool cd_march(const PlaneView &p, double psx, double psy, double pex, double pey) {
    if (rt_probe(p, __double2int_rz(psy), __double2int_rz(psx)) | rt_probe(p, __double2int_rz(pey), __double2int_rz(pex))) return true;
    const float len = __double2float_rn(rt_dist64(psx, psy, pex, pey));
    const int P = rt_pow2_floor(rt_march_M(len, 10.0f));
    const float inv = 1.0f / (float)P;
    for (int j = 1; j < P; j++)
        if (cd_probe_interior(p, psx, psy, pex, pey, __fmul_rn((float)j, inv))) return true;
    return false;
}

// RTRT_CrashDetection, CDRT.cpp:507-603 (callable by a single lane)
__device__ bool cd_crash(const CdView &v, const CdI &in, double psx, double psy, int invS, double pex, double pey, int invE) {
    const int is = invS & 0x3FFFFFFF, ie = invE & 0x3FFFFFFF;
    const unsigned ks = (unsigned)invS & 0xC0000000u, ke = (unsigned)invE & 0xC0000000u;
    int island = -1; // the island whose kd-tree decides whether the dynamic plane is consulted
    if (ks == 0u) {
        if (ke == 0u) {
            if (!csr_has(v.cc_off, v.cc, is, ie)) return true;
            island = v.cp[v.cp_off[is]];
            if (!csr_has(v.cp_off, v.cp, ie, island)) island = v.cp[v.cp_off[is + 1] - 1];
        } else if (ke == 0x40000000u) {
            if (!csr_has(v.pc_off, v.pc, in.agent_island, is)) return true;
            island = in.agent_island;
        } else if (ke == 0x80000000u) {
            if (!csr_has(v.cp_off, v.cp, is, ie)) return true;
            island = ie;
        }
    } else if (ks == 0x40000000u) {
        if (ke == 0u) {
            if (!csr_has(v.pc_off, v.pc, in.agent_island, ie)) return true;
        } else if (ke == 0x80000000u) {
            if (in.agent_island != ie) return true;
        }
        island = in.agent_island;
    } else if (ks == 0x80000000u) {
        if (ke == 0u) {
            if (!csr_has(v.pc_off, v.pc, is, ie)) return true;
        } else if (ke == 0x40000000u) {
            if (in.agent_island != is) return true;
        } else if (ke == 0x80000000u) {
            if (is != ie) return true;
        }
        island = is;
    }
    if (island < 0 || !in.isl_tree[island]) return false;
    return cd_march(in.dpl, psx, psy, pex, pey);
}

// ---------------------------------------------------------------------------------------------
// eQr: 32-ary min-heap on (cost word, S point, E point)   (CDRT.h:263-281,328)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool pt_less(double2 a, double2 b) { return a.x != b.x ? a.x < b.x : a.y < b.y; }
__device__ bool q_tie_less(const CdI &in, int2 a, int2 b) {
    const double2 as = in.xy[a.x], bs = in.xy[b.x];
    if (as.x != bs.x || as.y != bs.y) return pt_less(as, bs);
    return pt_less(in.xy[a.y], in.xy[b.y]);
}
__device__ __forceinline__ bool q_less(const CdI &in, u64 ka, int2 a, u64 kb, int2 b) {
    return ka < kb || (ka == kb && (a.x != b.x || a.y != b.y) && q_tie_less(in, a, b));
}
// warp-collective, uniform arguments: sift the entry up from slot i (slots beyond i may already hold valid entries)
__device__ void q_sift_up(CdI &in, int i, u64 key, int S, int E) {
    const int lane = threadIdx.x & 31;
    int idx = i, steps = 0;
    while (idx > 0 && steps <= lane) idx = (idx - 1) >> 5, steps++;
    const bool valid = steps == lane + 1; // lane l holds the (l+1)-th ancestor
    u64 ak = 0;
    int2 ase = make_int2(0, 0);
    if (valid) ak = in.qkey[idx], ase = in.qse[idx];
    const bool up = valid && q_less(in, key, make_int2(S, E), ak, ase);
    const unsigned bal = __ballot_sync(FULL, up);
    const int m = __ffs(~bal) - 1; // ancestors are non-increasing upwards: the ones that move down are a prefix
    const int below = __shfl_up_sync(FULL, idx, 1);
    if (lane < m) {
        const int dst = lane == 0 ? i : below;
        in.qkey[dst] = ak, in.qse[dst] = ase;
    }
    const int top = __shfl_sync(FULL, idx, m > 0 ? m - 1 : 0);
    if (lane == 0) {
        const int dst = m == 0 ? i : top;
        in.qkey[dst] = key, in.qse[dst] = make_int2(S, E);
    }
    __syncwarp();
}
// warp-collective; queue must be non-empty
__device__ void q_pop(const CdView &v, CdI &in, u64 *key, int2 *se) {
    const int lane = threadIdx.x & 31;
    *key = in.qkey[0], *se = in.qse[0];
    const int n = --in.q_n;
    if (n == 0) return;
    const u64 lk = in.qkey[n];
    const int2 lse = in.qse[n];
    int i = 0;
    for (;;) {
        const int c0 = (i << 5) + 1;
        if (c0 >= n) break;
        const int c = c0 + lane;
        const bool valid = c < n;
        u64 ck = ~0ull;
        int2 cse = make_int2(-1, -1);
        if (valid) ck = in.qkey[c], cse = in.qse[c];
        const u64 mk = warp_min_u64(ck);
        unsigned tied = __ballot_sync(FULL, valid && ck == mk);
        int w = __ffs(tied) - 1;
        tied &= tied - 1;
        while (tied) { // equal cost words: the edge points decide
            const int t = __ffs(tied) - 1;
            tied &= tied - 1;
            const int2 a = make_int2(__shfl_sync(FULL, cse.x, t), __shfl_sync(FULL, cse.y, t));
            const int2 b = make_int2(__shfl_sync(FULL, cse.x, w), __shfl_sync(FULL, cse.y, w));
            if ((a.x != b.x || a.y != b.y) && q_tie_less(in, a, b)) w = t;
        }
        const int2 wse = make_int2(__shfl_sync(FULL, cse.x, w), __shfl_sync(FULL, cse.y, w));
        if (!q_less(in, mk, wse, lk, lse)) break;
        if (lane == 0) in.qkey[i] = mk, in.qse[i] = wse;
        i = c0 + w;
    }
    if (lane == 0) in.qkey[i] = lk, in.qse[i] = lse;
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// RTRT_Neighbor, CDRT.cpp:177-274: fills in.xnear in the reference's list order, returns its length
// ---------------------------------------------------------------------------------------------
struct TopK { // lane j holds the j-th best (d2, pos) so far, ascending; ties by position (= insertion order)
    double d2;
    int pos;
};
__device__ __forceinline__ bool tk_less(double a, int ia, double b, int ib) { return a < b || (a == b && ia < ib); }
// k nearest points of one island to (qx, qy); on return lane j < count holds the j-th
__device__ int cd_island_knn(const CdView &v, const CdI &in, int island, double qx, double qy, int k, TopK &mine) {
    const int lane = threadIdx.x & 31;
    mine.d2 = CUDART_INF, mine.pos = 0x7fffffff;
    double kd2 = CUDART_INF;
    int kpos = 0x7fffffff, found = 0;
    const int cnt = in.isl_cnt[island];
    const int *ids = in.isl_nodes + v.isl_cap_off[island];
    for (int i0 = 0; i0 < cnt; i0 += 32) {
        const int i = i0 + lane;
        double d2 = CUDART_INF;
        if (i < cnt) {
            const double2 p = in.xy[ids[i]];
            const double dx = __dsub_rn(qx, p.x), dy = __dsub_rn(qy, p.y);
            d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
        }
        unsigned m = __ballot_sync(FULL, i < cnt && tk_less(d2, i, kd2, kpos));
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const double cd2 = __shfl_sync(FULL, d2, src);
            const int cpos = i0 + src;
            if (!tk_less(cd2, cpos, kd2, kpos)) continue;
            const int at = __popc(__ballot_sync(FULL, tk_less(mine.d2, mine.pos, cd2, cpos)));
            const double ud2 = __shfl_up_sync(FULL, mine.d2, 1);
            const int upos = __shfl_up_sync(FULL, mine.pos, 1);
            if (lane == at) mine.d2 = cd2, mine.pos = cpos;
            else if (lane > at) mine.d2 = ud2, mine.pos = upos;
            if (lane >= k) mine.d2 = CUDART_INF, mine.pos = 0x7fffffff;
            found = min(found + 1, k);
            kd2 = __shfl_sync(FULL, mine.d2, k - 1);
            kpos = __shfl_sync(FULL, mine.pos, k - 1);
        }
    }
    return found;
}
__device__ int cd_append_cut_nodes(const CdView &v, CdI &in, const int *off, const int *val, int row, int m) {
    const int lane = threadIdx.x & 31;
    for (int j = off[row]; j < off[row + 1]; j++) {
        const int c = val[j], cnt = in.cut_cnt[c];
        const int *src = in.cut_nodes + v.cut_cap_off[c];
        for (int t = lane; t < cnt; t += 32) in.xnear[m + t] = src[t];
        m += cnt;
    }
    return m;
}
__device__ int cd_append_island_knn(const CdView &v, CdI &in, int island, double qx, double qy, bool drop_self, int m) {
    const int lane = threadIdx.x & 31;
    if (!in.isl_tree[island]) return m;
    TopK tk;
    const int found = cd_island_knn(v, in, island, qx, qy, v.nearK, tk);
    const bool keep = lane < found && !(drop_self && tk.d2 < 1e-16);
    const unsigned bal = __ballot_sync(FULL, keep);
    if (keep) in.xnear[m + __popc(bal & ((1u << lane) - 1u))] = in.isl_nodes[v.isl_cap_off[island] + tk.pos];
    return m + __popc(bal);
}
__device__ int cd_neighbors(const CdView &v, CdI &in, double px, double py, int invM) {
    const int lane = threadIdx.x & 31;
    int m = 0;
    if (invM == kOrigin) {
        m = cd_append_cut_nodes(v, in, v.pc_off, v.pc, in.agent_island, m);
        const double2 a = in.xy[in.agent_id];
        m = cd_append_island_knn(v, in, in.agent_island, a.x, a.y, false, m);
    } else if ((unsigned)invM & 0xC0000000u) {
        const int island = invM & 0x3FFFFFFF;
        m = cd_append_cut_nodes(v, in, v.pc_off, v.pc, island, m);
        m = cd_append_island_knn(v, in, island, px, py, true, m);
        if (island == in.agent_island) {
            if (lane == 0) in.xnear[m] = in.agent_id;
            m++;
        }
    } else {
        const int c = invM & 0x3FFFFFFF;
        m = cd_append_cut_nodes(v, in, v.cc_off, v.cc, c, m);
        for (int j = v.cp_off[c]; j < v.cp_off[c + 1]; j++) {
            const int island = v.cp[j];
            m = cd_append_island_knn(v, in, island, px, py, true, m);
            if (island == in.agent_island) {
                if (lane == 0) in.xnear[m] = in.agent_id;
                m++;
            }
        }
    }
    __syncwarp();
    return m;
}

// choose-parent over in.xnear[0..m)  (CDRT.cpp:848-864, 903-919, 1094-1108): strict '>' so the first of equals wins
__device__ void cd_best_parent(const CdView &v, CdI &in, double px, double py, int invM, int m, u64 *cmin, int *pmin) {
    const int lane = threadIdx.x & 31;
    const u64 init = cw_make(true, kDoubleMax);
    u64 best = init;
    int bestq = -1;
    for (int base = 0; base < m; base += 32) {
        const int j = base + lane;
        u64 c = ~0ull;
        int q = -1;
        if (j < m) {
            q = in.xnear[j];
            const double2 p = in.xy[q];
            c = cw_add(in.cw[q], cd_crash(v, in, px, py, invM, p.x, p.y, in.inv[q]), rt_dist64(px, py, p.x, p.y));
        }
        const u64 mk = warp_min_u64(c);
        const int w = __ffs(__ballot_sync(FULL, j < m && c == mk)) - 1;
        if (w >= 0 && mk < best) best = mk, bestq = __shfl_sync(FULL, q, w);
    }
    if (lane == 0) in.g->checks += m;
    *cmin = best, *pmin = bestq;
}

// pushes the qualifying (h, S, E) of a chunk one by one
// Batched push: the qualifying entries take the next leaf slots; an entry that is not smaller than its slot's parent
// (the usual case: eQr mostly receives entries far from its head) is final where it lands — one parent load and one
// coalesced store for the whole chunk; the few that are smaller are sifted up one by one afterwards (ancestor slots
// only ever decrease, so the entries already placed stay valid).
__device__ void cd_push_chunk(const CdView &v, CdI &in, bool ok, u64 h, int S, int E) {
    const int lane = threadIdx.x & 31;
    const unsigned bal = __ballot_sync(FULL, ok);
    const int cnt = __popc(bal);
    if (cnt == 0) return;
    const int n0 = in.q_n;
    if (n0 + cnt > v.qcap) {
        in.status = ST_QUEUE_FULL;
        return;
    }
    const int pos = n0 + __popc(bal & ((1u << lane) - 1u));
    bool slow = ok;
    if (n0 >= 64 && ok) { // (below that a slot's parent could itself be a slot of this chunk)
        const int par = (pos - 1) >> 5;
        slow = q_less(in, h, make_int2(S, E), in.qkey[par], in.qse[par]);
        if (!slow) in.qkey[pos] = h, in.qse[pos] = make_int2(S, E);
    }
    in.q_n = n0 + cnt;
    __syncwarp();
    unsigned sb = __ballot_sync(FULL, slow);
    while (sb) {
        const int src = __ffs(sb) - 1;
        sb &= sb - 1;
        q_sift_up(in, __shfl_sync(FULL, pos, src), __shfl_sync(FULL, h, src), __shfl_sync(FULL, S, src), __shfl_sync(FULL, E, src));
    }
    if (lane == 0) in.g->pushes += cnt;
}
// candidate links of node p against in.xnear[0..m): `larger` = p may get a cheaper parent among them
// (RTRT_AddPotentialLinks CDRT.cpp:797-830); otherwise p may become their cheaper parent (also the
// loop after a successful sample, CDRT.cpp:1017-1031)
__device__ void cd_links(const CdView &v, CdI &in, int p, int m, bool larger) {
    const int lane = threadIdx.x & 31;
    const double2 pp = in.xy[p];
    const u64 pc = in.cw[p];
    const bool goal_set = in.g->goal_set;
    const double2 goal = in.g->goal;
    for (int base = 0; base < m; base += 32) {
        const int j = base + lane;
        bool ok = false;
        u64 h = 0;
        int S = 0, E = 0;
        if (j < m) {
            const int q = in.xnear[j];
            const double2 qp = in.xy[q];
            const u64 qc = in.cw[q];
            const double d = rt_dist64(pp.x, pp.y, qp.x, qp.y);
            if (larger) {
                h = cw_add(qc, false, d);
                ok = pc > h;
                if (ok && goal_set) h = cw_add(h, false, rt_dist64(pp.x, pp.y, goal.x, goal.y));
                S = q, E = p;
            } else {
                h = cw_add(pc, false, d);
                ok = qc > h;
                if (ok && goal_set) h = cw_add(h, false, rt_dist64(qp.x, qp.y, goal.x, goal.y));
                S = p, E = q;
            }
        }
        cd_push_chunk(v, in, ok, h, S, E);
    }
}
__device__ void cd_add_potential_links(const CdView &v, CdI &in, int p, bool larger) {
    const double2 pp = in.xy[p];
    const int m = cd_neighbors(v, in, pp.x, pp.y, in.inv[p]);
    cd_links(v, in, p, m, larger);
}

// ---------------------------------------------------------------------------------------------
// RTRT_CrashDObsIsland, CDRT.cpp:638-762
// ---------------------------------------------------------------------------------------------
__device__ void cd_crash_dobs_island(const CdView &v, CdI &in, const double4 *pose, const double2 *chk, const int *chk_obs) {
    const int lane = threadIdx.x & 31;
    const int epoch = in.g->epoch + 1;
    __syncwarp();
    if (lane == 0) in.g->epoch = epoch;
    __syncwarp();
    // islands under the obstacles' check points (:644-668)
    for (int i = lane; i < in.g->n_chk; i += 32) {
        const double4 o = pose[chk_obs[i]];
        const double2 p = chk[i];
        const int ix = __double2int_rz(__dadd_rn(__dsub_rn(__dmul_rn(o.w, p.x), __dmul_rn(o.z, p.y)), o.x));
        const int iy = __double2int_rz(__dadd_rn(__dadd_rn(__dmul_rn(o.z, p.x), __dmul_rn(o.w, p.y)), o.y));
        if (ix >= v.W || iy >= v.H || ix < 0 || iy < 0) continue;
        const int isl = v.isl[(size_t)iy * v.W + ix];
        if (isl != 0xFFFF && !(isl & 0xC000)) in.isl_mark[isl] = epoch;
    }
    __syncwarp();
    // gone = has a tree but no longer covered; their nodes die.  First re-hang the cut-line nodes whose parent dies (:679-705)
    bool any_gone = false;
    for (int i0 = 0; i0 < v.npoly; i0 += 32) {
        const int i = i0 + lane;
        const bool gone = i < v.npoly && in.isl_tree[i] && in.isl_mark[i] != epoch;
        if (__any_sync(FULL, gone)) any_gone = true;
    }
    if (any_gone) {
        for (int island = 0; island < v.npoly; island++) {
            if (!(in.isl_tree[island] && in.isl_mark[island] != epoch)) continue;
            for (int j = v.pc_off[island]; j < v.pc_off[island + 1]; j++) {
                const int c = v.pc[j], cnt = in.cut_cnt[c];
                const int *src = in.cut_nodes + v.cut_cap_off[c];
                for (int t = lane; t < cnt; t += 32) {
                    const int x = src[t];
                    int a = in.par[x];
                    if (a < 0) continue;
                    const int ai = in.inv[a];
                    if (((unsigned)ai & 0xC0000000u) != 0x80000000u) continue;
                    const int pi = ai & 0x3FFFFFFF;
                    if (!(in.isl_tree[pi] && in.isl_mark[pi] != epoch)) continue;
                    while (a >= 0 && ((unsigned)in.inv[a] & 0xC0000000u) == 0x80000000u) a = in.par[a];
                    in.par[x] = a;
                }
            }
        }
        __syncwarp();
        for (int island = 0; island < v.npoly; island++) { // erase (:707-716)
            if (!(in.isl_tree[island] && in.isl_mark[island] != epoch)) continue;
            const int cnt = in.isl_cnt[island];
            const int *ids = in.isl_nodes + v.isl_cap_off[island];
            for (int t = lane; t < cnt; t += 32) in.flag[ids[t]] = 0;
            __syncwarp();
            if (lane == 0) in.isl_cnt[island] = 0, in.isl_tree[island] = 0;
            __syncwarp();
        }
    }
    // newly covered islands get an (empty) tree (:759-761)
    int trees = 0;
    for (int i0 = 0; i0 < v.npoly; i0 += 32) {
        const int i = i0 + lane;
        bool has = false;
        if (i < v.npoly) {
            if (!in.isl_tree[i] && in.isl_mark[i] == epoch) in.isl_tree[i] = 1;
            has = in.isl_tree[i];
        }
        trees += __popc(__ballot_sync(FULL, has));
    }
    in.n_trees = trees;
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// RTRT_UpdateTreeCost, CDRT.cpp:764-796
// ---------------------------------------------------------------------------------------------
__device__ void cd_set_add(const CdView &v, CdI &in, int *list, int *count, int id) {
    const int at = atomicAdd(count, 1);
    if (at < v.set_cap) list[at] = id;
    else in.status = ST_SET_FULL;
}
__device__ void cd_update_tree_cost(const CdView &v, CdI &in) {
    const int lane = threadIdx.x & 31;
    const int n = in.n;
    int *cs = in.child_start;
    // children CSR of the live nodes (counting sort by parent)
    for (int i = lane; i <= n + 1; i += 32) cs[i] = 0;
    __syncwarp();
    for (int i = lane; i < n; i += 32)
        if ((in.flag[i] & F_ALIVE) && in.par[i] >= 0) atomicAdd(&cs[in.par[i] + 2], 1);
    __syncwarp();
    int carry = 0;
    for (int base = 0; base <= n + 1; base += 32) { // inclusive scan; cs[p+1] becomes the cursor of parent p
        const int i = base + lane;
        const int val = i <= n + 1 ? cs[i] : 0;
        int x = val;
        for (int off = 1; off < 32; off <<= 1) {
            const int y = __shfl_up_sync(FULL, x, off);
            if (lane >= off) x += y;
        }
        if (i <= n + 1) cs[i] = carry + x;
        carry += __shfl_sync(FULL, x, 31);
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32)
        if ((in.flag[i] & F_ALIVE) && in.par[i] >= 0) in.children[atomicAdd(&cs[in.par[i] + 1], 1)] = i;
    __syncwarp();
    // now children of p = children[cs[p] .. cs[p+1])
    int *cur = in.frontier, *nxt = in.frontier + v.cap;
    if (lane == 0) cur[0] = in.agent_id;
    int fn = 1;
    __syncwarp();
    long long checks = 0;
    while (fn > 0) {
        int produced = 0;
        for (int base = 0; base < fn; base += 32) {
            const int f = base + lane;
            int p = -1, c0 = 0, nc = 0;
            if (f < fn) p = cur[f], c0 = cs[p], nc = cs[p + 1] - c0;
            int x = nc;
            for (int off = 1; off < 32; off <<= 1) {
                const int y = __shfl_up_sync(FULL, x, off);
                if (lane >= off) x += y;
            }
            int at = produced + x - nc;
            produced += __shfl_sync(FULL, x, 31);
            if (nc > 0) {
                const double2 pp = in.xy[p];
                const u64 pc = in.cw[p];
                const int pinv = in.inv[p];
                for (int t = 0; t < nc; t++) {
                    const int s = in.children[c0 + t];
                    const double2 sp = in.xy[s];
                    const u64 c = cw_add(pc, cd_crash(v, in, pp.x, pp.y, pinv, sp.x, sp.y, in.inv[s]), rt_dist64(pp.x, pp.y, sp.x, sp.y));
                    const u64 old = in.cw[s];
                    unsigned char fl = in.flag[s];
                    if (old < c) {
                        if (!(fl & F_LARGER)) cd_set_add(v, in, in.larger, &in.g->n_larger, s);
                        fl = (fl | F_LARGER) & ~F_SMALLER;
                    } else if (c < old) {
                        if (!(fl & F_SMALLER)) cd_set_add(v, in, in.smaller, &in.g->n_smaller, s);
                        fl |= F_SMALLER;
                    }
                    in.flag[s] = fl;
                    in.cw[s] = c;
                    nxt[at++] = s;
                    checks++;
                }
            }
        }
        __syncwarp();
        int *t = cur;
        cur = nxt, nxt = t;
        fn = produced;
    }
    for (int off = 16; off > 0; off >>= 1) checks += __shfl_xor_sync(FULL, checks, off);
    if (lane == 0) in.g->checks += checks;
    in.status = __reduce_max_sync(FULL, in.status);
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// nodes
// ---------------------------------------------------------------------------------------------
__device__ int cd_new_node(const CdView &v, CdI &in, double x, double y, int invM) {
    const int lane = threadIdx.x & 31;
    if (in.n >= v.cap) {
        in.status = ST_NODES_FULL;
        return -1;
    }
    const int id = in.n++;
    if (lane == 0) {
        in.xy[id] = make_double2(x, y);
        in.cw[id] = cw_make(false, 0.0);
        in.par[id] = -1, in.inv[id] = invM, in.flag[id] = F_ALIVE;
    }
    __syncwarp();
    return id;
}
// insert id into Mnode[c], kept in std::set<BIpoint> order
__device__ void cd_cut_insert(const CdView &v, CdI &in, int c, int id) {
    const int lane = threadIdx.x & 31;
    const int cnt = in.cut_cnt[c], capc = v.cut_cap_off[c + 1] - v.cut_cap_off[c];
    if (cnt >= capc) {
        in.status = ST_CUT_FULL;
        return;
    }
    int *lst = in.cut_nodes + v.cut_cap_off[c];
    const double2 p = in.xy[id];
    int before = 0;
    for (int base = 0; base < cnt; base += 32) {
        const int j = base + lane;
        before += __popc(__ballot_sync(FULL, j < cnt && pt_less(in.xy[lst[j]], p)));
    }
    for (int hi = cnt; hi > before; hi -= 32) { // shift [before, cnt) up by one, from the top
        const int j = hi - 1 - lane;
        int val = 0;
        if (j >= before) val = lst[j];
        __syncwarp();
        if (j >= before) lst[j + 1] = val;
        __syncwarp();
    }
    if (lane == 0) lst[before] = id, in.cut_cnt[c] = cnt + 1;
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// RTRT_Sampling, CDRT.cpp:832-925.  Returns the new node id (or -1) and the length of in.xnear.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int cd_rand_at(CdI &in, int pos) { return pos < in.rand_n ? in.rand[pos] : 0; }
__device__ int cd_sampling(const CdView &v, CdI &in, int *m_out) {
    const int lane = threadIdx.x & 31;
    *m_out = 0;
    const float p1 = __fdiv_rn(__int2float_rn(cd_rand_at(in, in.rand_pos)), 2147483648.0f);
    in.rand_pos++;
    if ((double)p1 < v.p_th || in.n_trees == 0) {
        if (in.samp_pos >= in.samp_n) return -1;
        const int id = 1 + v.ncut + in.samp_pos; // nodes of the shuffled sample list were laid out at start
        in.samp_pos++;
        const double2 p = in.xy[id];
        const int c = in.inv[id];
        const int m = cd_neighbors(v, in, p.x, p.y, c);
        u64 cmin;
        int pmin;
        cd_best_parent(v, in, p.x, p.y, c, m, &cmin, &pmin);
        if (pmin < 0) in.status = ST_REF_UB; // the reference would create a node at (-1,-1)
        if (lane == 0) in.cw[id] = cmin, in.par[id] = pmin, in.flag[id] = F_ALIVE;
        __syncwarp();
        cd_cut_insert(v, in, c, id);
        *m_out = m;
        return id;
    }
    const float fw = __uint2float_rn((unsigned)v.W), fh = __uint2float_rn((unsigned)v.H);
    for (int err = 0; err < 20; err++) {
        bool ok = false;
        double x = 0, y = 0;
        int island = 0;
        for (int rej0 = 0; rej0 < 100 && !ok; rej0 += 32) { // up to 32 rejection rounds evaluated at once
            const int r = rej0 + lane;
            bool acc = false;
            double cx = 0, cy = 0;
            int ci = 0;
            if (r < 100) {
                const int at = in.rand_pos + 2 * r;
                cx = (double)__fmul_rn(__fdiv_rn(__int2float_rn(cd_rand_at(in, at)), 2147483648.0f), fw);
                cy = (double)__fmul_rn(__fdiv_rn(__int2float_rn(cd_rand_at(in, at + 1)), 2147483648.0f), fh);
                const int ix = __double2int_rz(cx), iy = __double2int_rz(cy);
                if (!rt_probe(in.dpl, iy, ix)) {
                    ci = cd_island_at(v, iy, ix);
                    acc = ci != 0xFFFF && ci < v.npoly && in.isl_tree[ci];
                }
            }
            const unsigned bal = __ballot_sync(FULL, acc);
            if (bal) {
                const int w = __ffs(bal) - 1;
                x = __shfl_sync(FULL, cx, w), y = __shfl_sync(FULL, cy, w), island = __shfl_sync(FULL, ci, w);
                in.rand_pos += 2 * (rej0 + w + 1);
                ok = true;
            } else if (rej0 + 32 >= 100)
                in.rand_pos += 200;
        }
        if (!ok) continue;
        TopK tk;
        const int found = cd_island_knn(v, in, island, x, y, 1, tk);
        const double d0 = __shfl_sync(FULL, tk.d2, 0);
        if (found > 0 && d0 < v.gate2) continue;
        const int inv = (int)(0x80000000u | (unsigned)island);
        const int id = cd_new_node(v, in, x, y, inv);
        if (id < 0) return -1;
        const int m = cd_neighbors(v, in, x, y, inv);
        u64 cmin;
        int pmin;
        cd_best_parent(v, in, x, y, inv, m, &cmin, &pmin);
        if (pmin < 0) in.status = ST_REF_UB;
        const int cnt = in.isl_cnt[island];
        if (cnt >= v.isl_cap_off[island + 1] - v.isl_cap_off[island]) {
            in.status = ST_ISLAND_FULL;
            return -1;
        }
        if (lane == 0) {
            in.cw[id] = cmin, in.par[id] = pmin;
            in.isl_nodes[v.isl_cap_off[island] + cnt] = id;
            in.isl_cnt[island] = cnt + 1;
        }
        __syncwarp();
        *m_out = m;
        return id;
    }
    return -1;
}

// ---------------------------------------------------------------------------------------------
// GetPath(point, path), CDRT.cpp:1076-1123: best entry node for `point`; the chain is read by the callers
// ---------------------------------------------------------------------------------------------
__device__ void cd_entry_node(const CdView &v, CdI &in, double2 pt, u64 *cmin, int *pmin) {
    const int inv = (int)(0x80000000u | (unsigned)cd_island_at(v, __double2int_rz(pt.y), __double2int_rz(pt.x)));
    const int m = cd_neighbors(v, in, pt.x, pt.y, inv);
    cd_best_parent(v, in, pt.x, pt.y, inv, m, cmin, pmin);
}
__device__ void cd_get_path(const CdView &v, CdI &in) {
    const int lane = threadIdx.x & 31;
    CdInst *g = in.g;
    if (!g->goal_set) {
        g->path_len = 0, g->path_col = 1, g->path_cost = kDoubleMax;
        return;
    }
    u64 cmin;
    int pmin;
    cd_entry_node(v, in, g->goal, &cmin, &pmin);
    int hops = 0;
    for (int a = pmin; a >= 0 && hops < 2000; a = in.par[a]) hops++;
    if (hops >= 2000 || hops + 1 > v.path_cap) { // path.size() > 2000 once the 2000th chain node is pushed (:1113-1117)
        g->path_len = 0, g->path_col = 1, g->path_cost = (double)3.4e38f;
        return;
    }
    if (lane == 0) {
        int at = hops;
        in.path[at--] = -2; // the goal point itself
        for (int a = pmin; a >= 0; a = in.par[a]) in.path[at--] = a;
    }
    g->path_len = hops + 1, g->path_col = (int)(cmin >> 63), g->path_cost = cw_cost(cmin);
    __syncwarp();
}

// the agent hand-over inside RTRT_task, CDRT.cpp:946-996
__device__ void cd_move_agent(const CdView &v, CdI &in) {
    const int lane = threadIdx.x & 31;
    const double2 tag = in.g->agent_tag;
    const int tag_island = cd_island_at(v, __double2int_rz(tag.y), __double2int_rz(tag.x));
    const int old = in.agent_id;
    if (in.agent_island == tag_island) {
        const int id = cd_new_node(v, in, tag.x, tag.y, kOrigin);
        if (id < 0) return;
        for (int i = lane; i < in.n; i += 32)
            if (in.par[i] == old && (in.flag[i] & F_ALIVE)) in.par[i] = id;
        if (lane == 0) in.flag[old] = 0;
        in.agent_id = id;
    } else {
        u64 cmin;
        int pmin;
        cd_entry_node(v, in, tag, &cmin, &pmin); // GetPath(agent_tag, path) with the OLD agent island
        if (pmin < 0) { // reference: path.pop_front() / front() on a one-element list
            in.status = ST_REF_UB;
            return;
        }
        const int id = cd_new_node(v, in, tag.x, tag.y, kOrigin);
        if (id < 0) return;
        // reverse the chain pmin -> ... -> old agent so that it hangs under the new root
        int prev = id, cur = pmin, hops = 0;
        while (cur != old && cur >= 0 && hops <= 2000) {
            const int next = in.par[cur];
            __syncwarp();
            if (lane == 0) in.par[cur] = prev;
            __syncwarp();
            prev = cur, cur = next, hops++;
        }
        if (cur != old) { // chain did not end at the agent, or > 2000 hops (path.clear())
            in.status = ST_REF_UB;
            return;
        }
        // prev = x1, the path's second point: the old agent's other children move under it
        for (int i = lane; i < in.n; i += 32)
            if (in.par[i] == old && i != prev && (in.flag[i] & F_ALIVE)) in.par[i] = prev;
        if (lane == 0) in.flag[old] = 0;
        in.agent_id = id;
    }
    in.agent_island = tag_island;
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// RTRT_task, CDRT.cpp:927-1068
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) cdrt_tick_kernel(CdView v, long long dt_ns, int do_get_path) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= v.n_inst) return;
    CdInst *g = v.inst + w;
    if (g->status != ST_OK || g->n == 0 || !v.hdr[w].started) return;
    if (lane == 0) {
        const CdHdr h = v.hdr[w];
        g->agent_tag = h.tag, g->goal = h.goal, g->goal_set = h.goal_set, g->n_obs = h.n_obs, g->n_chk = h.n_chk;
    }
    __syncwarp();
    CdI in;
    in.g = g;
    const size_t nb = (size_t)w * v.cap;
    in.xy = v.xy + nb, in.cw = v.cw + nb, in.par = v.par + nb, in.inv = v.inv + nb, in.flag = v.flag + nb;
    in.cut_cnt = v.cut_cnt + (size_t)w * v.ncut, in.cut_nodes = v.cut_nodes + (size_t)w * v.cut_cap_total;
    in.isl_cnt = v.isl_cnt + (size_t)w * v.npoly, in.isl_nodes = v.isl_nodes + (size_t)w * v.isl_cap_total;
    in.isl_tree = v.isl_tree + (size_t)w * v.npoly, in.isl_mark = v.isl_mark + (size_t)w * v.npoly;
    in.qkey = v.qkey + (size_t)w * v.qcap, in.qse = v.qse + (size_t)w * v.qcap;
    in.larger = v.larger + (size_t)w * v.set_cap, in.smaller = v.smaller + (size_t)w * v.set_cap;
    in.xnear = v.xnear + (size_t)w * (v.cap + 64);
    in.child_start = v.child_start + (size_t)w * (v.cap + 2), in.children = v.children + nb, in.frontier = v.frontier + 2 * nb;
    in.path = v.path + (size_t)w * v.path_cap;
    in.rand = v.rand + (size_t)w * v.rand_cap;
    in.dpl.w = v.dplane + (size_t)w * v.plane_words, in.dpl.W = v.W, in.dpl.H = v.H, in.dpl.tilesX = v.tilesX;
    in.agent_id = g->agent_id, in.agent_island = g->agent_island, in.n = g->n, in.q_n = g->q_n, in.status = ST_OK;
    in.clk_n = g->clk_n, in.dt = dt_ns, in.rand_pos = g->rand_pos, in.rand_n = g->rand_n;
    in.samp_pos = g->samp_pos, in.samp_n = g->samp_n, in.n_trees = g->n_trees;
    __syncwarp();
    // every lane keeps identical copies of the scalar state; lane 0 owns the global copy
#define G0(stmt)                                                                                                                            \
    do {                                                                                                                                    \
        if (lane == 0) { stmt; }                                                                                                           \
        __syncwarp();                                                                                                                       \
    } while (0)

    long long update_old = g->update_old, smaller_old = g->smaller_old;
    if (!g->timers_init) { // function-local statics (:929-930)
        update_old = cd_now(in);
        smaller_old = cd_now(in);
    }
    cd_now(in); // Time0
    long long tc0 = clock64();
    // (the dynamic plane of this instance was drawn before the launch: RTRT_DrawDObsmap)
    __syncwarp();
    cd_crash_dobs_island(v, in, v.pose + (size_t)w * v.max_obs, v.chk + (size_t)w * v.max_chk, v.chk_obs + (size_t)w * v.max_chk);

    const double2 ag = in.xy[in.agent_id];
    const double2 tag = g->agent_tag;
    bool gate = rt_dist64(ag.x, ag.y, tag.x, tag.y) > v.r_th;
    if (!gate) gate = (double)(cd_now(in) - update_old) > 50e6;
    if (gate) {
        update_old = cd_now(in);
        if (ag.x != tag.x || ag.y != tag.y) cd_move_agent(v, in);
        if (in.status == ST_OK) cd_update_tree_cost(v, in);
    }
    long long tc1 = clock64();
    if (in.status == ST_OK) {
        // largerSet (:998-1003)
        const int nl = g->n_larger;
        for (int i = 0; i < nl && in.status == ST_OK; i++) {
            const int p = in.larger[i];
            const unsigned char fl = in.flag[p];
            __syncwarp();
            if (!(fl & F_LARGER)) continue;
            if (lane == 0) in.flag[p] = fl & ~F_LARGER;
            __syncwarp();
            if (fl & F_ALIVE) cd_add_potential_links(v, in, p, true);
        }
        G0(g->n_larger = 0);
        // smallerSet every 100 ms (:1004-1013)
        if ((double)(cd_now(in) - smaller_old) > 100e6) {
            smaller_old = cd_now(in);
            const int ns = g->n_smaller;
            for (int i = 0; i < ns && in.status == ST_OK; i++) {
                const int p = in.smaller[i];
                const unsigned char fl = in.flag[p];
                __syncwarp();
                if (!(fl & F_SMALLER)) continue;
                if (lane == 0) in.flag[p] = fl & ~F_SMALLER;
                __syncwarp();
                if (fl & F_ALIVE) cd_add_potential_links(v, in, p, false);
            }
            G0(g->n_smaller = 0);
        }
    }
    long long tc2 = clock64();
    // RewireN samples (:1014-1033)
    for (int s = 0; s < v.RewireN && in.status == ST_OK; s++) {
        int m;
        const int x = cd_sampling(v, in, &m);
        if (x < 0) continue;
        if (lane == 0) g->sampled++;
        cd_links(v, in, x, m, false);
    }
    long long tc3 = clock64();
    long long fine[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // the eQr loop (:1035-1067)
    if (in.status == ST_OK) {
        const long long t0 = cd_now(in);
        while (in.q_n > 0 && (double)(cd_now(in) - t0) < 20e6 && in.status == ST_OK) {
            u64 key;
            int2 se;
            const long long f0 = clock64();
            q_pop(v, in, &key, &se);
            const long long f1 = clock64();
            fine[0] += f1 - f0;
            if (lane == 0) g->pops++;
            const int S = se.x, E = se.y;
            if (!(in.flag[S] & F_ALIVE) || !(in.flag[E] & F_ALIVE)) continue;
            const double2 sp = in.xy[S], ep = in.xy[E];
            const u64 hat = cw_add(in.cw[S], false, rt_dist64(sp.x, sp.y, ep.x, ep.y));
            const bool stale = !(in.cw[E] > hat);
            const long long f2 = clock64();
            fine[1] += f2 - f1;
            if (stale) continue;
            const u64 tc = cw_add(hat, cd_crash(v, in, sp.x, sp.y, in.inv[S], ep.x, ep.y, in.inv[E]), 0.0);
            if (lane == 0) g->checks++;
            const bool keep = in.cw[E] > tc;
            const long long f3 = clock64();
            fine[2] += f3 - f2, fine[5]++;
            if (!keep) continue;
            __syncwarp();
            if (lane == 0) in.cw[E] = tc, in.par[E] = S, g->rewired++;
            __syncwarp();
            const double2 pp = in.xy[E];
            const int m = cd_neighbors(v, in, pp.x, pp.y, in.inv[E]);
            const long long f4 = clock64();
            cd_links(v, in, E, m, false);
            fine[3] += f4 - f3, fine[4] += clock64() - f4, fine[6] += m, fine[7]++;
        }
    }
    long long tc4 = clock64();
    if (do_get_path && in.status == ST_OK) cd_get_path(v, in);
    __syncwarp();
    if (lane == 0) g->t_phase[0] = tc1 - tc0, g->t_phase[1] = tc2 - tc1, g->t_phase[2] = tc3 - tc2, g->t_phase[3] = tc4 - tc3;
    if (lane == 0)
        for (int k = 0; k < 8; k++) g->t_fine[k] = fine[k];
    if (lane == 0) {
        g->agent_id = in.agent_id, g->agent_island = in.agent_island, g->n = in.n, g->q_n = in.q_n;
        g->clk_n = in.clk_n, g->update_old = update_old, g->smaller_old = smaller_old, g->timers_init = 1;
        g->rand_exhausted |= in.rand_pos > in.rand_n;
        g->rand_pos = min(in.rand_pos, in.rand_n);
        g->samp_pos = in.samp_pos, g->n_trees = in.n_trees;
        g->status = in.status;
    }
#undef G0
}

// =============================================================================================
// host side
// =============================================================================================
struct CdObstacle { // CTrtrt_DynamicObstacle, CDRT.h:250-261
    int type = -1;
    double x = 0, y = 0, angle = 0, radius = 0;
    std::vector<double> outline, check; // (x, y) pairs
};
struct CdHostInst {
    std::map<int, CdObstacle> obs;
    double agent_tag[2] = {0, 0}, goal[2] = {0, 0};
    int goal_set = 0;
    bool started = false, check_dirty = true;
};
struct CdCut {
    double S[2], E[2], core[2];
};
struct rtpp_cdrt_engine {
    rtpp_ctx *ctx;
    CdView v;
    double resolution;
    std::vector<CdCut> cut;
    std::vector<int> cp_off, cp, cc_off, cc, pc_off, pc, cut_cap_off, isl_cap_off;
    std::vector<uint16_t> isl; // host copy of the island plane
    std::vector<CdHostInst> inst;
    std::vector<void *> allocs;
    std::vector<double> h_pose, h_chk;
    std::vector<int> h_chk_obs;
    bool chk_dirty = true;
    RasterShapes drawn;       // what the instances' dynamic planes hold (erased, not memset, at the next tick)
    bool drawn_valid = false; // false until the first tick has cleared every plane
};

#define CD_CK(e, call) RTPP_CK((e)->ctx, call)
template <class T> static int cd_alloc(rtpp_cdrt_engine *e, T **p, size_t count, bool zero = true) {
    *p = nullptr;
    if (cudaMalloc((void **)p, std::max<size_t>(count, 1) * sizeof(T)) != cudaSuccess) return RTPP_ERR_NOMEM;
    e->allocs.push_back(*p);
    if (zero && cudaMemset(*p, 0, std::max<size_t>(count, 1) * sizeof(T)) != cudaSuccess) return RTPP_ERR_CUDA;
    return RTPP_OK;
}
template <class T> static int cd_upload_const(rtpp_cdrt_engine *e, const T **dst, const std::vector<T> &src) {
    T *d = nullptr;
    int rc = cd_alloc(e, &d, src.size(), false);
    if (rc) return rc;
    if (!src.empty() && cudaMemcpy(d, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) return RTPP_ERR_CUDA;
    *dst = d;
    return RTPP_OK;
}
static double h_dist64(const double *a, const double *b) {
    double dx = a[0] - b[0], dy = a[1] - b[1];
    return sqrt(dx * dx + dy * dy);
}
static u64 h_cw(bool col, double c) {
    u64 b;
    memcpy(&b, &c, 8);
    return ((u64)col << 63) | (b & 0x7fffffffffffffffull);
}
// the cut-line sample points of RTRT_start (:309-329) in generation order
static void cd_cut_samples(const rtpp_cdrt_engine *e, std::vector<double> &xy, std::vector<int> &cut_of) {
    for (size_t j = 0; j < e->cut.size(); j++) {
        const double *S = e->cut[j].S, *E = e->cut[j].E;
        const double dlen = h_dist64(S, E);
        const int dnum = (int)(dlen / e->resolution + 4);
        auto add = [&](double a, double b) { // S*a + E*b
            xy.push_back(a * S[0] + b * E[0]), xy.push_back(a * S[1] + b * E[1]);
            cut_of.push_back((int)j);
        };
        if (20 < dlen) {
            const double t = 8.0 / dlen;
            add(t, 1 - t);
            add(1 - t, t);
        }
        for (int k = 1; k < dnum; k++) {
            if (2 * k == dnum) continue;
            const double t = (double)k / dnum, t1 = 1.0 - t;
            add(t1, t);
        }
    }
}
static int cd_island_host(const rtpp_cdrt_engine *e, double x, double y) { // at<uint16_t>((int)y, (int)x)
    const long long lin = (long long)(int)y * e->v.W + (int)x;
    if (lin < 0 || lin >= (long long)e->v.W * e->v.H) return 0xFFFF;
    return e->isl[lin];
}
static bool cd_placeable(const rtpp_cdrt_engine *e, double x, double y) { // CDRT.cpp:386-398
    if (x >= (double)(unsigned)e->v.W || y >= (double)(unsigned)e->v.H) return false;
    const int i = cd_island_host(e, x, y);
    return !(i == 0xFFFF || (i & 0xC000));
}

extern "C" {

int rtpp_cdrt_create(rtpp_ctx *ctx, int n_instances, int node_capacity, int queue_capacity, int W, int H, int ncut, const double *cut_se,
                     const double *cut_core, const int32_t *cp_off, const int32_t *cp, const int32_t *cc_off, const int32_t *cc, int npoly,
                     const int32_t *pv_off, const double *pv, const int32_t *pc_off, const int32_t *pc, const uint16_t *island_plane,
                     double resolution, double r_th, double p_th, int nearK, int RewireN, int rand_capacity, rtpp_cdrt_engine **out) {
    if (!ctx || !out || n_instances <= 0 || ncut <= 0 || npoly <= 0 || !cut_se || !cut_core || !cp_off || !cp || !cc_off || !cc || !pv_off || !pv ||
        !pc_off || !pc || !island_plane || nearK < 1 || nearK > 32 || W <= 0 || H <= 0 || npoly >= 0x4000)
        return RTPP_ERR_INVALID;
    RTPP_CK(ctx, cudaSetDevice(ctx->device));
    rtpp_cdrt_engine *e = new rtpp_cdrt_engine();
    e->ctx = ctx, e->resolution = resolution;
    CdView &v = e->v;
    memset(&v, 0, sizeof(v));
    v.n_inst = n_instances, v.qcap = queue_capacity, v.ncut = ncut, v.npoly = npoly, v.W = W, v.H = H;
    v.nearK = nearK, v.RewireN = RewireN, v.r_th = r_th, v.p_th = p_th, v.gate2 = 10 * r_th * r_th;
    v.max_obs = 16, v.max_chk = 8192, v.path_cap = 2048, v.rand_cap = rand_capacity;
    e->cut.resize(ncut);
    for (int i = 0; i < ncut; i++) {
        memcpy(e->cut[i].S, cut_se + 4 * i, 16), memcpy(e->cut[i].E, cut_se + 4 * i + 2, 16), memcpy(e->cut[i].core, cut_core + 2 * i, 16);
    }
    e->cp_off.assign(cp_off, cp_off + ncut + 1), e->cp.assign(cp, cp + cp_off[ncut]);
    e->cc_off.assign(cc_off, cc_off + ncut + 1), e->cc.assign(cc, cc + cc_off[ncut]);
    e->pc_off.assign(pc_off, pc_off + npoly + 1), e->pc.assign(pc, pc + pc_off[npoly]);
    e->isl.assign(island_plane, island_plane + (size_t)W * H);
    // slice capacities: a cut line holds its core + its samples; an island holds at most one node per
    // sqrt(10)*r_th disc (the density gate of RTRT_Sampling, :891), bounded through its bounding box
    std::vector<double> sxy;
    std::vector<int> scut;
    cd_cut_samples(e, sxy, scut);
    e->cut_cap_off.assign(ncut + 1, 0);
    for (int c : scut) e->cut_cap_off[c + 1]++;
    for (int i = 0; i < ncut; i++) e->cut_cap_off[i + 1] += e->cut_cap_off[i] + 1;
    e->isl_cap_off.assign(npoly + 1, 0);
    const double spacing = std::max(1.0, sqrt(10.0) * r_th);
    for (int i = 0; i < npoly; i++) {
        double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
        for (int j = pv_off[i]; j < pv_off[i + 1]; j++)
            x0 = std::min(x0, pv[2 * j]), x1 = std::max(x1, pv[2 * j]), y0 = std::min(y0, pv[2 * j + 1]), y1 = std::max(y1, pv[2 * j + 1]);
        const double nx = (x1 - x0) / spacing * 1.16 + 2, ny = (y1 - y0) / spacing * 1.16 + 2;
        e->isl_cap_off[i + 1] = e->isl_cap_off[i] + (int)(nx * ny) + 4;
    }
    v.cut_cap_total = e->cut_cap_off[ncut], v.isl_cap_total = e->isl_cap_off[npoly];
    const int n_fixed = 1 + ncut + (int)scut.size();
    v.cap = std::max(node_capacity, n_fixed + 64);
    v.set_cap = 2 * v.cap;
    v.tilesX = (W + 31) >> 5;
    v.plane_words = (size_t)v.tilesX * ((H + 31) >> 5) * 32;
    const size_t N = n_instances;
    int rc = 0;
    rc |= cd_upload_const(e, &v.cp_off, e->cp_off), rc |= cd_upload_const(e, &v.cp, e->cp);
    rc |= cd_upload_const(e, &v.cc_off, e->cc_off), rc |= cd_upload_const(e, &v.cc, e->cc);
    rc |= cd_upload_const(e, &v.pc_off, e->pc_off), rc |= cd_upload_const(e, &v.pc, e->pc);
    rc |= cd_upload_const(e, &v.cut_cap_off, e->cut_cap_off), rc |= cd_upload_const(e, &v.isl_cap_off, e->isl_cap_off);
    rc |= cd_upload_const(e, &v.isl, e->isl);
    rc |= cd_alloc(e, &v.inst, N), rc |= cd_alloc(e, &v.xy, N * v.cap), rc |= cd_alloc(e, &v.cw, N * v.cap);
    rc |= cd_alloc(e, &v.par, N * v.cap), rc |= cd_alloc(e, &v.inv, N * v.cap), rc |= cd_alloc(e, &v.flag, N * v.cap);
    rc |= cd_alloc(e, &v.cut_cnt, N * ncut), rc |= cd_alloc(e, &v.cut_nodes, N * v.cut_cap_total);
    rc |= cd_alloc(e, &v.isl_cnt, N * npoly), rc |= cd_alloc(e, &v.isl_nodes, N * v.isl_cap_total);
    rc |= cd_alloc(e, &v.isl_tree, N * npoly), rc |= cd_alloc(e, &v.isl_mark, N * npoly);
    rc |= cd_alloc(e, &v.qkey, N * v.qcap, false), rc |= cd_alloc(e, &v.qse, N * v.qcap, false);
    rc |= cd_alloc(e, &v.larger, N * v.set_cap), rc |= cd_alloc(e, &v.smaller, N * v.set_cap);
    rc |= cd_alloc(e, &v.xnear, N * (v.cap + 64)), rc |= cd_alloc(e, &v.child_start, N * (v.cap + 2));
    rc |= cd_alloc(e, &v.children, N * v.cap), rc |= cd_alloc(e, &v.frontier, 2 * N * v.cap), rc |= cd_alloc(e, &v.path, N * v.path_cap);
    int *drand = nullptr;
    rc |= cd_alloc(e, &drand, N * v.rand_cap);
    v.rand = drand;
    rc |= cd_alloc(e, &v.pose, N * v.max_obs), rc |= cd_alloc(e, &v.chk, N * v.max_chk), rc |= cd_alloc(e, &v.chk_obs, N * v.max_chk);
    rc |= cd_alloc(e, &v.dplane, N * v.plane_words);
    CdHdr *dh = nullptr;
    rc |= cd_alloc(e, &dh, N);
    v.hdr = dh;
    if (rc) {
        for (void *p : e->allocs) cudaFree(p);
        delete e;
        return RTPP_ERR_NOMEM;
    }
    e->inst.resize(N);
    *out = e;
    return RTPP_OK;
}

int rtpp_cdrt_destroy(rtpp_cdrt_engine *e) {
    if (!e) return RTPP_ERR_INVALID;
    cudaStreamSynchronize(e->ctx->stream);
    for (void *p : e->allocs) cudaFree(p);
    delete e;
    return RTPP_OK;
}

// RTRT_SetAgent / RTRT_SetGoal, CDRT.cpp:384-423 (an unplaceable point is ignored, as there: returns 1)
int rtpp_cdrt_set_agent(rtpp_cdrt_engine *e, int inst, double x, double y) {
    if (!e || inst < 0 || inst >= e->v.n_inst) return RTPP_ERR_INVALID;
    if (!cd_placeable(e, x, y)) return 1;
    e->inst[inst].agent_tag[0] = x, e->inst[inst].agent_tag[1] = y;
    return RTPP_OK;
}
int rtpp_cdrt_set_goal(rtpp_cdrt_engine *e, int inst, double x, double y) {
    if (!e || inst < 0 || inst >= e->v.n_inst) return RTPP_ERR_INVALID;
    if (!cd_placeable(e, x, y)) return 1;
    e->inst[inst].goal[0] = x, e->inst[inst].goal[1] = y, e->inst[inst].goal_set = 1;
    return RTPP_OK;
}

// RTRT_AddDynamicObstacle, CDRT.cpp:425-500: stores the obstacle and derives its check points with
// the binary32 intermediates written there
int rtpp_cdrt_add_obstacle(rtpp_cdrt_engine *e, int inst, int num, int type, double x, double y, double angle, double radius,
                           const double *outline_xy, int npts, double check_len) {
    if (!e || inst < 0 || inst >= e->v.n_inst || npts < 0 || (npts && !outline_xy)) return RTPP_ERR_INVALID;
    CdObstacle &o = e->inst[inst].obs[num];
    o.type = type, o.x = x, o.y = y, o.angle = angle, o.radius = radius;
    o.outline.assign(outline_xy, outline_xy + 2 * npts);
    e->inst[inst].check_dirty = true;
    if (!(check_len > 0)) return RTPP_OK;
    o.check.clear();
    if (type == 0) {
        const float r = (float)radius, two_pi = 2 * 3.14159265f;
        const int n = (int)((double)(two_pi * r) / check_len);
        const float dangle = two_pi / (float)n;
        for (int p = 0; p < npts; p++)
            for (int i = 0; i < n; i++) {
                const float a = dangle * (float)i;
                const float cx = (float)((double)(sinf(a) * r) + outline_xy[2 * p]);
                const float cy = (float)((double)(cosf(a) * r) + outline_xy[2 * p + 1]);
                o.check.push_back((double)cx), o.check.push_back((double)cy);
            }
    } else if (type == 1) {
        for (int k = 0; k < npts; k++) {
            const double *a = outline_xy + 2 * k, *b = outline_xy + 2 * ((k + 1) % npts);
            const float dx = (float)(b[0] - a[0]), dy = (float)(b[1] - a[1]);
            const float len = sqrtf(dx * dx + dy * dy);
            const unsigned n = (unsigned)((double)len / check_len) + 1;
            for (unsigned i = 0; i < n; i++) {
                const float t = (float)i / (float)n;
                o.check.push_back(a[0] + (double)(t * dx)), o.check.push_back(a[1] + (double)(t * dy));
            }
        }
    }
    return RTPP_OK;
}
int rtpp_cdrt_del_obstacle(rtpp_cdrt_engine *e, int inst, int num) { // RTRT_DelDynamicObstacle :502-505
    if (!e || inst < 0 || inst >= e->v.n_inst) return RTPP_ERR_INVALID;
    e->inst[inst].obs.erase(num);
    e->inst[inst].check_dirty = true;
    return RTPP_OK;
}
// the demos move an obstacle by writing DObstacles[num].x/.y/.angle (CDRT2/test/testD.cpp:189-190)
int rtpp_cdrt_move_obstacle(rtpp_cdrt_engine *e, int inst, int num, double x, double y, double angle) {
    if (!e || inst < 0 || inst >= e->v.n_inst || !e->inst[inst].obs.count(num)) return RTPP_ERR_INVALID;
    CdObstacle &o = e->inst[inst].obs[num];
    o.x = x, o.y = y, o.angle = angle;
    return RTPP_OK;
}

// RTRT_start, CDRT.cpp:276-383.  rand: the injected rand() stream of this instance (the shuffle of the
// cut-line samples consumes the first draws here on the host, the tick kernel continues after them).
int rtpp_cdrt_start(rtpp_cdrt_engine *e, int inst, const int32_t *rand, int n_rand) {
    if (!e || inst < 0 || inst >= e->v.n_inst || n_rand < 0 || (n_rand && !rand) || n_rand > e->v.rand_cap) return RTPP_ERR_INVALID;
    rtpp_ctx *ctx = e->ctx;
    CdView &v = e->v;
    CdHostInst &hi = e->inst[inst];
    const double ax = hi.agent_tag[0], ay = hi.agent_tag[1];
    if (!cd_placeable(e, ax, ay)) RTPP_FAIL(ctx, RTPP_ERR_STATE, "RTRT_start: the agent is not on a free convex polygon");
    const int agent_island = cd_island_host(e, ax, ay);
    // samples + std::random_shuffle (libstdc++: swap(v[i], v[rand() % (i + 1)]), i = 1..n-1)
    std::vector<double> sxy;
    std::vector<int> scut;
    cd_cut_samples(e, sxy, scut);
    const int ns = (int)scut.size();
    std::vector<int> order(ns);
    for (int i = 0; i < ns; i++) order[i] = i;
    int rpos = 0;
    for (int i = 1; i < ns; i++) {
        const int r = rpos < n_rand ? rand[rpos] : 0;
        rpos++;
        const int j = r % (i + 1);
        if (i != j) std::swap(order[i], order[j]);
    }
    if (rpos > n_rand) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "rand stream too short for the sample shuffle");
    // node table: 0 = agent, 1..ncut = cut-line cores, then the samples in list order (not alive yet)
    const int n = 1 + v.ncut + ns;
    std::vector<double> xy(2 * (size_t)v.cap, 0.0);
    std::vector<u64> cw(v.cap, h_cw(false, 0.0));
    std::vector<int> par(v.cap, -1), inv(v.cap, 0);
    std::vector<unsigned char> flag(v.cap, 0);
    xy[0] = ax, xy[1] = ay, inv[0] = kOrigin, flag[0] = F_ALIVE;
    for (int c = 0; c < v.ncut; c++) {
        xy[2 * (1 + c)] = e->cut[c].core[0], xy[2 * (1 + c) + 1] = e->cut[c].core[1];
        cw[1 + c] = h_cw(true, kDoubleMax), inv[1 + c] = c, flag[1 + c] = F_ALIVE;
    }
    for (int s = 0; s < ns; s++) {
        const int src = order[s], id = 1 + v.ncut + s;
        xy[2 * id] = sxy[2 * src], xy[2 * id + 1] = sxy[2 * src + 1], inv[id] = scut[src];
    }
    // initial tree: Dijkstra over the cut-line cores from the agent (:343-382); no kd-trees exist yet, so
    // RTRT_Neighbor is: agent -> cores of its polygon's cut lines; core c -> cores of c's linked cut lines
    // (+ the agent when c borders the agent's polygon)
    struct QE {
        u64 c;
        int S, E;
    };
    auto pt_lt = [&](int a, int b) { return xy[2 * a] != xy[2 * b] ? xy[2 * a] < xy[2 * b] : xy[2 * a + 1] < xy[2 * b + 1]; };
    auto worse = [&](const QE &a, const QE &b) { // a has lower priority than b
        if (a.c != b.c) return a.c > b.c;
        if (a.S != b.S && (xy[2 * a.S] != xy[2 * b.S] || xy[2 * a.S + 1] != xy[2 * b.S + 1])) return pt_lt(b.S, a.S);
        return pt_lt(b.E, a.E);
    };
    std::priority_queue<QE, std::vector<QE>, decltype(worse)> pq(worse);
    std::vector<char> exist(n, 0);
    exist[0] = 1;
    auto cost_of = [&](u64 w) {
        u64 b = w & 0x7fffffffffffffffull;
        double d;
        memcpy(&d, &b, 8);
        return d;
    };
    auto neighbours = [&](int id, std::vector<int> &out) {
        out.clear();
        if (id == 0) {
            for (int j = e->pc_off[agent_island]; j < e->pc_off[agent_island + 1]; j++) out.push_back(1 + e->pc[j]);
        } else {
            const int c = id - 1;
            for (int j = e->cc_off[c]; j < e->cc_off[c + 1]; j++) out.push_back(1 + e->cc[j]);
            for (int j = e->cp_off[c]; j < e->cp_off[c + 1]; j++)
                if (e->cp[j] == agent_island) out.push_back(0);
        }
    };
    std::vector<int> near;
    neighbours(0, near);
    for (int q : near) pq.push({h_cw(false, h_dist64(&xy[0], &xy[2 * q])), 0, q});
    while (!pq.empty()) {
        const QE t = pq.top();
        pq.pop();
        if (exist[t.E]) continue;
        if (cw[t.E] > t.c) {
            par[t.E] = t.S, cw[t.E] = t.c;
            exist[t.E] = 1;
            neighbours(t.E, near);
            for (int q : near)
                if (!exist[q]) pq.push({h_cw(false, cost_of(t.c) + h_dist64(&xy[2 * t.E], &xy[2 * q])), t.E, q});
        }
    }
    // upload
    const size_t nb = (size_t)inst * v.cap;
    CD_CK(e, cudaMemcpyAsync(v.xy + nb, xy.data(), (size_t)v.cap * 16, cudaMemcpyHostToDevice, ctx->stream));
    CD_CK(e, cudaMemcpyAsync(v.cw + nb, cw.data(), (size_t)v.cap * 8, cudaMemcpyHostToDevice, ctx->stream));
    CD_CK(e, cudaMemcpyAsync(v.par + nb, par.data(), (size_t)v.cap * 4, cudaMemcpyHostToDevice, ctx->stream));
    CD_CK(e, cudaMemcpyAsync(v.inv + nb, inv.data(), (size_t)v.cap * 4, cudaMemcpyHostToDevice, ctx->stream));
    CD_CK(e, cudaMemcpyAsync(v.flag + nb, flag.data(), (size_t)v.cap, cudaMemcpyHostToDevice, ctx->stream));
    std::vector<int> cut_cnt(v.ncut, 1), cut_nodes(v.cut_cap_total, 0);
    for (int c = 0; c < v.ncut; c++) cut_nodes[e->cut_cap_off[c]] = 1 + c;
    CD_CK(e, cudaMemcpyAsync(v.cut_cnt + (size_t)inst * v.ncut, cut_cnt.data(), (size_t)v.ncut * 4, cudaMemcpyHostToDevice, ctx->stream));
    CD_CK(e, cudaMemcpyAsync(v.cut_nodes + (size_t)inst * v.cut_cap_total, cut_nodes.data(), (size_t)v.cut_cap_total * 4, cudaMemcpyHostToDevice,
                             ctx->stream));
    CD_CK(e, cudaMemsetAsync(v.isl_cnt + (size_t)inst * v.npoly, 0, (size_t)v.npoly * 4, ctx->stream));
    CD_CK(e, cudaMemsetAsync(v.isl_tree + (size_t)inst * v.npoly, 0, (size_t)v.npoly * 4, ctx->stream));
    CD_CK(e, cudaMemsetAsync(v.isl_mark + (size_t)inst * v.npoly, 0, (size_t)v.npoly * 4, ctx->stream));
    if (n_rand) CD_CK(e, cudaMemcpyAsync((int *)v.rand + (size_t)inst * v.rand_cap, rand, (size_t)n_rand * 4, cudaMemcpyHostToDevice, ctx->stream));
    CdInst g;
    memset(&g, 0, sizeof(g));
    g.n = n, g.agent_id = 0, g.agent_island = agent_island, g.goal_set = hi.goal_set;
    g.agent_tag = make_double2(ax, ay), g.goal = make_double2(hi.goal[0], hi.goal[1]);
    g.samp_n = ns, g.rand_pos = rpos, g.rand_n = n_rand;
    CD_CK(e, cudaMemcpyAsync(v.inst + inst, &g, sizeof(g), cudaMemcpyHostToDevice, ctx->stream));
    CD_CK(e, cudaStreamSynchronize(ctx->stream));
    hi.started = true;
    return RTPP_OK;
}

// a fresh rand() stream for a running instance (the next draw is rand[0])
int rtpp_cdrt_set_rand(rtpp_cdrt_engine *e, int inst, const int32_t *rand, int n_rand) {
    if (!e || inst < 0 || inst >= e->v.n_inst || n_rand < 0 || (n_rand && !rand) || n_rand > e->v.rand_cap) return RTPP_ERR_INVALID;
    rtpp_ctx *ctx = e->ctx;
    if (n_rand) CD_CK(e, cudaMemcpyAsync((int *)e->v.rand + (size_t)inst * e->v.rand_cap, rand, (size_t)n_rand * 4, cudaMemcpyHostToDevice, ctx->stream));
    const int hdr[3] = {0, n_rand, 0}; // rand_pos, rand_n, rand_exhausted are adjacent in CdInst
    CD_CK(e, cudaMemcpyAsync(&e->v.inst[inst].rand_pos, hdr, sizeof(hdr), cudaMemcpyHostToDevice, ctx->stream));
    CD_CK(e, cudaStreamSynchronize(ctx->stream));
    return RTPP_OK;
}

// RTRT_task for every started instance (+ GetPath when do_get_path): draws each instance's dynamic plane
// (RTRT_DrawDObsmap :605-636; the outline rotation uses the host's sinf/cosf and binary64 products as
// written at :617-618), ships the obstacle poses and agent/goal points, runs one kernel.
int rtpp_cdrt_tick(rtpp_cdrt_engine *e, long long dt_ns, int do_get_path) {
    if (!e || dt_ns <= 0) return RTPP_ERR_INVALID;
    rtpp_ctx *ctx = e->ctx;
    CdView &v = e->v;
    std::vector<double> &pose = e->h_pose, &chk = e->h_chk;
    std::vector<int> &chk_obs = e->h_chk_obs;
    pose.assign((size_t)v.n_inst * v.max_obs * 4, 0.0);
    chk.resize((size_t)v.n_inst * v.max_chk * 2);
    chk_obs.resize((size_t)v.n_inst * v.max_chk);
    std::vector<CdHdr> hdr(v.n_inst);
    std::vector<int32_t> circles, poly_xy, poly_sizes, poly_inst;
    for (int i = 0; i < v.n_inst; i++) {
        CdHostInst &hi = e->inst[i];
        CdHdr &h = hdr[i];
        memset(&h, 0, sizeof(h));
        if (!hi.started) continue;
        int k = 0;
        size_t nchk = 0;
        if ((int)hi.obs.size() > v.max_obs) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "too many dynamic obstacles");
        for (const auto &kv : hi.obs) {
            const CdObstacle &o = kv.second;
            const float s = sinf((float)o.angle), c = cosf((float)o.angle);
            double *ps = &pose[((size_t)i * v.max_obs + k) * 4];
            ps[0] = o.x, ps[1] = o.y, ps[2] = (double)s, ps[3] = (double)c;
            auto px = [&](double x, double y, int *ix, int *iy) {
                *ix = (int)((double)c * x - (double)s * y + o.x);
                *iy = (int)((double)s * x + (double)c * y + o.y);
            };
            if (o.type == 0) {
                for (size_t p = 0; p + 1 < o.outline.size(); p += 2) {
                    int ix, iy;
                    px(o.outline[p], o.outline[p + 1], &ix, &iy);
                    circles.push_back(ix), circles.push_back(iy), circles.push_back((int)o.radius), circles.push_back(i);
                }
            } else if (o.type == 1) {
                for (size_t p = 0; p + 1 < o.outline.size(); p += 2) {
                    int ix, iy;
                    px(o.outline[p], o.outline[p + 1], &ix, &iy);
                    poly_xy.push_back(ix), poly_xy.push_back(iy);
                }
                poly_sizes.push_back((int)o.outline.size() / 2), poly_inst.push_back(i);
            }
            if (nchk + o.check.size() / 2 > (size_t)v.max_chk) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "too many obstacle check points");
            for (size_t p = 0; p + 1 < o.check.size(); p += 2, nchk++) {
                chk[((size_t)i * v.max_chk + nchk) * 2] = o.check[p], chk[((size_t)i * v.max_chk + nchk) * 2 + 1] = o.check[p + 1];
                chk_obs[(size_t)i * v.max_chk + nchk] = k;
            }
            k++;
        }
        h.tag = make_double2(hi.agent_tag[0], hi.agent_tag[1]), h.goal = make_double2(hi.goal[0], hi.goal[1]);
        h.goal_set = hi.goal_set, h.n_obs = k, h.n_chk = (int)nchk, h.started = 1;
        if (hi.check_dirty) e->chk_dirty = true, hi.check_dirty = false;
    }
    int rc = rtpp_raster_batch(ctx, v.dplane, v.plane_words, v.n_inst, v.W, v.H, circles.data(), (int)circles.size() / 4, poly_xy.data(),
                               poly_sizes.data(), poly_inst.data(), (int)poly_sizes.size(), e->drawn_valid ? &e->drawn : nullptr);
    if (rc) {
        e->drawn_valid = false;
        return rc;
    }
    e->drawn.circles4.swap(circles), e->drawn.poly_xy.swap(poly_xy), e->drawn.poly_sizes.swap(poly_sizes), e->drawn.poly_inst.swap(poly_inst);
    e->drawn_valid = true;
    CD_CK(e, cudaMemcpyAsync((void *)v.hdr, hdr.data(), hdr.size() * sizeof(CdHdr), cudaMemcpyHostToDevice, ctx->stream));
    CD_CK(e, cudaMemcpyAsync(v.pose, pose.data(), pose.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (e->chk_dirty) { // the check points only change with RTRT_AddDynamicObstacle
        CD_CK(e, cudaMemcpyAsync(v.chk, chk.data(), chk.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
        CD_CK(e, cudaMemcpyAsync(v.chk_obs, chk_obs.data(), chk_obs.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
        e->chk_dirty = false;
    }
    cdrt_tick_kernel<<<(v.n_inst * 32 + 127) / 128, 128, 0, ctx->stream>>>(v, dt_ns, do_get_path);
    RTPP_LAUNCHED(ctx);
    CD_CK(e, cudaStreamSynchronize(ctx->stream));
    return RTPP_OK;
}

// 24 int64 per instance: n_nodes(ids), agent_id, agent_island, clock_reads, rand_pos, rand_exhausted, queue, larger, smaller,
// samples_left, n_trees, status, pops, rewired, sampled, pushes, edge_checks, path_len, path_collision, path_cost (binary64 bits), 0...
int rtpp_cdrt_read_stats(rtpp_cdrt_engine *e, long long *out24) {
    if (!e || !out24) return RTPP_ERR_INVALID;
    std::vector<CdInst> g(e->v.n_inst);
    CD_CK(e, cudaMemcpy(g.data(), e->v.inst, g.size() * sizeof(CdInst), cudaMemcpyDeviceToHost));
    for (int i = 0; i < e->v.n_inst; i++) {
        long long *o = out24 + 24 * (size_t)i;
        memset(o, 0, 24 * 8);
        const CdInst &s = g[i];
        o[0] = s.n, o[1] = s.agent_id, o[2] = s.agent_island, o[3] = s.clk_n, o[4] = s.rand_pos, o[5] = s.rand_exhausted, o[6] = s.q_n;
        o[7] = s.n_larger, o[8] = s.n_smaller, o[9] = s.samp_n - s.samp_pos, o[10] = s.n_trees, o[11] = s.status, o[12] = s.pops;
        o[13] = s.rewired, o[14] = s.sampled, o[15] = s.pushes, o[16] = s.checks, o[17] = s.path_len, o[18] = s.path_col;
        memcpy(&o[19], &s.path_cost, 8);
        for (int k = 0; k < 4; k++) o[20 + k] = s.t_phase[k];
    }
    return RTPP_OK;
}
// profiling read-out of the last tick's eQr loop (one instance): SM cycles in {pop, entry filter, edge check, neighbour list,
// pushes}, then the counts {edge checks, neighbours listed, rewires}
int rtpp_cdrt_read_cycles(rtpp_cdrt_engine *e, int inst, long long *out8) {
    if (!e || inst < 0 || inst >= e->v.n_inst || !out8) return RTPP_ERR_INVALID;
    CdInst g;
    CD_CK(e, cudaMemcpy(&g, e->v.inst + inst, sizeof(g), cudaMemcpyDeviceToHost));
    memcpy(out8, g.t_fine, sizeof(g.t_fine));
    return RTPP_OK;
}
// nodes by id (dead ones included: alive = 0)
int rtpp_cdrt_read_tree(rtpp_cdrt_engine *e, int inst, int max_nodes, double *xy, int32_t *parent, double *cost, uint8_t *collision,
                        int32_t *invM, uint8_t *alive, int *n_out) {
    if (!e || inst < 0 || inst >= e->v.n_inst || !n_out) return RTPP_ERR_INVALID;
    CdInst g;
    CD_CK(e, cudaMemcpy(&g, e->v.inst + inst, sizeof(g), cudaMemcpyDeviceToHost));
    *n_out = g.n;
    const int n = std::min(g.n, max_nodes);
    if (n <= 0) return RTPP_OK;
    const size_t nb = (size_t)inst * e->v.cap;
    if (xy) CD_CK(e, cudaMemcpy(xy, e->v.xy + nb, (size_t)n * 16, cudaMemcpyDeviceToHost));
    if (parent) CD_CK(e, cudaMemcpy(parent, e->v.par + nb, (size_t)n * 4, cudaMemcpyDeviceToHost));
    if (invM) CD_CK(e, cudaMemcpy(invM, e->v.inv + nb, (size_t)n * 4, cudaMemcpyDeviceToHost));
    if (cost || collision) {
        std::vector<u64> w(n);
        CD_CK(e, cudaMemcpy(w.data(), e->v.cw + nb, (size_t)n * 8, cudaMemcpyDeviceToHost));
        for (int i = 0; i < n; i++) {
            const u64 b = w[i] & 0x7fffffffffffffffull;
            if (cost) memcpy(&cost[i], &b, 8);
            if (collision) collision[i] = (uint8_t)(w[i] >> 63);
        }
    }
    if (alive) {
        CD_CK(e, cudaMemcpy(alive, e->v.flag + nb, (size_t)n, cudaMemcpyDeviceToHost));
        for (int i = 0; i < n; i++) alive[i] &= F_ALIVE;
    }
    return RTPP_OK;
}
// agent_path of the last GetPath as points (root ... entry node, goal); *len_out = number of points
int rtpp_cdrt_read_path(rtpp_cdrt_engine *e, int inst, int max_len, double *xy, int *len_out) {
    if (!e || inst < 0 || inst >= e->v.n_inst || !len_out) return RTPP_ERR_INVALID;
    CdInst g;
    CD_CK(e, cudaMemcpy(&g, e->v.inst + inst, sizeof(g), cudaMemcpyDeviceToHost));
    *len_out = g.path_len;
    const int n = std::min(g.path_len, max_len);
    if (n <= 0 || !xy) return RTPP_OK;
    std::vector<int> ids(n);
    CD_CK(e, cudaMemcpy(ids.data(), e->v.path + (size_t)inst * e->v.path_cap, (size_t)n * 4, cudaMemcpyDeviceToHost));
    std::vector<double> all(2 * (size_t)g.n); // one copy of the node coordinates, gathered on the host
    CD_CK(e, cudaMemcpy(all.data(), e->v.xy + (size_t)inst * e->v.cap, (size_t)g.n * 16, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; i++) {
        if (ids[i] == -2) xy[2 * i] = g.goal.x, xy[2 * i + 1] = g.goal.y;
        else xy[2 * i] = all[2 * (size_t)ids[i]], xy[2 * i + 1] = all[2 * (size_t)ids[i] + 1];
    }
    return RTPP_OK;
}
// islands that currently hold a kd-tree (obsKdtreeMap keys) and their point counts
int rtpp_cdrt_read_islands(rtpp_cdrt_engine *e, int inst, int32_t *has_tree, int32_t *count) {
    if (!e || inst < 0 || inst >= e->v.n_inst) return RTPP_ERR_INVALID;
    if (has_tree) CD_CK(e, cudaMemcpy(has_tree, e->v.isl_tree + (size_t)inst * e->v.npoly, (size_t)e->v.npoly * 4, cudaMemcpyDeviceToHost));
    if (count) CD_CK(e, cudaMemcpy(count, e->v.isl_cnt + (size_t)inst * e->v.npoly, (size_t)e->v.npoly * 4, cudaMemcpyDeviceToHost));
    return RTPP_OK;
}

} // extern "C"
