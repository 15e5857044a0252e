// engine_fmt.cu — sequential-exact RT-FMT engine: one warp runs one planner instance.
//
// Replaces (reference RTFMT/lib/RTFMT/RTFMT.cpp):
//   F1  ExpandFMT            :423-529   -> expand_fmt()
//   F2a RewireFromObstacles  :531-595   -> rewire_from_obstacles()
//   F2b RewireFromRoot       :665-703   -> rewire_from_root()
//   R3  UpdateTree           :92-176    -> fmt_update_tree()   (re-root, BFS re-cost, N_b capture)
//   P1  GetPath              :285-391   -> fmt_get_path()      (goal-entry branch; depth-5 fallback: path_fallback.cuh)
// The node set is fixed after Start (:393-422), so every Orthogonal_k_neighbor_search(.., 25) of
// the loops (:427,440,477,549) is a row of a k-NN table built once per instance (:677's k = 20 is
// the first 20 entries of the same sorted row).  The std::set<RTpoint> containers V_unv / V_open /
// V_opennew / V_closed / V_toopen become a state byte per node plus compact lists; wherever the
// reference iterates a std::set (argmin over V_open :490-497, N_b -> Qo :535-536, subs :573-577)
// its (x, y)-lexicographic order is reproduced through a precomputed rank.
// Known reference quirk kept: cost(x) = cost(y_min) + {0, cost_min} counts cost(y_min) twice
// (:448,460) until the next UpdateTree.
#include <float.h>
#include <math_constants.h>

#include <algorithm>
#include <numeric>
#include <vector>

#include "nodes.cuh"
#include "path_fallback.cuh"

#define FULL 0xffffffffu
static constexpr int FK = 25;

enum : uint8_t { ST_UNV = 0, ST_OPEN = 1, ST_OPENNEW = 2, ST_CLOSED = 3, ST_MASK = 3, ST_TOOPEN = 4 };

struct FmtInst {
    int map_slot, n, n_sorted;
    int root, z;
    float2 agent, goal;
    int n_unv, open_n, opennew_n, toopen_n;
    int xnear_head, xnear_n;
    int nb_n;
    int qo_head, qo_tail;
    int qr_head, qr_tail, qr_epoch;
    long long n_expand;
    int status; // 0 ok, 2 queue overflow
    int path_len, path_ok;
};

struct FmtView {
    int n_inst, cap, qo_cap, path_cap;
    int gw, gh, shift, ncells;
    FmtInst *inst;
    float2 *xy, *sxy;
    cpword *cp;
    uint8_t *state, *nb_flag;
    int *rank, *by_rank, *knn;
    int *open_list, *open_pos, *opennew_list, *toopen_list, *xnear;
    int *qo, *qr, *qr_mark;
    int *sid, *cell_of, *cell_start, *cursor;
    int *child_start, *children, *frontier;
    int *path;
    PlaneView planes_sd[16], planes_s[16], planes_d[16];
};

struct rtpp_fmt_engine {
    rtpp_ctx *ctx;
    FmtView v;
    uint32_t *plane_s[16], *plane_d[16], *plane_sd[16];
    int W[16], H[16];
    MarchThresholds mt;
};

struct FI { // per-warp view of one instance
    FmtInst s;
    float2 *xy, *sxy;
    cpword *cp;
    uint8_t *state, *nb_flag;
    int *rank, *by_rank, *knn, *open_list, *open_pos, *opennew_list, *toopen_list, *xnear, *qo, *qr, *qr_mark;
    int *sid, *cell_of, *cell_start, *cursor, *child_start, *children, *frontier, *path;
    PlaneView sd, st, dy;
};

__device__ __forceinline__ bool fmt_load(const FmtView &ev, int w, FI &in) {
    if (w >= ev.n_inst) return false;
    in.s = ev.inst[w];
    const size_t c = (size_t)w * ev.cap, g = (size_t)w * (ev.ncells + 1);
    in.xy = ev.xy + c, in.sxy = ev.sxy + c, in.cp = ev.cp + c, in.state = ev.state + c, in.nb_flag = ev.nb_flag + c;
    in.rank = ev.rank + c, in.by_rank = ev.by_rank + c, in.knn = ev.knn + c * FK;
    in.open_list = ev.open_list + c, in.open_pos = ev.open_pos + c, in.opennew_list = ev.opennew_list + (size_t)w * 32;
    in.toopen_list = ev.toopen_list + c, in.xnear = ev.xnear + (size_t)w * 32;
    in.qo = ev.qo + (size_t)w * ev.qo_cap, in.qr = ev.qr + c, in.qr_mark = ev.qr_mark + c;
    in.sid = ev.sid + c, in.cell_of = ev.cell_of + c, in.cell_start = ev.cell_start + g, in.cursor = ev.cursor + g;
    in.child_start = ev.child_start + (size_t)w * (ev.cap + 1), in.children = ev.children + c, in.frontier = ev.frontier + 2 * c;
    in.path = ev.path + (size_t)w * ev.path_cap;
    in.sd = ev.planes_sd[in.s.map_slot & 15], in.st = ev.planes_s[in.s.map_slot & 15], in.dy = ev.planes_d[in.s.map_slot & 15];
    return true;
}

__device__ __forceinline__ GridView fmt_grid(const FmtView &ev, const FI &in) {
    GridView g;
    g.sxy = in.sxy, g.sid = in.sid, g.cell_start = in.cell_start, g.xy = in.xy;
    g.gw = ev.gw, g.gh = ev.gh, g.shift = ev.shift, g.n_sorted = in.s.n_sorted, g.n = in.s.n;
    return g;
}

// lane j < FK gets entry j of node x's neighbour row (-1 = none)
__device__ __forceinline__ int knn_row(const FI &in, int x) {
    const int lane = threadIdx.x & 31;
    return lane < FK ? in.knn[(size_t)x * FK + lane] : -1;
}

__device__ __forceinline__ void open_insert(FI &in, int id) { // V_open.insert(id)
    in.open_list[in.s.open_n] = id;
    in.open_pos[id] = in.s.open_n;
}
__device__ __forceinline__ void open_erase_lane0(FI &in, int id) { // swap-remove
    const int pos = in.open_pos[id], last = in.open_list[in.s.open_n - 1];
    in.open_list[pos] = last;
    in.open_pos[last] = pos;
}

// z = argmin over V_open of (collision, cost), ties in std::set order (RTFMT.cpp:488-499)
__device__ int argmin_open(FI &in) {
    const int lane = threadIdx.x & 31;
    unsigned long long best = ~0ull;
    for (int i = lane; i < in.s.open_n; i += 32) {
        const int id = in.open_list[i];
        const unsigned long long key = ((unsigned long long)cp_hi(in.cp[id]) << 32) | (unsigned)in.rank[id];
        best = key < best ? key : best;
    }
    for (int off = 16; off; off >>= 1) {
        const unsigned long long o = __shfl_xor_sync(FULL, best, off);
        best = o < best ? o : best;
    }
    if ((uint32_t)(best >> 32) >= cost_pack(true, 3.4e38f)) return -1; // nothing below {1, 3.4e38}
    return in.by_rank[(int)(uint32_t)best];
}

// ---------------------------------------------------------------------------------------------
// F1 ExpandFMT (RTFMT.cpp:423-529)
// ---------------------------------------------------------------------------------------------
__device__ void expand_fmt(FI &in, const MarchThresholds &mt) {
    const int lane = threadIdx.x & 31;
    in.s.n_expand++;
    if (in.s.n_unv > 0 && in.s.xnear_n == 0 && in.s.z >= 0) { // :425-434
        const int s = knn_row(in, in.s.z);
        const bool unv = s >= 0 && (in.state[s] & ST_MASK) == ST_UNV;
        const unsigned m = __ballot_sync(FULL, unv);
        if (unv) in.xnear[__popc(m & ((1u << lane) - 1))] = s;
        in.s.xnear_head = 0, in.s.xnear_n = __popc(m);
        __syncwarp();
    }
    if (in.s.xnear_n > 0) { // :436-464
        const int x = in.xnear[in.s.xnear_head];
        in.s.xnear_head++, in.s.xnear_n--;
        const float2 px = in.xy[x];
        const int y = knn_row(in, x);
        unsigned long long key = ~0ull;
        if (y >= 0 && (in.state[y] & ST_MASK) == ST_OPEN) {
            const float2 py = in.xy[y];
            const float cost_now = __fadd_rn(cost_val(cp_hi(in.cp[y])), rt_dist(py.x, py.y, px.x, px.y));
            if (3.4e38f > cost_now) key = ((unsigned long long)__float_as_uint(cost_now) << 32) | (unsigned)lane;
        }
        for (int off = 16; off; off >>= 1) {
            const unsigned long long o = __shfl_xor_sync(FULL, key, off);
            key = o < key ? o : key;
        }
        if (key != ~0ull) {
            const int y_min = __shfl_sync(FULL, y, (int)(key & 31u));
            const float cost_min = __uint_as_float((uint32_t)(key >> 32));
            const float2 py = in.xy[y_min];
            const uint32_t hy = cp_hi(in.cp[y_min]);
            if (!rt_edge_blocked_warp(in.sd, px.x, px.y, py.x, py.y, mt) && !cost_col(hy)) {
                if (lane == 0) {
                    in.state[x] = ST_OPENNEW | (in.state[x] & ST_TOOPEN);
                    in.opennew_list[in.s.opennew_n] = x;
                    in.cp[x] = cp_make(cost_add(hy, cost_pack(false, cost_min)), y_min); // quirk: cost(y) counted twice
                }
                in.s.opennew_n++, in.s.n_unv--;
                __syncwarp();
            }
        }
    }
    if (in.s.xnear_n == 0 && in.s.z >= 0) { // :467-500
        const int z = in.s.z;
        if (lane == 0) {
            for (int i = 0; i < in.s.opennew_n; i++) { // V_open.insert(V_opennew)
                const int id = in.opennew_list[i];
                in.state[id] = ST_OPEN | (in.state[id] & ST_TOOPEN);
                in.open_list[in.s.open_n + i] = id;
                in.open_pos[id] = in.s.open_n + i;
            }
        }
        in.s.open_n += in.s.opennew_n, in.s.opennew_n = 0;
        __syncwarp();
        if ((in.state[z] & ST_MASK) == ST_OPEN) { // V_open.erase(z)
            if (lane == 0) open_erase_lane0(in, z);
            in.s.open_n--;
        }
        __syncwarp();
        if (lane == 0) in.state[z] = ST_CLOSED | (in.state[z] & ST_TOOPEN); // V_closed.insert(z)
        __syncwarp();
        // z stays interesting if an unvisited neighbour is reachable ignoring moving obstacles (:477-486)
        const float2 pz = in.xy[z];
        const int s = knn_row(in, z);
        bool reach = false;
        if (s >= 0 && (in.state[s] & ST_MASK) == ST_UNV) {
            const float2 q = in.xy[s];
            reach = !rt_edge_blocked(in.st, pz.x, pz.y, q.x, q.y, mt);
        }
        if (__any_sync(FULL, reach) && !(in.state[z] & ST_TOOPEN)) {
            if (lane == 0) {
                in.state[z] |= ST_TOOPEN;
                in.toopen_list[in.s.toopen_n] = z;
            }
            in.s.toopen_n++;
        }
        __syncwarp();
        in.s.z = argmin_open(in);
    }
    if (in.s.z < 0) { // :502-524
        if (lane == 0) {
            for (int i = 0; i < in.s.toopen_n; i++) {
                const int id = in.toopen_list[i];
                if ((in.state[id] & ST_MASK) != ST_OPEN) {
                    in.open_list[in.s.open_n] = id;
                    in.open_pos[id] = in.s.open_n;
                    in.s.open_n++;
                }
                in.state[id] = ST_OPEN; // leaves V_closed, flag cleared
            }
        }
        in.s.open_n = __shfl_sync(FULL, in.s.open_n, 0);
        in.s.toopen_n = 0;
        __syncwarp();
        in.s.z = argmin_open(in);
    }
}

__device__ __forceinline__ void qo_push(const FmtView &ev, FI &in, bool want, int id) {
    const int lane = threadIdx.x & 31;
    const unsigned m = __ballot_sync(FULL, want);
    if (want) in.qo[(unsigned)(in.s.qo_tail + __popc(m & ((1u << lane) - 1))) % (unsigned)ev.qo_cap] = id;
    in.s.qo_tail += __popc(m);
    if (in.s.qo_tail - in.s.qo_head > ev.qo_cap) in.s.status = 2;
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// F2a RewireFromObstacles (RTFMT.cpp:531-595)
// ---------------------------------------------------------------------------------------------
__device__ void rewire_from_obstacles(const FmtView &ev, FI &in, const MarchThresholds &mt) {
    const int lane = threadIdx.x & 31;
    if (in.s.qo_tail == in.s.qo_head) {
        if (in.s.nb_n > 0) { // N_b is a std::set: pushed in (x,y) order = rank order
            for (int r0 = 0; r0 < in.s.n; r0 += 32) {
                const int r = r0 + lane;
                int id = -1;
                if (r < in.s.n) id = in.by_rank[r];
                const bool f = id >= 0 && in.nb_flag[id];
                if (f) in.nb_flag[id] = 0;
                qo_push(ev, in, f, id);
            }
            in.s.nb_n = 0;
        }
        return;
    }
    const int xb = in.qo[(unsigned)in.s.qo_head % (unsigned)ev.qo_cap];
    in.s.qo_head++;
    const float2 pb = in.xy[xb];
    if (rt_probe(in.dy, __float2int_rz(pb.y), __float2int_rz(pb.x))) return; // DObsmap(x_b) != 0
    const int y = knn_row(in, xb);
    unsigned long long key = ~0ull;
    if (y >= 0 && y != xb) {
        const uint8_t st = in.state[y] & ST_MASK;
        const uint32_t hy = cp_hi(in.cp[y]);
        if ((st == ST_OPEN || st == ST_CLOSED) && !cost_col(hy)) {
            const float2 py = in.xy[y];
            const float cost_now = __fadd_rn(cost_val(hy), rt_dist(py.x, py.y, pb.x, pb.y));
            if (cost_now < 3.4e38f) key = ((unsigned long long)__float_as_uint(cost_now) << 32) | (unsigned)lane;
        }
    }
    for (int off = 16; off; off >>= 1) {
        const unsigned long long o = __shfl_xor_sync(FULL, key, off);
        key = o < key ? o : key;
    }
    if (key == ~0ull) return; // y_min = {-1,-1}: the reference's CrashDetection reads outside the buffer -> blocked
    const int y_min = __shfl_sync(FULL, y, (int)(key & 31u));
    const float cost_min = __uint_as_float((uint32_t)(key >> 32));
    const float2 py = in.xy[y_min];
    if (rt_edge_blocked_warp(in.sd, py.x, py.y, pb.x, pb.y, mt)) return;
    if (!(cost_min < cost_val(cp_hi(in.cp[xb])))) return;
    const uint32_t pc = cost_add(cp_hi(in.cp[y_min]), cost_pack(false, rt_dist(pb.x, pb.y, py.x, py.y)));
    if (lane == 0) in.cp[xb] = cp_make(pc, y_min);
    __syncwarp();
    // direct children of x_b in std::set order: re-cost and queue them (:572-577)
    for (int r0 = 0; r0 < in.s.n; r0 += 32) {
        const int r = r0 + lane;
        int id = -1;
        if (r < in.s.n) id = in.by_rank[r];
        const bool child = id >= 0 && id != xb && cp_parent(in.cp[id]) == xb;
        if (child) {
            const float2 q = in.xy[id];
            const uint32_t w = cost_pack(rt_edge_blocked(in.sd, pb.x, pb.y, q.x, q.y, mt), rt_dist(pb.x, pb.y, q.x, q.y));
            in.cp[id] = cp_make(cost_add(pc, w), xb);
        }
        qo_push(ev, in, child, id);
    }
}

// ---------------------------------------------------------------------------------------------
// F2b RewireFromRoot (RTFMT.cpp:665-703), k = 20
// ---------------------------------------------------------------------------------------------
__device__ void rewire_from_root(FI &in, const MarchThresholds &mt) {
    const int lane = threadIdx.x & 31;
    if (in.s.qr_tail == in.s.qr_head) {
        in.s.qr_epoch++;
        in.s.qr_head = 0, in.s.qr_tail = 1;
        if (lane == 0) {
            in.qr[0] = in.s.root;
            in.qr_mark[in.s.root] = in.s.qr_epoch;
        }
        __syncwarp();
        return;
    }
    const int xb = in.qr[in.s.qr_head];
    in.s.qr_head++;
    const float2 pb = in.xy[xb];
    const int s = lane < 20 ? knn_row(in, xb) : -1;
    bool member = false;
    if (s >= 0 && s != xb) {
        const uint8_t st = in.state[s] & ST_MASK;
        member = st == ST_OPEN || st == ST_CLOSED;
    }
    if (member) {
        const float2 q = in.xy[s];
        const uint32_t w = cost_pack(rt_edge_blocked(in.sd, pb.x, pb.y, q.x, q.y, mt), rt_dist(pb.x, pb.y, q.x, q.y));
        const uint32_t c_new = cost_add(cp_hi(in.cp[xb]), w);
        if (c_new < cp_hi(in.cp[s])) in.cp[s] = cp_make(c_new, xb);
    }
    const bool fresh = member && in.qr_mark[s] != in.s.qr_epoch;
    const unsigned m = __ballot_sync(FULL, fresh);
    if (fresh) {
        in.qr[in.s.qr_tail + __popc(m & ((1u << lane) - 1))] = s;
        in.qr_mark[s] = in.s.qr_epoch;
    }
    in.s.qr_tail += __popc(m);
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// R3 UpdateTree with N_b capture (RTFMT.cpp:92-176)
// ---------------------------------------------------------------------------------------------
__device__ void fmt_update_tree(const FmtView &ev, FI &in, const MarchThresholds &mt) {
    const int lane = threadIdx.x & 31;
    const int n = in.s.n;
    const float2 rootp = in.xy[in.s.root];
    if (in.s.agent.x != rootp.x || in.s.agent.y != rootp.y) {
        KnnEntry e;
        knn_warp(fmt_grid(ev, in), (double)in.s.agent.x, (double)in.s.agent.y, 1, e);
        const int near = __shfl_sync(FULL, e.id, 0);
        if (near != in.s.root) {
            if (lane == 0) {
                int x = near, x_old = -1, guard = 0;
                while (x != in.s.root && x >= 0 && guard++ < n) {
                    const cpword w = in.cp[x];
                    in.cp[x] = cp_make(cp_hi(w), x_old);
                    x_old = x;
                    x = cp_parent(w);
                }
                if (x >= 0) in.cp[x] = cp_make(cp_hi(in.cp[x]), x_old);
                in.cp[near] = cp_make(cost_pack(false, 0.0f), -1);
            }
            __syncwarp();
            in.s.root = near;
        }
    }
    for (int i = lane; i <= n; i += 32) in.child_start[i] = 0;
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
        const int p = cp_parent(in.cp[i]);
        if (i != in.s.root && p >= 0) atomicAdd(&in.child_start[p], 1);
    }
    __syncwarp();
    int carry = 0;
    for (int base = 0; base <= n; base += 32) {
        const int c = base + lane;
        const int v = c <= n ? in.child_start[c] : 0;
        int x = v;
        for (int off = 1; off < 32; off <<= 1) {
            const int y = __shfl_up_sync(FULL, x, off);
            if (lane >= off) x += y;
        }
        if (c <= n) in.child_start[c] = carry + x - v;
        carry += __shfl_sync(FULL, x, 31);
    }
    __syncwarp();
    int *fill = in.cell_of;
    for (int i = lane; i < n; i += 32) fill[i] = 0;
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
        const int p = cp_parent(in.cp[i]);
        if (i != in.s.root && p >= 0) in.children[in.child_start[p] + atomicAdd(&fill[p], 1)] = i;
    }
    __syncwarp();
    int *cur = in.frontier, *nxt = in.frontier + ev.cap;
    if (lane == 0) cur[0] = in.s.root;
    int cur_n = 1, nb_add = 0;
    __syncwarp();
    while (cur_n > 0) {
        int nxt_n = 0;
        for (int f0 = 0; f0 < cur_n; f0 += 32) {
            const int f = f0 + lane;
            int s = 0, e = 0, p = -1;
            uint32_t hp = 0;
            float2 pp = make_float2(0.f, 0.f);
            if (f < cur_n) {
                p = cur[f];
                s = in.child_start[p], e = in.child_start[p + 1];
                hp = cp_hi(in.cp[p]);
                pp = in.xy[p];
            }
            const int cnt = e - s;
            int x = cnt;
            for (int off = 1; off < 32; off <<= 1) {
                const int y = __shfl_up_sync(FULL, x, off);
                if (lane >= off) x += y;
            }
            const int slot = nxt_n + x - cnt;
            for (int j = 0; j < cnt; j++) {
                const int c = in.children[s + j];
                const float2 pc = in.xy[c];
                const uint32_t w = cost_pack(rt_edge_blocked(in.sd, pp.x, pp.y, pc.x, pc.y, mt), rt_dist(pp.x, pp.y, pc.x, pc.y));
                const uint32_t hn = cost_add(hp, w);
                if (cost_col(hn) && !cost_col(cp_hi(in.cp[c])) && !in.nb_flag[c]) { // N_b.insert(sub), :170-171
                    in.nb_flag[c] = 1;
                    nb_add++;
                }
                in.cp[c] = cp_make(hn, p);
                nxt[slot + j] = c;
            }
            nxt_n += __shfl_sync(FULL, x, 31);
        }
        __syncwarp();
        int *t = cur;
        cur = nxt, nxt = t;
        cur_n = nxt_n;
    }
    for (int off = 16; off; off >>= 1) nb_add += __shfl_xor_sync(FULL, nb_add, off);
    in.s.nb_n += nb_add;
}

// P1 GetPath, goal-entry branch (RTFMT.cpp:285-320): best entry among the 25-NN of the goal with
// c_min < {0, 20000}; path_ok = 0 leaves the depth-5 fallback (:321-357) to the host.
__device__ void fmt_get_path(const FmtView &ev, FI &in, const MarchThresholds &mt) {
    const int lane = threadIdx.x & 31;
    in.s.path_len = 0, in.s.path_ok = 0;
    if (!(in.s.goal.x > 0.0f)) return;
    const float2 g = in.s.goal;
    KnnEntry e;
    const int found = knn_warp(fmt_grid(ev, in), (double)g.x, (double)g.y, 25, e);
    unsigned long long key = ((unsigned long long)cost_pack(true, FLT_MAX) << 32) | 63u;
    if (lane < found) {
        const float2 q = in.xy[e.id];
        const uint32_t w = cost_pack(rt_edge_blocked(in.sd, g.x, g.y, q.x, q.y, mt), rt_dist(g.x, g.y, q.x, q.y));
        const uint32_t c = cost_add(cp_hi(in.cp[e.id]), w);
        if (c < cost_pack(true, FLT_MAX)) key = ((unsigned long long)c << 32) | (unsigned)lane;
    }
    for (int off = 16; off; off >>= 1) {
        const unsigned long long o = __shfl_xor_sync(FULL, key, off);
        key = o < key ? o : key;
    }
    if ((uint32_t)(key >> 32) < cost_pack(false, 20000.0f)) {
        const int x_min = __shfl_sync(FULL, e.id, (int)(key & 31u));
        int len = 0, bad = 0;
        if (lane == 0) {
            in.path[len++] = -2;
            int x = x_min;
            while (x >= 0) {
                if (len < ev.path_cap) in.path[len] = x;
                len++;
                x = cp_parent(in.cp[x]);
                if (len > 2000) {
                    bad = 1;
                    break;
                }
            }
        }
        len = __shfl_sync(FULL, len, 0), bad = __shfl_sync(FULL, bad, 0);
        in.s.path_ok = 1;
        in.s.path_len = bad ? 0 : len;
    } else { // RTFMT.cpp:322-391: the chain of the reachable node nearest the goal (path_ok stays 0: no goal point in it)
        const int len = path_fallback_warp(in.xy, in.cp, in.s.n, in.s.root, g, in.path, ev.path_cap);
        in.s.path_len = len > 0 ? len : 0;
    }
}

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) fmt_tick_kernel(FmtView ev, MarchThresholds mt, int do_update, int iters, int phases,
                                                      int do_path) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    FI in;
    if (!fmt_load(ev, w, in)) return;
    if (in.s.n > 0) {
        if (do_update) fmt_update_tree(ev, in, mt);
        for (int i = 0; i < iters && !in.s.status; i++) { // RTFMT.cpp:714-719
            if (phases & 1) expand_fmt(in, mt);
            if (phases & 2) rewire_from_obstacles(ev, in, mt);
            if (phases & 4) rewire_from_root(in, mt);
        }
        if (do_path) fmt_get_path(ev, in, mt);
    }
    if ((threadIdx.x & 31) == 0) ev.inst[w] = in.s;
}

// Start (RTFMT.cpp:393-422) on the device: node 0 = root (cost {0,0}, open), the rest unvisited
// {1, 3.4e38}; per-instance hash built by the instance's warp.
__global__ void __launch_bounds__(128) fmt_start_kernel(FmtView ev, const int *__restrict__ n_nodes, const int *__restrict__ slots,
                                                       const float2 *__restrict__ goal) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= ev.n_inst) return;
    FI in;
    ev.inst[w] = FmtInst{};
    fmt_load(ev, w, in);
    const int n = n_nodes[w];
    in.s.map_slot = slots[w], in.s.n = n, in.s.root = 0, in.s.z = 0;
    in.s.agent = in.xy[0], in.s.goal = goal[w];
    in.s.n_unv = n - 1, in.s.open_n = 1;
    for (int i = lane; i < n; i += 32) {
        in.cp[i] = i == 0 ? cp_make(cost_pack(false, 0.0f), -1) : cp_make(cost_pack(true, 3.4e38f), -1);
        in.state[i] = i == 0 ? ST_OPEN : ST_UNV;
        in.nb_flag[i] = 0, in.qr_mark[i] = 0;
        in.by_rank[in.rank[i]] = i;
    }
    if (lane == 0) in.open_list[0] = 0, in.open_pos[0] = 0;
    // hash: counting sort by cell
    const int m = ev.ncells;
    for (int c = lane; c <= m; c += 32) in.cursor[c] = 0;
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
        const float2 p = in.xy[i];
        const int c = grid_cell_of((double)p.y, ev.shift, ev.gh - 1) * ev.gw + grid_cell_of((double)p.x, ev.shift, ev.gw - 1);
        in.cell_of[i] = c;
        atomicAdd(&in.cursor[c], 1);
    }
    __syncwarp();
    int carry = 0;
    for (int base = 0; base <= m; base += 32) {
        const int c = base + lane;
        const int v = c <= m ? in.cursor[c] : 0;
        int x = v;
        for (int off = 1; off < 32; off <<= 1) {
            const int y = __shfl_up_sync(FULL, x, off);
            if (lane >= off) x += y;
        }
        if (c <= m) in.cell_start[c] = carry + x - v, in.cursor[c] = carry + x - v;
        carry += __shfl_sync(FULL, x, 31);
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
        const int pos = atomicAdd(&in.cursor[in.cell_of[i]], 1);
        in.sxy[pos] = in.xy[i];
        in.sid[pos] = i;
    }
    __syncwarp();
    in.s.n_sorted = n;
    if (lane == 0) ev.inst[w] = in.s;
}

// the static neighbour table: one warp per (instance, node)
__global__ void __launch_bounds__(256) fmt_knn_table_kernel(FmtView ev) {
    const int lane = threadIdx.x & 31;
    const size_t warp0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    const size_t total = (size_t)ev.n_inst * ev.cap;
    for (size_t t = warp0; t < total; t += nwarps) {
        const int w = (int)(t / ev.cap), i = (int)(t % ev.cap);
        FI in;
        fmt_load(ev, w, in);
        if (i >= in.s.n) continue;
        const float2 p = in.xy[i];
        KnnEntry e;
        const int found = knn_warp(fmt_grid(ev, in), (double)p.x, (double)p.y, FK, e);
        if (lane < FK) in.knn[(size_t)i * FK + lane] = lane < found ? e.id : -1;
    }
}

__global__ void fmt_set_agents_kernel(FmtView ev, const float2 *__restrict__ agent, const float2 *__restrict__ goal) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= ev.n_inst) return;
    if (agent) ev.inst[w].agent = agent[w];
    if (goal) ev.inst[w].goal = goal[w];
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
#define FMT_CK(e, call) RTPP_CK((e)->ctx, call)
template <class T> static int fmt_alloc(rtpp_fmt_engine *e, T **p, size_t count) {
    FMT_CK(e, cudaMalloc((void **)p, count * sizeof(T)));
    FMT_CK(e, cudaMemsetAsync(*p, 0, count * sizeof(T), e->ctx->stream));
    return RTPP_OK;
}

extern "C" {

int rtpp_fmt_create(rtpp_ctx *ctx, int n_instances, int node_capacity, int cell_log2, int domain_w, int domain_h,
                    rtpp_fmt_engine **out) {
    if (!ctx || !out) return RTPP_ERR_INVALID;
    *out = nullptr;
    if (n_instances < 1 || node_capacity < 32 || cell_log2 < 2 || cell_log2 > 10 || domain_w < 1 || domain_h < 1)
        RTPP_FAIL(ctx, RTPP_ERR_INVALID, "bad engine geometry");
    RTPP_CK(ctx, cudaSetDevice(ctx->device));
    rtpp_fmt_engine *e = new (std::nothrow) rtpp_fmt_engine();
    if (!e) return RTPP_ERR_NOMEM;
    memset(&e->v, 0, sizeof(e->v));
    for (int i = 0; i < 16; i++) e->plane_s[i] = e->plane_d[i] = e->plane_sd[i] = nullptr, e->W[i] = e->H[i] = 0;
    e->ctx = ctx;
    FmtView &v = e->v;
    v.n_inst = n_instances, v.cap = node_capacity, v.qo_cap = 8 * node_capacity, v.path_cap = 2048, v.shift = cell_log2;
    v.gw = ((domain_w - 1) >> cell_log2) + 1, v.gh = ((domain_h - 1) >> cell_log2) + 1, v.ncells = v.gw * v.gh;
    rtpp_march_thresholds(5, &e->mt);
    const size_t N = (size_t)n_instances, C = N * node_capacity, G = N * (v.ncells + 1);
    int rc = 0;
    rc |= fmt_alloc(e, &v.inst, N), rc |= fmt_alloc(e, &v.xy, C), rc |= fmt_alloc(e, &v.sxy, C), rc |= fmt_alloc(e, &v.cp, C);
    rc |= fmt_alloc(e, &v.state, C), rc |= fmt_alloc(e, &v.nb_flag, C), rc |= fmt_alloc(e, &v.rank, C), rc |= fmt_alloc(e, &v.by_rank, C);
    rc |= fmt_alloc(e, &v.knn, C * FK), rc |= fmt_alloc(e, &v.open_list, C), rc |= fmt_alloc(e, &v.open_pos, C);
    rc |= fmt_alloc(e, &v.opennew_list, N * 32), rc |= fmt_alloc(e, &v.toopen_list, C), rc |= fmt_alloc(e, &v.xnear, N * 32);
    rc |= fmt_alloc(e, &v.qo, N * v.qo_cap), rc |= fmt_alloc(e, &v.qr, C), rc |= fmt_alloc(e, &v.qr_mark, C);
    rc |= fmt_alloc(e, &v.sid, C), rc |= fmt_alloc(e, &v.cell_of, C), rc |= fmt_alloc(e, &v.cell_start, G), rc |= fmt_alloc(e, &v.cursor, G);
    rc |= fmt_alloc(e, &v.child_start, N * (node_capacity + 1)), rc |= fmt_alloc(e, &v.children, C), rc |= fmt_alloc(e, &v.frontier, 2 * C);
    rc |= fmt_alloc(e, &v.path, N * v.path_cap);
    if (rc) {
        rtpp_fmt_destroy(e); // frees what was allocated (cudaFree(nullptr) is a no-op)
        RTPP_FAIL(ctx, RTPP_ERR_NOMEM, "engine buffers do not fit the device");
    }
    *out = e;
    return RTPP_OK;
}

int rtpp_fmt_destroy(rtpp_fmt_engine *e) {
    if (!e) return RTPP_ERR_INVALID;
    cudaStreamSynchronize(e->ctx->stream);
    FmtView &v = e->v;
    void *ptrs[] = {v.inst, v.xy, v.sxy, v.cp, v.state, v.nb_flag, v.rank, v.by_rank, v.knn, v.open_list, v.open_pos, v.opennew_list,
                    v.toopen_list, v.xnear, v.qo, v.qr, v.qr_mark, v.sid, v.cell_of, v.cell_start, v.cursor, v.child_start, v.children,
                    v.frontier, v.path};
    for (void *p : ptrs) cudaFree(p);
    for (int i = 0; i < 16; i++) cudaFree(e->plane_s[i]), cudaFree(e->plane_d[i]), cudaFree(e->plane_sd[i]);
    delete e;
    return RTPP_OK;
}

int rtpp_fmt_set_map(rtpp_fmt_engine *e, int slot, int plane, const uint8_t *src, int W, int H) {
    if (!e || slot < 0 || slot >= 16 || (plane != 1 && plane != 2)) return RTPP_ERR_INVALID;
    rtpp_ctx *ctx = e->ctx;
    if (plane == 1 && !src) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "static plane needs data");
    if (plane == 2 && !e->plane_s[slot]) RTPP_FAIL(ctx, RTPP_ERR_STATE, "set the static plane of the slot first");
    if (plane == 2) W = e->W[slot], H = e->H[slot];
    const int tilesX = (W + 31) >> 5, tilesY = (H + 31) >> 5;
    const size_t words = (size_t)tilesX * tilesY * 32;
    if (plane == 1) {
        cudaFree(e->plane_s[slot]), cudaFree(e->plane_d[slot]), cudaFree(e->plane_sd[slot]);
        FMT_CK(e, cudaMalloc(&e->plane_s[slot], words * 4));
        FMT_CK(e, cudaMalloc(&e->plane_d[slot], words * 4));
        FMT_CK(e, cudaMalloc(&e->plane_sd[slot], words * 4));
        FMT_CK(e, cudaMemsetAsync(e->plane_d[slot], 0, words * 4, ctx->stream));
        e->W[slot] = W, e->H[slot] = H;
    }
    uint32_t *dst = plane == 1 ? e->plane_s[slot] : e->plane_d[slot];
    if (src) {
        void *d = nullptr;
        int rc = rtpp_scratch(ctx, 0, (size_t)W * H, &d);
        if (rc) return rc;
        FMT_CK(e, cudaMemcpyAsync(d, src, (size_t)W * H, cudaMemcpyHostToDevice, ctx->stream));
        rc = rtpp_plane_pack(ctx, (const uint8_t *)d, W, H, dst);
        if (rc) return rc;
    } else
        FMT_CK(e, cudaMemsetAsync(dst, 0, words * 4, ctx->stream));
    {
        const int rcm = rtpp_plane_merge(ctx, e->plane_s[slot], e->plane_d[slot], e->plane_sd[slot], words);
        if (rcm) return rcm;
    }
    FMT_CK(e, cudaStreamSynchronize(ctx->stream));
    PlaneView pv;
    pv.W = W, pv.H = H, pv.tilesX = tilesX;
    pv.w = e->plane_sd[slot], e->v.planes_sd[slot] = pv;
    pv.w = e->plane_s[slot], e->v.planes_s[slot] = pv;
    pv.w = e->plane_d[slot], e->v.planes_d[slot] = pv;
    return RTPP_OK;
}

// SetAgent on an empty planner + Start (RTFMT.cpp:11-36, 393-422): nodes_xy holds, per instance,
// node_capacity (x,y) slots of which n_nodes[i] are used; node 0 is the agent/root, the rest are the
// free-space samples the host-side sampler drew.  Builds ranks, hash and the 25-NN table.
int rtpp_fmt_start(rtpp_fmt_engine *e, const int32_t *map_slot, const int32_t *n_nodes, const float *nodes_xy, const float *goal_xy) {
    if (!e || !map_slot || !n_nodes || !nodes_xy || !goal_xy) return RTPP_ERR_INVALID;
    rtpp_ctx *ctx = e->ctx;
    FmtView &v = e->v;
    const int N = v.n_inst;
    std::vector<int> rank((size_t)N * v.cap, 0), order;
    for (int w = 0; w < N; w++) {
        const int n = n_nodes[w];
        if (n < 1 || n > v.cap) RTPP_FAIL(ctx, RTPP_ERR_CAPACITY, "instance node count out of range");
        const float *xy = nodes_xy + (size_t)w * v.cap * 2;
        order.resize(n);
        std::iota(order.begin(), order.end(), 0);
        std::sort(order.begin(), order.end(), [&](int a, int b) { // RTpoint::operator< (RTFMT.h:36-42)
            return xy[2 * a] != xy[2 * b] ? xy[2 * a] < xy[2 * b] : (xy[2 * a + 1] != xy[2 * b + 1] ? xy[2 * a + 1] < xy[2 * b + 1] : a < b);
        });
        for (int r = 0; r < n; r++) rank[(size_t)w * v.cap + order[r]] = r;
    }
    FMT_CK(e, cudaMemcpyAsync(v.xy, nodes_xy, (size_t)N * v.cap * 8, cudaMemcpyHostToDevice, ctx->stream));
    FMT_CK(e, cudaMemcpyAsync(v.rank, rank.data(), rank.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    void *dn = nullptr, *ds = nullptr, *dg = nullptr;
    int rc = rtpp_scratch(ctx, 0, (size_t)N * 4, &dn);
    rc |= rtpp_scratch(ctx, 1, (size_t)N * 4, &ds);
    rc |= rtpp_scratch(ctx, 2, (size_t)N * 8, &dg);
    if (rc) return rc;
    FMT_CK(e, cudaMemcpyAsync(dn, n_nodes, (size_t)N * 4, cudaMemcpyHostToDevice, ctx->stream));
    FMT_CK(e, cudaMemcpyAsync(ds, map_slot, (size_t)N * 4, cudaMemcpyHostToDevice, ctx->stream));
    FMT_CK(e, cudaMemcpyAsync(dg, goal_xy, (size_t)N * 8, cudaMemcpyHostToDevice, ctx->stream));
    fmt_start_kernel<<<(N + 3) / 4, 128, 0, ctx->stream>>>(v, (const int *)dn, (const int *)ds, (const float2 *)dg);
    RTPP_LAUNCHED(ctx);
    size_t blocks = ((size_t)N * v.cap + 7) / 8, cap = (size_t)ctx->sm_count * 16;
    fmt_knn_table_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, ctx->stream>>>(v);
    RTPP_LAUNCHED(ctx);
    FMT_CK(e, cudaStreamSynchronize(ctx->stream));
    return RTPP_OK;
}

int rtpp_fmt_set_inputs(rtpp_fmt_engine *e, const float *agent_xy, const float *goal_xy) {
    if (!e) return RTPP_ERR_INVALID;
    rtpp_ctx *ctx = e->ctx;
    const int n = e->v.n_inst;
    void *da = nullptr, *dg = nullptr;
    if (agent_xy) {
        int rc = rtpp_scratch(ctx, 0, (size_t)n * 8, &da);
        if (rc) return rc;
        FMT_CK(e, cudaMemcpyAsync(da, agent_xy, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (goal_xy) {
        int rc = rtpp_scratch(ctx, 1, (size_t)n * 8, &dg);
        if (rc) return rc;
        FMT_CK(e, cudaMemcpyAsync(dg, goal_xy, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    fmt_set_agents_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(e->v, (const float2 *)da, (const float2 *)dg);
    RTPP_LAUNCHED(ctx);
    FMT_CK(e, cudaStreamSynchronize(ctx->stream));
    return RTPP_OK;
}

// MainTask minus DrawDObsmap (RTFMT.cpp:704-728): optional UpdateTree, `iters` rounds of the phases
// selected by phase_mask (1 ExpandFMT, 2 RewireFromObstacles, 4 RewireFromRoot; MainTask = 7),
// optional GetPath.
int rtpp_fmt_tick(rtpp_fmt_engine *e, int do_update_tree, int iters, int phase_mask, int do_get_path) {
    if (!e || iters < 0 || phase_mask < 0 || phase_mask > 7) return RTPP_ERR_INVALID;
    rtpp_ctx *ctx = e->ctx;
    const int wpb = e->v.n_inst >= 4 * ctx->sm_count ? 4 : 1;
    fmt_tick_kernel<<<(e->v.n_inst + wpb - 1) / wpb, 32 * wpb, 0, ctx->stream>>>(e->v, e->mt, do_update_tree, iters, phase_mask, do_get_path);
    RTPP_LAUNCHED(ctx);
    return RTPP_OK;
}

// 16 int64 per instance: n_nodes, root, z, n_unv, n_open, n_opennew, n_toopen, x_near, n_b, qo_size,
// qr_size, expand_calls, status, path_len, path_ok, 0
int rtpp_fmt_read_stats(rtpp_fmt_engine *e, long long *out16) {
    if (!e || !out16) return RTPP_ERR_INVALID;
    rtpp_ctx *ctx = e->ctx;
    const int n = e->v.n_inst;
    std::vector<FmtInst> s(n);
    FMT_CK(e, cudaMemcpyAsync(s.data(), e->v.inst, (size_t)n * sizeof(FmtInst), cudaMemcpyDeviceToHost, ctx->stream));
    FMT_CK(e, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < n; i++) {
        long long *o = out16 + (size_t)i * 16;
        o[0] = s[i].n, o[1] = s[i].root, o[2] = s[i].z, o[3] = s[i].n_unv, o[4] = s[i].open_n, o[5] = s[i].opennew_n, o[6] = s[i].toopen_n;
        o[7] = s[i].xnear_n, o[8] = s[i].nb_n, o[9] = s[i].qo_tail - s[i].qo_head, o[10] = s[i].qr_tail - s[i].qr_head;
        o[11] = s[i].n_expand, o[12] = s[i].status, o[13] = s[i].path_len, o[14] = s[i].path_ok, o[15] = 0;
    }
    return RTPP_OK;
}

// nodes of one instance: parent, cost, collision, state (0 unvisited, 1 open, 2 open-new, 3 closed, +4 in V_toopen)
int rtpp_fmt_read_tree(rtpp_fmt_engine *e, int inst, int max_nodes, int32_t *parent, float *cost, uint8_t *collision, uint8_t *state,
                       int *n_out) {
    if (!e || inst < 0 || inst >= e->v.n_inst || !n_out) return RTPP_ERR_INVALID;
    rtpp_ctx *ctx = e->ctx;
    FmtInst s;
    FMT_CK(e, cudaMemcpyAsync(&s, e->v.inst + inst, sizeof(s), cudaMemcpyDeviceToHost, ctx->stream));
    FMT_CK(e, cudaStreamSynchronize(ctx->stream));
    *n_out = s.n;
    const int n = s.n < max_nodes ? s.n : max_nodes;
    if (n <= 0) return RTPP_OK;
    std::vector<cpword> w(n);
    FMT_CK(e, cudaMemcpyAsync(w.data(), e->v.cp + (size_t)inst * e->v.cap, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (state) FMT_CK(e, cudaMemcpyAsync(state, e->v.state + (size_t)inst * e->v.cap, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    FMT_CK(e, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < n; i++) {
        const uint32_t hi = cp_hi(w[i]), cb = hi & 0x7fffffffu;
        if (parent) parent[i] = cp_parent(w[i]);
        if (cost) memcpy(&cost[i], &cb, 4);
        if (collision) collision[i] = hi >> 31;
    }
    return RTPP_OK;
}

// node ids root..entry then -2 (= the goal) when the last GetPath found an entry (path_ok)
int rtpp_fmt_read_path(rtpp_fmt_engine *e, int inst, int max_len, int32_t *ids, int *len_out) {
    if (!e || inst < 0 || inst >= e->v.n_inst || !ids || !len_out) return RTPP_ERR_INVALID;
    rtpp_ctx *ctx = e->ctx;
    FmtInst s;
    FMT_CK(e, cudaMemcpyAsync(&s, e->v.inst + inst, sizeof(s), cudaMemcpyDeviceToHost, ctx->stream));
    FMT_CK(e, cudaStreamSynchronize(ctx->stream));
    *len_out = s.path_len;
    int n = s.path_len < e->v.path_cap ? s.path_len : e->v.path_cap;
    if (n > max_len) n = max_len;
    if (n <= 0) return RTPP_OK;
    std::vector<int> p(n);
    FMT_CK(e, cudaMemcpyAsync(p.data(), e->v.path + (size_t)inst * e->v.path_cap, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    FMT_CK(e, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < n; i++) ids[i] = p[n - 1 - i];
    return RTPP_OK;
}

} // extern "C"
