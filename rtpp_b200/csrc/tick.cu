// tick.cu — K1w: one replanning tick's collision pass over the RESIDENT tree, with the map window in shared memory.
//
// The tick's edges (reference AMRRT.cpp:185-199 UpdateTree: parent -> child; AMRRT.cpp:496-500 RewireRoot: node -> each of
// its k nearest; RTRRT.cpp:128-134, 304-320 likewise) are implied by (parents, k-NN table), so no edge list is streamed.
// A static PLAN, rebuilt only with the k-NN table (or a new divisor), orders the nodes by 64x64-pixel supercell (2x2 bit
// tiles) in groups of 32:
//   * a warp owns a group: lane = source node, the loop runs over its k neighbours (the j-th neighbours of 32 nearby nodes
//     have similar lengths, so the dyadic march levels of RTRRT.cpp:210-218 are nearly warp-uniform);
//   * the neighbour's coordinates come from the plan, transposed so that step j is one coalesced 256-byte load;
//   * the 4x4 bit tiles around the supercell (supercell + 32-pixel halo) are staged in shared memory per warp, column-major
//     by tile column, so a probe is one 32-bit shared load whose bank is the pixel row: the 32 lanes of a step probe around
//     32 nearby nodes and mostly hit distinct banks.  Staging is either one coalesced 128-byte load + conflict-free store
//     per tile (RTPP_TICK_BULK=0) or, by default, one cp.async.bulk per tile completing on the warp's mbarrier;
//   * the interpolation ps*t + (1-t)*pe of both coordinates is one packed binary32x2 sequence (FMUL2, FMUL2, FFMA2 by a
//     run-time 1.0 so that ptxas cannot contract the product into the sum, FADD2.RZ for the truncation): the same roundings
//     as the reference's scalar code, half the instructions;
//   * the source pixel is probed once per node, not once per edge; the verdict bits of a node leave as ONE word
//     (bit j = edge node -> knn[node][j] blocked).
// Edges the window cannot serve — neighbour outside the halo, an end point on the map's last row / column or outside it, a
// march finer than 16 steps for the plan's divisor — are known when the plan is built: they form a static FAR list (explicit
// end points) checked with the general warp routine of edges.cu by the same launch; a window warp picks their verdict bits
// up when it writes its rows.  The tree edges (parent -> child, any length) take the same general routine.
// The launch is ONE resident wave of persistent warps over one list of work items: the far list's 32-edge words first (so
// nobody waits for them), then the tree's words (long edges are pre-filtered by their own lane, see edge_march.cuh, and the
// survivors marched warp-wide), then the window groups, longest first (a cost estimate computed with the plan) — no block-sized
// tail and no warp idling inside a block.  The first four fifths of the items are dealt round-robin (no atomic); the rest is
// drawn from a counter by whichever warp is free, because item times differ.
//
// Multi-GPU (SURVEY 8(e) grain 2): with kPeers the tree words and window groups of ONE tick are dealt round-robin over the
// ranks (the small far list is checked by every rank); a warp stores its verdict words into every rank's result buffers
// over NVLink (CUDA IPC mappings); every block publishes its stores with one system-scope fence and counts itself, and the
// LAST block of the launch to retire signals the peers and waits — bounded, failing closed — for their signals: the
// exchange is part of the launch, no collective call, no second kernel.
// Profiling hooks (off unless the environment asks): RTPP_TICK_TIMING (per-block time stamps, longest item), RTPP_TICK_ITEMS
// (per-item log for tools/tick_items.py), RTPP_TICK_PARTS / _STATIC_PCT / _LPT / _BULK (ablations).
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "edge_march.cuh"
#include "nodes.cuh"
#include "scan.cuh"

typedef unsigned long long u64;

// ------------------------------------------------------------------------------------------------ packed binary32x2
__device__ __forceinline__ u64 pk2(float a, float b) {
    u64 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
    u64 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ u64 add2rz(u64 a, u64 b) {
    u64 r;
    asm("add.rz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// ------------------------------------------------------------------------------------------------ mbarrier / bulk copy
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// false = the barrier never completed (a lost copy): the caller traps instead of hanging the GPU
__device__ __forceinline__ bool mbar_wait(unsigned bar, unsigned parity) {
#pragma unroll 1
    for (int it = 0; it < (1 << 20); it++) {
        unsigned ok;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok)
                     : "r"(bar), "r"(parity)
                     : "memory");
        if (ok) return true;
    }
    return false;
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// ------------------------------------------------------------------------------------------------ the plan
static constexpr int kSCT = 2;            // supercell = kSCT x kSCT bit tiles (64 x 64 pixels)
static constexpr int kWT = kSCT + 2;      // window = supercell + one tile of halo on every side (128 x 128 pixels)
static constexpr int kWinWords = kWT * kWT * 32;
static constexpr int kTickWarps = 8;

struct TickPlan {
    // slot order (a function of the node set alone)
    int order_n = -1, order_k = 0, W = 0, H = 0;
    bool applies = false;                 // hash cells <= supercell
    bool canonical = false;               // no cell was too large to rank: the order is the same on every rank
    int scw = 0, sch = 0, n_groups = 0, n_slots = 0;
    int *sc_start = nullptr;   // [scw*sch + 1] first slot of a supercell (slots are padded to groups of 32 per supercell)
    int *group_sc = nullptr;   // [n_groups] supercell of a group
    int *slot_id = nullptr;    // [n_slots] node id (-1 = padding)
    int *slot_of_id = nullptr; // [capacity] node id -> slot
    // edges (depend on the divisor: a march finer than 16 steps is a far edge)
    int edges_divisor = 0;
    bool usable = false;                  // applies, and at most RTPP_TICK_MAX_FAR of the edges are far
    long long far_edges = 0, valid_edges = 0;
    int *group_cost = nullptr; // [n_groups] march steps a warp spends on the group (sum over j of the largest level among its lanes)
    std::vector<int> h_cost; // host copy of group_cost (profiling dump)
    int *item_group = nullptr; // [n_groups] groups by decreasing cost: the order they are dealt in (longest first, so that the
                               // items drawn last are the short ones); a function of the plan alone: identical on every rank
    float2 *slot_xy = nullptr; // [n_slots] source coordinates (a stand-in inside the supercell for padding / sources the window cannot serve)
    uint32_t *far_mask = nullptr, *valid_mask = nullptr; // [n_slots] bit j: edge j is in the far list / exists
    int *far_off = nullptr;    // [n_slots + 1] first far-list entry of a slot
    float2 *ent = nullptr;     // [n_groups * k * 32] neighbour coordinates of (group, j, lane); the source itself where the edge is far / absent
    float4 *far_list = nullptr; // [far_edges] explicit end points
    uint32_t *far_bits = nullptr; // [far_edges / 32 + 1] their verdicts, rewritten by every launch
    unsigned *far_done = nullptr; // far words finished so far (monotonic across launches; a launch waits for its own target)
    unsigned far_target = 0, work_base = 0; // far_done[1] is the tail's ticket counter
    unsigned long long *stats = nullptr; // [3] far edges, valid edges, cells too large to rank
    uint32_t *row_scratch = nullptr, *tree_scratch = nullptr; // for the legacy packed-bit entry points / pageable host buffers
    size_t cap_sc = 0, cap_slot_id = 0, cap_slot_xy = 0, cap_far = 0, cap_valid = 0, cap_far_off = 0, cap_group = 0, cap_ent = 0,
           cap_far_list = 0, cap_far_bits = 0, cap_row_scratch = 0, cap_tree_scratch = 0, cap_group_cost = 0, cap_item_group = 0;
};

struct PlanGeom {
    int shift, gw, gh;   // hash grid
    int c;               // hash cells per supercell side
    int scw, sch;
    int W, H, tilesX, tilesY;
    int k;
    float t3;            // squared length from which a march has more than 16 steps (MarchThresholds::t[3])
    MarchThresholds mt;  // for the groups' cost estimate
};

__global__ void __launch_bounds__(256) plan_count_kernel(PlanGeom g, const int *__restrict__ cell_start, int *__restrict__ cnt) {
    const int sc = blockIdx.x * blockDim.x + threadIdx.x;
    if (sc >= g.scw * g.sch) return;
    const int sx = sc % g.scw, sy = sc / g.scw;
    const int x0 = sx * g.c, x1 = min(x0 + g.c, g.gw);
    int total = 0;
    for (int r = 0; r < g.c; r++) {
        const int cy = sy * g.c + r;
        if (cy >= g.gh || x0 >= g.gw) break;
        total += cell_start[cy * g.gw + x1] - cell_start[cy * g.gw + x0];
    }
    cnt[sc] = (total + 31) & ~31;
}

// One block per supercell.  Slot order inside a supercell = (hash cell, row-major; node id ascending inside the cell): the hash's
// own order inside a cell comes from atomics and differs from run to run and from rank to rank, the slot order must not (the
// multi-GPU tick exchanges verdict words by slot).  A warp takes a cell and ranks its ids (cells hold tens of nodes).
static constexpr int kPlanRankMax = 8192; // a cell beyond this keeps the hash order (flagged: no multi-GPU exchange)
__global__ void __launch_bounds__(256) plan_slots_kernel(PlanGeom g, const int *__restrict__ cell_start, const int *__restrict__ sid,
                                                        const int *__restrict__ sc_start, int *__restrict__ slot_id, int *__restrict__ group_sc,
                                                        int *__restrict__ slot_of_id, unsigned long long *__restrict__ stats) {
    const int sc = blockIdx.x;
    const int base = sc_start[sc], total = sc_start[sc + 1] - base;
    if (total == 0) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int sx = sc % g.scw, sy = sc / g.scw;
    const int x0 = sx * g.c, x1 = min(x0 + g.c, g.gw);
    const int ncx = x1 - x0, nrows = min(g.c, g.gh - sy * g.c);
    for (int ci = warp; ci < ncx * nrows; ci += nwarps) {
        const int r = ci / ncx, cx = x0 + ci % ncx, cy = sy * g.c + r;
        int off = cell_start[cy * g.gw + cx] - cell_start[cy * g.gw + x0]; // cells before this one in its row
        for (int rr = 0; rr < r; rr++) off += cell_start[(sy * g.c + rr) * g.gw + x1] - cell_start[(sy * g.c + rr) * g.gw + x0];
        const int a = cell_start[cy * g.gw + cx], m = cell_start[cy * g.gw + cx + 1] - a;
        if (m > kPlanRankMax && lane == 0) atomicAdd(stats + 2, 1ull);
        for (int e = lane; e < m; e += 32) {
            const int id = sid[a + e];
            int rank = e;
            if (m <= kPlanRankMax) {
                rank = 0;
                for (int f = 0; f < m; f++) rank += sid[a + f] < id;
            }
            const int slot = base + off + rank;
            slot_id[slot] = id;
            slot_of_id[id] = slot;
        }
    }
    int count = 0;
    for (int r = 0; r < nrows; r++) count += cell_start[(sy * g.c + r) * g.gw + x1] - cell_start[(sy * g.c + r) * g.gw + x0];
    for (int q = count + threadIdx.x; q < total; q += blockDim.x) slot_id[base + q] = -1;
    for (int q = threadIdx.x; q < (total >> 5); q += blockDim.x) group_sc[(base >> 5) + q] = sc;
}

__device__ __forceinline__ bool plan_pt_safe(const PlanGeom &g, float2 p) {
    return p.x >= 0.f && p.y >= 0.f && p.x < (float)(g.W - 1) && p.y < (float)(g.H - 1);
}

// One warp per group: classifies the k edges of its 32 nodes.  Pass 0 writes the masks and the transposed neighbour
// coordinates and counts the far edges per slot (far_cnt); pass 1 (after the prefix sum far_off) writes the far list.
__global__ void __launch_bounds__(256) plan_entries_kernel(PlanGeom g, int n_groups, int pass, const int *__restrict__ group_sc,
                                                          const int *__restrict__ slot_id, const float2 *__restrict__ xy, float2 *__restrict__ slot_xy,
                                                          const int *__restrict__ knn_pos, const float2 *__restrict__ sxy, float2 *__restrict__ ent,
                                                          uint32_t *__restrict__ far_mask, uint32_t *__restrict__ valid_mask, int *__restrict__ far_cnt,
                                                          const int *__restrict__ far_off, float4 *__restrict__ far_list,
                                                          unsigned long long *__restrict__ stats, int *__restrict__ group_cost) {
    const int lane = threadIdx.x & 31;
    const int grp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (grp >= n_groups) return;
    int cost = 0;
    const int sc = group_sc[grp], sx = sc % g.scw, sy = sc / g.scw;
    const int slot = grp * 32 + lane, id = slot_id[slot];
    const int px0 = sx * kSCT * 32, py0 = sy * kSCT * 32; // supercell origin in pixels
    const float2 pad = make_float2((float)px0 + 1.5f, (float)py0 + 1.5f);
    const float2 real = id >= 0 ? xy[id] : pad;
    const int ix = __float2int_rz(real.x), iy = __float2int_rz(real.y);
    const bool src_ok = id >= 0 && plan_pt_safe(g, real) && ix >= px0 && ix < px0 + kSCT * 32 && iy >= py0 && iy < py0 + kSCT * 32;
    // a source the window cannot serve (outside its supercell's pixels: a point outside the hash domain; on the map's last row /
    // column) is replaced by a stand-in inside the supercell: all its edges are far and are checked with the real coordinates
    const float2 src = src_ok ? real : pad;
    // the window spans [px0 - 32, px0 + (kSCT + 1) * 32) pixels; a neighbour is served by it when its pixel lies at least one
    // pixel inside (an interpolated point can round one pixel beyond the end points' box, see common.cuh rt_edge_safe)
    const int lox = px0 - 32 + 1, hix = px0 + (kSCT + 1) * 32 - 2, loy = py0 - 32 + 1, hiy = py0 + (kSCT + 1) * 32 - 2;
    uint32_t fm = 0, vm = 0;
    float2 *out = ent + (size_t)grp * g.k * 32 + lane;
    int f = pass ? far_off[slot] : 0;
    for (int j = 0; j < g.k; j++) {
        const int bp = id >= 0 ? knn_pos[(size_t)id * g.k + j] : -1;
        float2 e = src;
        if (bp >= 0) {
            const float2 d = sxy[bp];
            const int dx = __float2int_rz(d.x), dy = __float2int_rz(d.y);
            const bool near = src_ok && plan_pt_safe(g, d) && dx >= lox && dx <= hix && dy >= loy && dy <= hiy &&
                              rt_sqlen(d.x, d.y, real.x, real.y) < g.t3; // the kernel's own squared length (symmetric: squares of +-dx)
            vm |= 1u << j;
            if (near) e = d;
            else {
                fm |= 1u << j;
                if (pass) far_list[f++] = make_float4(real.x, real.y, d.x, d.y);
            }
        }
        if (!pass) {
            out[(size_t)j * 32] = e;
            // the kernel marches step j to the deepest level any lane needs
            const int P = (e.x != src.x || e.y != src.y) ? rt_march_P(g.mt, rt_sqlen(e.x, e.y, src.x, src.y)) : 0;
            cost += 2 + __reduce_max_sync(0xffffffffu, P);
        }
    }
    if (pass) return;
    if (lane == 0 && group_cost) group_cost[grp] = cost;
    slot_xy[slot] = src;
    far_mask[slot] = fm;
    valid_mask[slot] = vm;
    far_cnt[slot] = __popc(fm);
    unsigned nf = __popc(fm), nv = __popc(vm);
    for (int o = 16; o; o >>= 1) nf += __shfl_xor_sync(0xffffffffu, nf, o), nv += __shfl_xor_sync(0xffffffffu, nv, o);
    if (lane == 0) {
        atomicAdd(stats, (unsigned long long)nf);
        atomicAdd(stats + 1, (unsigned long long)nv);
    }
}

// ------------------------------------------------------------------------------------------------ the kernel
struct TickPeers {
    uint32_t *row[RTPP_MAX_PEERS - 1];
    uint32_t *tree[RTPP_MAX_PEERS - 1];
    int *flags[RTPP_MAX_PEERS]; // [r]: rank r's flag array as mapped here (own included)
    int n, rank, world, epoch;
    uint32_t *own_row, *own_tree; // for the fail-closed fill
    size_t row_words, tree_words;
    long long spin_limit;
};
// flag array layout (ints): [0 .. RTPP_MAX_PEERS) epoch last signalled by rank r; then
static constexpr int kFlagTimeouts = RTPP_MAX_PEERS;     // number of exchanges that timed out (sticky)
static constexpr int kFlagDone = RTPP_MAX_PEERS + 1;     // blocks of the current launch that have retired
static_assert(kFlagDone < RTPP_PEER_FLAG_INTS, "flag array too small");

struct TickArgs {
    PlaneView pl;
    MarchThresholds mt;
    float one; // 1.0f at run time (see fma2 above)
    // work items: [0, n_far_words) = 32-edge words of the far list, then this rank's share of the tick's tree words and window
    // groups (shared item s: s < n_tree_words tree word s, else group s - n_tree_words), dealt round-robin to the launch's warps
    unsigned n_items;
    // the first n_static items are dealt round-robin; the rest (the last fifth) is drawn from a counter by whichever warp is free:
    // item = atomicAdd(work) - work_base, monotonic across launches
    unsigned n_static, work_base;
    unsigned *work;
    // far list
    const float4 *far_list;
    uint32_t *far_bits;
    unsigned *far_done;      // far words finished so far (monotonic; this launch's window warps wait for far_target)
    unsigned far_target;
    int n_far, n_far_words;
    // tree edges
    const float2 *xy;
    const cpword *cp;
    int n, n_tree_words;
    uint32_t *tree_bits;
    // rewire edges (the rest)
    int n_groups, k, scw;
    const int *group_sc, *far_off, *item_group;
    const float2 *slot_xy, *ent;
    const uint32_t *far_mask, *valid_mask;
    uint32_t *row_bits;
    int stride, first;         // multi-GPU: this launch's shared item i is item i * stride + first of the tick
    int parts;                 // bit 0: tree edges, bit 1: rewire edges, bit 3: skip the far merge, bit 4: skip the peer row stores (RTPP_TICK_PARTS, profiling only; default 3)
    unsigned long long *item_log; // RTPP_TICK_ITEMS=<file> with RTPP_TICK_TIMING (multi-GPU kernels): per item start (low 32 bits of ns) << 32 | duration ns
    unsigned long long *timing; // RTPP_TICK_TIMING=1 (profiling only): per block {start, work done, exit} in globaltimer ns + its longest item (ns << 32 | item)
};

// One probe of the shared window.  q = packed (x, y) + 2^23 rounded toward zero: the low mantissa bits are the pixel.
// cblk folds the window's shared address and origin: addr = (col >> 5 - tx0) * kWT*128 + (row - wy0) * 4 + base.
__device__ __forceinline__ uint32_t win_probe(u64 p, unsigned cblk) {
    const u64 q = add2rz(p, 0x4B0000004B000000ull);
    const unsigned xb = (unsigned)q, yb = (unsigned)(q >> 32);
    // (xb >> 5) * kWT*128 == (xb & ~31) << 4 for the 4-tile window: one mask, two shift-adds
    unsigned addr;
    asm("{\n.reg .b32 t, u;\nand.b32 t, %1, 0xffffffe0;\nshl.b32 t, t, 4;\nadd.u32 u, t, %3;\nshl.b32 t, %2, 2;\nadd.u32 %0, t, u;\n}"
        : "=r"(addr)
        : "r"(xb), "r"(yb), "r"(cblk));
    uint32_t w;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(addr) : "memory");
    return __funnelshift_r(w, 0u, xb); // w >> (col & 31)
}
static_assert(kWT == 4, "win_probe's shift assumes a 4-tile-wide window");

// 32 consecutive edges of an explicit list (far list) or of the tree (child e+1 and its parent): one warp, one word
template <bool kPeers, bool kTree>
__device__ __forceinline__ void tick_list_word(const TickArgs &a, const TickPeers &peers, size_t w) {
    const int lane = threadIdx.x & 31;
    const size_t n_edges = kTree ? (size_t)(a.n - 1) : (size_t)a.n_far;
    const size_t e = (w << 5) + lane;
    bool valid = e < n_edges;
    float4 ed = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) {
        if (kTree) {
            const int b = (int)e + 1, p = cp_parent(a.cp[b]);
            valid = p >= 0;
            if (valid) {
                const float2 pa = a.xy[p], pb = a.xy[b];
                ed = make_float4(pa.x, pa.y, pb.x, pb.y);
            }
        } else {
            ed = a.far_list[e];
        }
    }
    // tree words: the warp routine of edges.cu (a hub's thousand-pixel edges are marched by the whole warp).  Far words: every lane
    // marches its own edge — far edges are tens of pixels long (32 to 64 steps), all of a word's edges are like that, and marching
    // them one after the other warp-wide made these few words the launch's critical path
    const unsigned bits = kTree ? warp_edges_verdict(a.pl, ed, valid, a.mt)
                                : __ballot_sync(0xffffffffu, valid && rt_edge_blocked(a.pl, ed.x, ed.y, ed.z, ed.w, a.mt));
    if (lane == 0) {
        if (kTree) {
            a.tree_bits[w] = bits;
            if (kPeers) {
#pragma unroll
                for (int r = 0; r < RTPP_MAX_PEERS - 1; r++)
                    if (r < peers.n) peers.tree[r][w] = bits;
            }
        } else {
            a.far_bits[w] = bits;
        }
    }
}

// (5 blocks per SM at 48 registers measured the same 0.097 ms as 4 at 62: the kernel is not short of warps)
template <bool kPeers, bool kBulk>
__global__ void __launch_bounds__(32 * kTickWarps, 4) tick_check_kernel(const TickArgs a, const TickPeers peers) {
    __shared__ __align__(128) uint32_t win[kTickWarps][kWinWords];
    __shared__ __align__(8) unsigned long long bars[kTickWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned full = 0xffffffffu;
    if (a.timing && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        a.timing[4 * blockIdx.x] = t;
    }
    const unsigned wbase = (unsigned)__cvta_generic_to_shared(&win[warp][0]);
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&bars[warp]);
    unsigned parity = 0;
    if (kBulk) {
        if (lane == 0) {
            mbar_init(bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }
    int cur_sc = -1;
    unsigned cblk = 0;
    const u64 one2 = pk2(a.one, a.one);
    // persistent warps: one resident wave.  Warp w takes items w, w + W, w + 2W, ... of the first four fifths (far words first, then
    // the tree — whose long edges sit at its head and so spread over all warps — then the window groups) and draws the rest from a
    // counter.  (Drawing EVERY item from one counter made that atomic 30 % of the stall samples; dealing ALL items round-robin left
    // the slowest block 25 % behind the mean.)
    const unsigned n_warps = gridDim.x * kTickWarps;
    unsigned next_static = blockIdx.x * kTickWarps + warp;
    unsigned long long t_item = 0;
    unsigned prev_item = 0;
    for (;;) {
        unsigned item;
        if (next_static < a.n_static) {
            item = next_static, next_static += n_warps;
        } else { // the tail: equal item counts are not equal times (march levels, window reloads, long tree edges)
            unsigned t = 0;
            if (lane == 0) t = atomicAdd(a.work, 1u) - a.work_base;
            item = a.n_static + __shfl_sync(full, t, 0);
            if (item >= a.n_items) break;
        }
        if (kPeers && a.timing) { // profiling (multi-GPU launches): the longest item of the block (closed at the next draw)
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (lane == 0 && t_item) {
                atomicMax(a.timing + 4 * blockIdx.x + 3, ((t - t_item) << 32) | prev_item);
                if (a.item_log) a.item_log[prev_item] = (t_item << 32) | ((t - t_item) & 0xffffffffull);
            }
            t_item = t, prev_item = item;
        }
        if (item < (unsigned)a.n_far_words) {
            // far list: every rank checks all of it; window warps wait for these words (taken first, and waiting for nobody)
            tick_list_word<false, false>(a, peers, item);
            __threadfence();
            __syncwarp();
            if (lane == 0) atomicAdd(a.far_done, 1u);
            continue;
        }
        const size_t rel = (size_t)item - a.n_far_words;
        const size_t sh = kPeers ? rel * a.stride + a.first : rel;
        if (sh < (size_t)a.n_tree_words) {
            if (a.parts & 1) tick_list_word<kPeers, true>(a, peers, sh);
            continue;
        }
        if (!(a.parts & 2)) continue;
        {
            {
                const size_t grp = a.item_group ? (size_t)a.item_group[sh - a.n_tree_words] : sh - a.n_tree_words;
                const int sc = a.group_sc[grp];
                if (sc != cur_sc) { // warp-uniform: bring the supercell's window in
                    cur_sc = sc;
                    const int sx = sc % a.scw, sy = sc / a.scw;
                    const int tx0 = sx * kSCT - 1, ty0 = sy * kSCT - 1; // first tile of the window (may lie outside the plane)
                    const int tilesY = (a.pl.H + 31) >> 5;
                    __syncwarp(); // every lane is done with the previous window
                    if (kBulk) {
                        // one bulk asynchronous copy per 128-byte tile, completion counted in bytes on the warp's mbarrier
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        int nvalid = 0;
                        for (int t = 0; t < kWT * kWT; t++) {
                            const int tx = tx0 + t / kWT, ty = ty0 + t % kWT;
                            nvalid += (tx >= 0 && tx < a.pl.tilesX && ty >= 0 && ty < tilesY);
                        }
                        if (lane == 0) mbar_expect_tx(bar, (unsigned)nvalid * 128u);
                        __syncwarp();
                        if (lane < kWT * kWT) {
                            const int tx = tx0 + lane / kWT, ty = ty0 + lane % kWT; // slot t = txr * kWT + tyr: column-major by tile column
                            if (tx >= 0 && tx < a.pl.tilesX && ty >= 0 && ty < tilesY) {
                                bulk_g2s(wbase + (unsigned)lane * 128u, a.pl.w + (((size_t)ty * a.pl.tilesX + tx) << 5), 128u, bar);
                            } else { // outside the plane: never probed (plan), kept defined
                                uint4 *z = (uint4 *)&win[warp][lane * 32];
                                for (int q = 0; q < 8; q++) z[q] = make_uint4(0u, 0u, 0u, 0u);
                            }
                        }
                        if (!mbar_wait(bar, parity)) __trap();
                        parity ^= 1;
                    } else {
                        // lane l carries word l (= pixel row l) of every tile: each tile is one coalesced 128-byte load and one
                        // conflict-free shared store; tiles outside the plane read as free (never probed: plan)
                        uint32_t wv[kWT * kWT];
#pragma unroll
                        for (int t = 0; t < kWT * kWT; t++) {
                            const int tx = tx0 + t / kWT, ty = ty0 + t % kWT;
                            wv[t] = 0u;
                            if (tx >= 0 && tx < a.pl.tilesX && ty >= 0 && ty < tilesY) wv[t] = __ldg(a.pl.w + (((size_t)ty * a.pl.tilesX + tx) << 5) + lane);
                        }
#pragma unroll
                        for (int t = 0; t < kWT * kWT; t++) win[warp][t * 32 + lane] = wv[t];
                    }
                    cblk = wbase - ((0x4B000000u >> 5) + (unsigned)tx0) * (unsigned)(kWT * 128) - ((0x4B000000u + (unsigned)(ty0 * 32)) << 2);
                    __syncwarp();
                }
                const size_t slot = grp * 32 + lane;
                const float2 src = a.slot_xy[slot];
                const uint32_t fmask = a.far_mask[slot], vmask = a.valid_mask[slot];
                const u64 A = pk2(src.x, src.y);
                const bool srcb = win_probe(A, cblk) & 1u; // the node's own pixel, once per node
                const float2 *ep = a.ent + grp * (size_t)a.k * 32 + lane;
                uint32_t row = 0;
                // neighbours in chunks of 5 (k = 25: five chunks): a chunk's five coalesced 256-byte loads are issued together, one
                // chunk ahead of the march
                float2 nx[5];
#pragma unroll
                for (int u = 0; u < 5; u++)
                    if (u < a.k) nx[u] = ep[(size_t)u * 32];
                for (int j0 = 0; j0 < a.k; j0 += 5) {
                    float2 dd[5];
#pragma unroll
                    for (int u = 0; u < 5; u++) dd[u] = nx[u];
#pragma unroll
                    for (int u = 0; u < 5; u++)
                        if (j0 + 5 + u < a.k) nx[u] = ep[(size_t)(j0 + 5 + u) * 32];
#pragma unroll
                    for (int u = 0; u < 5; u++) {
                        if (j0 + u >= a.k) break; // warp-uniform
                        const float2 d = dd[u];
                        const u64 B = pk2(d.x, d.y);
                        // squared length exactly as rt_sqlen: two rounded differences, two rounded squares, one rounded sum
                        const u64 df = sub2(B, A), sq = mul2(df, df);
                        const float s = __fmaf_rn(__uint_as_float((unsigned)sq), a.one, __uint_as_float((unsigned)(sq >> 32)));
                        const bool n4 = s >= a.mt.t[0], n8 = s >= a.mt.t[1], n16 = s >= a.mt.t[2]; // s < t[3]: the plan
                        auto at = [&](const float t, const float omt) -> uint32_t {
                            // fl(fl(ps*t) + fl((1-t)*pe)) for x and y at once; the verdict is an OR over the level, and t, 1-t are
                            // both in it, so which end point takes t does not matter
                            return win_probe(fma2(mul2(A, pk2(t, t)), one2, mul2(B, pk2(omt, omt))), cblk);
                        };
                        uint32_t acc = win_probe(B, cblk) | at(0.5f, 0.5f);
                        if (__any_sync(full, n4)) {
                            const uint32_t v4 = at(0.25f, 0.75f) | at(0.75f, 0.25f);
                            if (n4) acc |= v4;
                            if (__any_sync(full, n8)) {
                                const uint32_t v8 = at(0.125f, 0.875f) | at(0.375f, 0.625f) | at(0.625f, 0.375f) | at(0.875f, 0.125f);
                                if (n8) acc |= v8;
                                if (__any_sync(full, n16)) {
                                    const uint32_t v16 = at(0.0625f, 0.9375f) | at(0.1875f, 0.8125f) | at(0.3125f, 0.6875f) | at(0.4375f, 0.5625f) |
                                                         at(0.5625f, 0.4375f) | at(0.6875f, 0.3125f) | at(0.8125f, 0.1875f) | at(0.9375f, 0.0625f);
                                    if (n16) acc |= v16;
                                }
                            }
                        }
                        row = __funnelshift_r(row, acc, 1); // bit 0 of acc enters at the top; after k steps: >> (32 - k)
                    }
                }
                row >>= (32 - a.k);
                // A blocked source blocks every edge of the node (RTRRT.cpp:204-207).  Nodes whose edges are all far (also: sources the
                // window cannot serve, whose window coordinates are a stand-in) skip this shortcut: the far list has their verdicts.
                if (srcb && fmask != vmask) row = vmask;
                else if (fmask && !(a.parts & 8)) {
                    // verdicts of this node's far edges: the queue's first items, long taken (and usually finished) by now
                    if (a.n_far_words) {
                        for (long long it = 0; (int)(*(volatile unsigned *)a.far_done - a.far_target) < 0; it++) {
                            __nanosleep(32);
                            if (it > (1ll << 26)) __trap(); // cannot happen: the far words precede every group in the queue
                        }
                        __threadfence();
                    }
                    // the slot's far verdicts are consecutive bits of the far list's words (at most k <= 32 of them: two words)
                    const int f = a.far_off[slot];
                    const uint32_t w0 = __ldcg(a.far_bits + (f >> 5)), w1 = __ldcg(a.far_bits + (f >> 5) + 1);
                    uint32_t fb = __funnelshift_r(w0, w1, f & 31);
                    row &= ~fmask;
                    for (uint32_t m = fmask; m; m &= m - 1, fb >>= 1) row |= (fb & 1u) << (__ffs(m) - 1);
                }
                row &= vmask;
                a.row_bits[slot] = row;
                if (kPeers && !(a.parts & 16)) {
#pragma unroll
                    for (int r = 0; r < RTPP_MAX_PEERS - 1; r++)
                        if (r < peers.n) peers.row[r][slot] = row;
                }
            }
        }
    }
    if (a.timing) {
        if (kPeers) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (lane == 0 && t_item) {
                atomicMax(a.timing + 4 * blockIdx.x + 3, ((t - t_item) << 32) | prev_item);
                if (a.item_log) a.item_log[prev_item] = (t_item << 32) | ((t - t_item) & 0xffffffffull);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            a.timing[4 * blockIdx.x + 1] = t;
        }
    }
    if (kPeers) {
        // last-block-done exchange: every block publishes its stores system-wide and counts itself; the last one signals the
        // peers (their flag slot `rank` := epoch) and waits until every peer has signalled this rank.  The wait is bounded:
        // on a timeout the result buffers are filled with "blocked" (fail closed) and the sticky timeout counter is raised,
        // which rtpp_peer_status reports.
        // (one fence per block: the barrier orders the block's stores before thread 0's fence, which is cumulative)
        __syncthreads();
        __shared__ int is_last;
        int *mine = peers.flags[peers.rank];
        if (threadIdx.x == 0) {
            __threadfence_system();
            is_last = atomicAdd(mine + kFlagDone, 1) == (int)gridDim.x - 1;
        }
        __syncthreads();
        if (is_last) {
            // every block published its stores system-wide before it counted itself; the signalling threads order that count (read by
            // thread 0, handed over by the barrier) before their flag store at system scope.  Nothing has to be ordered after the
            // flag store on this side.
            bool timed_out = false;
            if (threadIdx.x < (unsigned)peers.world && (int)threadIdx.x != peers.rank) {
                __threadfence_system();
                *(volatile int *)(peers.flags[threadIdx.x] + peers.rank) = peers.epoch;
                volatile int *src = mine + threadIdx.x;
                long long it = 0;
                while (*src < peers.epoch && it < peers.spin_limit) {
                    __nanosleep(64);
                    it++;
                }
                timed_out = *src < peers.epoch;
            }
            const int any_timeout = __syncthreads_or(timed_out);
            if (any_timeout) {
                for (size_t i = threadIdx.x; i < peers.row_words; i += blockDim.x) peers.own_row[i] = 0xffffffffu;
                for (size_t i = threadIdx.x; i < peers.tree_words; i += blockDim.x) peers.own_tree[i] = 0xffffffffu;
                if (threadIdx.x == 0) atomicAdd(mine + kFlagTimeouts, 1);
            }
            if (threadIdx.x == 0) mine[kFlagDone] = 0;
        }
    }
    if (a.timing && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        a.timing[4 * blockIdx.x + 2] = t;
    }
}

// legacy packed-bit order of rtpp_tick_edges_check: edge e < n-1 = tree edge of child e+1, then (node id, j) row-major
__global__ void __launch_bounds__(256) tick_repack_kernel(const uint32_t *__restrict__ tree_bits, const uint32_t *__restrict__ row_bits,
                                                         const int *__restrict__ slot_of_id, int n, int k, size_t n_words,
                                                         uint32_t *__restrict__ out) {
    const size_t n_tree = (size_t)(n - 1), n_edges = n_tree + (size_t)n * k;
    for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += (size_t)gridDim.x * blockDim.x) {
        uint32_t v = 0;
        size_t e = w << 5;
        int b = 0;
        while (b < 32 && e < n_tree) { // tree part
            v |= ((tree_bits[e >> 5] >> (e & 31)) & 1u) << b;
            b++, e++;
        }
        if (b < 32 && e < n_edges) {
            size_t q = e - n_tree;
            int i = (int)(q / (unsigned)k), j = (int)(q - (size_t)i * k);
            uint32_t row = row_bits[slot_of_id[i]];
            while (b < 32 && i < n) {
                v |= ((row >> j) & 1u) << b;
                b++;
                if (++j == k) {
                    j = 0;
                    if (++i < n) row = row_bits[slot_of_id[i]];
                }
            }
        }
        out[w] = v;
    }
}

// ------------------------------------------------------------------------------------------------ host side
void rtpp_tick_plan_free(NodeStore *s) {
    TickPlan *p = s->plan;
    if (!p) return;
    cudaFree(p->sc_start), cudaFree(p->group_sc), cudaFree(p->slot_id), cudaFree(p->slot_of_id), cudaFree(p->slot_xy);
    cudaFree(p->far_mask), cudaFree(p->valid_mask), cudaFree(p->far_off), cudaFree(p->ent), cudaFree(p->far_list), cudaFree(p->far_bits);
    cudaFree(p->far_done), cudaFree(p->stats), cudaFree(p->row_scratch), cudaFree(p->tree_scratch), cudaFree(p->group_cost), cudaFree(p->item_group);
    delete p;
    s->plan = nullptr;
}
void rtpp_tick_plan_invalidate(NodeStore *s) {
    if (s->plan) s->plan->order_n = -1, s->plan->edges_divisor = 0;
}

template <class T> static int plan_grow(rtpp_ctx *ctx, T **ptr, size_t *cap, size_t want) {
    if (*cap >= want && *ptr) return RTPP_OK;
    if (*ptr) {
        RTPP_CK(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(*ptr);
        *ptr = nullptr, *cap = 0;
    }
    const size_t n = want + want / 8 + 64;
    RTPP_CK(ctx, cudaMalloc(ptr, n * sizeof(T)));
    *cap = n;
    return RTPP_OK;
}

static void plan_geom(rtpp_ctx *ctx, PlanGeom *g) {
    NodeStore *s = ctx->nodes;
    const int cellpx = 1 << s->shift, scpx = kSCT * 32;
    g->shift = s->shift, g->gw = s->gw, g->gh = s->gh, g->c = scpx >= cellpx ? scpx / cellpx : 1;
    g->scw = (s->gw + g->c - 1) / g->c, g->sch = (s->gh + g->c - 1) / g->c;
    g->W = ctx->W, g->H = ctx->H, g->tilesX = ctx->tilesX, g->tilesY = ctx->tilesY, g->k = s->knn_k;
    g->t3 = 0.f;
}

// The slot order for the current nodes / table / plane shape.
static int plan_order(rtpp_ctx *ctx) {
    NodeStore *s = ctx->nodes;
    if (!s || !s->knn_table) RTPP_FAIL(ctx, RTPP_ERR_STATE, "rtpp_knn_table_build has not been called");
    if (s->knn_n != s->n || s->n_sorted != s->n) RTPP_FAIL(ctx, RTPP_ERR_STATE, "nodes were appended after rtpp_knn_table_build");
    if (!ctx->plane[0]) RTPP_FAIL(ctx, RTPP_ERR_STATE, "no map plane has been set");
    if (!s->plan) {
        s->plan = new (std::nothrow) TickPlan();
        if (!s->plan) return RTPP_ERR_NOMEM;
    }
    TickPlan *p = s->plan;
    if (p->order_n == s->n && p->order_k == s->knn_k && p->W == ctx->W && p->H == ctx->H) return RTPP_OK;
    p->order_n = -1, p->edges_divisor = 0, p->usable = false, p->applies = false;
    p->order_k = s->knn_k, p->W = ctx->W, p->H = ctx->H;
    p->n_slots = p->n_groups = 0;
    if ((1 << s->shift) > kSCT * 32 || s->n < 1) { // hash cells larger than a supercell: the plan does not apply
        p->order_n = s->n;
        return RTPP_OK;
    }
    PlanGeom g;
    plan_geom(ctx, &g);
    p->scw = g.scw, p->sch = g.sch;
    const int nsc = g.scw * g.sch;
    int rc;
    if ((rc = plan_grow(ctx, &p->sc_start, &p->cap_sc, (size_t)nsc + 1))) return rc;
    if (!p->stats) RTPP_CK(ctx, cudaMalloc(&p->stats, 32));
    if (!p->far_done) {
        RTPP_CK(ctx, cudaMalloc(&p->far_done, 16));
        RTPP_CK(ctx, cudaMemsetAsync(p->far_done, 0, 16, ctx->stream));
        p->far_target = 0, p->work_base = 0;
    }
    if (!p->slot_of_id) RTPP_CK(ctx, cudaMalloc(&p->slot_of_id, s->capacity * sizeof(int)));
    // padded counts -> first slots (the scan's scratch of the hash build is large enough: nsc <= cells)
    int *cnt = s->cursor; // free between hash builds (gw*gh + 1 ints >= nsc)
    plan_count_kernel<<<(nsc + 255) / 256, 256, 0, ctx->stream>>>(g, s->cell_start, cnt);
    RTPP_LAUNCHED(ctx);
    if ((rc = rtpp_exclusive_scan(ctx, cnt, nsc, p->sc_start, cnt, s->scan_totals))) return rc;
    int n_slots = 0;
    RTPP_CK(ctx, cudaMemcpyAsync(&n_slots, p->sc_start + nsc, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    RTPP_CK(ctx, cudaStreamSynchronize(ctx->stream));
    p->n_slots = n_slots, p->n_groups = n_slots / 32;
    if ((rc = plan_grow(ctx, &p->slot_id, &p->cap_slot_id, (size_t)n_slots + 32))) return rc;
    if ((rc = plan_grow(ctx, &p->group_sc, &p->cap_group, (size_t)n_slots / 32 + 1))) return rc;
    RTPP_CK(ctx, cudaMemsetAsync(p->stats, 0, 32, ctx->stream));
    plan_slots_kernel<<<nsc, 256, 0, ctx->stream>>>(g, s->cell_start, s->sid, p->sc_start, p->slot_id, p->group_sc, p->slot_of_id, p->stats);
    RTPP_LAUNCHED(ctx);
    unsigned long long st[3] = {0, 0, 0};
    RTPP_CK(ctx, cudaMemcpyAsync(st, p->stats, 24, cudaMemcpyDeviceToHost, ctx->stream));
    RTPP_CK(ctx, cudaStreamSynchronize(ctx->stream));
    p->canonical = st[2] == 0;
    p->applies = p->n_groups > 0;
    p->order_n = s->n;
    return RTPP_OK;
}

// The plan's edges for `divisor` (neighbour coordinates, masks, far list).  plan->usable tells whether the windowed kernel
// applies; when it does not, callers use the gathering kernels of edges.cu.
static int plan_build(rtpp_ctx *ctx, int divisor) {
    if (divisor <= 0) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "divisor must be positive");
    int rc = plan_order(ctx);
    if (rc) return rc;
    NodeStore *s = ctx->nodes;
    TickPlan *p = s->plan;
    if (!p->applies || p->edges_divisor == divisor) return RTPP_OK;
    p->usable = false, p->edges_divisor = 0;
    PlanGeom g;
    plan_geom(ctx, &g);
    MarchThresholds mt;
    rtpp_march_thresholds(divisor, &mt);
    g.t3 = mt.t[3], g.mt = mt;
    const size_t n_slots = (size_t)p->n_slots;
    if ((rc = plan_grow(ctx, &p->slot_xy, &p->cap_slot_xy, n_slots + 32))) return rc;
    if ((rc = plan_grow(ctx, &p->far_mask, &p->cap_far, n_slots + 32))) return rc;
    if ((rc = plan_grow(ctx, &p->valid_mask, &p->cap_valid, n_slots + 32))) return rc;
    if ((rc = plan_grow(ctx, &p->far_off, &p->cap_far_off, n_slots + 32))) return rc;
    if ((rc = plan_grow(ctx, &p->ent, &p->cap_ent, n_slots * s->knn_k + 32))) return rc;
    if ((rc = plan_grow(ctx, &p->group_cost, &p->cap_group_cost, (size_t)p->n_groups + 32))) return rc;
    if ((rc = plan_grow(ctx, &p->item_group, &p->cap_item_group, (size_t)p->n_groups + 32))) return rc;
    // per-slot far counts live in far_off until the scan turns them into offsets; the scan's block totals need their own
    // scratch sized for n_slots elements (n_slots can exceed the node capacity the hash's scan_totals is sized for)
    void *totals = nullptr;
    if ((rc = rtpp_scratch(ctx, 3, (n_slots / kScanTile + 8) * sizeof(int), &totals))) return rc;
    RTPP_CK(ctx, cudaMemsetAsync(p->stats, 0, 32, ctx->stream));
    const int eb = (p->n_groups + 7) / 8;
    plan_entries_kernel<<<eb, 256, 0, ctx->stream>>>(g, p->n_groups, 0, p->group_sc, p->slot_id, s->xy, p->slot_xy, s->knn_pos, s->sxy, p->ent, p->far_mask,
                                                     p->valid_mask, p->far_off, nullptr, nullptr, p->stats, p->group_cost);
    RTPP_LAUNCHED(ctx);
    if ((rc = rtpp_exclusive_scan(ctx, p->far_off, (int)n_slots, p->far_off, p->far_off, (int *)totals))) return rc;
    unsigned long long st[2] = {0, 0};
    RTPP_CK(ctx, cudaMemcpyAsync(st, p->stats, 16, cudaMemcpyDeviceToHost, ctx->stream));
    RTPP_CK(ctx, cudaStreamSynchronize(ctx->stream));
    p->far_edges = (long long)st[0], p->valid_edges = (long long)st[1];
    if ((rc = plan_grow(ctx, &p->far_list, &p->cap_far_list, (size_t)p->far_edges + 32))) return rc;
    if ((rc = plan_grow(ctx, &p->far_bits, &p->cap_far_bits, (size_t)p->far_edges / 32 + 8))) return rc;
    if (p->far_edges > 0) {
        plan_entries_kernel<<<eb, 256, 0, ctx->stream>>>(g, p->n_groups, 1, p->group_sc, p->slot_id, s->xy, p->slot_xy, s->knn_pos, s->sxy, p->ent,
                                                         p->far_mask, p->valid_mask, nullptr, p->far_off, p->far_list, p->stats, nullptr);
        RTPP_LAUNCHED(ctx);
    }
    { // deal order: longest groups first (ties by index: the same permutation on every rank)
        std::vector<int> &cost = p->h_cost, order((size_t)p->n_groups);
        cost.resize((size_t)p->n_groups);
        RTPP_CK(ctx, cudaMemcpyAsync(cost.data(), p->group_cost, cost.size() * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        RTPP_CK(ctx, cudaStreamSynchronize(ctx->stream));
        for (int i = 0; i < p->n_groups; i++) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return cost[x] > cost[y]; });
        RTPP_CK(ctx, cudaMemcpyAsync(p->item_group, order.data(), order.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        RTPP_CK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    // a plan whose edges mostly leave the window (a sparse tree: k-NN radius beyond the 32-pixel halo) is not worth running
    static const double max_far = getenv("RTPP_TICK_MAX_FAR") ? atof(getenv("RTPP_TICK_MAX_FAR")) : 0.25;
    p->usable = (double)p->far_edges <= max_far * (double)(p->valid_edges > 0 ? p->valid_edges : 1);
    p->edges_divisor = divisor;
    return RTPP_OK;
}

// one resident wave of persistent blocks
static int tick_grid(rtpp_ctx *ctx, bool peers, bool bulk) {
    int per_sm = 0;
    cudaError_t e;
    if (peers) e = bulk ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tick_check_kernel<true, true>, 32 * kTickWarps, 0)
                        : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tick_check_kernel<true, false>, 32 * kTickWarps, 0);
    else e = bulk ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tick_check_kernel<false, true>, 32 * kTickWarps, 0)
                  : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tick_check_kernel<false, false>, 32 * kTickWarps, 0);
    if (e != cudaSuccess || per_sm < 1) {
        (void)cudaGetLastError();
        per_sm = 2;
    }
    return ctx->sm_count * per_sm;
}

// fills the launch arguments and advances the plan's monotonic far counter for ONE launch
static void tick_fill_args(rtpp_ctx *ctx, const PlaneView &pv, int divisor, uint32_t *d_tree_bits, uint32_t *d_row_bits, int rank, int world, int grid,
                           TickArgs *a) {
    NodeStore *s = ctx->nodes;
    TickPlan *p = s->plan;
    static const int parts = getenv("RTPP_TICK_PARTS") ? atoi(getenv("RTPP_TICK_PARTS")) : 3;
    a->pl = pv;
    rtpp_march_thresholds(divisor, &a->mt);
    volatile float one = 1.0f;
    a->one = one;
    a->far_list = p->far_list, a->far_bits = p->far_bits, a->far_done = p->far_done, a->n_far = (int)p->far_edges;
    a->n_far_words = (int)(((size_t)p->far_edges + 31) / 32);
    p->far_target += (unsigned)a->n_far_words; // this launch's far words bring the counter here
    a->far_target = p->far_target;
    a->xy = s->xy, a->cp = s->cp, a->n = s->n;
    a->n_tree_words = (int)(((size_t)(s->n - 1) + 31) / 32);
    a->tree_bits = d_tree_bits;
    a->n_groups = p->n_groups, a->k = p->order_k, a->scw = p->scw;
    a->group_sc = p->group_sc, a->far_off = p->far_off, a->slot_xy = p->slot_xy, a->ent = p->ent;
    static const bool lpt = !(getenv("RTPP_TICK_LPT") && atoi(getenv("RTPP_TICK_LPT")) == 0);
    a->item_group = lpt ? p->item_group : nullptr;
    a->far_mask = p->far_mask, a->valid_mask = p->valid_mask;
    a->row_bits = d_row_bits;
    a->stride = world, a->first = rank, a->parts = parts, a->timing = nullptr, a->item_log = nullptr;
    const size_t shared = (size_t)a->n_tree_words + (size_t)a->n_groups;
    const size_t mine = shared > (size_t)rank ? (shared - rank + world - 1) / world : 0;
    a->n_items = (unsigned)((size_t)a->n_far_words + mine);
    // four fifths of the items are dealt round-robin (whole rounds of the launch's warps), the rest is drawn dynamically
    static const int static_pct = getenv("RTPP_TICK_STATIC_PCT") ? std::min(100, std::max(0, atoi(getenv("RTPP_TICK_STATIC_PCT")))) : 80;
    const unsigned n_warps = (unsigned)grid * kTickWarps;
    a->n_static = (unsigned)((unsigned long long)a->n_items * static_pct / 100 / n_warps) * n_warps;
    a->work = p->far_done + 1;
    a->work_base = p->work_base;
    p->work_base += (a->n_items - a->n_static) + n_warps; // every warp draws one ticket past the end
}
// Last line of defence before a launch: every pointer the kernel dereferences unconditionally must be set.  (A refactoring once
// dropped five assignments of tick_fill_args; the kernels ran on uninitialised pointers and faulted every GPU of a 4-GPU box.)
static int tick_args_ok(rtpp_ctx *ctx, const TickArgs &a) {
    const bool ok = a.pl.w && a.work && a.far_done && a.xy && a.cp && a.tree_bits && a.row_bits && a.group_sc && a.far_off && a.slot_xy && a.ent &&
                    a.far_mask && a.valid_mask && (a.n_far == 0 || (a.far_list && a.far_bits)) && a.stride >= 1 && a.first >= 0 && a.first < a.stride &&
                    a.k >= 1 && a.k <= 32 && a.n >= 1 && a.one == 1.0f;
    if (!ok) RTPP_FAIL(ctx, RTPP_ERR_STATE, "internal: incomplete tick launch arguments");
    return RTPP_OK;
}
static bool tick_use_bulk() {
    static const bool bulk = !(getenv("RTPP_TICK_BULK") && atoi(getenv("RTPP_TICK_BULK")) == 0);
    return bulk;
}

// RTPP_TICK_TIMING=1 (profiling only): every launch is followed by a synchronisation and a one-line summary of the blocks' time
// stamps on stderr — when the first / last block started, when the blocks finished their items, when they left the kernel
static int tick_timing_begin(rtpp_ctx *ctx, int grid, TickArgs *a) {
    static const bool on = getenv("RTPP_TICK_TIMING") != nullptr;
    if (!on) return RTPP_OK;
    void *d = nullptr;
    int rc = rtpp_scratch(ctx, 2, (size_t)grid * 32, &d);
    if (rc) return rc;
    RTPP_CK(ctx, cudaMemsetAsync(d, 0, (size_t)grid * 32, ctx->stream));
    a->timing = (unsigned long long *)d;
    if (getenv("RTPP_TICK_ITEMS")) {
        if ((rc = rtpp_scratch(ctx, 3, (size_t)(a->n_items + 1) * 8, &d))) return rc;
        RTPP_CK(ctx, cudaMemsetAsync(d, 0, (size_t)(a->n_items + 1) * 8, ctx->stream));
        a->item_log = (unsigned long long *)d;
    }
    return RTPP_OK;
}
static int tick_timing_end(rtpp_ctx *ctx, int grid, TickArgs *a) {
    if (!a->timing) return RTPP_OK;
    std::vector<unsigned long long> h((size_t)grid * 4);
    RTPP_CK(ctx, cudaMemcpyAsync(h.data(), a->timing, h.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    RTPP_CK(ctx, cudaStreamSynchronize(ctx->stream));
    unsigned long long t0 = ~0ull, s1 = 0, w0 = ~0ull, w1 = 0, e1 = 0, longest = 0;
    double wsum = 0, lsum = 0;
    for (int b = 0; b < grid; b++) {
        t0 = std::min(t0, h[4 * b]), s1 = std::max(s1, h[4 * b]);
        w0 = std::min(w0, h[4 * b + 1]), w1 = std::max(w1, h[4 * b + 1]), e1 = std::max(e1, h[4 * b + 2]);
        longest = std::max(longest, h[4 * b + 3]), lsum += (double)(h[4 * b + 3] >> 32);
    }
    for (int b = 0; b < grid; b++) wsum += (double)(h[4 * b + 1] - t0);
    if (a->item_log) { // raw dump for tools/tick_items.py: header, per-item records, the group item codes, the groups' cost estimates
        TickPlan *p = ctx->nodes->plan;
        std::vector<unsigned long long> it((size_t)a->n_items);
        RTPP_CK(ctx, cudaMemcpyAsync(it.data(), a->item_log, it.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
        const int ncodes = a->item_group ? p->n_groups : 0;
        std::vector<int> codes((size_t)ncodes);
        if (ncodes) RTPP_CK(ctx, cudaMemcpyAsync(codes.data(), a->item_group, codes.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
        RTPP_CK(ctx, cudaStreamSynchronize(ctx->stream));
        if (FILE *f = fopen(getenv("RTPP_TICK_ITEMS"), "wb")) {
            const long long hdr[8] = {(long long)a->n_items, a->n_far_words, a->n_tree_words, a->stride, a->first, ncodes, (long long)p->h_cost.size(), (long long)a->n_static};
            fwrite(hdr, 8, 8, f);
            fwrite(it.data(), 8, it.size(), f);
            fwrite(codes.data(), 4, codes.size(), f);
            fwrite(p->h_cost.data(), 4, p->h_cost.size(), f);
            fclose(f);
        }
    }
    const unsigned li = (unsigned)longest;
    const char *kind = li < (unsigned)a->n_far_words ? "far word" : ((size_t)(li - a->n_far_words) * a->stride + a->first < (size_t)a->n_tree_words ? "tree word" : "window group");
    fprintf(stderr, "[rtpp tick timing] blocks %d items %u: last start +%.1f us, work done first +%.1f mean +%.1f last +%.1f us, exit +%.1f us", grid, a->n_items,
            (s1 - t0) * 1e-3, (w0 - t0) * 1e-3, wsum / grid * 1e-3, (w1 - t0) * 1e-3, (e1 - t0) * 1e-3);
    if (longest) // items are only timed by the multi-GPU kernels
        fprintf(stderr, "; longest item %.1f us (item %u, %s), mean of the blocks' longest %.1f us", (longest >> 32) * 1e-3, li, kind, lsum / grid * 1e-3);
    fprintf(stderr, "\n");
    return RTPP_OK;
}

// shared with edges.cu: runs the windowed tick when the plan applies.  *ran = false: the caller takes its own path.
int rtpp_tick_window_run(rtpp_ctx *ctx, int divisor, int plane_mask, uint32_t *d_tree_bits, uint32_t *d_row_bits, bool *ran) {
    *ran = false;
    PlaneView pv;
    int rc = rtpp_select_plane(ctx, plane_mask, &pv);
    if (rc) return rc;
    rc = plan_build(ctx, divisor);
    if (rc) return rc;
    TickPlan *p = ctx->nodes->plan;
    if (!p->usable) return RTPP_OK;
    TickArgs a{}; // zero first: a field the filler forgets is then null and caught by tick_args_ok, not a wild pointer
    const int grid = tick_grid(ctx, false, tick_use_bulk());
    tick_fill_args(ctx, pv, divisor, d_tree_bits, d_row_bits, 0, 1, grid, &a);
    TickPeers none;
    memset(&none, 0, sizeof(none));
    if ((rc = tick_args_ok(ctx, a))) return rc;
    rc = tick_timing_begin(ctx, grid, &a);
    if (rc) return rc;
    if (tick_use_bulk()) tick_check_kernel<false, true><<<grid, 32 * kTickWarps, 0, ctx->stream>>>(a, none);
    else tick_check_kernel<false, false><<<grid, 32 * kTickWarps, 0, ctx->stream>>>(a, none);
    RTPP_LAUNCHED(ctx);
    rc = tick_timing_end(ctx, grid, &a);
    if (rc) return rc;
    *ran = true;
    return RTPP_OK;
}

// legacy order for rtpp_tick_edges_check[_dev]: windowed kernel into plan scratch + repack
int rtpp_tick_window_run_packed(rtpp_ctx *ctx, int divisor, int plane_mask, uint32_t *d_bits, bool *ran) {
    *ran = false;
    int rc = plan_build(ctx, divisor);
    if (rc) return rc;
    NodeStore *s = ctx->nodes;
    TickPlan *p = s->plan;
    if (!p->usable) return RTPP_OK;
    const size_t tree_words = ((size_t)(s->n - 1) + 31) / 32 + 1;
    if ((rc = plan_grow(ctx, &p->row_scratch, &p->cap_row_scratch, (size_t)p->n_slots))) return rc;
    if ((rc = plan_grow(ctx, &p->tree_scratch, &p->cap_tree_scratch, tree_words))) return rc;
    rc = rtpp_tick_window_run(ctx, divisor, plane_mask, p->tree_scratch, p->row_scratch, ran);
    if (rc || !*ran) return rc;
    const size_t ne = (size_t)(s->n - 1) + (size_t)s->n * s->knn_k, nwords = (ne + 31) / 32;
    size_t blocks = (nwords + 255) / 256, cap = (size_t)ctx->sm_count * 16;
    tick_repack_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, ctx->stream>>>(p->tree_scratch, p->row_scratch, p->slot_of_id, s->n, s->knn_k, nwords, d_bits);
    RTPP_LAUNCHED(ctx);
    return RTPP_OK;
}

extern "C" {

int rtpp_tick_plan_info(rtpp_ctx *ctx, int divisor, size_t *n_slots, const int32_t **d_slot_id, const int32_t **d_slot_of_id, int *usable,
                        double *far_fraction) {
    if (!ctx) return RTPP_ERR_INVALID;
    int rc = plan_build(ctx, divisor);
    if (rc) return rc;
    TickPlan *p = ctx->nodes->plan;
    if (n_slots) *n_slots = (size_t)p->n_slots;
    if (d_slot_id) *d_slot_id = p->slot_id;
    if (d_slot_of_id) *d_slot_of_id = p->slot_of_id;
    if (usable) *usable = p->usable ? 1 : 0;
    if (far_fraction) *far_fraction = p->applies && p->valid_edges > 0 ? (double)p->far_edges / (double)p->valid_edges : (p->applies ? 0.0 : 1.0);
    return RTPP_OK;
}

int rtpp_tick_check_dev(rtpp_ctx *ctx, int divisor, int plane_mask, uint32_t *d_tree_bits, uint32_t *d_row_bits) {
    if (!ctx || !d_tree_bits || !d_row_bits) return RTPP_ERR_INVALID;
    bool ran = false;
    int rc = rtpp_tick_window_run(ctx, divisor, plane_mask, d_tree_bits, d_row_bits, &ran);
    if (rc) return rc;
    if (!ran) RTPP_FAIL(ctx, RTPP_ERR_STATE, "the windowed tick plan does not apply to this node store (see rtpp_tick_plan_info); use rtpp_tick_edges_check");
    return RTPP_OK;
}

// Host buffers.  Page-locked (cudaHostAlloc / rtpp_host_alloc) result buffers are written by the kernel itself, one
// coalesced 128-byte store per group as the warps retire; pageable ones go through device scratch and a copy.
int rtpp_tick_check(rtpp_ctx *ctx, int divisor, int plane_mask, uint32_t *tree_bits, uint32_t *row_bits) {
    if (!ctx || !tree_bits || !row_bits) return RTPP_ERR_INVALID;
    int rc = plan_build(ctx, divisor);
    if (rc) return rc;
    NodeStore *s = ctx->nodes;
    TickPlan *p = s->plan;
    if (!p->usable) RTPP_FAIL(ctx, RTPP_ERR_STATE, "the windowed tick plan does not apply to this node store (see rtpp_tick_plan_info); use rtpp_tick_edges_check");
    const size_t tree_words = ((size_t)(s->n - 1) + 31) / 32;
    cudaPointerAttributes at1, at2;
    const bool m1 = cudaPointerGetAttributes(&at1, tree_bits) == cudaSuccess && at1.type == cudaMemoryTypeHost && at1.devicePointer;
    const bool m2 = cudaPointerGetAttributes(&at2, row_bits) == cudaSuccess && at2.type == cudaMemoryTypeHost && at2.devicePointer;
    (void)cudaGetLastError();
    bool ran = false;
    if (m1 && m2) {
        rc = rtpp_tick_window_run(ctx, divisor, plane_mask, (uint32_t *)at1.devicePointer, (uint32_t *)at2.devicePointer, &ran);
        if (rc) return rc;
    } else {
        if ((rc = plan_grow(ctx, &p->row_scratch, &p->cap_row_scratch, (size_t)p->n_slots))) return rc;
        if ((rc = plan_grow(ctx, &p->tree_scratch, &p->cap_tree_scratch, tree_words + 1))) return rc;
        rc = rtpp_tick_window_run(ctx, divisor, plane_mask, p->tree_scratch, p->row_scratch, &ran);
        if (rc) return rc;
        RTPP_CK(ctx, cudaMemcpyAsync(tree_bits, p->tree_scratch, tree_words * 4, cudaMemcpyDeviceToHost, ctx->stream));
        RTPP_CK(ctx, cudaMemcpyAsync(row_bits, p->row_scratch, (size_t)p->n_slots * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    RTPP_CK(ctx, cudaStreamSynchronize(ctx->stream));
    return RTPP_OK;
}

// ONE tick over the GPUs of a node: blocks round-robin over the ranks, verdict words stored into every rank's buffers,
// last-block-done flag exchange inside the launch.  row_all[r] / tree_all[r] / flags_all[r]: rank r's buffers as mapped in
// this process (rtpp_peer_alloc / rtpp_peer_open; flags: RTPP_PEER_FLAG_INTS ints, zero-initialised).
int rtpp_tick_check_allgather_dev(rtpp_ctx *ctx, int divisor, int plane_mask, uint32_t *const *tree_all, uint32_t *const *row_all,
                                  int32_t *const *flags_all, int rank, int world, int epoch) {
    if (!ctx || !tree_all || !row_all || !flags_all) return RTPP_ERR_INVALID;
    if (world < 1 || world > RTPP_MAX_PEERS || rank < 0 || rank >= world) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "bad rank / world size");
    PlaneView pv;
    int rc = rtpp_select_plane(ctx, plane_mask, &pv);
    if (rc) return rc;
    rc = plan_build(ctx, divisor);
    if (rc) return rc;
    NodeStore *s = ctx->nodes;
    TickPlan *p = s->plan;
    if (!p->usable) RTPP_FAIL(ctx, RTPP_ERR_STATE, "the windowed tick plan does not apply to this node store");
    if (!p->canonical) RTPP_FAIL(ctx, RTPP_ERR_STATE, "a hash cell holds too many nodes for a rank-independent slot order");
    TickArgs a{}; // zero first: a field the filler forgets is then null and caught by tick_args_ok, not a wild pointer
    const int grid = tick_grid(ctx, true, tick_use_bulk());
    tick_fill_args(ctx, pv, divisor, tree_all[rank], row_all[rank], rank, world, grid, &a);
    TickPeers peers;
    memset(&peers, 0, sizeof(peers));
    peers.rank = rank, peers.world = world, peers.epoch = epoch;
    for (int r = 0; r < world; r++) {
        peers.flags[r] = flags_all[r];
        if (r != rank) peers.row[peers.n] = row_all[r], peers.tree[peers.n] = tree_all[r], peers.n++;
    }
    peers.own_row = row_all[rank], peers.own_tree = tree_all[rank];
    peers.row_words = (size_t)p->n_slots, peers.tree_words = ((size_t)(s->n - 1) + 31) / 32;
    peers.spin_limit = (long long)(ctx->peer_timeout_ms * 1e6 / 100.0) + 1; // one poll is ~100 ns (64 ns sleep + a system-scope load)
    for (int r = 0; r < world; r++)
        if (!tree_all[r] || !row_all[r] || !flags_all[r]) RTPP_FAIL(ctx, RTPP_ERR_INVALID, "a rank's result / flag buffer is NULL");
    if ((rc = tick_args_ok(ctx, a))) return rc;
    rc = tick_timing_begin(ctx, grid, &a);
    if (rc) return rc;
    if (tick_use_bulk()) tick_check_kernel<true, true><<<grid, 32 * kTickWarps, 0, ctx->stream>>>(a, peers);
    else tick_check_kernel<true, false><<<grid, 32 * kTickWarps, 0, ctx->stream>>>(a, peers);
    RTPP_LAUNCHED(ctx);
    rc = tick_timing_end(ctx, grid, &a);
    if (rc) return rc;
    return RTPP_OK;
}

int rtpp_peer_set_timeout_ms(rtpp_ctx *ctx, double ms) {
    if (!ctx || !(ms > 0.0)) return RTPP_ERR_INVALID;
    ctx->peer_timeout_ms = ms;
    return RTPP_OK;
}

// Synchronises the stream and reports how many fused exchanges on this rank's flag array timed out since the last call
// (their result buffers read "blocked" everywhere: fail closed).  Clears the counter.
int rtpp_peer_status(rtpp_ctx *ctx, int32_t *d_flags, int *timeouts) {
    if (!ctx || !d_flags || !timeouts) return RTPP_ERR_INVALID;
    int v = 0;
    RTPP_CK(ctx, cudaMemcpyAsync(&v, d_flags + kFlagTimeouts, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    RTPP_CK(ctx, cudaStreamSynchronize(ctx->stream));
    if (v) RTPP_CK(ctx, cudaMemsetAsync(d_flags + kFlagTimeouts, 0, sizeof(int), ctx->stream));
    *timeouts = v;
    return RTPP_OK;
}

} // extern "C"
