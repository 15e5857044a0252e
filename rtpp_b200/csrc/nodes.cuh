// nodes.cuh — the GPU-resident node store shared by grid.cu (hash + k-NN), tree.cu (costs, parents,
// propagate, relax) and the planner engines.
#pragma once
#include "common.cuh"
#include "knn.cuh"

// C1: RTrrt_Cost {bool collision; float cost} (reference RTRRT.h:55-91) + the parent link of
// RTrrt_NodeAttribute (RTRRT.h:108-113) packed into one 64-bit word:
//   bits 63     collision
//   bits 62..32 binary32 bit pattern of cost (cost >= 0, so the pattern orders like the value)
//   bits 31..0  parent node id (0xffffffff = none)
// Unsigned comparison of the high half is exactly the reference's lexicographic operator<, so a
// relaxation is one 64-bit atomicMin.
typedef unsigned long long cpword;
__host__ __device__ __forceinline__ uint32_t cp_hi(cpword w) { return (uint32_t)(w >> 32); }
__host__ __device__ __forceinline__ int cp_parent(cpword w) { return (int)(uint32_t)w; }
__host__ __device__ __forceinline__ cpword cp_make(uint32_t hi, int parent) { return ((cpword)hi << 32) | (uint32_t)parent; }
__device__ __forceinline__ uint32_t cost_pack(bool col, float c) { return ((uint32_t)col << 31) | (__float_as_uint(c) & 0x7fffffffu); }
__device__ __forceinline__ bool cost_col(uint32_t hi) { return hi >> 31; }
__device__ __forceinline__ float cost_val(uint32_t hi) { return __uint_as_float(hi & 0x7fffffffu); }
// operator+ of RTrrt_Cost (RTRRT.h:59-62): OR the flags, one binary32 add
__device__ __forceinline__ uint32_t cost_add(uint32_t a, uint32_t b) {
    return ((a | b) & 0x80000000u) | (__float_as_uint(__fadd_rn(cost_val(a), cost_val(b))) & 0x7fffffffu);
}

struct NodeStore {
    int shift = 4, gw = 0, gh = 0;
    size_t capacity = 0;
    int n = 0, n_sorted = 0;
    float2 *xy = nullptr;       // [capacity] coordinates by node id
    float2 *sxy = nullptr;      // [capacity] cell-sorted copy
    int *sid = nullptr;         // [capacity] sorted position -> id
    int *cell_of = nullptr;     // [capacity] scratch
    int *cell_start = nullptr;  // [gw*gh + 1]
    int *cursor = nullptr;      // [gw*gh + 1]
    int *scan_totals = nullptr; // scratch of the prefix sums
    unsigned char *gap = nullptr; // [gw*gh] Chebyshev distance in cells to the nearest non-empty cell (ring-walk start)
    cpword *cp = nullptr;       // [capacity] cost|parent words
    // tree.cu scratch (allocated on first use)
    uint32_t *edge_w = nullptr; // per node: weight of the edge (parent -> node)
    int *child_start = nullptr; // [capacity + 2]
    int *children = nullptr;    // [capacity]
    int *frontier = nullptr;    // [2 * capacity]
    int *counters = nullptr;    // [8]
    int *knn_table = nullptr;   // resident all-k-NN table [n * knn_k] (rtpp_knn_table_build)
    int knn_k = 0, knn_n = 0;
    int *knn_pos = nullptr;     // the same table in cell-sorted positions (coherent coordinate gathers)
    int *pos_of = nullptr;      // node id -> cell-sorted position
    float4 *edge_list = nullptr; // materialised tick edge list (tree edges, then node -> k-NN edges), NaN = no edge
    size_t edge_list_cap = 0;
    int edge_list_n = -1;       // node count the node -> k-NN part was built for (-1 = stale)
    struct TickPlan *plan = nullptr; // tick.cu: supercell-ordered plan of the windowed tick check
};

int rtpp_nodes_build_internal(rtpp_ctx *ctx);
void rtpp_tick_plan_free(NodeStore *s);       // tick.cu
void rtpp_tick_plan_invalidate(NodeStore *s); // tick.cu
// tick.cu: the windowed tick; *ran = false when the plan does not apply (the caller takes the gathering kernels)
int rtpp_tick_window_run(rtpp_ctx *ctx, int divisor, int plane_mask, uint32_t *d_tree_bits, uint32_t *d_row_bits, bool *ran);
int rtpp_tick_window_run_packed(rtpp_ctx *ctx, int divisor, int plane_mask, uint32_t *d_bits, bool *ran);
int rtpp_grid_view(rtpp_ctx *ctx, GridView *g);
