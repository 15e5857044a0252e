/* rtpp_gpu.h — C ABI of librtpp_gpu.so, the B200 (sm_100a) implementation of the data-parallel hot
 * path shared by the four planners of TZY-H/RTPP (RT-RRT*, RT-FMT, AM-RRT*, CDRT-RRT*).
 *
 * The reference has no plugin/FFI seam: its hot path is a set of member functions of the planner
 * classes.  This header is the seam a maintainer would bind; every entry point names the
 * reference routine (file:line, relative to the reference root) whose work it takes over.
 * INTEGRATION.md shows the planner-side call sites.
 *
 * Conventions
 *   - extern "C", opaque rtpp_ctx*, plain pointers and sizes, no C++/torch types.
 *   - every call returns 0 (RTPP_OK) or a negative rtpp_status; rtpp_last_error() has the text.
 *     Nothing throws; nothing falls back to the CPU: without a CUDA device rtpp_create fails.
 *   - one ctx per host thread; calls on one ctx are ordered on the ctx stream.
 *   - functions without suffix take HOST buffers (copy in, run, copy out, synchronise);
 *     functions ending in _dev take DEVICE pointers, are stream-ordered and do not synchronise.
 *   - coordinates are pixels, x = column in [0,W), y = row in [0,H); finite, |v| < 2^20.
 *   - "blocked bits": edge i -> bit (i & 31) of word (i >> 5); 1 = the reference's
 *     CrashDetection would return true.  Unused high bits of the last word are 0.
 *   - arithmetic is bit-exact with the reference's scalar code (binary32/binary64, round to
 *     nearest even, no FMA contraction, truncating float->int), see DESIGN.md.
 */
#ifndef RTPP_GPU_H
#define RTPP_GPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define RTPP_API __attribute__((visibility("default")))
#else
#define RTPP_API
#endif

typedef struct rtpp_ctx rtpp_ctx;

enum rtpp_status {
    RTPP_OK = 0,
    RTPP_ERR_INVALID = -1,     /* bad argument */
    RTPP_ERR_CUDA = -2,        /* CUDA runtime error (text in rtpp_last_error) */
    RTPP_ERR_NOMEM = -3,       /* device or host allocation failed */
    RTPP_ERR_STATE = -4,       /* call order: e.g. edge check before any plane was set */
    RTPP_ERR_CAPACITY = -5     /* a fixed-capacity store (nodes, queue) is full */
};

/* plane ids; also the bits of a plane_mask */
enum { RTPP_PLANE_STATIC = 1, RTPP_PLANE_DYNAMIC = 2 };

#define RTPP_ABI_VERSION 1

/* ---------------------------------------------------------------- lifecycle ------------------ */
RTPP_API int rtpp_abi_version(void);
RTPP_API const char *rtpp_status_string(int status);
RTPP_API int rtpp_create(int device, rtpp_ctx **out);
RTPP_API int rtpp_destroy(rtpp_ctx *ctx);
RTPP_API const char *rtpp_last_error(rtpp_ctx *ctx);
/* Run on a caller-owned cudaStream_t (e.g. torch's current stream); NULL restores the ctx stream. */
RTPP_API int rtpp_set_stream(rtpp_ctx *ctx, void *cuda_stream);
RTPP_API int rtpp_sync(rtpp_ctx *ctx);
/* Number of kernels this ctx has launched so far (bench.py's gpu_launches). */
RTPP_API long long rtpp_kernel_launches(rtpp_ctx *ctx);
RTPP_API int rtpp_device_sm_count(rtpp_ctx *ctx, int *sm_count);
/* Roofline denominators measured in the same run as the kernels they bound (L2 bandwidth and the FP32
 * issue peak are not in MEASURED_PEAKS.json).  l2_bytes = size of the L2-resident test buffer (power of two; use the
 * bit-plane size class, e.g. 8 MiB).  out[0] = L2 random 32-byte-sector gather GB/s (L1 bypassed: the cost of a
 * bit-plane probe that misses L1), out[1] = L2 coalesced-line read GB/s, out[2] = FP32 FFMA issue TFLOP/s,
 * out[3] = HBM copy GB/s (read + write bytes over a 1 GiB buffer).  Best of `reps` timed launches after warm-up. */
RTPP_API int rtpp_measure_peaks(rtpp_ctx *ctx, size_t l2_bytes, int reps, double out[4]);

/* device / pinned-host buffers for resident workflows (bench.py, planner host code) */
RTPP_API int rtpp_dev_alloc(rtpp_ctx *ctx, size_t bytes, void **dptr);
RTPP_API int rtpp_dev_free(rtpp_ctx *ctx, void *dptr);
RTPP_API int rtpp_host_alloc(rtpp_ctx *ctx, size_t bytes, void **hptr); /* pinned */
RTPP_API int rtpp_host_free(rtpp_ctx *ctx, void *hptr);
RTPP_API int rtpp_h2d(rtpp_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes); /* async on the ctx stream */
RTPP_API int rtpp_d2h(rtpp_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes); /* async on the ctx stream */

/* ---------------------------------------------------------------- maps ----------------------- */
/* Replaces the cv::Mat planes the planners probe: Obsmap (RTRRT.cpp:69-79) = RTPP_PLANE_STATIC,
 * DObsmap (RTRRT.cpp:138-169, rebuilt every tick) = RTPP_PLANE_DYNAMIC.  `rowmajor` is the H*W
 * uint8 plane (non-zero = obstacle).  The device keeps bit-packed 32x32-pixel tiles (one 128-byte
 * line per tile) for S, D and the merged S|D.  Both planes must have the same W,H. */
RTPP_API int rtpp_map_set_plane(rtpp_ctx *ctx, int plane, const uint8_t *rowmajor, int W, int H);
RTPP_API int rtpp_map_set_plane_dev(rtpp_ctx *ctx, int plane, const uint8_t *d_rowmajor, int W, int H);
/* DObsmap = cv::Mat::zeros (RTRRT.cpp:140) without an upload. */
RTPP_API int rtpp_map_clear_plane(rtpp_ctx *ctx, int plane);
/* Unpack the plane selected by plane_mask (1 = S, 2 = D, 3 = S|D) back to H*W bytes (0/1). */
RTPP_API int rtpp_map_get_plane(rtpp_ctx *ctx, int plane_mask, uint8_t *out_rowmajor);
RTPP_API int rtpp_map_shape(rtpp_ctx *ctx, int *W, int *H);
/* D1 DrawDObsmap (RTRRT.cpp:138-169, RTFMT.cpp:178-209, AMRRT.cpp:202-233, CDRT.cpp:605-636) on the
 * device: clears the dynamic plane, draws n_circles filled circles — (cx, cy, r) int triples, one per
 * outline point of a type-0 obstacle, = cv::circle(.., FILLED) — and n_polys filled polygons — poly_xy
 * concatenated (x, y) int pairs, poly_sizes[p] vertices each (<= 64), = cv::fillPoly — then refreshes
 * the merged S|D plane.  The outline rotation (sinf/cosf, truncation: RTRRT.cpp:145-152) is done by the
 * caller on the host, so a tick ships a few hundred bytes instead of an H*W byte plane. */
RTPP_API int rtpp_map_draw_dynamic(rtpp_ctx *ctx, const int32_t *circles, int n_circles, const int32_t *poly_xy,
                                   const int32_t *poly_sizes, int n_polys);

/* ---------------------------------------------------------------- edge checks (K1) ----------- */
/* E1 RTRRTplanner::CrashDetection RTRRT.cpp:202-222, RTFMTplanner::CrashDetection RTFMT.cpp:242-262
 *    (divisor 5, plane_mask 3);
 * E2 RTFMTplanner::CrashDetectionNDObs RTFMT.cpp:264-283 (divisor 5, plane_mask 1);
 * E3 AMRRTplanner::CrashDetection AMRRT.cpp:266-286 (divisor 2, plane_mask 3).
 * xyxy: n edges as (ps.x, ps.y, pe.x, pe.y) binary32. */
RTPP_API int rtpp_edges_check(rtpp_ctx *ctx, const float *xyxy, size_t n, int divisor, int plane_mask,
                              uint32_t *blocked_bits);
/* --- one candidate-edge batch over the GPUs of a node (SURVEY 8(e) grain 2; one process per GPU) ---
 * rtpp_peer_alloc: device memory other ranks can map (CUDA IPC; zero-filled); exchange the 64-byte handle with any
 * host-side transport (rtpp_b200/shard.py uses torch.distributed) and map it with rtpp_peer_open.
 * rtpp_edges_check_allgather_dev: every rank holds the whole batch of n edges (the tree is replicated) and checks its
 * share — the 2048-edge blocks g with g mod world == rank, round-robin because an edge's cost varies with its length by
 * two orders of magnitude — and every block stores its verdict words into the same word range of every rank's result
 * buffer (bits_all[r], as mapped in this process; plain stores over NVLink), then the rank signals its peers and waits
 * for theirs (flags_all[r]: RTPP_PEER_FLAG_INTS zero-initialised int32 per rank, `epoch` must grow by one per call): when the
 * stream reaches the end of the call, bits_all[rank] holds the verdict words of the whole batch, as rtpp_edges_check_dev on one
 * GPU writes them.  Alternate two result buffers (epoch parity) so that a fast peer never overwrites words still being read.
 * The wait for the peers is BOUNDED (rtpp_peer_set_timeout_ms, default 1000 ms).  A rank whose peer never arrives fails closed: its
 * result buffer is filled with 1-bits (every edge reads blocked) and rtpp_peer_status reports the timeout; callers must check it
 * before trusting a batch.  d_xyxy must be 16-byte aligned and n must not exceed 2^32 - 1 (RTPP_ERR_INVALID otherwise). */
#define RTPP_MAX_PEERS 8
RTPP_API int rtpp_peer_alloc(rtpp_ctx *ctx, size_t bytes, void **dptr, unsigned char handle[64]);
RTPP_API int rtpp_peer_open(rtpp_ctx *ctx, const unsigned char handle[64], void **peer_ptr);
RTPP_API int rtpp_peer_close(rtpp_ctx *ctx, void *peer_ptr);
RTPP_API int rtpp_edges_check_allgather_dev(rtpp_ctx *ctx, const float *d_xyxy, size_t n, int divisor, int plane_mask,
                                            uint32_t *const *bits_all, int32_t *const *flags_all, int rank, int world,
                                            int epoch);
RTPP_API int rtpp_edges_check_dev(rtpp_ctx *ctx, const float *d_xyxy, size_t n, int divisor, int plane_mask,
                                  uint32_t *d_blocked_bits);
/* E5b, the ray-march inside BImap::RTRT_CrashDetection CDRT.cpp:584-601: binary64 points,
 * binary32 interpolation parameter, divisor 10, plane_mask 2 in the reference. */
RTPP_API int rtpp_edges_check_f64(rtpp_ctx *ctx, const double *xyxy, size_t n, int divisor, int plane_mask,
                                  uint32_t *blocked_bits);
RTPP_API int rtpp_edges_check_f64_dev(rtpp_ctx *ctx, const double *d_xyxy, size_t n, int divisor, int plane_mask,
                                      uint32_t *d_blocked_bits);
/* E4 AMRRTplanner::CrashDetectionStart AMRRT.cpp:287-308 (linear march, spacing len/2+2, both
 * planes): blocked bit + the last free point pstart (2 floats per edge). */
RTPP_API int rtpp_edges_first_hit(rtpp_ctx *ctx, const float *xyxy, size_t n, uint32_t *blocked_bits,
                                  float *pstart_xy);
RTPP_API int rtpp_edges_first_hit_dev(rtpp_ctx *ctx, const float *d_xyxy, size_t n, uint32_t *d_blocked_bits,
                                      float *d_pstart_xy);
/* One replanning tick's collision pass over the RESIDENT tree (nodes + parents + k-NN table on the
 * device): edge e < n-1 is the tree edge (parent(e+1) -> e+1) that UpdateTree re-costs (RTRRT.cpp:132,
 * AMRRT.cpp:196), edge n-1 + i*k + j is the rewire edge (node i -> knn[i][j]) (RTRRT.cpp:310,339,
 * AMRRT.cpp:500).  What crosses the boundary per tick is what changes per tick: the new dynamic plane
 * in (H*W bytes as DrawDObsmap produces it; NULL = unchanged), the verdict words out. */
RTPP_API int rtpp_tick_edges_check(rtpp_ctx *ctx, const uint8_t *dynamic_plane, int divisor, int plane_mask,
                                   uint32_t *blocked_bits, size_t *n_edges);
RTPP_API int rtpp_tick_edges_check_dev(rtpp_ctx *ctx, int divisor, int plane_mask, uint32_t *d_blocked_bits, size_t *n_edges);

/* The same tick with the verdicts in the layout the windowed kernel produces (csrc/tick.cu; north star: bit-packed tiles with
 * TMA / shared-memory staging of the per-block map window).  The nodes are ordered by 64x64-pixel supercell into SLOTS (groups
 * of 32, padded per supercell); the order is fixed by rtpp_knn_table_build and read with rtpp_tick_plan_info:
 *   d_slot_id[slot]     node id held by a slot (-1 = padding),   d_slot_of_id[id]  the slot of a node,
 *   usable              0 when the plan does not apply (hash cells larger than 64 pixels, or more than a quarter of the edges
 *                       leave the 32-pixel halo): rtpp_tick_check* then fail with RTPP_ERR_STATE and rtpp_tick_edges_check uses
 *                       its gathering kernels,   far_fraction  share of the edges (for `divisor`: a march finer than 16 steps
 *                       counts) that the same launch checks with the general routine instead of the window.
 * rtpp_tick_check[_dev]: tree_bits[(n-1+31)/32]: bit e = CrashDetection(parent(e+1), e+1) (UpdateTree, AMRRT.cpp:185-199;
 * RTRRT.cpp:128-134); row_bits[n_slots]: bit j of word slot = CrashDetection(node, knn[node][j]) (RewireRoot, AMRRT.cpp:496-500;
 * RTRRT.cpp:304-320), 0 where the node has fewer than k neighbours.  Host variant: page-locked buffers are written by the kernel. */
RTPP_API int rtpp_tick_plan_info(rtpp_ctx *ctx, int divisor, size_t *n_slots, const int32_t **d_slot_id, const int32_t **d_slot_of_id,
                                 int *usable, double *far_fraction);
RTPP_API int rtpp_tick_check_dev(rtpp_ctx *ctx, int divisor, int plane_mask, uint32_t *d_tree_bits, uint32_t *d_row_bits);
RTPP_API int rtpp_tick_check(rtpp_ctx *ctx, int divisor, int plane_mask, uint32_t *tree_bits, uint32_t *row_bits);
/* ONE tick over the GPUs of a node (SURVEY 8(e) grain 2; tree, table and planes replicated): the launch's blocks are dealt
 * round-robin over the ranks, every block stores its verdict words into EVERY rank's buffers (tree_all[r], row_all[r]: rank r's
 * buffers as mapped in this process, rtpp_peer_alloc / rtpp_peer_open), and the last block to retire signals the peers and waits
 * for theirs (flags_all[r]: RTPP_PEER_FLAG_INTS zero-initialised ints per rank; `epoch` must increase by one per call; alternate
 * two result buffers by epoch parity).  The wait is bounded (rtpp_peer_set_timeout_ms, default 1000 ms): when a peer never arrives
 * this rank's buffers are filled with 1-bits (every edge reads BLOCKED: fail closed) and rtpp_peer_status reports the timeout. */
#define RTPP_PEER_FLAG_INTS 16
RTPP_API int rtpp_tick_check_allgather_dev(rtpp_ctx *ctx, int divisor, int plane_mask, uint32_t *const *tree_all, uint32_t *const *row_all,
                                           int32_t *const *flags_all, int rank, int world, int epoch);
/* Synchronises the ctx stream; *timeouts = fused exchanges that timed out on this rank's flag array since the last call. */
RTPP_API int rtpp_peer_status(rtpp_ctx *ctx, int32_t *d_flags, int *timeouts);
RTPP_API int rtpp_peer_set_timeout_ms(rtpp_ctx *ctx, double ms);
/* C2 RTpoint::operator% RTRRT.h:47-52 on n point pairs (exposed for parity tests). */
RTPP_API int rtpp_dist(rtpp_ctx *ctx, const float *xyxy, size_t n, float *out);

/* ---------------------------------------------------------------- nodes + k-NN (K2/K3) ------- */
/* The node store replaces `Kd_tree NeighborSearchTree` + the coordinate keys of
 * `std::map<RTpoint, ...> Node_Map` (RTRRT.h:129-130): nodes are float2 points, node id = insertion
 * index.  The spatial hash uses 2^cell_log2-pixel cells over [0,domain_w) x [0,domain_h) (points
 * outside are binned into the border cells).  reset drops all nodes. */
RTPP_API int rtpp_nodes_reset(rtpp_ctx *ctx, int cell_log2, size_t capacity, int domain_w, int domain_h);
/* Kd_tree::insert (CGAL/Kd_tree.h:461-466; RTRRT.cpp:29,291, RTFMT.cpp:29,417, AMRRT.cpp:31,474,
 * CDRT.cpp:918): append n points (x,y float pairs); *first_id receives the id of the first one.
 * Later queries see every point appended so far. */
RTPP_API int rtpp_nodes_append(rtpp_ctx *ctx, const float *xy, size_t n, int32_t *first_id);
RTPP_API int rtpp_nodes_append_dev(rtpp_ctx *ctx, const float *d_xy, size_t n, int32_t *first_id);
RTPP_API int rtpp_nodes_count(rtpp_ctx *ctx, size_t *n);
/* Force the re-sort of the spatial hash now (otherwise lazy). */
RTPP_API int rtpp_nodes_build(rtpp_ctx *ctx);
/* Device pointer to the resident coordinates (float2 by node id). */
RTPP_API int rtpp_nodes_xy_dev(rtpp_ctx *ctx, const float **d_xy);
/* Orthogonal_k_neighbor_search(tree, q, k) (CGAL/Orthogonal_k_neighbor_search.h:51-79; call sites
 * RTRRT.cpp:96,265,304,333,368; RTFMT.cpp:97,291,427,440,477,549,677; AMRRT.cpp:160,316,444,449,496,
 * 546; CDRT.cpp:206,231,261,890): the min(k,N) nodes of smallest squared distance evaluated in
 * binary64, ascending; ties by ascending node id.  qxy: nq (x,y) binary64 pairs.  Rows of k:
 * idx (-1 padded), d2 (+inf padded; may be NULL), count = min(k,N) (may be NULL).  1 <= k <= 32. */
RTPP_API int rtpp_knn(rtpp_ctx *ctx, const double *qxy, size_t nq, int k, int32_t *idx, double *d2, int32_t *count);
RTPP_API int rtpp_knn_dev(rtpp_ctx *ctx, const double *d_qxy, size_t nq, int k, int32_t *d_idx, double *d_d2,
                          int32_t *d_count);
/* All-k-NN table of the node set itself, row i = neighbours of node i (includes i at d2 = 0, as
 * the reference's queries at tree nodes do: RTRRT.cpp:304, RTFMT.cpp:440). */
RTPP_API int rtpp_knn_nodes_dev(rtpp_ctx *ctx, int k, int32_t *d_idx, double *d_d2);

/* Resident all-k-NN table inside the ctx (rows of k ids per node), for the tick-level calls. */
RTPP_API int rtpp_knn_table_build(rtpp_ctx *ctx, int k);
RTPP_API int rtpp_knn_table_dev(rtpp_ctx *ctx, const int32_t **d_idx, int *k);

/* ---------------------------------------------------------------- tree costs (K4/K5) --------- */
/* Per-node cost and parent link: RTrrt_NodeAttribute {cost, par} (RTRRT.h:108-113) with
 * RTrrt_Cost {bool collision; float cost} (RTRRT.h:55-91), kept on the device as one 64-bit word
 * per node (collision | cost bits | parent id) so that a relaxation is one atomicMin.
 * set/get copy entries 0..n-1 (parent -1 = none). */
RTPP_API int rtpp_tree_set(rtpp_ctx *ctx, const int32_t *parent, const float *cost, const uint8_t *collision, size_t n);
RTPP_API int rtpp_tree_get(rtpp_ctx *ctx, int32_t *parent, float *cost, uint8_t *collision, size_t n);
RTPP_API int rtpp_tree_words_dev(rtpp_ctx *ctx, unsigned long long **d_words);
/* UpdateTree part (a), RTRRT.cpp:98-119: make new_root the root by reversing the parent chain
 * new_root -> ... -> old_root; cost[new_root] = {0,0}. */
RTPP_API int rtpp_tree_reroot(rtpp_ctx *ctx, int old_root, int new_root);
/* UpdateTree part (b), RTRRT.cpp:121-135 (RTFMT.cpp:152-174, AMRRT.cpp:185-199, CDRT.cpp:764-796):
 * BFS from root, cost[child] = cost[parent] + {CrashDetection(parent,child), |parent-child|} for
 * every tree edge.  flags (n bytes, may be NULL): bit0 = collision went 0->1 (RT-FMT N_b,
 * RTFMT.cpp:170-171), bit1 = cost grew, bit2 = cost shrank (CDRT.cpp:781-790).  *reached (may be
 * NULL) = nodes reached from root. */
RTPP_API int rtpp_tree_propagate(rtpp_ctx *ctx, int root, int divisor, int plane_mask, uint8_t *flags, int *reached);
RTPP_API int rtpp_tree_propagate_dev(rtpp_ctx *ctx, int root, int divisor, int plane_mask, uint8_t *d_flags, int *reached);
/* Rewire relaxations RTRRT.cpp:305-319,334-346 (AMRRT.cpp:497-510 with free_only = 1) over the
 * directed edges x -> knn[x][j] of a resident k-NN table (n*k ids, -1 = none), drained as a work list ordered by cost
 * bucket until no cost word changes (at most max_sweeps buckets; *sweeps_done = buckets drained).  The reference
 * relaxes one queue element at a time; both reach the same collision-free costs at quiescence; among
 * parents that give a node the same cost the smaller id is kept. */
RTPP_API int rtpp_relax_dev(rtpp_ctx *ctx, const int32_t *d_knn_idx, int k, int divisor, int plane_mask, int free_only,
                            int max_sweeps, int *sweeps_done);

/* ---------------------------------------------------------------- planner engines ----------- */
/* Sequential-exact replanning loops on the GPU: one warp runs one planner instance with the
 * reference's own queue order (lanes = the 25 neighbours of the popped node), thousands of
 * instances (map x start/goal x seed scenarios) run side by side.  planner_kind 1 = RT-RRT*:
 *   rtpp_engine_init_instances  = SetAgent on an empty planner + SetGoal   RTRRT.cpp:11-55
 *   rtpp_engine_set_inputs      = SetAgent/SetGoal on a live planner + the sample points the
 *                                 host-side sampler (RTRRT.cpp:234-262) drew for the next call
 *   rtpp_engine_tick            = UpdateTree RTRRT.cpp:92-136, ExpandRewiring(run_ms)
 *                                 RTRRT.cpp:226-355, GetPath RTRRT.cpp:356-400 (each optional)
 * The wall-clock gates of the reference become a virtual clock: the k-th clock read of an instance
 * returns k*dt_ns (the oracle uses the same rule, so runs are reproducible and comparable).
 * Each instance owns node_capacity nodes, a 2^cell_log2-pixel hash over domain_w x domain_h, its
 * queues and max_samples sample slots.  Up to 16 map slots hold the planes. */
typedef struct rtpp_engine rtpp_engine;
RTPP_API int rtpp_engine_create(rtpp_ctx *ctx, int planner_kind, int n_instances, int node_capacity, int cell_log2,
                                int domain_w, int domain_h, int max_samples, rtpp_engine **out);
RTPP_API int rtpp_engine_destroy(rtpp_engine *e);
/* plane 1 = static (Obsmap), 2 = dynamic (DObsmap; src NULL = all free) of map slot 0..15 */
RTPP_API int rtpp_engine_set_map(rtpp_engine *e, int slot, int plane, const uint8_t *rowmajor, int W, int H);
/* DrawDObsmap (RTRRT.cpp:138-169, AMRRT.cpp:202-233) of a map slot on the device, arguments as rtpp_map_draw_dynamic: the
 * slot's dynamic plane is cleared, the circles and polygons are drawn into its bit tiles, S|D is refreshed; stream-ordered. */
RTPP_API int rtpp_engine_draw_dynamic(rtpp_engine *e, int slot, const int32_t *circles, int n_circles, const int32_t *poly_xy,
                                      const int32_t *poly_sizes, int n_polys);
RTPP_API int rtpp_engine_init_instances(rtpp_engine *e, const int32_t *map_slot, const float *agent_xy, const float *goal_xy);
RTPP_API int rtpp_engine_set_inputs(rtpp_engine *e, const float *agent_xy, const float *goal_xy, const float *samples_xy,
                                    const int32_t *n_samples);
/* RT-RRT* (planner_kind 1): the rand() stream of the next tick(s), n_inst rows of max_samples raw rand() values (n_rand[i] valid);
 * the engine then draws its samples ON THE DEVICE — SamplePoint RTRRT.cpp:234-262, uniform or the informed ellipse of the last
 * GetPath's c_base, with glibc's atan2f / cosf / sinf bits (csrc/glibc_math.cuh) — instead of taking the injected points of
 * rtpp_engine_set_inputs.  rand == NULL switches back.  Stats slot 14 = draws consumed, 15 = the stream ran out. */
RTPP_API int rtpp_engine_set_rand(rtpp_engine *e, const int32_t *rand, const int32_t *n_rand);
RTPP_API int rtpp_engine_tick(rtpp_engine *e, int do_update_tree, long run_ms, long long dt_ns, int do_get_path);
/* planner_kind 2 = AM-RRT* (divisor 2): the tick runs UpdateTree AMRRT.cpp:156-200, then the MainTask loop
 * AMRRT.cpp:598-613 = { Expand :395-481 (sampler, Steer :354-390 with CrashDetectionStart :287-308 and
 * DiffusivityMetric :94-116, choose-parent), RewireRoot :483-519 } under the virtual clock (run_ms > 0
 * enables the loop; its budgets are the reference's 80 / 20 / 5 ms), and GetPath :310-351.  rand() is
 * consumed on the device from an injected stream (`max_samples` of rtpp_engine_create = draws per
 * instance and call).  RewireGoal (:521-585) runs, as in the reference, only while the goal point itself is a
 * collision-free tree node (a caller can put the goal onto a node; Expand never does). */
RTPP_API int rtpp_engine_set_diffusion(rtpp_engine *e, const int32_t *cells_ij, int ncells, const float *matrix, int dim,
                                       float ratio);
/* rand: n_inst rows of `max_samples` raw rand() values, n_rand[i] valid; ellipse: 6 floats per instance
 * {c_base/2, sqrtf(c_base^2-c_min^2)/2, cosf(angle), sinf(angle), (goal.x+root.x)/2, (goal.y+root.y)/2}
 * (AMRRT.cpp:418-432; host glibc); outpath_nonempty: OutPath.size() != 0.  NULL = unchanged. */
RTPP_API int rtpp_engine_set_am_inputs(rtpp_engine *e, const int32_t *rand, const int32_t *n_rand, const float *ellipse,
                                       const int32_t *outpath_nonempty);
/* 16 int64 per instance: n_nodes, root, samples_used, clock_reads, iterations, pops_r (AM-RRT*: RewireGoal pops), pops_s,
 * qr_size (AM-RRT*: |Sgoal| + |Qgoal|), qs_size, status (0 ok, 1 node store full, 2 queue overflow),
 * path_len, path_collision, path_cost (float bits), c_base (float bits), rand_pos, rand_exhausted */
RTPP_API int rtpp_engine_read_stats(rtpp_engine *e, long long *out16);
RTPP_API int rtpp_engine_read_tree(rtpp_engine *e, int inst, int max_nodes, float *xy, int32_t *parent, float *cost,
                                   uint8_t *collision, int *n_out);
/* path of the last GetPath (RTRRT.cpp:356-437), *len_out = length:
 *   path_collision = 0: node ids root..entry, then -2 for the goal point itself (cost = path_cost);
 *   path_collision = 1, path_len > 0: no collision-free entry to the goal exists — the reference's fallback (:401-437): the chain
 *     root..x of the node x nearest the goal among those within five collision-free levels of the root; no goal point in it,
 *     path_cost = 3.4e38 as the reference returns;
 *   path_len = 0 with path_cost = 2.4e38: a chain longer than 2000 nodes (the reference clears the path). */
RTPP_API int rtpp_engine_read_path(rtpp_engine *e, int inst, int max_len, int32_t *ids, int *len_out);

/* RT-FMT engine (one warp per planner instance, sequential-exact):
 *   rtpp_fmt_start      = SetAgent on an empty planner + Start     RTFMT.cpp:11-36, 393-422
 *                         (the free-space samples are drawn by the host-side sampler; node 0 = agent)
 *   rtpp_fmt_set_inputs = SetAgent / SetGoal on a live planner      RTFMT.cpp:32-35, 38-55
 *   rtpp_fmt_tick       = UpdateTree RTFMT.cpp:92-176 (with N_b capture), then `iters` rounds of the
 *                         phases in phase_mask (1 ExpandFMT :423-529, 2 RewireFromObstacles :531-595,
 *                         4 RewireFromRoot :665-703), GetPath :285-320 (each optional): MainTask :704-728
 *                         is tick(1, 500, 7, x).
 * nodes_xy: per instance node_capacity (x,y) slots, n_nodes[i] used. */
typedef struct rtpp_fmt_engine rtpp_fmt_engine;
RTPP_API int rtpp_fmt_create(rtpp_ctx *ctx, int n_instances, int node_capacity, int cell_log2, int domain_w, int domain_h,
                             rtpp_fmt_engine **out);
RTPP_API int rtpp_fmt_destroy(rtpp_fmt_engine *e);
RTPP_API int rtpp_fmt_set_map(rtpp_fmt_engine *e, int slot, int plane, const uint8_t *rowmajor, int W, int H);
RTPP_API int rtpp_fmt_start(rtpp_fmt_engine *e, const int32_t *map_slot, const int32_t *n_nodes, const float *nodes_xy,
                            const float *goal_xy);
RTPP_API int rtpp_fmt_set_inputs(rtpp_fmt_engine *e, const float *agent_xy, const float *goal_xy);
RTPP_API int rtpp_fmt_tick(rtpp_fmt_engine *e, int do_update_tree, int iters, int phase_mask, int do_get_path);
/* 16 int64 per instance: n_nodes, root, z, |V_unv|, |V_open|, |V_opennew|, |V_toopen|, |X_near|, |N_b|,
 * |Qo|, |Qr|, ExpandFMT calls, status, path_len, path_ok, 0 */
RTPP_API int rtpp_fmt_read_stats(rtpp_fmt_engine *e, long long *out16);
/* state: 0 unvisited, 1 open, 2 open-new, 3 closed, +4 when also in V_toopen */
RTPP_API int rtpp_fmt_read_tree(rtpp_fmt_engine *e, int inst, int max_nodes, int32_t *parent, float *cost, uint8_t *collision,
                                uint8_t *state, int *n_out);
/* RTFMT.cpp:285-391: path_ok = 1: ids root..entry then -2 (the goal); path_ok = 0 with path_len > 0: the depth-5 fallback chain
 * root..nearest reachable node, as for rtpp_engine_read_path */
RTPP_API int rtpp_fmt_read_path(rtpp_fmt_engine *e, int inst, int max_len, int32_t *ids, int *len_out);

/* CDRT-RRT* engine (one warp per BImap planner, sequential-exact; reference CDRT2/lib/CDRT/CDRT.cpp).
 * The convex decomposition (BIgraph: cut lines, free convex polygons and their links, CDRT.h:240-248,
 * built by StartCut CDRTmapcd.cpp:191-410) and the island id plane (DrawBImap CDRT.cpp:121-175:
 * island_plane[y*W+x] = BIimap where BIamap == the agent's region, else 0xFFFF) are inputs.  Link
 * lists are CSR (offsets + values, ascending like the reference's std::set): cp = cut line -> its
 * polygons, cc = cut line -> adjacent cut lines, pc = polygon -> its cut lines, pv = polygon vertices.
 *   rtpp_cdrt_create        BImap + the tunables r_th, p_th, nearK, RewireN (CDRT.h:319-322) and the
 *                           cut-line sampling Resolution of RTRT_start
 *   rtpp_cdrt_set_agent / set_goal    RTRT_SetAgent / RTRT_SetGoal   CDRT.cpp:384-423 (1 = rejected point)
 *   rtpp_cdrt_add_obstacle / del_obstacle / move_obstacle
 *                           RTRT_AddDynamicObstacle :425-500 (derives CheckPoints), RTRT_DelDynamicObstacle,
 *                           and the demos' in-place writes to DObstacles[num].x/.y/.angle
 *   rtpp_cdrt_start         RTRT_start :276-383 (cut-line samples, std::random_shuffle on the injected
 *                           rand() stream, initial tree over the cut-line cores)
 *   rtpp_cdrt_tick          RTRT_task :927-1068 for every started instance — RTRT_DrawDObsmap,
 *                           RTRT_CrashDObsIsland, the agent hand-over, RTRT_UpdateTreeCost, larger/smaller
 *                           sets, RewireN x RTRT_Sampling, the eQr rewiring loop — then GetPath :1076-1123.
 * Wall-clock gates read the virtual clock (k-th read = k*dt_ns).  status: 0 ok, 1 node store full,
 * 2 queue full, 3 island slice full, 4 the reference would have hit undefined behaviour (empty
 * neighbour list / broken path in the agent hand-over), 5 cut-line slice full, 6 set list full. */
typedef struct rtpp_cdrt_engine rtpp_cdrt_engine;
RTPP_API int rtpp_cdrt_create(rtpp_ctx *ctx, int n_instances, int node_capacity, int queue_capacity, int W, int H, int ncut,
                              const double *cut_se, const double *cut_core, const int32_t *cp_off, const int32_t *cp,
                              const int32_t *cc_off, const int32_t *cc, int npoly, const int32_t *pv_off, const double *pv,
                              const int32_t *pc_off, const int32_t *pc, const uint16_t *island_plane, double resolution,
                              double r_th, double p_th, int nearK, int RewireN, int rand_capacity, rtpp_cdrt_engine **out);
RTPP_API int rtpp_cdrt_destroy(rtpp_cdrt_engine *e);
RTPP_API int rtpp_cdrt_set_agent(rtpp_cdrt_engine *e, int inst, double x, double y);
RTPP_API int rtpp_cdrt_set_goal(rtpp_cdrt_engine *e, int inst, double x, double y);
RTPP_API int rtpp_cdrt_add_obstacle(rtpp_cdrt_engine *e, int inst, int num, int type, double x, double y, double angle,
                                    double radius, const double *outline_xy, int npts, double check_len);
RTPP_API int rtpp_cdrt_del_obstacle(rtpp_cdrt_engine *e, int inst, int num);
RTPP_API int rtpp_cdrt_move_obstacle(rtpp_cdrt_engine *e, int inst, int num, double x, double y, double angle);
RTPP_API int rtpp_cdrt_start(rtpp_cdrt_engine *e, int inst, const int32_t *rand, int n_rand);
/* a fresh rand() stream for a running instance (the next draw is rand[0]) */
RTPP_API int rtpp_cdrt_set_rand(rtpp_cdrt_engine *e, int inst, const int32_t *rand, int n_rand);
RTPP_API int rtpp_cdrt_tick(rtpp_cdrt_engine *e, long long dt_ns, int do_get_path);
/* 24 int64 per instance: node ids used, agent id, agent island, clock reads, rand_pos, rand_exhausted, |eQr|,
 * |largerSet| list, |smallerSet| list, cut-line samples left, islands with a kd-tree, status, pops, rewired,
 * sampled, pushes, edge checks, path length, path collision, path cost (binary64 bits), SM cycles of the last tick spent in
 * {environment + tree update, larger/smaller sets, sampling, eQr loop} */
RTPP_API int rtpp_cdrt_read_stats(rtpp_cdrt_engine *e, long long *out24);
/* profiling read-out of one instance's last tick, eQr loop (CDRT.cpp:1035-1067): SM cycles in {pop, entry filter, edge
 * check, neighbour list, pushes}, then the counts {edge checks, neighbours listed, rewires} */
RTPP_API int rtpp_cdrt_read_cycles(rtpp_cdrt_engine *e, int inst, long long *out8);
/* nodes by id, erased ones included (alive = 0); parent -1 = none; invM as CDRT.h:226-228 */
RTPP_API int rtpp_cdrt_read_tree(rtpp_cdrt_engine *e, int inst, int max_nodes, double *xy, int32_t *parent, double *cost,
                                 uint8_t *collision, int32_t *invM, uint8_t *alive, int *n_out);
/* agent_path of the last GetPath: root ... entry node, goal */
RTPP_API int rtpp_cdrt_read_path(rtpp_cdrt_engine *e, int inst, int max_len, double *xy, int *len_out);
/* obsKdtreeMap: per free polygon, whether it holds a kd-tree and how many points */
RTPP_API int rtpp_cdrt_read_islands(rtpp_cdrt_engine *e, int inst, int32_t *has_tree, int32_t *count);

/* --- glibc's binary32 transcendentals on the device (csrc/glibc_math.cuh) ---
 * The informed-ellipse samplers (RTRRT.cpp:241-262, AMRRT.cpp:415-432) call glibc's sinf / cosf / atan2f / sqrtf; a device sampler
 * must return the same bits.  rtpp_glibc_host_fma: 1 when the running host's libm uses its fused-multiply-add build of sinf / cosf
 * (the two builds differ on 12 of the 2^32 arguments); the engines follow the host they run next to.
 * rtpp_glibc_checksum: sum over the binary32 bit patterns b in [first, first + count) of canon(f(b)) * (2b + 1) mod 2^64 (NaNs
 * canonicalised), fn 0 sinf, 1 cosf, 2 atanf; rtpp_glibc_eval: out[i] = f(in[i]); fn 3 = atan2f(in[2i], in[2i+1]).  Test hooks. */
RTPP_API int rtpp_glibc_host_fma(void);
RTPP_API int rtpp_glibc_checksum(rtpp_ctx *ctx, int fn, int fma, uint32_t first_bits, unsigned long long count, unsigned long long *out);
RTPP_API int rtpp_glibc_eval(rtpp_ctx *ctx, int fn, int fma, const float *in, size_t n, float *out);

#ifdef __cplusplus
}
#endif
#endif /* RTPP_GPU_H */
