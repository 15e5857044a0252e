// edge_march.cuh — the warp-level body of K1 (shared by edges.cu and tick.cu): 32 lanes bring 32 edges and get the ballot
// word of their CrashDetection verdicts (reference RTRRT.cpp:202-222, AMRRT.cpp:266-286).
#pragma once
#include "common.cuh"

static constexpr int kShortMaxP = 16; // P <= 16 -> thread marches alone; P >= 32 -> warp marches

// The warp-level body of K1: every lane brings one edge (e = ps.x, ps.y, pe.x, pe.y; valid = false
// for padding lanes) and gets the ballot word of the 32 verdicts.
// defer_list != nullptr (multi-GPU batches): edges with >= kDeferP march steps are not marched here but appended to the list
// (their verdict bit stays 0) and marched one warp per edge by edge_deferred_kernel, which ORs their bits in afterwards —
// a block that meets a run of thousand-pixel edges otherwise becomes the kernel's critical path once the batch is split.
static constexpr int kDeferP = 128;
// kRound = 32-probe chunks issued together per round of the warp-wide march of a long edge
// kPre = probe positions of a long edge sampled by its own lane before the warp-wide march (0 = none)
template <int kRound = 1, int kPre = 16>
__device__ __forceinline__ unsigned warp_edges_verdict(const PlaneView &pl, const float4 e, const bool valid, const MarchThresholds &mt,
                                                       uint32_t *defer_list = nullptr, unsigned *defer_count = nullptr,
                                                       uint32_t edge_index = 0) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    uint32_t acc = 0; // bit 0 = blocked
    int P = 1;
    if (valid) {
        P = rt_march_P(mt, rt_sqlen(e.x, e.y, e.z, e.w));
        if (rt_edge_safe(pl, e.x, e.y, e.z, e.w)) {
            // The reference's probe set is the dyadic levels t = 1/2 | 1/4, 3/4 | odd/8 | odd/16 ... up to P
            // (RTRRT.cpp:210-218); the verdict is their OR, so the levels are issued as straight-line groups
            // with immediate t and 1-t (both exact): no loop counter, no t accumulation, and the first level
            // (every edge has it: M >= 2) goes out together with the two end-point probes.
            auto at = [&](const float t, const float omt) -> uint32_t {
                const int x = rt_trunc_pos(__fadd_rn(__fmul_rn(e.x, t), __fmul_rn(omt, e.z)));
                const int y = rt_trunc_pos(__fadd_rn(__fmul_rn(e.y, t), __fmul_rn(omt, e.w)));
                return rt_probe_word(pl, y, x);
            };
            acc = rt_probe_word(pl, rt_trunc_pos(e.y), rt_trunc_pos(e.x)) | rt_probe_word(pl, rt_trunc_pos(e.w), rt_trunc_pos(e.z)) |
                  at(0.5f, 0.5f);
            if (P >= 4 && P <= kShortMaxP) {
                acc |= at(0.25f, 0.75f) | at(0.75f, 0.25f);
                if (P >= 8) {
                    acc |= at(0.125f, 0.875f) | at(0.375f, 0.625f) | at(0.625f, 0.375f) | at(0.875f, 0.125f);
                    if (P >= 16)
                        acc |= at(0.0625f, 0.9375f) | at(0.1875f, 0.8125f) | at(0.3125f, 0.6875f) | at(0.4375f, 0.5625f) |
                               at(0.5625f, 0.4375f) | at(0.6875f, 0.3125f) | at(0.8125f, 0.1875f) | at(0.9375f, 0.0625f);
                }
            }
            acc &= 1u;
        } else {
            acc = rt_probe(pl, __float2int_rz(e.y), __float2int_rz(e.x)) | rt_probe(pl, __float2int_rz(e.w), __float2int_rz(e.z));
            if (!acc && P <= kShortMaxP) {
                const float inv = rt_inv_pow2(P);
                for (int j = 1; j < P; j++) acc |= rt_probe_interior(pl, e.x, e.y, e.z, e.w, __fmul_rn((float)j, inv));
            }
        }
    }
    // Long edges (P >= 32) are marched by the whole warp below, one edge after the other, each round a dependent L2 latency.
    // Most long edges through a cluttered map are blocked somewhere, so every lane first samples kPre of its OWN edge's probe
    // positions, spread evenly over the edge and issued together (one latency for all 32 edges): a hit settles the edge — the
    // verdict is the OR over the probe set, any subset may be looked at first — and only the survivors are marched in full.
    if (kPre) {
        const bool mine = valid && !acc && P > kShortMaxP;
        if (__any_sync(full, mine)) {
            if (mine) {
                const float inv = rt_inv_pow2(P);
                const int step = (P - 1) / kPre; // P >= 32 > kPre: step >= 1, positions 1 + i*step < P are distinct probe indices
                bool hit = false;
#pragma unroll
                for (int i = 0; i < kPre; i++) hit |= rt_probe_interior(pl, e.x, e.y, e.z, e.w, __fmul_rn((float)(1 + i * step), inv));
                acc = hit;
            }
        }
    }
    unsigned longmask = __ballot_sync(full, !acc && P > kShortMaxP);
    if (defer_list) {
        const unsigned dm = __ballot_sync(full, !acc && P >= kDeferP);
        if (dm) {
            unsigned at = 0;
            if (lane == 0) at = atomicAdd(defer_count, (unsigned)__popc(dm));
            at = __shfl_sync(full, at, 0);
            if (dm >> lane & 1u) defer_list[at + __popc(dm & ((1u << lane) - 1u))] = edge_index;
            longmask &= ~dm;
        }
    }
    while (longmask) {
        const int src = __ffs(longmask) - 1;
        longmask &= longmask - 1;
        const float sx = __shfl_sync(full, e.x, src), sy = __shfl_sync(full, e.y, src);
        const float tx = __shfl_sync(full, e.z, src), ty = __shfl_sync(full, e.w, src);
        const int Ps = __shfl_sync(full, P, src);
        const float inv = rt_inv_pow2(Ps);
        bool hit = false;
        // warp-uniform trip count, early exit per round.  A round is kRound 32-probe chunks issued together: the exit test makes
        // the rounds a chain of dependent L2 / L1 latencies, and a thousand-pixel edge (a hub's child) is tens of chunks long
        for (int j0 = 1; j0 < Ps; j0 += 32 * kRound) {
#pragma unroll
            for (int c = 0; c < kRound; c++) {
                const int j = j0 + 32 * c + lane;
                if (j < Ps) hit |= rt_probe_interior(pl, sx, sy, tx, ty, __fmul_rn((float)j, inv));
            }
            if (__any_sync(full, hit)) {
                hit = true;
                break;
            }
        }
        if (lane == src) acc = hit;
    }
    return __ballot_sync(full, acc != 0);
}

