// ref_common.hpp — shared by the oracle/_ref harnesses (TEST INFRASTRUCTURE ONLY).
// Virtual clock: the k-th clock read returns k*dt (same rule as orc::VClock).
// Injected rand(): values handed in by the test (same rule as orc::RandStream).
#pragma once
#include <cstdint>
#include <time.h>

static long long g_clk_n = 0, g_clk_dt = 1000000;
static const int32_t *g_rand = nullptr;
static long long g_rand_n = 0, g_rand_pos = 0;
static int g_rand_exhausted = 0;

extern "C" int rtpp_ref_clock_gettime(clockid_t, struct timespec *ts) {
    long long t = g_clk_dt * (g_clk_n++);
    ts->tv_sec = t / 1000000000LL;
    ts->tv_nsec = t % 1000000000LL;
    return 0;
}
extern "C" int rtpp_ref_rand(void) {
    if (g_rand_pos >= g_rand_n) {
        g_rand_exhausted = 1;
        return 0;
    }
    return g_rand[g_rand_pos++];
}
#define REF_EXPORT extern "C" __attribute__((visibility("default")))
REF_EXPORT void ref_set_clock(long long dt_ns, long long n0) { g_clk_dt = dt_ns, g_clk_n = n0; }
REF_EXPORT long long ref_clock_reads(void) { return g_clk_n; }
REF_EXPORT void ref_set_rand(const int32_t *r, long long n) { g_rand = r, g_rand_n = n, g_rand_pos = 0, g_rand_exhausted = 0; }
REF_EXPORT long long ref_rand_pos(void) { return g_rand_pos; }
REF_EXPORT int ref_rand_exhausted(void) { return g_rand_exhausted; }
