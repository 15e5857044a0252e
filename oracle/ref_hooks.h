// ref_hooks.h — force-included (-include) in front of the reference's planner sources when they are
// compiled in place into oracle/_ref (TEST INFRASTRUCTURE ONLY).  Redirects the reference's two
// sources of non-determinism — wall-clock reads (utime_ns() -> clock_gettime, e.g. RTRRT.cpp:3-9)
// and un-seeded glibc rand() (e.g. RTRRT.cpp:234) — to harness-controlled functions.  The system
// headers are included first so that only the reference's own code sees the renamed calls.
#pragma once
#include <cstdlib>
#ifdef REF_HOOK_STD_RAND
// CDRT.cpp:330 shuffles the cut-line samples with std::random_shuffle, whose libstdc++ body calls
// std::rand(): define the redirect before <algorithm> is parsed so the shuffle too draws from the
// injected stream (first + rand() % (i + 1) for i = 1..n-1).
extern "C" int rtpp_ref_rand(void);
namespace std { using ::rtpp_ref_rand; }
#define rand rtpp_ref_rand
#include <algorithm>
#undef rand
#endif
#include <ctime>
#include <time.h>
#include <math.h>
#include <cmath>
#include <cfloat>
#include <algorithm>
#include <fstream>
#include <iostream>
#include <list>
#include <map>
#include <queue>
#include <random>
#include <set>
#include <thread>
#include <vector>

extern "C" int rtpp_ref_rand(void);
extern "C" int rtpp_ref_clock_gettime(clockid_t, struct timespec *);
#define rand rtpp_ref_rand
#define clock_gettime rtpp_ref_clock_gettime
