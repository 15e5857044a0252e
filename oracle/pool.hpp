// pool.hpp — persistent host worker pool for the CPU baselines (TEST INFRASTRUCTURE ONLY).
// The reference is single-threaded; its edge routine only reads the two planes, so bench.py's CPU arm fans a
// batch out over the host cores.  Threads are started once and parked on a condition variable (thread start-up is
// not inside a timed step); a batch is dealt in fixed-size chunks through an atomic cursor, so a run of long edges
// at the head of the batch does not leave one thread with most of the work.
#pragma once
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

struct OrcPool {
    std::vector<std::thread> th;
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    std::function<void(long long, long long)> job;
    std::atomic<long long> cursor{0};
    long long n = 0, chunk = 1;
    unsigned long long generation = 0;
    int working = 0;
    bool stop = false;

    explicit OrcPool(int nthreads) {
        for (int t = 0; t < nthreads; t++) th.emplace_back([this] { loop(); });
    }
    ~OrcPool() {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv_go.notify_all();
        for (auto &t : th) t.join();
    }
    void loop() {
        unsigned long long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_go.wait(lk, [&] { return stop || generation != seen; });
                if (stop) return;
                seen = generation;
            }
            for (;;) {
                const long long a = cursor.fetch_add(chunk);
                if (a >= n) break;
                job(a, a + chunk < n ? a + chunk : n);
            }
            {
                std::lock_guard<std::mutex> lk(mu);
                if (--working == 0) cv_done.notify_all();
            }
        }
    }
    // runs f(a, b) over [0, n) in chunks on all pool threads; returns when every chunk is done
    void run(long long n_, long long chunk_, std::function<void(long long, long long)> f) {
        if (n_ <= 0) return;
        std::unique_lock<std::mutex> lk(mu);
        job = std::move(f), n = n_, chunk = chunk_ > 0 ? chunk_ : 1;
        cursor.store(0);
        working = (int)th.size();
        generation++;
        cv_go.notify_all();
        cv_done.wait(lk, [&] { return working == 0; });
    }
};

// one pool per library, resized on demand (callers are single-threaded test / bench code)
static inline OrcPool &orc_pool(int nthreads) {
    static OrcPool *pool = nullptr;
    if (nthreads < 1) nthreads = 1;
    if (!pool || (int)pool->th.size() != nthreads) {
        delete pool;
        pool = new OrcPool(nthreads);
    }
    return *pool;
}
