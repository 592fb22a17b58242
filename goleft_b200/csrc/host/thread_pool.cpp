// thread_pool.cpp — see thread_pool.h
#include "thread_pool.h"
#include <sched.h>
#include <stdlib.h>
#include <stdio.h>
#include <chrono>
#if defined(__x86_64__)
#include <immintrin.h>
#define GL_CPU_PAUSE() _mm_pause()
#else
#define GL_CPU_PAUSE() ((void)0)
#endif

namespace glhost {

namespace {
const bool g_trace = getenv("GL_POOL_TRACE") != nullptr;
std::atomic<int64_t> g_trace_max_us{0};
constexpr int kLimitBits = 16;                         // gen word = (generation << 16) | participating threads
// how long an idle worker spins before it sleeps (GL_POOL_SPIN_US)
const auto kSpin = std::chrono::microseconds([] { const char* e = getenv("GL_POOL_SPIN_US"); const long v = e ? atol(e) : 200; return v < 0 ? 0L : v; }());
}  // namespace

int ThreadPool::default_threads() {
    if (const char* e = getenv("GL_THREADS")) {
        const int v = atoi(e);
        if (v > 0) return v > 1024 ? 1024 : v;
    }
    cpu_set_t set;                                     // the CPUs this process may run on (gl_bind_numa_for_device narrows it)
    if (sched_getaffinity(0, sizeof set, &set) == 0 && CPU_COUNT(&set) > 0) return CPU_COUNT(&set);
    const unsigned hc = std::thread::hardware_concurrency();
    return hc ? (int)hc : 1;
}

ThreadPool& ThreadPool::global() {
    static ThreadPool pool(default_threads());
    return pool;
}

ThreadPool::ThreadPool(int threads) {
    if (threads < 1) threads = 1;
    for (int i = 1; i < threads; i++) workers_.emplace_back([this, i] { worker_main(i); });
}

ThreadPool::~ThreadPool() {
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_ = true;
    }
    cv_[0].notify_all();
    cv_[1].notify_all();
    for (auto& t : workers_) t.join();
}

void ThreadPool::work(int id, uint32_t generation) {
    for (;;) {
        uint64_t cur = job_.load(std::memory_order_acquire);
        if ((uint32_t)(cur >> 32) != generation) return;                 // this job is over (a later one may be running: not ours)
        const int64_t i = (int64_t)(cur & 0xffffffffull);
        if (i >= n_tasks_) return;                                       // n_tasks_ / fn_ were published before job_ (release / acquire)
        if (!job_.compare_exchange_weak(cur, cur + 1, std::memory_order_acq_rel, std::memory_order_acquire)) continue;
        if (!g_trace) (*fn_)(i, id);
        else {                                                           // GL_POOL_TRACE: which task (if any) made a run slow
            const auto t0 = std::chrono::steady_clock::now();
            (*fn_)(i, id);
            const int64_t us = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
            int64_t prev = g_trace_max_us.load(std::memory_order_relaxed);
            while (us > prev && !g_trace_max_us.compare_exchange_weak(prev, us)) {}
            if (us > 2000) fprintf(stderr, "[pool] task %lld on worker %d (cpu %d) took %lld us\n", (long long)i, id, sched_getcpu(), (long long)us);
        }
        done_.fetch_add(1, std::memory_order_release);
    }
}

void ThreadPool::worker_main(int id) {
    uint64_t seen = 0;
    bool spin_ok = true;                                  // a worker that sat out the last job goes straight to sleep
    std::condition_variable& cv = cv_[id < kSmallGroup ? 0 : 1];
    for (;;) {
        uint64_t g = gen_.load(std::memory_order_acquire);
        if (g == seen) {
            const auto t0 = std::chrono::steady_clock::now();
            while ((g = gen_.load(std::memory_order_acquire)) == seen && !stop_.load(std::memory_order_relaxed)) {
                GL_CPU_PAUSE();
                if (!spin_ok || std::chrono::steady_clock::now() - t0 > kSpin) {
                    std::unique_lock<std::mutex> lk(mu_);
                    cv.wait(lk, [&] { return gen_.load(std::memory_order_acquire) != seen || stop_.load(); });
                    g = gen_.load(std::memory_order_acquire);
                    break;
                }
            }
        }
        if (stop_.load()) return;
        if (g == seen) continue;
        seen = g;
        const int limit = (int)(g & ((1u << kLimitBits) - 1));
        spin_ok = id < limit;
        if (id < limit) work(id, (uint32_t)(g >> kLimitBits));
    }
}

void ThreadPool::run(int64_t n_tasks, const std::function<void(int64_t, int)>& fn, int max_threads) {
    if (n_tasks <= 0) return;
    std::lock_guard<std::mutex> run_lk(run_mu_);
    int limit = size();
    if (max_threads > 0 && max_threads < limit) limit = max_threads;
    if ((int64_t)limit > n_tasks) limit = (int)n_tasks;
    if (limit <= 1 || n_tasks >= (int64_t(1) << 31)) {
        for (int64_t i = 0; i < n_tasks; i++) fn(i, 0);
        return;
    }
    const uint64_t generation = (gen_.load(std::memory_order_relaxed) >> kLimitBits) + 1;
    fn_ = &fn;
    n_tasks_ = n_tasks;
    done_.store(0, std::memory_order_relaxed);
    job_.store((uint64_t)(uint32_t)generation << 32, std::memory_order_release);
    {
        std::lock_guard<std::mutex> lk(mu_);
        gen_.store((generation << kLimitBits) | (uint64_t)limit, std::memory_order_release);
    }
    const auto tr0 = std::chrono::steady_clock::now();
    if (g_trace) g_trace_max_us.store(0);
    cv_[0].notify_all();
    if (limit > kSmallGroup) cv_[1].notify_all();         // workers >= limit that are asleep stay asleep
    const auto tr1 = std::chrono::steady_clock::now();
    work(0, (uint32_t)generation);
    const auto tr2 = std::chrono::steady_clock::now();
    while (done_.load(std::memory_order_acquire) != n_tasks) GL_CPU_PAUSE();
    fn_ = nullptr;
    if (g_trace) {
        const auto tr3 = std::chrono::steady_clock::now();
        auto us = [](std::chrono::steady_clock::duration d) { return (long long)std::chrono::duration_cast<std::chrono::microseconds>(d).count(); };
        if (us(tr3 - tr0) > 2000)
            fprintf(stderr, "[pool] run of %lld tasks on %d threads: notify %lld us, caller worked %lld us, then waited %lld us; longest task %lld us\n",
                    (long long)n_tasks, limit, us(tr1 - tr0), us(tr2 - tr1), us(tr3 - tr2), (long long)g_trace_max_us.load());
    }
}

}  // namespace glhost

// size of the library's host pool (creates it on first use: call before narrowing a thread's CPU affinity)
extern "C" int glhost_pool_size(void) { return glhost::ThreadPool::global().size(); }
