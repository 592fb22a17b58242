// thread_pool.cpp — see thread_pool.h
#include "thread_pool.h"
#include <sched.h>
#include <stdlib.h>
#include <chrono>
#if defined(__x86_64__)
#include <immintrin.h>
#define GL_CPU_PAUSE() _mm_pause()
#else
#define GL_CPU_PAUSE() ((void)0)
#endif

namespace glhost {

namespace {
constexpr int kLimitBits = 16;                         // gen word = (generation << 16) | participating threads
constexpr auto kSpin = std::chrono::microseconds(200); // how long an idle worker spins before it sleeps
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
    cv_.notify_all();
    for (auto& t : workers_) t.join();
}

void ThreadPool::work(int id, uint32_t generation) {
    for (;;) {
        uint64_t cur = job_.load(std::memory_order_acquire);
        if ((uint32_t)(cur >> 32) != generation) return;                 // this job is over (a later one may be running: not ours)
        const int64_t i = (int64_t)(cur & 0xffffffffull);
        if (i >= n_tasks_) return;                                       // n_tasks_ / fn_ were published before job_ (release / acquire)
        if (!job_.compare_exchange_weak(cur, cur + 1, std::memory_order_acq_rel, std::memory_order_acquire)) continue;
        (*fn_)(i, id);
        done_.fetch_add(1, std::memory_order_release);
    }
}

void ThreadPool::worker_main(int id) {
    uint64_t seen = 0;
    for (;;) {
        uint64_t g = gen_.load(std::memory_order_acquire);
        if (g == seen) {
            const auto t0 = std::chrono::steady_clock::now();
            while ((g = gen_.load(std::memory_order_acquire)) == seen && !stop_.load(std::memory_order_relaxed)) {
                GL_CPU_PAUSE();
                if (std::chrono::steady_clock::now() - t0 > kSpin) {
                    std::unique_lock<std::mutex> lk(mu_);
                    cv_.wait(lk, [&] { return gen_.load(std::memory_order_acquire) != seen || stop_.load(); });
                    g = gen_.load(std::memory_order_acquire);
                    break;
                }
            }
        }
        if (stop_.load()) return;
        if (g == seen) continue;
        seen = g;
        const int limit = (int)(g & ((1u << kLimitBits) - 1));
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
    cv_.notify_all();
    work(0, (uint32_t)generation);
    while (done_.load(std::memory_order_acquire) != n_tasks) GL_CPU_PAUSE();
    fn_ = nullptr;
}

}  // namespace glhost

// size of the library's host pool (creates it on first use: call before narrowing a thread's CPU affinity)
extern "C" int glhost_pool_size(void) { return glhost::ThreadPool::global().size(); }
