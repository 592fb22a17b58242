// thread_pool.h — a small persistent worker pool for the host side of the feeder (segment packing, BGZF inflate,
// record parsing, row assembly).  Work is a counted loop: run(n, fn) calls fn(i, worker) for i in [0, n), tasks claimed
// in index order (so a task may wait for lower-numbered tasks without deadlock), the calling thread takes part.
// A job is complete when its TASKS are done, not when every worker has checked in: on a host whose every logical CPU runs a
// pool thread some worker is always descheduled for a moment, and waiting for it cost ~0.3 ms per run on the 2 x 32-core box.
// Workers spin briefly after a job before they sleep, so back-to-back calls (one per contig) do not pay a futex wake-up.
#pragma once
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace glhost {

class ThreadPool {
public:
    explicit ThreadPool(int threads);
    ~ThreadPool();
    int size() const { return (int)workers_.size() + 1; }                    // workers + the caller
    // fn(task, worker): worker in [0, size()).  Uses at most max_threads threads (0 = all).  Not re-entrant.
    void run(int64_t n_tasks, const std::function<void(int64_t, int)>& fn, int max_threads = 0);
    static ThreadPool& global();                                             // sized by GL_THREADS or hardware_concurrency
    static int default_threads();
    static constexpr int kSmallGroup = 16;

private:
    void worker_main(int id);
    void work(int id, uint32_t generation);
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_[2];          // workers 1..kSmallGroup-1 sleep on [0], the rest on [1]: a small job does not wake the whole pool
    std::atomic<uint64_t> gen_{0};
    std::atomic<uint64_t> job_{0};          // (generation << 32) | next task index: a worker that arrives late for a finished job cannot claim from the next one by accident
    std::atomic<int64_t> done_{0};          // tasks of the current job that have finished
    std::atomic<bool> stop_{false};
    int64_t n_tasks_ = 0;
    int limit_ = 0;
    const std::function<void(int64_t, int)>* fn_ = nullptr;
    std::mutex run_mu_;
};

}  // namespace glhost
