// hostpool.h -- the library's own host worker threads.
//
// The host halves of the C ABI move bytes between caller memory and pinned staging (packing 24-byte Points, copying
// result rows out of the bounce ring): memory-bound loops that need a handful of threads to reach the PCIe rate.  They
// used to be `omp parallel` regions, which obey OMP_NUM_THREADS -- and launchers such as torchrun export
// OMP_NUM_THREADS=1 to every rank, which made the Point-list path 4x slower at N >= 2.
// This pool is sized by AUVI_HOST_THREADS, else by the CPUs this process may run on divided by the ranks sharing
// the box (LOCAL_WORLD_SIZE), capped at 16.  Workers are created on first use and sleep on a condition variable.
#pragma once
#include <condition_variable>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include <sched.h>

namespace auvi {

class HostPool {
  public:
    static HostPool& get() {
        static HostPool* pool = new HostPool;      // never destroyed: workers may be asleep at process exit
        return *pool;
    }
    int size() const { return n_; }                // threads a job may use, the caller included

    // fn(t) for t in [0, nt); the caller runs t = 0.  One job at a time (callers queue on run_mu_).
    void run(int nt, const std::function<void(int)>& fn) {
        if (nt > n_) nt = n_;
        if (nt <= 1) { fn(0); return; }
        std::lock_guard<std::mutex> job_lock(run_mu_);
        {
            std::lock_guard<std::mutex> lk(mu_);
            spawn_locked();
            job_ = &fn; job_nt_ = nt; pending_ = nt - 1; ++gen_;
        }
        cv_.notify_all();
        fn(0);
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [&] { return pending_ == 0; });
        job_ = nullptr;
    }

    // [0,n) cut into `nt` contiguous pieces whose boundaries are multiples of `grain`.
    template <typename F>
    void for_range(int64_t n, int nt, int64_t grain, F&& body) {
        if (nt > n_) nt = n_;
        if (nt < 1) nt = 1;
        const int64_t piece = ((n + nt - 1) / nt + grain - 1) / grain * grain;
        run(nt, [&](int t) {
            const int64_t lo = t * piece, hi = lo + piece < n ? lo + piece : n;
            if (lo < hi) body(lo, hi);
        });
    }

  private:
    HostPool() {
        int n = 0;
        if (const char* e = getenv("AUVI_HOST_THREADS")) n = atoi(e);
        if (n < 1) {
            cpu_set_t set;
            int cpus = 0;
            if (sched_getaffinity(0, sizeof set, &set) == 0) cpus = CPU_COUNT(&set);
            if (cpus < 1) cpus = static_cast<int>(std::thread::hardware_concurrency());
            int ranks = 1;
            if (const char* e = getenv("LOCAL_WORLD_SIZE")) ranks = atoi(e);
            if (ranks < 1) ranks = 1;
            n = cpus / ranks;
            if (n > 16) n = 16;
        }
        n_ = n < 1 ? 1 : (n > 64 ? 64 : n);
    }
    void spawn_locked() {
        while (static_cast<int>(workers_.size()) < n_ - 1) {
            const int id = static_cast<int>(workers_.size()) + 1;
            workers_.emplace_back([this, id] { work(id); });
            workers_.back().detach();
        }
    }
    void work(int id) {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int)>* job = nullptr;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (id < job_nt_) job = job_;
            }
            if (job) {
                (*job)(id);
                std::lock_guard<std::mutex> lk(mu_);
                if (--pending_ == 0) done_cv_.notify_one();
            }
        }
    }

    int n_ = 1;
    std::vector<std::thread> workers_;
    std::mutex mu_, run_mu_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(int)>* job_ = nullptr;
    int job_nt_ = 0, pending_ = 0;
    uint64_t gen_ = 0;
};

}  // namespace auvi
