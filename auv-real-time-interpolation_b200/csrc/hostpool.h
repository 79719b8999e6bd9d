// hostpool.h -- the library's own host worker threads.
//
// The host halves of the C ABI move bytes between caller memory and pinned staging (packing 24-byte Points, copying
// result rows out of the bounce ring): memory-bound loops that need a handful of threads to reach the PCIe rate.  They
// used to be `omp parallel` regions, which obey OMP_NUM_THREADS -- and launchers such as torchrun export
// OMP_NUM_THREADS=1 to every rank, which made the Point-list path 4x slower at N >= 2.
// This pool is sized by AUVI_HOST_THREADS, else by the CPUs this process may run on divided by the ranks sharing
// the box (LOCAL_WORLD_SIZE) minus one, capped at 16.  Workers are created on first use; between jobs they spin on their
// own mailbox for 2-3 ms, then sleep on a condition variable.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include <pthread.h>
#include <sched.h>

namespace auvi {

class HostPool {
  public:
    static HostPool& get() {
        static HostPool* pool = new HostPool;      // never destroyed: workers may be asleep at process exit
        return *pool;
    }
    int size() const { return n_; }                // threads a job may use, the caller included

    // fn(t) for t in [0, nt); the caller runs t = 0.  One job at a time (callers queue on run_mu_).  Only the workers of
    // the job are signalled -- each through its own mailbox -- and only they are waited for, so a worker that is slow to
    // be scheduled delays the jobs it takes part in and no others.
    void run(int nt, const std::function<void(int)>& fn) {
        if (nt > n_) nt = n_;
        if (nt <= 1) { fn(0); return; }
        std::lock_guard<std::mutex> job_lock(run_mu_);
        {
            std::lock_guard<std::mutex> lk(mu_);
            spawn_locked();
        }
        job_ = &fn;
        pending_.store(nt - 1, std::memory_order_relaxed);
        const uint64_t g = ++gen_;
        for (int id = 1; id < nt; ++id) box_[id].go.store(g, std::memory_order_release);
        if (sleepers_.load(std::memory_order_acquire) > 0) {
            { std::lock_guard<std::mutex> lk(mu_); }   // a worker between its last check and its wait holds mu_
            cv_.notify_all();
        }
        fn(0);
        for (int spins = 0; pending_.load(std::memory_order_acquire) != 0; ++spins) {
            if (spins < kSpins) { cpu_relax(); continue; }
            std::unique_lock<std::mutex> lk(mu_);
            done_cv_.wait(lk, [&] { return pending_.load(std::memory_order_acquire) == 0; });
        }
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
    // A worker spins this many pause instructions (2-3 ms) for the next job before it sleeps on the condition
    // variable: the stages of one pipelined call follow each other within that time, and a futex wake-up costs
    // 50-100 us per worker -- more than the copy it is woken for (measured: 5 M points 4.9 -> 6.3 ms with sleeping workers).
    static constexpr int kSpins = 50000;
    static void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#else
        std::this_thread::yield();
#endif
    }
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
            if (n > 2) --n;                        // spinning workers on every core starve the thread that feeds them
            if (n > 16) n = 16;
        }
        n_ = n < 1 ? 1 : (n > 64 ? 64 : n);
        pthread_atfork(nullptr, nullptr, &HostPool::forget_workers_in_child);
    }
    // A forked child has this object but none of its threads: forget them, the next job spawns new ones.
    static void forget_workers_in_child() {
        HostPool& p = get();
        for (std::thread& t : p.workers_) new (&t) std::thread();      // the handles belong to threads of the parent
        p.workers_.clear();
        for (Mailbox& b : p.box_) b.go.store(0, std::memory_order_relaxed);
        p.pending_.store(0, std::memory_order_relaxed);
        p.sleepers_.store(0, std::memory_order_relaxed);
        p.gen_ = 0;
        p.job_ = nullptr;
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
            uint64_t g;
            for (int spins = 0; (g = box_[id].go.load(std::memory_order_acquire)) == seen; ++spins) {
                if (spins < kSpins) { cpu_relax(); continue; }
                std::unique_lock<std::mutex> lk(mu_);
                sleepers_.fetch_add(1, std::memory_order_acq_rel);
                cv_.wait(lk, [&] { return box_[id].go.load(std::memory_order_acquire) != seen; });
                sleepers_.fetch_sub(1, std::memory_order_acq_rel);
            }
            seen = g;
            (*job_)(id);                           // job_ was written before this worker's mailbox (release / acquire)
            if (pending_.fetch_sub(1, std::memory_order_acq_rel) == 1) {
                std::lock_guard<std::mutex> lk(mu_);
                done_cv_.notify_one();
            }
        }
    }

    struct alignas(64) Mailbox { std::atomic<uint64_t> go{0}; };
    int n_ = 1;
    std::vector<std::thread> workers_;
    std::mutex mu_, run_mu_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(int)>* job_ = nullptr;
    Mailbox box_[65];
    alignas(64) std::atomic<int> pending_{0};
    std::atomic<int> sleepers_{0};
    uint64_t gen_ = 0;                             // only touched under run_mu_
};

}  // namespace auvi
