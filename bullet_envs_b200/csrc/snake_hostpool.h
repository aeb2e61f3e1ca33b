// snake_hostpool.h -- persistent host thread pool of the float64 host entry point (snk_step_host_f64).  Plain C++: also compiled by
// tests/test_hostpool.py with g++ to exercise it without a GPU.
#pragma once
#include <stdlib.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

// Host-side worker threads of the float64 entry point: a persistent pool (created on first use, joined at process exit), sized
// SNK_HOST_THREADS or min(16, hardware threads / ranks on this node) -- under torchrun every rank of the node runs its own pool
// (LOCAL_WORLD_SIZE), so the pools together do not oversubscribe the host cores.
class HostPool {
public:
    static HostPool& get() { static HostPool p; return p; }
    int size() const { return (int)workers_.size() + 1; }
    // f(begin, end) over [0, n): the calling thread takes the first chunk (begin == 0), the workers the others -- every chunk on
    // its own thread at the same time (there are never more chunks than threads).  Fewer than `serial_below` items: the caller alone.
    void run(size_t n, const std::function<void(size_t, size_t)>& f, size_t serial_below = (size_t)1 << 16) {
        const int nt = (n < serial_below) ? 1 : size();
        if (nt == 1) { f((size_t)0, n); return; }
        std::lock_guard<std::mutex> one_job(run_mu_); // callers on several threads (one handle each) take turns
        const size_t per = (n + nt - 1) / nt;
        {
            std::unique_lock<std::mutex> lk(mu_);
            fn_ = &f; n_ = n; per_ = per; next_ = 1; pending_ = 0;
            for (int t = 1; t < nt; t++) if ((size_t)t * per < n) pending_++;
        }
        cv_.notify_all();
        f((size_t)0, per < n ? per : n);
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [&] { return pending_ == 0; });
        fn_ = nullptr;
    }

private:
    HostPool() {
        const char* e = getenv("SNK_HOST_THREADS");
        int hw = (int)std::thread::hardware_concurrency();
        const char* lw = getenv("LOCAL_WORLD_SIZE");
        const int ranks = (lw && atoi(lw) > 0) ? atoi(lw) : 1;
        int v = e ? atoi(e) : hw / ranks;
        if (!e && v > 16) v = 16;
        if (v < 1) v = 1;
        for (int t = 1; t < v; t++) workers_.emplace_back([this] { loop(); });
    }
    ~HostPool() {
        { std::unique_lock<std::mutex> lk(mu_); stop_ = true; }
        cv_.notify_all();
        for (auto& w : workers_) w.join();
    }
    void loop() {
        for (;;) {
            size_t b, e;
            const std::function<void(size_t, size_t)>* f;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || (fn_ && next_ * per_ < n_); });
                if (stop_) return;
                const size_t t = next_++;
                b = t * per_; e = b + per_ < n_ ? b + per_ : n_;
                f = fn_;
            }
            (*f)(b, e);
            {
                std::unique_lock<std::mutex> lk(mu_);
                if (--pending_ == 0) done_.notify_all();
            }
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_, run_mu_;
    std::condition_variable cv_, done_;
    const std::function<void(size_t, size_t)>* fn_ = nullptr;
    size_t n_ = 0, per_ = 1, next_ = 0;
    int pending_ = 0;
    bool stop_ = false;
};
