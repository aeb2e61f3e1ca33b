// snake_rowflags.h -- host side of the per-row ready words of snk_step_host_f64 (HandOut::flag_rows in snake_exact.cu).  Plain C++:
// also compiled by tests/test_rowflags_host.py with g++, where a thread plays the kernel.
//
// The producer (the env-step kernel, writing into mapped page-locked buffers) posts an environment's observation row, reward and
// done byte and THEN, behind a system-wide fence, its ticks word; the consumer set every ticks word to -1 before the launch.  One
// consumer thread per contiguous range of environments widens each row into the caller's float64 arrays as soon as its word is
// non-negative.  Exactly one of the threads is the `leader`: it watches the launch (query / wait callbacks wrapping cudaStreamQuery
// and cudaStreamSynchronize) so that a launch that fails ends the wait of all threads instead of hanging them.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <time.h>

#include <atomic>
#include <vector>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

// One observation row float32 -> float64.  The caller's array is written once and not read again by these threads: with SSE2 and a
// 16-byte aligned destination the doubles go out as streaming stores (no read-for-ownership of the destination lines -- the widening
// is bound by host memory traffic when eight ranks share a box).  `stream` = false or no SSE2: plain stores.
static inline void row_widen(const float* src, double* dst, int n, bool stream) {
#if defined(__SSE2__)
    if (stream && (((uintptr_t)dst & 15u) == 0) && (n % 4) == 0) {
        for (int c = 0; c < n; c += 4) {
            const __m128 v = _mm_loadu_ps(src + c);
            _mm_stream_pd(dst + c, _mm_cvtps_pd(v));
            _mm_stream_pd(dst + c + 2, _mm_cvtps_pd(_mm_movehl_ps(v, v)));
        }
        return;
    }
#endif
    for (int c = 0; c < n; c++) dst[c] = (double)src[c];
}
static inline void rows_widen_done() {
#if defined(__SSE2__)
    _mm_sfence(); // streaming stores become visible before the thread reports its range as done
#endif
}

enum { ROWS_IN_FLIGHT = 0, ROWS_ALL_POSTED = 1, ROWS_LAUNCH_FAILED = 2, ROWS_MISSING = 3 };

struct RowSink {
    const float* obs_src; const float* rew_src; const uint8_t* done_src; const int32_t* ticks_src; // the mapped buffers the producer writes
    double* obs; double* rew; uint8_t* done; int32_t* ticks;                                       // the caller's arrays (ticks may be null)
    int obs_dim;
    bool stream; // streaming stores for the observation rows (row_widen)
};

// Widen the rows [b, e).  `state` is shared by all consumer threads of the call (ROWS_IN_FLIGHT at the start).  query(): 0 while the
// launch runs, 1 once it has finished, < 0 when it failed; wait(): blocks until it has finished, 1 or < 0.  Only the leader calls them.
template <class Query, class Wait>
static inline void rows_widen_as_posted(const RowSink& s, size_t b, size_t e, std::atomic<int>& state, bool leader, bool prefault, Query query, Wait wait) {
    if (prefault) { // fresh numpy arrays are untouched anonymous memory: take the page faults now, not row by row
        volatile char* p0 = (volatile char*)(s.obs + b * s.obs_dim);
        for (size_t k = 0; k < (e - b) * s.obs_dim * sizeof(double); k += 4096) p0[k] = 0;
        volatile char* p1 = (volatile char*)(s.rew + b);
        for (size_t k = 0; k < (e - b) * sizeof(double); k += 4096) p1[k] = 0;
    }
    std::vector<uint32_t> pend(e - b);
    for (size_t k = 0; k < e - b; k++) pend[k] = (uint32_t)(b + k);
    size_t np = e - b;
    while (np) {
        const int before = state.load(std::memory_order_acquire);
        size_t w = 0;
        for (size_t k = 0; k < np; k++) {
            const size_t i = pend[k];
            const int32_t t = *(volatile const int32_t*)(s.ticks_src + i);
            if (t < 0) { pend[w++] = (uint32_t)i; continue; }
            std::atomic_thread_fence(std::memory_order_acquire); // the row was posted before its ticks word
            row_widen(s.obs_src + i * s.obs_dim, s.obs + i * s.obs_dim, s.obs_dim, s.stream);
            s.rew[i] = (double)s.rew_src[i];
            s.done[i] = s.done_src[i];
            if (s.ticks) s.ticks[i] = t;
        }
        const size_t converted = np - w;
        np = w;
        if (!np) break;
        if (before == ROWS_LAUNCH_FAILED || before == ROWS_MISSING) { rows_widen_done(); return; }
        if (before == ROWS_ALL_POSTED) { rows_widen_done(); state.store(ROWS_MISSING, std::memory_order_release); return; } // finished before this pass, and a word is still -1
        if (leader) {
            const int q = query();
            if (q > 0) state.store(ROWS_ALL_POSTED, std::memory_order_release);
            else if (q < 0) state.store(ROWS_LAUNCH_FAILED, std::memory_order_release);
        }
        if (converted == 0) { struct timespec ts = {0, 30000}; nanosleep(&ts, nullptr); } // nothing new: leave the memory bus alone for 30 us
    }
    rows_widen_done();
    if (leader && state.load(std::memory_order_acquire) == ROWS_IN_FLIGHT) // the other threads rely on the leader to notice a failed launch
        state.store(wait() > 0 ? ROWS_ALL_POSTED : ROWS_LAUNCH_FAILED, std::memory_order_release);
}
