"""Host side of the per-row ready words of snk_step_host_f64 (bullet_envs_b200/csrc/snake_rowflags.h) without a GPU: a g++ harness in
which a producer thread plays the env-step kernel -- it posts rows in a shuffled order, each row before its ticks word -- while the
pool's threads widen them; then a launch that fails half way (every consumer must return, state LAUNCH_FAILED) and a launch that
"finishes" with a word still at -1 (state MISSING instead of a hang)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HARNESS = r'''
#include "snake_hostpool.h"
#include "snake_rowflags.h"
#include <stdio.h>
#include <algorithm>
#include <random>
#include <thread>

static int run_case(size_t n, int mode, unsigned seed) { // mode 0: all rows posted; 1: the launch fails after half of them; 2: finishes with one row missing
    const int D = 56;
    std::vector<float> ho(n * D, -7.f), hr(n, -7.f); std::vector<uint8_t> hd(n, 9); std::vector<int32_t> ht(n, -1);
    std::vector<double> obs(n * D, -1.0), rew(n, -1.0); std::vector<uint8_t> done(n, 7); std::vector<int32_t> ticks(n, -5);
    std::vector<uint32_t> order(n);
    for (size_t i = 0; i < n; i++) order[i] = (uint32_t)i;
    std::mt19937 rng(seed);
    std::shuffle(order.begin(), order.end(), rng);
    std::atomic<int> producer{0}; // 0 running, 1 finished, -1 failed
    const size_t stop = mode == 1 ? n / 2 : mode == 2 ? n - 1 : n;
    std::thread gpu([&] {
        for (size_t k = 0; k < stop; k++) {
            const size_t i = order[k];
            for (int c = 0; c < D; c++) ho[i * D + c] = (float)(i % 1000) + 0.25f * c;
            hr[i] = -(float)(i % 333); hd[i] = (uint8_t)(i % 2);
            std::atomic_thread_fence(std::memory_order_release);
            *(volatile int32_t*)&ht[i] = (int32_t)(i % 42);
            if ((k & 1023) == 0) std::this_thread::yield();
        }
        producer.store(mode == 1 ? -1 : 1, std::memory_order_release);
    });
    RowSink s;
    s.obs_src = ho.data(); s.rew_src = hr.data(); s.done_src = hd.data(); s.ticks_src = ht.data();
    s.obs = obs.data(); s.rew = rew.data(); s.done = done.data(); s.ticks = ticks.data(); s.obs_dim = D; s.stream = (seed & 1) != 0;
    std::atomic<int> state{ROWS_IN_FLIGHT};
    HostPool::get().run(n, [&](size_t b, size_t e) {
        rows_widen_as_posted(s, b, e, state, b == 0, true,
            [&]() { return producer.load(std::memory_order_acquire); },
            [&]() { int p; while ((p = producer.load(std::memory_order_acquire)) == 0) std::this_thread::yield(); return p; });
    }, 4096);
    gpu.join();
    const int st = state.load();
    if (mode == 1) return st == ROWS_LAUNCH_FAILED ? 0 : 10 + st;
    if (mode == 2) return st == ROWS_MISSING ? 0 : 20 + st;
    if (st != ROWS_ALL_POSTED && st != ROWS_IN_FLIGHT) return 30 + st; // IN_FLIGHT: every thread had its rows before anyone looked at the launch
    for (size_t i = 0; i < n; i++) {
        for (int c = 0; c < D; c++) if (obs[i * D + c] != (double)((float)(i % 1000) + 0.25f * c)) return 1;
        if (rew[i] != -(double)(i % 333) || done[i] != (uint8_t)(i % 2) || ticks[i] != (int32_t)(i % 42)) return 2;
    }
    return 0;
}

int main() {
    const size_t sizes[] = {1, 97, 4095, 4096, 50001, 300000};
    for (int rep = 0; rep < 3; rep++)
        for (size_t n : sizes)
            for (int mode = 0; mode < 3; mode++) {
                if (mode == 2 && n < 2) continue;
                const int rc = run_case(n, mode, 17u * rep + (unsigned)n);
                if (rc) { printf("n=%zu mode=%d rc=%d\n", n, mode, rc); return 1; }
            }
    printf("ok\n");
    return 0;
}
'''


def test_rows_are_widened_as_they_are_posted_and_a_failed_launch_ends_the_wait(tmp_path):
    src = tmp_path / "harness.cpp"
    src.write_text(HARNESS)
    exe = tmp_path / "harness"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-I", os.path.join(ROOT, "bullet_envs_b200", "csrc"), "-o", str(exe), str(src)])
    out = subprocess.run([str(exe)], env=dict(os.environ, SNK_HOST_THREADS="4"), stdout=subprocess.PIPE, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stdout
