// snake_gae.cu -- generalised advantage estimation over a device-resident rollout as ONE kernel (snk_gae).
//
// Reference: compute_gae (ppo/agent.py:14-22), called on the lists the rollout loop appends to (ppo/train.py:131-136,170-173):
//     delta_t = r_t + gamma V_{t+1} m_t - V_t ;  gae_t = delta_t + gamma tau m_t gae_{t+1} ;  return_t = gae_t + V_t ,  m_t = 1 - done_t
// The reference runs it as T Python iterations over [N,1] tensors; with the rollout stored as [T, N] arrays (RolloutBuffer) it
// is a backward scan per environment: one thread per environment walks t = T-1 .. 0, neighbouring threads read neighbouring
// addresses of every [t] row (coalesced), nothing is re-read.  HBM bound: 13 B per (t, env) element (r 4 + V 4 + done 1 in,
// return 4 out).  The grid is a multiple of the SM count for large N.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

// UN consecutive time steps are loaded before they are consumed so that a thread has several independent loads in flight
template <int UN>
__global__ void __launch_bounds__(256) snk_gae_kernel(const float* __restrict__ rewards, const uint8_t* __restrict__ dones, const float* __restrict__ values,
                                                      const float* __restrict__ next_value, float gamma, float tau, float* __restrict__ returns,
                                                      float* __restrict__ advantages, int T, int64_t n) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        float nxt = next_value[e], gae = 0.f;
        int t = T - 1;
        for (; t - UN + 1 >= 0; t -= UN) {
            float r[UN], v[UN], m[UN];
#pragma unroll
            for (int u = 0; u < UN; u++) {
                const int64_t i = (int64_t)(t - u) * n + e;
                r[u] = rewards[i]; v[u] = values[i]; m[u] = dones[i] ? 0.f : 1.f;
            }
#pragma unroll
            for (int u = 0; u < UN; u++) {
                const int64_t i = (int64_t)(t - u) * n + e;
                const float delta = r[u] + gamma * nxt * m[u] - v[u];
                gae = delta + gamma * tau * m[u] * gae;
                returns[i] = gae + v[u];
                if (advantages) advantages[i] = gae;
                nxt = v[u];
            }
        }
        for (; t >= 0; t--) {
            const int64_t i = (int64_t)t * n + e;
            const float r = rewards[i], v = values[i], m = dones[i] ? 0.f : 1.f;
            const float delta = r + gamma * nxt * m - v;
            gae = delta + gamma * tau * m * gae;
            returns[i] = gae + v;
            if (advantages) advantages[i] = gae;
            nxt = v;
        }
    }
}

cudaError_t snk_launch_gae(const float* rewards, const uint8_t* dones, const float* values, const float* next_value, float gamma, float tau,
                           float* returns, float* advantages, int T, int64_t n, cudaStream_t st) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)sms * 8; // 8 resident CTAs of 256 threads per SM
    if (blocks > cap) blocks = cap;
    // loads in flight per thread: 10 time steps when the rollout length allows (num_steps = 20, ppo/params.py:10), else 4
    const char* un = getenv("SNK_GAE_UN");
    const bool deep = un ? atoi(un) >= 10 : (T % 10 == 0);
    if (deep) snk_gae_kernel<10><<<(unsigned)blocks, 256, 0, st>>>(rewards, dones, values, next_value, gamma, tau, returns, advantages, T, n);
    else snk_gae_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(rewards, dones, values, next_value, gamma, tau, returns, advantages, T, n);
    return cudaGetLastError();
}
