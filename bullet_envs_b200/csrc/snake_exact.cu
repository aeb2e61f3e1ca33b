// snake_exact.cu -- fused SnakeGymEnv.step() for N environments, one environment per THREAD (sm_100a).
//
// This is the kernel the reference configuration runs on (motor force = inf, kd = 1, so the motor
// rows are equalities and only the 6 rigid degrees of freedom of the chain are unknown; see
// snake_exact_core.cuh for the algorithm and oracle/snake_oracle.c:tick_exact for its CPU twin).
//
// One launch = one SubprocVecEnv.step(): clip + createAction, the data-dependent 0..41-tick loop,
// observation, reward, termination, auto-reset.  A CTA is one warp of 32 consecutive environments;
// its contact rows live in shared memory as [contact][thread] columns (conflict free), the base
// state in registers, the joint state in the handle's [slot][env] structure-of-arrays in global
// memory (every access of a warp is one fully used 128 B line, L1/L2 resident during the step).
// The 32 x 56 observation block of the CTA is contiguous in the caller's [N,56] buffer: it is
// transposed through shared memory and written with coalesced 128 B stores.
//
// Reference call sites replaced: SnakeGymEnv.py:33-50,82-103; snake.py:209-306,336-341;
// ppo/multiprocessing_env.py:11-16; snake_gait_test.py:96-104 (raw ticks).
#include <cuda_runtime.h>

#include "snake_exact_core.cuh"

__constant__ ExTables cT;

static_assert(sizeof(ExSmem) >= EB * (SNK_OBS_DIM + 1) * sizeof(float), "observation staging must fit in the row storage");

// RAW = false: one SubprocVecEnv.step.  RAW = true: n_ticks raw ticks with targets[N,16] (gait script).
template <bool RAW, bool CONE>
__global__ void __launch_bounds__(EB, 3)
snk_exact_kernel(const KParams P, float* __restrict__ state, int64_t npad, const float* __restrict__ in, float* __restrict__ obs,
                 float* __restrict__ rew, uint8_t* __restrict__ done, int32_t* __restrict__ ticks, unsigned long long* __restrict__ counters,
                 int64_t n, int n_ticks) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ExSmem& S = *reinterpret_cast<ExSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const int64_t env0 = (int64_t)blockIdx.x * EB;
    const int64_t env = env0 + tid;
    const bool live = env < n; // padded columns of the last CTA run on the (valid, reset) padding state
    ExEnv e;
    e.st = state + env;
    e.npad = npad;
    e.tid = tid;
    // ---- joint targets: checkBound (SnakeGymEnv.py:82-88) + createAction (snake.py:247-269) + scaling (snake.py:223-225)
    if (RAW) {
#pragma unroll
        for (int j = 0; j < NJ; j += 4) {
            float4 t = live ? *reinterpret_cast<const float4*>(in + env * NJ + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            S.tgt[j][tid] = t.x; S.tgt[j + 1][tid] = t.y; S.tgt[j + 2][tid] = t.z; S.tgt[j + 3][tid] = t.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < NJ; j++) S.tgt[j][tid] = 0.f;
        if (live) {
#pragma unroll 1
            for (int k = 0; k < P.actdim; k++) {
                float a = in[env * P.actdim + k];
                a = (a < -1.f) ? -1.f : a; // checkBound's comparisons: a NaN passes through (SnakeGymEnv.py:84-87)
                a = (a > 1.f) ? 1.f : a;
                const int j = (P.gait == 0) ? 2 * k : (P.gait == 1) ? 2 * k + 1 : k;
                S.tgt[j][tid] = a * P.sf;
            }
        }
    }
    ex_load_base(e);
    if (RAW) {
        int iters = 0;
        for (int t = 0; t < n_ticks; t++) {
            bool ab;
            ExTickOut to;
            ex_tick<CONE>(cT, P, S, e, false, &ab, &to);
            iters += to.iterations;
        }
        ex_store_base(e);
        if (counters && live) {
            atomicAdd(&counters[0], (unsigned long long)n_ticks);
            atomicAdd(&counters[1], (unsigned long long)iters);
        }
        return;
    }
    ExStepOut o;
    ex_env_step<CONE>(cT, P, S, e, &o);
    // ---- outputs: rew/done/ticks are one coalesced store per warp; obs goes through shared memory
    if (live) {
        rew[env] = o.rew;
        done[env] = (uint8_t)o.done;
        if (ticks) ticks[env] = o.ticks;
    }
    __syncwarp();
    float* stage = reinterpret_cast<float*>(smem_raw); // [EB][57]
#pragma unroll 1
    for (int k = 0; k < SNK_OBS_DIM; k++) stage[tid * (SNK_OBS_DIM + 1) + k] = ex_obs_of(e, k);
    __syncwarp();
    {
        const int64_t nrow = (n - env0 < EB) ? (n - env0) : EB;
        const int total = (int)nrow * SNK_OBS_DIM;
        float* go = obs + env0 * SNK_OBS_DIM;
#pragma unroll 1
        for (int idx = tid; idx < total; idx += EB) {
            const int r = idx / SNK_OBS_DIM, c = idx - r * SNK_OBS_DIM;
            go[idx] = stage[r * (SNK_OBS_DIM + 1) + c];
        }
    }
    if (counters) { // one atomic per warp and counter
        unsigned long long t = live ? (unsigned long long)o.ticks : 0ull, it = live ? (unsigned long long)o.iters : 0ull;
        unsigned dn = __ballot_sync(0xffffffffu, live && o.done), bd = __ballot_sync(0xffffffffu, live && o.bad);
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) { t += __shfl_xor_sync(0xffffffffu, t, s); it += __shfl_xor_sync(0xffffffffu, it, s); }
        if (tid == 0) {
            atomicAdd(&counters[0], t);
            atomicAdd(&counters[1], it);
            if (dn) atomicAdd(&counters[2], (unsigned long long)__popc(dn));
            if (bd) atomicAdd(&counters[3], (unsigned long long)__popc(bd));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// launch wrappers used by the C-ABI host code (snake_abi.cu)
// ---------------------------------------------------------------------------------------------
size_t snk_exact_smem_bytes() { return sizeof(ExSmem); }

cudaError_t snk_exact_configure(const ExTables* host_tables) {
    cudaError_t e = cudaMemcpyToSymbol(cT, host_tables, sizeof(ExTables));
    const void* kernels[4] = {(const void*)snk_exact_kernel<false, true>, (const void*)snk_exact_kernel<false, false>,
                              (const void*)snk_exact_kernel<true, true>, (const void*)snk_exact_kernel<true, false>};
    for (int i = 0; i < 4 && e == cudaSuccess; i++)
        e = cudaFuncSetAttribute(kernels[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ExSmem));
    return e;
}

cudaError_t snk_exact_launch_step(const KParams& P, float* state, int64_t npad, const float* actions, float* obs, float* rew, uint8_t* done,
                                  int32_t* ticks, unsigned long long* counters, int64_t n, cudaStream_t st) {
    dim3 grid((unsigned)(npad / EB)), block(EB);
    if (P.cone) snk_exact_kernel<false, true><<<grid, block, sizeof(ExSmem), st>>>(P, state, npad, actions, obs, rew, done, ticks, counters, n, 0);
    else snk_exact_kernel<false, false><<<grid, block, sizeof(ExSmem), st>>>(P, state, npad, actions, obs, rew, done, ticks, counters, n, 0);
    return cudaGetLastError();
}

cudaError_t snk_exact_launch_tick(const KParams& P, float* state, int64_t npad, const float* targets, unsigned long long* counters, int64_t n,
                                  int n_ticks, cudaStream_t st) {
    dim3 grid((unsigned)(npad / EB)), block(EB);
    if (P.cone) snk_exact_kernel<true, true><<<grid, block, sizeof(ExSmem), st>>>(P, state, npad, targets, nullptr, nullptr, nullptr, nullptr, counters, n, n_ticks);
    else snk_exact_kernel<true, false><<<grid, block, sizeof(ExSmem), st>>>(P, state, npad, targets, nullptr, nullptr, nullptr, nullptr, counters, n, n_ticks);
    return cudaGetLastError();
}
