// snake_exact.cu -- fused SnakeGymEnv.step() for N environments, one environment per THREAD (sm_100a).
//
// This is the kernel the reference configuration runs on (motor force = inf, kd = 1, so the motor
// rows are equalities and only the 6 rigid degrees of freedom of the chain are unknown; see
// snake_exact_core.cuh for the algorithm and oracle/snake_oracle.c:tick_exact for its CPU twin).
//
// One launch = one SubprocVecEnv.step(): clip + createAction, the data-dependent 0..41-tick loop,
// observation, reward, termination, auto-reset.  The grid is persistent: 3 one-warp CTAs per SM (the
// shared-memory limit), each lane owning one environment at a time.  The lock-step unit of a warp is
// ONE PHYSICS TICK, not one env-step: a lane whose environment has finished its tick loop writes its
// outputs and takes the next environment from a global counter while the other lanes keep ticking, so
// the 0..41 spread of tick counts costs no idle lanes (only the last partial wave of the launch does).
//
// Data placement: contact rows in shared memory as [contact][thread] columns (conflict free), base
// state and loop progress in registers, joint state in the environment's 256 B record of the handle's
// [N][64] state array (L1 resident while the lane owns the environment), model tables in constant
// memory at warp-uniform addresses.
//
// Reference call sites replaced: SnakeGymEnv.py:33-50,82-103; snake.py:209-306,336-341;
// ppo/multiprocessing_env.py:11-16; snake_gait_test.py:96-104 (raw ticks).
#include <cuda_runtime.h>

#include "snake_exact_core.cuh"

__constant__ ExTables cT;

#define FULL 0xffffffffu

// joint targets of one environment: checkBound (SnakeGymEnv.py:82-88) + createAction (snake.py:247-269)
// + scaling (snake.py:223-225)
__device__ __forceinline__ void load_targets(const KParams& P, ExSmem& S, int tid, const float* __restrict__ act) {
#pragma unroll
    for (int j = 0; j < NJ; j++) S.tgt[j][tid] = 0.f;
#pragma unroll 1
    for (int k = 0; k < P.actdim; k++) {
        float a = act[k];
        a = (a < -1.f) ? -1.f : a; // checkBound's comparisons: a NaN passes through (SnakeGymEnv.py:84-87)
        a = (a > 1.f) ? 1.f : a;
        const int j = (P.gait == 0) ? 2 * k : (P.gait == 1) ? 2 * k + 1 : k;
        S.tgt[j][tid] = a * P.sf;
    }
}

// counters: [0] ticks, [1] PGS sweeps, [2] dones, [3] non-finite resets, [4] next environment to hand out
template <bool CONE>
__global__ void __launch_bounds__(EB, 3)
snk_exact_step_kernel(const KParams P, float* __restrict__ state, const float* __restrict__ actions, float* __restrict__ obs,
                      float* __restrict__ rew, uint8_t* __restrict__ done, int32_t* __restrict__ ticks, unsigned long long* __restrict__ counters,
                      int64_t n) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ExSmem& S = *reinterpret_cast<ExSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const unsigned lt_mask = (1u << tid) - 1u;
    ExEnv e;
    e.st = state;
    e.tid = tid;
    ExRun run;
    int64_t env = -1;
    bool have = false;
    unsigned long long c_ticks = 0, c_iters = 0;
    unsigned c_done = 0, c_bad = 0;
#pragma unroll 1
    for (;;) {
        // ---- lanes without an environment take the next ones from the global counter
        const unsigned need = __ballot_sync(FULL, !have);
        if (need) {
            unsigned long long base = 0;
            if (tid == 0) base = atomicAdd(&counters[4], (unsigned long long)__popc(need));
            base = __shfl_sync(FULL, base, 0);
            if (!have) {
                const int64_t cand = (int64_t)base + __popc(need & lt_mask);
                if (cand < n) {
                    env = cand; have = true;
                    e.st = state + env * SNK_STATE_STRIDE;
                    load_targets(P, S, tid, actions + env * P.actdim);
                    ex_load_base(e);
                    ex_step_begin(P, S, e, &run);
                }
            }
        }
        if (!__any_sync(FULL, have)) break;
        if (have) {
            if (ex_step_advance<CONE>(cT, P, S, e, &run)) {
                ExStepOut o;
                ex_step_end(cT, P, e, run, &o);
                rew[env] = o.rew;
                done[env] = (uint8_t)o.done;
                if (ticks) ticks[env] = o.ticks;
                float* go = obs + env * SNK_OBS_DIM; // 224 B row, 16 B aligned: 14 full-sector vector stores
#pragma unroll 1
                for (int k = 0; k < SNK_OBS_DIM; k += 4)
                    *reinterpret_cast<float4*>(go + k) = make_float4(ex_obs_of(e, k), ex_obs_of(e, k + 1), ex_obs_of(e, k + 2), ex_obs_of(e, k + 3));
                c_ticks += (unsigned long long)o.ticks; c_iters += (unsigned long long)o.iters; c_done += o.done; c_bad += o.bad;
                have = false;
            }
        }
    }
    if (counters) { // one atomic per warp and counter
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) {
            c_ticks += __shfl_xor_sync(FULL, c_ticks, sft); c_iters += __shfl_xor_sync(FULL, c_iters, sft);
            c_done += __shfl_xor_sync(FULL, c_done, sft); c_bad += __shfl_xor_sync(FULL, c_bad, sft);
        }
        if (tid == 0) {
            atomicAdd(&counters[0], c_ticks);
            atomicAdd(&counters[1], c_iters);
            if (c_done) atomicAdd(&counters[2], (unsigned long long)c_done);
            if (c_bad) atomicAdd(&counters[3], (unsigned long long)c_bad);
        }
    }
}

// n_ticks raw ticks with explicit targets[N,16] (gait script): every environment runs the same number of
// ticks, so the assignment is static (thread = environment)
template <bool CONE>
__global__ void __launch_bounds__(EB, 3)
snk_exact_tick_kernel(const KParams P, float* __restrict__ state, const float* __restrict__ targets, unsigned long long* __restrict__ counters,
                      int64_t n, int n_ticks) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ExSmem& S = *reinterpret_cast<ExSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const int64_t env = (int64_t)blockIdx.x * EB + tid;
    if (env >= n) return;
    ExEnv e;
    e.st = state + env * SNK_STATE_STRIDE;
    e.tid = tid;
#pragma unroll
    for (int j = 0; j < NJ; j += 4) {
        const float4 t = *reinterpret_cast<const float4*>(targets + env * NJ + j);
        S.tgt[j][tid] = t.x; S.tgt[j + 1][tid] = t.y; S.tgt[j + 2][tid] = t.z; S.tgt[j + 3][tid] = t.w;
    }
    ex_load_base(e);
    int iters = 0;
#pragma unroll 1
    for (int t = 0; t < n_ticks; t++) {
        bool ab;
        ExTickOut to;
        ex_tick<CONE>(cT, P, S, e, false, &ab, &to);
        iters += to.iterations;
    }
    ex_store_base(e);
    if (counters) {
        atomicAdd(&counters[0], (unsigned long long)n_ticks);
        atomicAdd(&counters[1], (unsigned long long)iters);
    }
}

// ---------------------------------------------------------------------------------------------
// launch wrappers used by the C-ABI host code (snake_abi.cu)
// ---------------------------------------------------------------------------------------------
size_t snk_exact_smem_bytes() { return sizeof(ExSmem); }

static int g_step_ctas = 0; // persistent grid of the step kernel: CTAs per SM (occupancy) x SMs

cudaError_t snk_exact_configure(const ExTables* host_tables) {
    cudaError_t e = cudaMemcpyToSymbol(cT, host_tables, sizeof(ExTables));
    const void* kernels[4] = {(const void*)snk_exact_step_kernel<true>, (const void*)snk_exact_step_kernel<false>,
                              (const void*)snk_exact_tick_kernel<true>, (const void*)snk_exact_tick_kernel<false>};
    for (int i = 0; i < 4 && e == cudaSuccess; i++)
        e = cudaFuncSetAttribute(kernels[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ExSmem));
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 0, per_sm = 0;
    e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, snk_exact_step_kernel<true>, EB, sizeof(ExSmem));
    if (e == cudaSuccess) g_step_ctas = sms * (per_sm > 0 ? per_sm : 1);
    return e;
}

cudaError_t snk_exact_launch_step(const KParams& P, float* state, const float* actions, float* obs, float* rew, uint8_t* done,
                                  int32_t* ticks, unsigned long long* counters, int64_t n, cudaStream_t st) {
    const int64_t warps = (n + EB - 1) / EB;
    dim3 grid((unsigned)(warps < g_step_ctas ? warps : g_step_ctas)), block(EB);
    if (P.cone) snk_exact_step_kernel<true><<<grid, block, sizeof(ExSmem), st>>>(P, state, actions, obs, rew, done, ticks, counters, n);
    else snk_exact_step_kernel<false><<<grid, block, sizeof(ExSmem), st>>>(P, state, actions, obs, rew, done, ticks, counters, n);
    return cudaGetLastError();
}

cudaError_t snk_exact_launch_tick(const KParams& P, float* state, const float* targets, unsigned long long* counters, int64_t n,
                                  int n_ticks, cudaStream_t st) {
    dim3 grid((unsigned)((n + EB - 1) / EB)), block(EB);
    if (P.cone) snk_exact_tick_kernel<true><<<grid, block, sizeof(ExSmem), st>>>(P, state, targets, counters, n, n_ticks);
    else snk_exact_tick_kernel<false><<<grid, block, sizeof(ExSmem), st>>>(P, state, targets, counters, n, n_ticks);
    return cudaGetLastError();
}
