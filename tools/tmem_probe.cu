// tmem_probe.cu -- microbenchmark: tensor memory (TMEM) as a per-lane scratchpad through tcgen05.ld/st
// (32x32b shape: thread i of warp w owns TMEM lane 32*(w%4)+i; columns are its private 32-bit words),
// against the same access pattern in shared memory.  Development probe for the env-step kernel's row
// storage; build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe tmem_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_st4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t addr, float& a, float& b, float& c, float& d) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "r"(addr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, float* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]), "=f"(v[10]),
                   "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
                 : "r"(addr) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t addr, float* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void tmem_st2(uint32_t addr, float a, float b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// mode 0: TMEM dependent loads (latency); 1: TMEM independent loads (throughput); 2/3: same in shared memory
__global__ void __launch_bounds__(128, 1) probe(int mode, int iters, long long* cycles, float* sink, int* ok) {
    __shared__ uint32_t tbase_s;
    extern __shared__ float4 sm[]; // [64 float4 rows][blockDim] columns
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tbase_s);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tbase = tbase_s;
    const uint32_t my = tbase + ((uint32_t)(32 * (warp & 3)) << 16); // lane field in bits 31:16
    // fill: column c of this thread holds a recognisable value
    for (int c = 0; c < 512; c += 4) tmem_st4(my + c, threadIdx.x * 1000.f + c, threadIdx.x * 1000.f + c + 1, threadIdx.x * 1000.f + c + 2, threadIdx.x * 1000.f + c + 3);
    tmem_wait_st();
    for (int c = 0; c < 64; c++) sm[c * blockDim.x + threadIdx.x] = make_float4(c, c + 1, c + 2, c + 3);
    __syncthreads();
    // correctness: read everything back
    int good = 1;
    for (int c = 0; c < 512; c += 4) {
        float a, b, d, e;
        tmem_ld4(my + c, a, b, d, e);
        tmem_wait_ld();
        good &= (a == threadIdx.x * 1000.f + c) && (b == threadIdx.x * 1000.f + c + 1) && (d == threadIdx.x * 1000.f + c + 2) && (e == threadIdx.x * 1000.f + c + 3);
    }
    // store-then-load of the same word by the same thread (the solver's impulse update)
    tmem_st4(my + 8, 1.f, 2.f, 3.f, 4.f);
    tmem_wait_st();
    {
        float a, b, d, e;
        tmem_ld4(my + 8, a, b, d, e);
        tmem_wait_ld();
        good &= (a == 1.f && e == 4.f);
    }
    atomicAnd(ok, good);
    __syncthreads();
    float acc = 0.f;
    int col = 0;
    long long t0 = clock64();
    if (mode == 0) {
        for (int i = 0; i < iters; i++) {
            float a, b, c, d;
            tmem_ld4(my + col, a, b, c, d);
            tmem_wait_ld();
            acc += a;
            col = ((int)a) & 0x1fc & 0; // dependent address (always 0 here, but unknown to the compiler)
            col += (i * 4) & 0x1fc;
        }
    } else if (mode == 1) {
        for (int i = 0; i < iters; i += 4) {
            float a[16];
#pragma unroll
            for (int u = 0; u < 4; u++) tmem_ld4(my + (((i + u) * 4) & 0x1fc), a[4 * u], a[4 * u + 1], a[4 * u + 2], a[4 * u + 3]);
            tmem_wait_ld();
#pragma unroll
            for (int u = 0; u < 16; u++) acc += a[u];
        }
    } else if (mode == 4) { // one x16 load per iteration, consumed by a short dependent FMA chain (prefetch distance 1)
        float v[16], w[16];
        tmem_ld16(my + 0, v);
        for (int i = 0; i < iters; i++) {
            tmem_wait_ld();
#pragma unroll
            for (int u = 0; u < 16; u++) w[u] = v[u];
            tmem_ld16(my + (((i + 1) * 16) & 0x1f0), v);
#pragma unroll
            for (int u = 0; u < 16; u++) acc = fmaf(acc, 0.999f, w[u]);
        }
        tmem_wait_ld();
    } else if (mode == 5) { // x16 load + x2 store + wait::st per iteration (the friction row of the solver)
        float v[16], w[16];
        tmem_ld16(my + 0, v);
        for (int i = 0; i < iters; i++) {
            tmem_wait_ld();
#pragma unroll
            for (int u = 0; u < 16; u++) w[u] = v[u];
            tmem_ld16(my + (((i + 1) * 16) & 0x1f0), v);
#pragma unroll
            for (int u = 0; u < 16; u++) acc = fmaf(acc, 0.999f, w[u]);
            tmem_st2(my + ((i * 16) & 0x1f0), acc, w[1]);
        }
        tmem_wait_ld(); tmem_wait_st();
    } else if (mode == 6) { // x8 loads back to back
        float v[8];
        for (int i = 0; i < iters; i++) {
            tmem_ld8(my + ((i * 8) & 0x1f8), v);
            tmem_wait_ld();
#pragma unroll
            for (int u = 0; u < 8; u++) acc += v[u];
        }
    } else if (mode == 7) { // only the FMA chain of mode 4 (no memory)
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int u = 0; u < 16; u++) acc = fmaf(acc, 0.999f, (float)u);
        }
    } else if (mode == 2) {
        for (int i = 0; i < iters; i++) {
            float4 v = sm[(col & 63) * blockDim.x + threadIdx.x];
            acc += v.x;
            col = ((int)v.x) & 0;
            col += i;
        }
    } else {
        for (int i = 0; i < iters; i += 4) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) v[u] = sm[((i + u) & 63) * blockDim.x + threadIdx.x];
#pragma unroll
            for (int u = 0; u < 4; u++) acc += v[u].x + v[u].y + v[u].z + v[u].w;
        }
    }
    long long t1 = clock64();
    if (lane == 0) cycles[blockIdx.x * 4 + warp] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
}

int main() {
    long long* cyc; float* sink; int* ok;
    cudaMalloc(&cyc, 148 * 4 * sizeof(long long)); cudaMalloc(&sink, 148 * 128 * sizeof(float)); cudaMalloc(&ok, sizeof(int));
    const int iters = 4096;
    size_t smem = 64 * 128 * sizeof(float4);
    if (cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { printf("smem attribute failed\n"); return 1; }
    const char* names[8] = {"TMEM dependent ld.x4", "TMEM independent ld.x4", "smem dependent LDS.128", "smem independent LDS.128", "TMEM ld.x16 prefetched+16 FMA", "TMEM ld.x16+st.x2+16 FMA", "TMEM ld.x8 + wait", "16-FMA chain only"};
    for (int threads = 32; threads <= 128; threads *= 2) {
        for (int mode = 0; mode < 8; mode++) {
            int one = 1; cudaMemcpy(ok, &one, sizeof one, cudaMemcpyHostToDevice);
            probe<<<148, threads, smem>>>(mode, iters, cyc, sink, ok);
            cudaError_t e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
            long long h[4]; int good;
            cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost); cudaMemcpy(&good, ok, sizeof good, cudaMemcpyDeviceToHost);
            printf("%d warps/SM  %-26s : %.1f cycles per iteration (warp 0), readback %s\n", threads / 32, names[mode], (double)h[0] / iters, good ? "ok" : "MISMATCH");
        }
    }
    return 0;
}
