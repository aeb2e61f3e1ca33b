// fma_probe.cu -- microbenchmark: issue rate of independent 3-register FFMA, FMUL and packed FFMA2 (fma.rn.f32x2)
// on one scheduler (1 warp) and with 2 warps per scheduler.  Development probe for the solver loops.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_probe fma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void probe(int iters, float seed, long long* cycles, float* sink) {
    float a[8], b[8], c[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = seed + i; b[i] = 1.0f + 1e-7f * (threadIdx.x + i); c[i] = 0.5f * i; }
    float2 p[4], q[4], r[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { p[i] = make_float2(a[2 * i], a[2 * i + 1]); q[i] = make_float2(b[2 * i], b[2 * i + 1]); r[i] = make_float2(c[2 * i], c[2 * i + 1]); }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {        // 8 independent 3-register FFMA chains
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = fmaf(a[i], b[i], c[i]);
        } else if (MODE == 1) { // 8 independent FMUL chains
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = a[i] * b[i];
        } else if (MODE == 2) { // 4 independent packed FFMA2 chains (= 8 FMAs)
#pragma unroll
            for (int i = 0; i < 4; i++) p[i] = __ffma2_rn(p[i], q[i], r[i]);
        } else {                // 8 FFMA with 4 distinct multiplicands shared pairwise (register reuse)
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = fmaf(a[i], b[i & 3], c[i & 1]);
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i];
#pragma unroll
    for (int i = 0; i < 4; i++) s += p[i].x + p[i].y;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
    long long* cyc; float* sink;
    cudaMalloc(&cyc, 148 * sizeof(long long)); cudaMalloc(&sink, 148 * 1024 * sizeof(float));
    const int iters = 20000;
    const char* names[4] = {"FFMA 3-reg x8", "FMUL x8", "FFMA2 x4 (8 FMAs)", "FFMA x8, shared operands"};
    for (int threads : {32, 128, 256}) { // 1 warp on one scheduler; 1 warp per scheduler; 2 warps per scheduler
        for (int mode = 0; mode < 4; mode++) {
            if (mode == 0) probe<0><<<148, threads>>>(iters, 1.f, cyc, sink);
            if (mode == 1) probe<1><<<148, threads>>>(iters, 1.f, cyc, sink);
            if (mode == 2) probe<2><<<148, threads>>>(iters, 1.f, cyc, sink);
            if (mode == 3) probe<3><<<148, threads>>>(iters, 1.f, cyc, sink);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            long long h; cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);
            printf("%3d threads/SM  %-26s : %.2f cycles per iteration (8 FMAs per thread)\n", threads, names[mode], (double)h / iters);
        }
    }
    return 0;
}
