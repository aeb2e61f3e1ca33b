// snake_dev.cuh -- device helpers shared by the step kernels (register 3-vectors, spatial
// transforms, forward kinematics).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "snake_step.cuh"

#define FULL 0xffffffffu

// ---------------------------------------------------------------------------------------------
// small helpers (register vectors)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cross3(const float* a, const float* b, float* o) {
    float x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z;
}
__device__ __forceinline__ float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ void m3v(const float* M, const float* v, float* o) {
    float x = M[0] * v[0] + M[1] * v[1] + M[2] * v[2], y = M[3] * v[0] + M[4] * v[1] + M[5] * v[2],
          z = M[6] * v[0] + M[7] * v[1] + M[8] * v[2];
    o[0] = x; o[1] = y; o[2] = z;
}
__device__ __forceinline__ void m3tv(const float* M, const float* v, float* o) {
    float x = M[0] * v[0] + M[3] * v[1] + M[6] * v[2], y = M[1] * v[0] + M[4] * v[1] + M[7] * v[2],
          z = M[2] * v[0] + M[5] * v[1] + M[8] * v[2];
    o[0] = x; o[1] = y; o[2] = z;
}
__device__ __forceinline__ void m3m3(const float* A, const float* B, float* o) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) o[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
    return x;
}
// motion vector parent -> child through a joint with child->parent rotation R and offset r
__device__ __forceinline__ void xmot(const float* R, const float* r, const float* vp, float* vc) {
    float t[3], u[3];
    cross3(vp, r, t);
    u[0] = vp[3] + t[0]; u[1] = vp[4] + t[1]; u[2] = vp[5] + t[2];
    m3tv(R, vp, vc);
    m3tv(R, u, vc + 3);
}
// force vector child -> parent
__device__ __forceinline__ void xfrc(const float* R, const float* r, const float* fc, float* fp) {
    float t[3];
    m3v(R, fc, fp);
    m3v(R, fc + 3, fp + 3);
    cross3(r, fp + 3, t);
    fp[0] += t[0]; fp[1] += t[1]; fp[2] += t[2];
}

// ---------------------------------------------------------------------------------------------
// forward kinematics; returns the checkSnakeHeight mean (snake.py:237-245), identical in all lanes
// ---------------------------------------------------------------------------------------------
template <class WM>
__device__ __forceinline__ float fk(WM& W, const DevTables* __restrict__ T, int lane) {
    if (lane < NJ) { // joint rotations, lane = joint
        float a[3] = {__ldg(&T->jax[lane][0]), __ldg(&T->jax[lane][1]), __ldg(&T->jax[lane][2])};
        float th = W.s[SNK_S_Q + lane], s, c;
        sincosf(th, &s, &c);
        float C = 1.f - c, Rq[9], R0[9], R[9];
        Rq[0] = c + a[0] * a[0] * C;        Rq[1] = a[0] * a[1] * C - a[2] * s; Rq[2] = a[0] * a[2] * C + a[1] * s;
        Rq[3] = a[1] * a[0] * C + a[2] * s; Rq[4] = c + a[1] * a[1] * C;        Rq[5] = a[1] * a[2] * C - a[0] * s;
        Rq[6] = a[2] * a[0] * C - a[1] * s; Rq[7] = a[2] * a[1] * C + a[0] * s; Rq[8] = c + a[2] * a[2] * C;
#pragma unroll
        for (int k = 0; k < 9; k++) R0[k] = __ldg(&T->jR0[lane][k]);
        m3m3(R0, Rq, R);
#pragma unroll
        for (int k = 0; k < 9; k++) W.Rj[lane + 1][k] = R[k];
    }
    __syncwarp();
    float R[9], p[3], Rm[9], pm[3];
    {
        float x = W.s[SNK_S_QUAT], y = W.s[SNK_S_QUAT + 1], z = W.s[SNK_S_QUAT + 2], w = W.s[SNK_S_QUAT + 3];
        R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w);     R[2] = 2 * (x * z + y * w);
        R[3] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
        R[6] = 2 * (x * z - y * w);     R[7] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
        p[0] = W.s[SNK_S_POS]; p[1] = W.s[SNK_S_POS + 1]; p[2] = W.s[SNK_S_POS + 2];
    }
#pragma unroll
    for (int k = 0; k < 9; k++) Rm[k] = R[k];
    pm[0] = p[0]; pm[1] = p[1]; pm[2] = p[2];
#pragma unroll 1
    for (int i = 1; i < NB; i++) { // every lane walks the chain; lane i keeps body i
        float Rn[9], t[3], r[3] = {__ldg(&T->jt[i - 1][0]), __ldg(&T->jt[i - 1][1]), __ldg(&T->jt[i - 1][2])};
        m3v(R, r, t);
        m3m3(R, W.Rj[i], Rn);
        p[0] += t[0]; p[1] += t[1]; p[2] += t[2];
#pragma unroll
        for (int k = 0; k < 9; k++) R[k] = Rn[k];
        if (lane == i) {
#pragma unroll
            for (int k = 0; k < 9; k++) Rm[k] = R[k];
            pm[0] = p[0]; pm[1] = p[1]; pm[2] = p[2];
        }
    }
    if (lane < NB) {
#pragma unroll
        for (int k = 0; k < 9; k++) W.Rw[lane][k] = Rm[k];
        W.pw[lane][0] = pm[0]; W.pw[lane][1] = pm[1]; W.pw[lane][2] = pm[2];
    }
    __syncwarp();
    float z = 0.f;
#pragma unroll 1
    for (int h = 0; h < NB; h++) {
        int b = __ldg(&T->hbody[h]);
        z += W.pw[b][2] + W.Rw[b][6] * __ldg(&T->hpt[h][0]) + W.Rw[b][7] * __ldg(&T->hpt[h][1]) + W.Rw[b][8] * __ldg(&T->hpt[h][2]);
    }
    return z / NB;
}

