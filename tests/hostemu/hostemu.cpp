// hostemu.cpp -- runs the fp32 core of the thread-per-env CUDA kernel (snake_exact_core.cuh) on the
// CPU, one environment after the other.  DEVELOPMENT CHECK ONLY: it lets the kernel arithmetic be
// compared with the fp64 oracle in the GPU-less authoring container before a B200 run.  It is built
// into tests/hostemu/_build/ by tests/hostemu/build.sh, loaded only by tests/test_hostemu.py, and is
// neither the oracle nor part of the product library (which contains the device code only).
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef EMU_DOUBLE /* same code in fp64: checks the formulation against the oracle to round-off */
#define float double
#define sqrtf sqrt
#define sinf sin
#define cosf cos
#define fminf fmin
#define fmaxf fmax
#define fabsf fabs
#define fmaf fma
#endif
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { float2 r = {x, y}; return r; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 r = {x, y, z, w}; return r; }

#include "../../bullet_envs_b200/csrc/snake_host.h"

struct Emu {
    ExTables T;
    KParams P;
    int64_t n;
    float* state; // [n][64]
    ExSmem S;
    float tgt[NJ];
    int64_t counters[4];
};

static void set_targets_from_actions(Emu* h, const float* a, int tid) {
    const KParams& P = h->P;
    for (int j = 0; j < NJ; j++) h->tgt[j] = 0.f;
    for (int k = 0; k < P.actdim; k++) {
        float v = a[k];
        v = (v < -1.f) ? -1.f : v; v = (v > 1.f) ? 1.f : v;
        int j = (P.gait == 0) ? 2 * k : (P.gait == 1) ? 2 * k + 1 : k;
        h->tgt[j] = v * P.sf;
    }
}

extern "C" {

int emu_create(const snk_model* M, const snk_params* p, int64_t n, Emu** out) {
    Emu* h = (Emu*)calloc(1, sizeof(Emu));
    if (snk_to_extables(M, &h->T)) { free(h); return -1; }
    snk_to_kparams(p, &h->P);
    h->n = n;
    h->state = (float*)calloc((size_t)64 * n, sizeof(float));
    for (int64_t e = 0; e < n; e++) h->state[e * 64 + SNK_S_QUAT + 3] = 1.f;
    *out = h;
    return 0;
}
int emu_destroy(Emu* h) { free(h->state); free(h); return 0; }
int emu_set_state(Emu* h, const float* aos) {
    memcpy(h->state, aos, (size_t)h->n * 64 * sizeof(float));
    return 0;
}
int emu_get_state(Emu* h, float* aos) {
    memcpy(aos, h->state, (size_t)h->n * 64 * sizeof(float));
    return 0;
}
int emu_tick(Emu* h, const float* targets, int n_ticks, int32_t* iters_out, int32_t* contacts_out, float* height_out) {
    for (int64_t e = 0; e < h->n; e++) {
        ExEnv env; env.st = h->state + e * 64; env.tid = (int)(e & 31);
        for (int j = 0; j < NJ; j++) h->tgt[j] = targets[e * NJ + j];
        ex_load_base(env);
        ExTickOut to = {0, 0, 0.f, 0.f};
        for (int t = 0; t < n_ticks; t++) {
            bool ab;
            RowsS R; R.s = &h->S; R.lane = env.tid; R.tg = h->tgt;
            if (h->P.cone) ex_tick<true>(h->T, h->P, R, env, true, false, &ab, &to); else ex_tick<false>(h->T, h->P, R, env, true, false, &ab, &to);
            h->counters[0]++; h->counters[1] += to.iterations;
        }
        ex_store_base(env);
        if (iters_out) iters_out[e] = to.iterations;
        if (contacts_out) contacts_out[e] = to.contacts;
        if (height_out) height_out[e] = ex_height(h->T, env);
    }
    return 0;
}
int emu_step(Emu* h, const float* actions, float* obs, float* rew, uint8_t* done, int32_t* ticks) {
    for (int64_t e = 0; e < h->n; e++) {
        ExEnv env; env.st = h->state + e * 64; env.tid = (int)(e & 31);
        set_targets_from_actions(h, actions + e * h->P.actdim, env.tid);
        ex_load_base(env);
        ExStepOut o;
        RowsS R; R.s = &h->S; R.lane = env.tid; R.tg = h->tgt;
        if (h->P.cone) ex_env_step<true>(h->T, h->P, R, env, &o); else ex_env_step<false>(h->T, h->P, R, env, &o);
        for (int k = 0; k < SNK_OBS_DIM; k++) obs[e * SNK_OBS_DIM + k] = ex_obs_of(env, k);
        rew[e] = o.rew; done[e] = (uint8_t)o.done;
        if (ticks) ticks[e] = o.ticks;
        h->counters[0] += o.ticks; h->counters[1] += o.iters; h->counters[2] += o.done; h->counters[3] += o.bad;
    }
    return 0;
}
int emu_link_positions(Emu* h, float* out /*[n][51]*/) {
    for (int64_t e = 0; e < h->n; e++) {
        ExEnv env; env.st = h->state + e * 64; env.tid = 0;
        ex_load_base(env);
        ex_link_positions(h->T, env, out + e * 3 * NB);
    }
    return 0;
}
int emu_counters(Emu* h, int64_t out[4], int clear) {
    for (int k = 0; k < 4; k++) { out[k] = h->counters[k]; if (clear) h->counters[k] = 0; }
    return 0;
}
}
