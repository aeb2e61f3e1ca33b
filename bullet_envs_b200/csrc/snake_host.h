// snake_host.h -- host-side conversion of the C-ABI structs (double) into the fp32 device tables.
// Shared by the C-ABI (snake_abi.cu) and by the CPU emulation harness of the kernel core
// (tests/hostemu, development check only).
#pragma once
#include <math.h>
#include <string.h>

#include "snake_exact_core.cuh"

static inline void snk_to_tables(const snk_model* M, DevTables* T) {
    memset(T, 0, sizeof *T);
    for (int i = 0; i < NJ; i++) {
        for (int k = 0; k < 9; k++) T->jR0[i][k] = (float)M->joint_R0[i][k];
        for (int k = 0; k < 3; k++) { T->jt[i][k] = (float)M->joint_t[i][k]; T->jax[i][k] = (float)M->joint_axis[i][k]; }
        T->jdamp[i] = (float)M->joint_damping[i];
    }
    for (int b = 0; b < NB; b++) {
        T->mass[b] = (float)M->body_mass[b];
        for (int k = 0; k < 3; k++) { T->com[b][k] = (float)M->body_com[b][k]; T->hpt[b][k] = (float)M->height_pt[b][k]; }
        for (int k = 0; k < 9; k++) T->Ic[b][k] = (float)M->body_inertia[b][k];
        T->hbody[b] = M->height_body[b];
    }
    for (int c = 0; c < NC; c++) {
        for (int k = 0; k < 3; k++) { T->ccen[c][k] = (float)M->cyl_center[c][k]; T->cax[c][k] = (float)M->cyl_axis[c][k]; }
        for (int k = 0; k < 9; k++) T->cfr[c][k] = (float)M->cyl_fric_R[c][k];
        T->crad[c] = (float)M->cyl_radius[c]; T->chl[c] = (float)M->cyl_halflen[c]; T->cend[c] = (float)M->cyl_end[c];
        T->cmar[c] = (float)M->cyl_margin[c]; T->cbrk[c] = (float)M->cyl_break[c];
        T->cbody[c] = M->cyl_body[c];
    }
    for (int k = 0; k < 3; k++) T->fzax[k] = (float)M->fz_axis[k];
    T->rootm = (float)M->root_mass;
}

// Tables of the thread-per-env kernel.  Returns 0, or -1 when the model does not have the layout
// that kernel assumes (cylinders grouped by body in chain order, height point h on body h).
static inline int snk_to_extables(const snk_model* M, ExTables* T) {
    memset(T, 0, sizeof *T);
    for (int i = 0; i < NJ; i++) {
        for (int k = 0; k < 9; k++) T->jR0[i][k] = (float)M->joint_R0[i][k];
        for (int k = 0; k < 3; k++) { T->jt[i][k] = (float)M->joint_t[i][k]; T->jax[i][k] = (float)M->joint_axis[i][k]; }
        T->jdamp[i] = (float)M->joint_damping[i];
    }
    double mtot = 0;
    for (int b = 0; b < NB; b++) {
        if (M->height_body[b] != b) return -1;
        T->mass[b] = (float)M->body_mass[b];
        mtot += M->body_mass[b];
        for (int k = 0; k < 3; k++) { T->com[b][k] = (float)M->body_com[b][k]; T->hpt[b][k] = (float)M->height_pt[b][k]; }
        const double* I = M->body_inertia[b];
        T->Ic[b][0] = (float)I[0]; T->Ic[b][1] = (float)(0.5 * (I[1] + I[3])); T->Ic[b][2] = (float)(0.5 * (I[2] + I[6]));
        T->Ic[b][3] = (float)I[4]; T->Ic[b][4] = (float)(0.5 * (I[5] + I[7])); T->Ic[b][5] = (float)I[8];
    }
    T->mtot = (float)mtot; T->inv_mtot = (float)(1.0 / mtot);
    int c = 0;
    for (int b = 0; b < NB; b++) {
        T->cstart[b] = c;
        while (c < NC && M->cyl_body[c] == b) c++;
    }
    T->cstart[NB] = c;
    if (c != NC) return -1;
    for (c = 0; c < NC; c++) {
        for (int k = 0; k < 3; k++) { T->ccen[c][k] = (float)M->cyl_center[c][k]; T->cax[c][k] = (float)M->cyl_axis[c][k]; }
        for (int k = 0; k < 9; k++) T->cfr[c][k] = (float)M->cyl_fric_R[c][k];
        T->crad[c] = (float)M->cyl_radius[c]; T->ceh[c] = (float)(M->cyl_end[c] * M->cyl_halflen[c]);
        T->cmar[c] = (float)M->cyl_margin[c]; T->cbrk[c] = (float)M->cyl_break[c];
    }
    for (int k = 0; k < 3; k++) T->fzax[k] = (float)M->fz_axis[k];
    T->rootm = (float)M->root_mass;
    return 0;
}

static inline void snk_to_kparams(const snk_params* p, KParams* P) {
    memset(P, 0, sizeof *P);
    P->dt = (float)p->dt; P->inv_dt = (float)(1.0 / p->dt);
    for (int k = 0; k < 3; k++) { P->g[k] = (float)p->gravity[k]; P->aniso[k] = (float)p->aniso[k]; }
    P->kp = (float)p->motor_kp; P->kd = (float)p->motor_kd;
    P->maximp = isinf(p->motor_max_force) ? INFINITY : (float)(p->motor_max_force * p->dt);
    P->sf = (float)p->scaling_factor; P->alpha = (float)p->alpha; P->beta = (float)p->beta; P->gamma = (float)p->gamma;
    P->edt = (float)p->energy_dt; P->mu = (float)p->friction; P->kl = (float)p->lin_damping; P->ka = (float)p->ang_damping;
    P->erp2 = (float)p->erp2; P->slop = (float)p->linear_slop; P->resthr = (float)p->residual_threshold; P->sthr = sqrtf(P->resthr);
    P->maxvel = (float)p->max_coord_vel; P->errthr = (float)p->err_threshold; P->hthr = (float)p->height_threshold;
    P->tang = (float)p->term_angle; P->donepen = (float)p->done_penalty; P->colf = (float)p->collision_force;
    P->colpen = (float)p->collision_penalty; P->iters = p->solver_iterations; P->maxticks = p->max_ticks;
    P->gait = p->gait_selection; P->cone = p->cone_friction; P->tjoint = p->term_joint; P->stale = p->stale_obs_on_reset;
    P->altmotor = p->alternate_motor_order;
    P->actdim = (p->gait_selection == 0 || p->gait_selection == 1) ? NJ / 2 : NJ; // SnakeGymEnv.py:72-76
    // motor_solver: 0 = Bullet-order PGS rows, 1 = exact elimination, 2 = auto (exact iff unlimited force and kd == 1)
    P->exact = (p->motor_solver == 1) || (p->motor_solver == 2 && isinf(p->motor_max_force) && p->motor_kd == 1.0);
}
