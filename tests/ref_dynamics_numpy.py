"""Independent numpy formulation of the snake's rigid-body dynamics (test helper).

World-frame Newton-Euler written from first principles (COM-based, no spatial algebra, no
articulated-body recursion) so that it shares no structure with oracle/snake_oracle.c:

* :func:`body_kinematics`  -- world pose, twist of every merged body from (state)
* :func:`momentum`         -- total linear / angular momentum about the world origin
* :func:`inverse_dynamics` -- generalized forces needed for a given generalized acceleration
* :func:`forward_dynamics` -- qdd = M^-1 (tau - h) with M, h built column-wise from the above

Generalized velocity = [omega_world(3), v_world of base origin(3), qd(16)].
"""
import numpy as np

NB, NJ = 17, 16


def quat_to_mat(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def rodrigues(a, th):
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def body_kinematics(model, pos, quat, vel, omega, q, qd):
    """Returns per body: R (world), p (origin, world), w (world), v (origin velocity, world),
    joint axis (world), and COM position."""
    R = [quat_to_mat(quat)]; p = [np.asarray(pos, float)]
    w = [np.asarray(omega, float)]; v = [np.asarray(vel, float)]
    ax = [None]
    for i in range(1, NB):
        R0 = model.joint_R0[i - 1].reshape(3, 3)
        Ri = R[i - 1] @ R0 @ rodrigues(model.joint_axis[i - 1], q[i - 1])
        pi = p[i - 1] + R[i - 1] @ model.joint_t[i - 1]
        a = Ri @ model.joint_axis[i - 1]
        R.append(Ri); p.append(pi); ax.append(a)
        w.append(w[i - 1] + a * qd[i - 1])
        v.append(v[i - 1] + np.cross(w[i - 1], pi - p[i - 1]))
    com = [p[i] + R[i] @ model.body_com[i] for i in range(NB)]
    return R, p, w, v, ax, com


def momentum(model, pos, quat, vel, omega, q, qd):
    R, p, w, v, ax, com = body_kinematics(model, pos, quat, vel, omega, q, qd)
    P = np.zeros(3); L = np.zeros(3)
    for i in range(NB):
        m = model.body_mass[i]
        vc = v[i] + np.cross(w[i], com[i] - p[i])
        Iw = R[i] @ model.body_inertia[i].reshape(3, 3) @ R[i].T
        P += m * vc
        L += Iw @ w[i] + np.cross(com[i], m * vc)
    return P, L


def kinetic_energy(model, pos, quat, vel, omega, q, qd):
    R, p, w, v, ax, com = body_kinematics(model, pos, quat, vel, omega, q, qd)
    T = 0.0
    for i in range(NB):
        m = model.body_mass[i]
        vc = v[i] + np.cross(w[i], com[i] - p[i])
        Iw = R[i] @ model.body_inertia[i].reshape(3, 3) @ R[i].T
        T += 0.5 * m * vc @ vc + 0.5 * w[i] @ Iw @ w[i]
    return T


def inverse_dynamics(model, pos, quat, vel, omega, q, qd, acc, gravity, ext_wrench=None):
    """Generalized forces [moment about base origin(3), force(3), joint torques(16)] that produce
    generalized acceleration ``acc`` = [domega_world, d/dt v_world(base origin), qdd] under gravity.
    ``ext_wrench`` : optional list of (force_world, torque_world about COM) applied to each body."""
    R, p, w, v, ax, com = body_kinematics(model, pos, quat, vel, omega, q, qd)
    dw = [np.asarray(acc[0:3], float)]; dv = [np.asarray(acc[3:6], float)]
    for i in range(1, NB):
        r = p[i] - p[i - 1]
        # acceleration of body i's origin (rigidly attached to body i-1)
        dv.append(dv[i - 1] + np.cross(dw[i - 1], r) + np.cross(w[i - 1], np.cross(w[i - 1], r)))
        dw.append(dw[i - 1] + ax[i] * acc[6 + i - 1] + np.cross(w[i - 1], ax[i] * qd[i - 1]))
    F = []; N = []
    for i in range(NB):
        m = model.body_mass[i]
        rc = com[i] - p[i]
        ac = dv[i] + np.cross(dw[i], rc) + np.cross(w[i], np.cross(w[i], rc))
        Iw = R[i] @ model.body_inertia[i].reshape(3, 3) @ R[i].T
        f = m * ac - m * np.asarray(gravity)
        n = Iw @ dw[i] + np.cross(w[i], Iw @ w[i])
        if ext_wrench is not None:
            f = f - ext_wrench[i][0]; n = n - ext_wrench[i][1]
        F.append(f); N.append(n)
    tau = np.zeros(6 + NJ)
    # accumulate from the tip: wrench transmitted through joint i, moment taken about p[i]
    fj = np.zeros(3); nj = np.zeros(3)  # about p[i+1] of the child
    for i in range(NB - 1, -1, -1):
        rc = com[i] - p[i]
        f_tot = F[i] + fj
        n_tot = N[i] + np.cross(rc, F[i])
        if i + 1 < NB:
            n_tot = n_tot + nj + np.cross(p[i + 1] - p[i], fj)
        if i > 0:
            tau[6 + i - 1] = ax[i] @ n_tot
        else:
            tau[0:3] = n_tot; tau[3:6] = f_tot
        fj, nj = f_tot, n_tot
    return tau


def damping_wrench(model, pos, quat, vel, omega, q, qd, kl, ka):
    """btMultiBody velocity damping per merged body: F = -m v_com (k + k|v_com|), T = -I w (k + k|w|)."""
    R, p, w, v, ax, com = body_kinematics(model, pos, quat, vel, omega, q, qd)
    out = []
    for i in range(NB):
        m = model.body_mass[i]
        vc = v[i] + np.cross(w[i], com[i] - p[i])
        Iw = R[i] @ model.body_inertia[i].reshape(3, 3) @ R[i].T
        out.append((-m * vc * (kl + kl * np.linalg.norm(vc)), -(Iw @ w[i]) * (ka + ka * np.linalg.norm(w[i]))))
    return out


def forward_dynamics(model, pos, quat, vel, omega, q, qd, tau_joint, gravity, kl=0.0, ka=0.0):
    n = 6 + NJ
    ext = damping_wrench(model, pos, quat, vel, omega, q, qd, kl, ka) if (kl or ka) else None
    h = inverse_dynamics(model, pos, quat, vel, omega, q, qd, np.zeros(n), gravity, ext)
    M = np.zeros((n, n))
    zero_ext = [(np.zeros(3), np.zeros(3))] * NB
    h0 = inverse_dynamics(model, pos, quat, np.zeros(3), np.zeros(3), q, np.zeros(NJ), np.zeros(n), np.zeros(3), zero_ext)
    for k in range(n):
        e = np.zeros(n); e[k] = 1
        M[:, k] = inverse_dynamics(model, pos, quat, np.zeros(3), np.zeros(3), q, np.zeros(NJ), e, np.zeros(3), zero_ext) - h0
    rhs = np.concatenate([np.zeros(6), tau_joint]) - h
    return np.linalg.solve(M, rhs), M
