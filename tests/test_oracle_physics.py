"""Known-answer physics tests that pin the CPU oracle (it has no reference golden vectors for the
tick: PyBullet is absent -- SURVEY.md section 8c)."""
import numpy as np
import pytest

import ref_dynamics_numpy as rd
from bullet_envs_b200 import default_params
from oracle.oracle_py import Oracle


def _random_state(rng, z=5.0):
    s = np.zeros((1, 64))
    s[0, 0:3] = [0.3, -0.2, z]
    q = rng.normal(size=4); s[0, 3:7] = q / np.linalg.norm(q)
    s[0, 7:10] = rng.normal(size=3) * 0.5
    s[0, 10:13] = rng.normal(size=3)
    s[0, 13:29] = rng.uniform(-0.6, 0.6, 16)
    s[0, 29:45] = rng.normal(size=16) * 2.0
    return s


def test_free_dynamics_match_independent_newton_euler(model):
    """ABA passes 1-3 (incl. Coriolis, joint damping, velocity damping) == numpy world-frame
    Newton-Euler with M built column-wise and solved densely."""
    rng = np.random.default_rng(0)
    p = default_params(motor_max_force=0.0)        # motors cannot push: pure forward dynamics
    o = Oracle(1, p, model)
    for _ in range(3):
        s = _random_state(rng)
        o.set_state(s)
        o.tick(np.zeros((1, 16)), 1)
        s2 = o.get_state()[0]
        acc = np.concatenate([(s2[10:13] - s[0, 10:13]), (s2[7:10] - s[0, 7:10]), (s2[29:45] - s[0, 29:45])]) / p.dt
        ref, _ = rd.forward_dynamics(model, s[0, 0:3], s[0, 3:7], s[0, 7:10], s[0, 10:13], s[0, 13:29], s[0, 29:45],
                                     -0.1 * s[0, 29:45], [0, 0, -9.8], 0.04, 0.04)
        assert np.abs(acc - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())
        # semi-implicit Euler: positions advance with the NEW velocities
        assert np.allclose(s2[0:3], s[0, 0:3] + p.dt * s2[7:10], atol=1e-15)
        assert np.allclose(s2[13:29], s[0, 13:29] + p.dt * s2[29:45], atol=1e-15)
        assert abs(np.linalg.norm(s2[3:7]) - 1) < 1e-14


def test_impulse_response_is_inverse_mass_matrix(model):
    import ctypes
    from oracle.oracle_py import _load
    rng = np.random.default_rng(1)
    o = Oracle(1, default_params(), model)
    s = _random_state(rng); s[0, 7:13] = 0; s[0, 29:45] = 0
    o.set_state(s)
    _, Mw = rd.forward_dynamics(model, s[0, 0:3], s[0, 3:7], np.zeros(3), np.zeros(3), s[0, 13:29], np.zeros(16), np.zeros(16), [0, 0, 0])
    R = rd.quat_to_mat(s[0, 3:7])
    Q = np.eye(22); Q[0:3, 0:3] = R; Q[3:6, 3:6] = R
    Minv = np.linalg.inv(Q.T @ Mw @ Q)
    f = _load().snk_cpu_debug_response
    out = np.zeros(22)
    for j in (1, 7, 16):      # unit torque impulse at joint j
        f(o._h, ctypes.c_int64(0), ctypes.c_int32(-1), None, ctypes.c_int32(j), ctypes.c_void_p(out.ctypes.data))
        assert np.abs(out - Minv[:, 5 + j]).max() < 1e-9 * np.abs(out).max()
    # unit force along body-0 x at the body-0 origin -> column 3 of M^-1
    f6 = np.array([0, 0, 0, 1.0, 0, 0])
    f(o._h, ctypes.c_int64(0), ctypes.c_int32(0), ctypes.c_void_p(f6.ctypes.data), ctypes.c_int32(0), ctypes.c_void_p(out.ctypes.data))
    assert np.abs(out - Minv[:, 3]).max() < 1e-9 * np.abs(out).max()


def test_motor_impulses_conserve_momentum(model):
    rng = np.random.default_rng(2)
    p = default_params(gravity=[0, 0, 0], lin_damping=0.0, ang_damping=0.0)
    o = Oracle(1, p, model)
    s = np.zeros((1, 64)); s[0, 2] = 5; s[0, 6] = 1; s[0, 13:29] = rng.uniform(-0.3, 0.3, 16)
    o.set_state(s)
    tgt = rng.uniform(-0.5, 0.5, (1, 16))
    P0, L0 = rd.momentum(model, s[0, 0:3], s[0, 3:7], s[0, 7:10], s[0, 10:13], s[0, 13:29], s[0, 29:45])
    o.tick(tgt, 1)
    x = o.get_state()[0]
    assert np.abs(x[29:45]).max() > 5.0                     # the motors really moved
    P, L = rd.momentum(model, s[0, 0:3], s[0, 3:7], x[7:10], x[10:13], s[0, 13:29], x[29:45])   # at the old configuration
    assert np.abs(P - P0).max() < 1e-10 and np.abs(L - L0).max() < 1e-10


def test_free_fall(model):
    p = default_params(lin_damping=0.0, ang_damping=0.0)
    o = Oracle(1, p, model)
    s = np.zeros((1, 64)); s[0, 2] = 10.0; s[0, 6] = 1
    o.set_state(s)
    n = 24
    o.tick(np.zeros((1, 16)), n)
    x = o.get_state()[0]
    assert x[9] == pytest.approx(-9.8 * n * p.dt, rel=1e-12)
    assert x[2] == pytest.approx(10.0 - 9.8 * p.dt ** 2 * n * (n + 1) / 2, rel=1e-12)   # semi-implicit Euler sum
    assert np.abs(x[13:45]).max() < 1e-9 and np.abs(x[10:13]).max() < 1e-9


def test_motor_law_infinite_force(model):
    """q+ = q + kp (q* - q) when the motor is not force limited (PGS-converged limit)."""
    rng = np.random.default_rng(3)
    p = default_params(solver_iterations=3000, residual_threshold=1e-16)
    o = Oracle(1, p, model)
    o.reset()
    o.tick(np.zeros((1, 16)), 20)
    q0 = o.get_state()[0, 13:29].copy()
    tgt = rng.uniform(-0.5, 0.5, (1, 16))
    o.tick(tgt, 1)
    q1 = o.get_state()[0, 13:29]
    assert np.abs(q1 - (q0 + 0.1 * (tgt[0] - q0))).max() < 5e-5   # PGS converges slowly on the pitch joints
    # default 50 iterations: same law within the solver's truncation error
    o2 = Oracle(1, default_params(), model); o2.reset(); o2.tick(np.zeros((1, 16)), 20)
    q0 = o2.get_state()[0, 13:29].copy(); o2.tick(tgt, 1)
    assert np.abs(o2.get_state()[0, 13:29] - (q0 + 0.1 * (tgt[0] - q0))).max() < 2e-2   # 50 Bullet-order iterations are far from converged


def test_force_limited_motor_is_clamped(model):
    p = default_params(motor_max_force=0.05)
    o = Oracle(1, p, model); o.reset()
    o.tick(np.full((1, 16), 0.5), 5)
    tau = o.get_state()[0, 45:61]
    assert np.abs(tau).max() <= 0.05 + 1e-12 and np.abs(tau).max() > 0.049


def test_rest_stays_at_rest_and_settles_on_margin(model):
    o = Oracle(1, default_params(), model); o.reset()
    o.tick(np.zeros((1, 16)), 400)
    x = o.get_state()[0]
    # hull margin 0.001 lifts the chain by ~1 mm minus the slop; no drift, no spin
    assert 0.0008 < x[2] < 0.0011
    assert np.abs(x[0:2]).max() < 1e-3 and np.abs(x[7:13]).max() < 1e-3 and np.abs(x[13:29]).max() < 1e-3
    o.tick(np.zeros((1, 16)), 1)
    assert abs(o.get_state()[0, 9]) < 2e-3


def test_friction_anisotropy_changes_the_gait_drift(model):
    """Same serpenoid joint motion, friction vector reversed => different planar drift."""
    def run(aniso):
        o = Oracle(1, default_params(aniso=aniso), model); o.reset()
        o.tick(np.zeros((1, 16)), 30)
        for t in range(240):
            tg = np.zeros((1, 16))
            n = np.arange(1, 16, 2)
            tg[0, 1::2] = -(np.pi / 6) * np.sin(4 * n + 2 * t / 60.0)
            o.tick(tg, 1)
        return o.get_state()[0, 0:2]
    a, b = run([1, 0.1, 0.01]), run([0.01, 0.1, 1])
    assert np.linalg.norm(a - b) > 5e-3


def test_fp32_build_tracks_fp64_without_friction(model):
    rng = np.random.default_rng(5)
    p = default_params(friction=0.0)
    a, b = Oracle(8, p, model), Oracle(8, p, model, f32=True)
    a.reset(); b.reset()
    for _ in range(3):
        act = rng.uniform(-1, 1, (8, 8))
        oa, ra, da, ta = a.step(act); ob, rb, db, tb = b.step(act)
        assert (ta == tb).all() and (da == db).all()
        assert np.abs(oa[:, :16] - ob[:, :16]).max() < 1e-4 and np.abs(ra - rb).max() < 1e-4


def test_exact_motor_elimination_matches_dense_hybrid_dynamics(model):
    """tick_exact (motor rows imposed) in free flight: the base acceleration solves
    M_bb a_b + M_bj qdd = F_b with the prescribed joint accelerations; M, F from the dense numpy Newton-Euler."""
    rng = np.random.default_rng(21)
    p = default_params(motor_solver=1)
    o = Oracle(1, p, model)
    for _ in range(4):
        s = _random_state(rng)
        tg = rng.uniform(-0.5, 0.5, (1, 16))
        o.set_state(s); o.tick(tg, 1)
        g = o.get_state()[0]
        acc_free, Mw = rd.forward_dynamics(model, s[0, 0:3], s[0, 3:7], s[0, 7:10], s[0, 10:13], s[0, 13:29], s[0, 29:45],
                                           np.zeros(16), [0, 0, -9.8], 0.04, 0.04)
        F = Mw @ acc_free
        qd_new = 0.1 * (tg[0] - s[0, 13:29]) / p.dt
        qdd = (qd_new - s[0, 29:45]) / p.dt
        ab = np.linalg.solve(Mw[:6, :6], F[:6] - Mw[:6, 6:] @ qdd)
        got = np.concatenate([g[10:13] - s[0, 10:13], g[7:10] - s[0, 7:10]]) / p.dt
        assert np.abs(got - ab).max() < 1e-8 * max(1.0, np.abs(ab).max())
        assert np.allclose(g[29:45], qd_new, atol=1e-12)
