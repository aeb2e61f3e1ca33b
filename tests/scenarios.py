"""Seeded inputs shared by the CPU and GPU parity tests (no reference files are read at run time)."""
import numpy as np

SLICES = (("pos", slice(0, 3)), ("quat", slice(3, 7)), ("vel", slice(7, 10)), ("omega", slice(10, 13)), ("q", slice(13, 29)),
          ("qd", slice(29, 45)), ("tau", slice(45, 61)), ("fz", slice(61, 62)))


def rollout_states(oracle, seed=3, env_steps=4, extra_ticks=7, gait=1):
    """Physically plausible mid-motion states: `env_steps` random-action env-steps from the reset pose
    plus `extra_ticks` raw ticks towards fresh targets, rounded to fp32.  Returns (states[n,64] f64, targets[n,16] f32)."""
    n = oracle.n
    rng = np.random.default_rng(seed)
    oracle.reset()
    for _ in range(env_steps):
        oracle.step(rng.uniform(-1, 1, (n, oracle.act_dim)), threads=8)
    tg = np.zeros((n, 16))
    if gait == 1:
        tg[:, 1::2] = rng.uniform(-1, 1, (n, 8)) * np.pi / 6
    else:
        tg[:] = rng.uniform(-1, 1, (n, 16)) * np.pi / 6
    oracle.tick(tg, extra_ticks)
    s = oracle.get_state().astype(np.float32).astype(np.float64)
    return s, tg.astype(np.float32)


def err_table(a, b):
    """per-field (scale, p50, p99, max) of the per-env max-abs error between two [n,64] state arrays"""
    out = {}
    for name, sl in SLICES:
        er = np.abs(a[:, sl] - b[:, sl]).max(1)
        out[name] = (float(np.abs(a[:, sl]).max()), float(np.percentile(er, 50)), float(np.percentile(er, 99)), float(er.max()))
    return out


def serpenoid_actions(steps, n_env=1, dt=0.3, phase=None):
    """serpenoid wave on the 8 driven (odd) joints: a_k(t) = -sin(4 n_k + 2 t), n_k = 2k+1 (snake_gait_test.py:65-89)"""
    t = np.arange(steps)[:, None, None] * dt
    nn = (2 * np.arange(8) + 1)[None, None, :]
    ph = np.zeros((1, n_env, 1)) if phase is None else np.asarray(phase).reshape(1, n_env, 1)
    return -np.sin(4 * nn + 2 * t + ph)
