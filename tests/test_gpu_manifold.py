"""Persistent contact manifolds + warm starting on the GPU (csrc/snake_manifold.cuh) against the oracle's
manifold mode (oracle/snake_oracle.c: manifold_update, tick_exact with `manifold`), through the C ABI (snk_set_manifold).

The contact caches are part of the state and have no set_state, so both sides start from the reset pose with empty caches and run
freely.  The fp32 build of the oracle itself separates from the fp64 one at this rate (measured, 128 environments): tick counts
and episode ends identical, joints 1e-7, base position median 1e-5 after the first env-step, 3e-3 after the third, 1.4e-2 after
the eighth, per-step batch-mean rewards within 1e-3, cached points per tick 32.98 vs 32.99 -- the bounds below are those with
head room."""
import numpy as np
import pytest

from bullet_envs_b200 import SnakeVecEnv
from bullet_envs_b200._abi import default_params
from oracle.oracle_py import Oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch as t
    if not t.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return t


@pytest.mark.parametrize("warm", [0.0, 0.1])
def test_manifold_steps_vs_oracle(torch, warm):
    n, steps = 128, 8
    g = torch.Generator().manual_seed(3)
    acts = (torch.rand((steps, n, 8), generator=g) * 2 - 1).numpy()
    p = default_params()
    env = SnakeVecEnv(num_envs=n, device=0, params=p)
    env.set_manifold(True, warm)
    o = Oracle(n, p)
    o.set_manifold(True, warm)
    env.reset(as_torch=True); o.reset()
    pts = tks = 0
    for t in range(steps):
        obs, rew, done, _ = env.step(torch.from_numpy(acts[t]).cuda())
        tk = env.last_ticks.cpu().numpy()
        a, b = env.manifold_stats(); pts += a; tks += b
        oo, orr, od, ot = o.step(acts[t].astype(np.float64), threads=8)
        og = obs.cpu().numpy().astype(np.float64); rg = rew.cpu().numpy().astype(np.float64); dg = done.cpu().numpy()
        assert (tk == ot).mean() >= 0.99, (t, (tk == ot).mean())
        assert (dg == od).mean() >= 0.98
        same = (tk == ot) & (dg == od)
        e = np.abs(og - oo)[same]
        assert e[:, :16].max() < 1e-5, (t, e[:, :16].max())                      # joints follow the motor law
        pos = np.median(e[:, 48:51].max(1))
        assert pos < (2e-4 if t == 0 else 1e-2 if t < 3 else 5e-2), (t, pos)      # base position: contact sensitive, free running
        assert np.median(np.abs(rg - orr)[same]) < (5e-4 if t == 0 else 1e-2), (t, np.median(np.abs(rg - orr)[same]))
        quiet = same & (rg > -4) & (orr > -4)   # without the -10 (|Fz| > 10) and -5 (episode end) events: one flip moves a 128-env mean by 0.08
        assert abs(rg[quiet].mean() - orr[quiet].mean()) < 5e-3, (t, rg[quiet].mean(), orr[quiet].mean())
        assert abs(int((rg < -4).sum()) - int((orr < -4).sum())) <= 3
    assert abs(tks - env_ticks_total(o)) <= 0.01 * tks, (tks, env_ticks_total(o))
    mean_pts, _ = o.row_stats()
    assert abs(pts / tks - mean_pts) < 0.02 * mean_pts, (pts / tks, mean_pts)     # ~33 cached points per tick while sliding
    env.close()


def env_ticks_total(o):
    return int(o.counters()["ticks"])


def test_manifold_switch_and_refusals(torch):
    """The switch changes the contact model (the returns move away from the one-point-per-cylinder tick), switching it off restores
    the default bit for bit, and the entry points that have no manifold variant refuse."""
    n = 64
    g = torch.Generator().manual_seed(5)
    acts = (torch.rand((4, n, 8), generator=g) * 2 - 1).cuda()

    def run(env):
        env.reset(as_torch=True)
        out = []
        for t in range(4):
            obs, rew, done, _ = env.step(acts[t])
            out.append((obs.clone(), rew.clone(), env.last_ticks.clone()))
        return out

    env = SnakeVecEnv(num_envs=n, device=0)
    base = run(env)
    env.set_manifold(True, 0.1)
    man = run(env)
    for (o0, r0, t0), (o1, r1, t1) in zip(base, man):
        assert torch.equal(t0, t1)                       # the joints are prescribed: tick counts do not depend on the contact model
        assert torch.allclose(o0[:, :16], o1[:, :16], atol=1e-6)
    assert (base[-1][0][:, 48:51] - man[-1][0][:, 48:51]).abs().max() > 1e-4   # ... the base motion does
    tg = torch.zeros((n, 16), device="cuda")
    assert env._lib.snk_tick(env._h, tg.data_ptr(), 1, None) == -1            # SNK_E_ARG: raw ticks have no manifold variant
    assert b"manifold" in env._lib.snk_last_error()
    env.set_manifold(False)
    again = run(env)
    for (o0, r0, t0), (o2, r2, t2) in zip(base, again):
        assert torch.equal(o0, o2) and torch.equal(r0, r2) and torch.equal(t0, t2)
    env.close()
    e0 = SnakeVecEnv(num_envs=8, device=0, params=default_params(motor_solver=0))
    with pytest.raises(RuntimeError):
        e0.set_manifold(True)
    e0.close()


def test_manifold_host_path_equals_device_path(torch):
    """SnakeVecEnv.step(numpy) goes through snk_step_host_f64 (the batch in pipelined chunks of environments, each chunk a launch at an
    environment offset): with the manifolds on it must return what the device path returns, bit for bit -- the caches and the row
    tables are indexed by environment and by resident thread, not by launch."""
    n = 1000
    g = torch.Generator().manual_seed(11)
    acts = (torch.rand((3, n, 8), generator=g) * 2 - 1)
    a = SnakeVecEnv(num_envs=n, device=0); b = SnakeVecEnv(num_envs=n, device=0)
    a.set_manifold(True, 0.1); b.set_manifold(True, 0.1)
    a.reset(); b.reset(as_torch=True)
    for t in range(3):
        oa, ra, da, _ = a.step(acts[t].numpy().astype(np.float64))
        ob, rb, db, _ = b.step(acts[t].cuda())
        assert np.array_equal(oa.astype(np.float32), ob.cpu().numpy())
        assert np.array_equal(ra.astype(np.float32), rb.cpu().numpy())
        assert np.array_equal(np.asarray(da, bool), db.cpu().numpy().astype(bool))
    a.close(); b.close()
