"""Branches of the env-step that random-action runs rarely or never take, each on the GPU against the oracle:
checkSnakeHeight terminations (snake.py:237-245,299-301; SnakeGymEnv.py:99-103), gaitSelection 0 and 2 (snake.py:247-269),
pyramid friction (cone_friction = 0), zeroed torque slots after a reset (stale_obs_on_reset = 0), plus stress tests of the
fused rollout's cross-CTA ready queue (no sanitizer on this pool: many repetitions, each compared bit for bit)."""
import numpy as np
import pytest

from bullet_envs_b200 import default_params
from oracle.oracle_py import Oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (there is no CPU fallback)")
    return torch


def make_env(n, params=None, **kw):
    from bullet_envs_b200 import SnakeVecEnv
    return SnakeVecEnv(num_envs=n, device=0, params=params, **kw)


def sync(env, o):
    s32 = o.get_state().astype(np.float32)
    o.set_state(s32.astype(np.float64)); env.set_state(s32)


def test_height_terminations_in_a_batch(torch):
    """One batch, four kinds of environment side by side (a warp mixes them, so the masked commits of aborted lanes are exercised):
    0 mod 4: untouched; 1 mod 4: lifted to z = 0.2 -> the height test breaks the loop after the first tick (snake.py:299-301);
    2 mod 4: thrown upwards from the ground at 2.5..3.5 m/s -> break in the middle of the loop; 3 mod 4: lifted with the action
    already reached -> zero ticks, done by checkTermination's own height test (SnakeGymEnv.py:99-103)."""
    n = 256
    p = default_params()
    env = make_env(n, p); o = Oracle(n, p)
    env.reset(as_torch=True); o.reset()
    rng = np.random.default_rng(12)
    a0 = rng.uniform(-1, 1, (n, 8)).astype(np.float32)
    env.step(torch.from_numpy(a0).cuda()); o.step(a0.astype(np.float64), threads=8)
    sync(env, o)
    s = o.get_state()
    kind = np.arange(n) % 4
    s[kind == 1, 2] += 0.2
    s[kind == 2, 9] += rng.uniform(2.5, 3.5, (kind == 2).sum())
    s[kind == 3, 2] += 0.2
    s = s.astype(np.float32).astype(np.float64)
    o.set_state(s); env.set_state(s.astype(np.float32))
    a1 = rng.uniform(-1, 1, (n, 8)).astype(np.float32)
    a1[kind == 3] = a0[kind == 3]                                   # already reached: checkFeedback is false, no tick
    obs, rew, done, _ = env.step(torch.from_numpy(a1).cuda())
    oo, orr, od, ot = o.step(a1.astype(np.float64), threads=8)
    tk = env.last_ticks.cpu().numpy(); dg = done.cpu().numpy(); rg = rew.cpu().numpy().astype(np.float64)
    assert np.array_equal(tk, ot) and np.array_equal(dg, od)
    assert (tk[kind == 1] == 1).all() and dg[kind == 1].all()
    assert ((tk[kind == 2] > 1) & (tk[kind == 2] < 20)).all() and dg[kind == 2].all()
    assert (tk[kind == 3] == 0).all() and dg[kind == 3].all()
    assert not dg[kind == 0].all() and (tk[kind == 0] > 20).mean() > 0.8
    lifted = kind != 0
    assert (rg[lifted] < -4.0).all() and (orr[lifted] < -4.0).all()   # the -5 of SnakeGymEnv.py:40 (plus a few cm of progress)
    assert np.abs(rg - orr)[kind == 3].max() < 1e-5                 # no tick ran: closed-form reward
    og = obs.cpu().numpy()
    assert (og[lifted][:, :32] == 0).all() and (og[lifted][:, 48:51] == 0).all() and (og[lifted][:, 54] == 1).all()   # post-reset observation
    assert np.abs(og[kind == 0][:, :16] - oo[kind == 0][:, :16]).max() < 1e-5
    env.close()


@pytest.mark.parametrize("gait", [0, 2])
def test_gait_selection_0_and_2(torch, gait):
    """createAction (snake.py:247-269): gaitSelection 0 drives the even (pitch) motors, anything else all 16 (action dim 16)."""
    n, steps = 128, 3
    p = default_params(gait_selection=gait)
    env = make_env(n, p); o = Oracle(n, p)
    ad = 8 if gait == 0 else 16
    assert env.act_dim == ad and env.action_space.shape == (ad,)
    env.reset(as_torch=True); o.reset()
    rng = np.random.default_rng(30 + gait)
    for t in range(steps):
        a = (rng.uniform(-1, 1, (n, ad)) * (0.5 if gait == 0 else 0.4)).astype(np.float32)   # pitching lifts the body: keep it moderate
        sync(env, o)
        obs, rew, done, _ = env.step(torch.from_numpy(a).cuda())
        oo, orr, od, ot = o.step(a.astype(np.float64), threads=8)
        tk = env.last_ticks.cpu().numpy(); og = obs.cpu().numpy().astype(np.float64)
        assert (tk == ot).mean() >= 0.99 and (done.cpu().numpy() == od).mean() >= 0.99
        same = (tk == ot) & (done.cpu().numpy() == od)
        assert np.abs(og - oo)[same][:, :16].max() < 1e-5
        live = same & ~od
        if gait == 0:   # only the even joints move
            assert np.abs(og[live][:, 1:16:2]).max() < 1e-6 and np.abs(og[live][:, 0:16:2]).max() > 0.05
        else:
            assert np.abs(og[live][:, 0:16:2]).max() > 0.05 and np.abs(og[live][:, 1:16:2]).max() > 0.05
        assert np.median(np.abs(og - oo)[same][:, 48:51].max(1)) < 5e-3
    env.close()


def test_pyramid_friction(torch):
    """cone_friction = 0: each friction row clamped on its own (the pre-2.87 Bullet rule) -- the CONE = false
    instantiation of the kernels.  One physics tick from mid-motion states against the oracle with the same switch, at the
    tolerances of the cone's one-tick test; and the switch really selects a different friction bound.  (Whole env-steps are not
    compared for this switch: with mu = 2 and the anisotropic rows the box bound makes the chain tumble within a few ticks in
    BOTH implementations, so 30-tick trajectories separate at once.)"""
    from scenarios import err_table, rollout_states
    n = 512
    p = default_params(cone_friction=0, motor_solver=1)
    o = Oracle(n, p)
    s, tg = rollout_states(Oracle(n, default_params(motor_solver=1)))
    env = make_env(n, p); cone = make_env(n, default_params(motor_solver=1))
    env.set_state(s); o.set_state(s); cone.set_state(s)
    env.tick(tg, 1); o.tick(tg.astype(np.float64), 1); cone.tick(tg, 1)
    g = env.get_state().cpu().numpy().astype(np.float64)
    t = err_table(o.get_state(), g)
    assert t["q"][3] <= 1e-6 and t["qd"][3] <= 1e-5 * max(1.0, t["qd"][0]), (t["q"], t["qd"])
    for f in ("vel", "omega"):
        assert t[f][1] < 1e-4 * t[f][0] and t[f][2] < 3e-2 * t[f][0], (f, t[f])
    assert t["pos"][1] < 1e-6 and t["quat"][1] < 1e-5 and t["pos"][3] < 1e-3
    assert t["tau"][1] < 1e-3 * t["tau"][0] and t["fz"][1] < 1e-3 * t["fz"][0]
    d = np.abs(g - cone.get_state().cpu().numpy())[:, 7:13].max(1)
    assert np.median(d) > 1e-4, np.median(d)                         # not the cone's result
    env.close(); cone.close()
    # one env-step from the reset pose through the step kernel's CONE = false instantiation: the prescribed joints agree wherever
    # the integer outputs do
    env = make_env(256, default_params(cone_friction=0)); o = Oracle(256, default_params(cone_friction=0))
    env.reset(as_torch=True); o.reset()
    a = (np.random.default_rng(41).uniform(-1, 1, (256, 8)) * 0.3).astype(np.float32)
    obs, rew, done, _ = env.step(torch.from_numpy(a).cuda())
    oo, orr, od, ot = o.step(a.astype(np.float64), threads=8)
    same = (env.last_ticks.cpu().numpy() == ot) & (done.cpu().numpy() == od)
    assert same.mean() > 0.5 and np.abs(obs.cpu().numpy() - oo)[same][:, :16].max() < 1e-5, same.mean()
    env.close()


def test_zeroed_torque_slots_after_reset(torch):
    """stale_obs_on_reset = 0 (the documented alternative to quirk Q9): after a done the torque / reaction-force slots of the
    post-reset observation read zero instead of the last tick's values; the default keeps them."""
    n = 64
    for stale in (0, 1):
        p = default_params(stale_obs_on_reset=stale)
        env = make_env(n, p); o = Oracle(n, p)
        env.reset(as_torch=True); o.reset()
        a = np.zeros((n, 8), np.float32); a[:, 4] = 1.0; a[:, 0:4] = 1.0      # drives |q9| past 0.5 (SnakeGymEnv.py:100)
        obs, rew, done, _ = env.step(torch.from_numpy(a).cuda())
        oo, orr, od, ot = o.step(a.astype(np.float64))
        og = obs.cpu().numpy()
        assert done.all() and od.all() and np.array_equal(env.last_ticks.cpu().numpy(), ot)
        assert (og[:, :32] == 0).all()
        if stale:
            assert np.abs(og[:, 32:48]).max() > 0 and np.abs(oo[:, 32:48]).max() > 0
        else:
            assert (og[:, 32:48] == 0).all() and (og[:, 55] == 0).all() and (oo[:, 32:48] == 0).all() and (oo[:, 55] == 0).all()
        env.close()


def test_rollout_ready_queue_stress(torch):
    """snk_rollout_linear hands env-steps between CTAs through a global ready queue (volatile polling + fences).  500 000
    environments x 50 steps (13 batch waves: every lane pops pushed environments thousands of times), the kernel repeated 20 times
    from the same start state; every repetition must reproduce the first bit for bit, and the first equals the stepwise loop."""
    from bullet_envs_b200 import SnakeVecEnv
    n, T, reps = 500_000, 50, 20
    g = torch.Generator().manual_seed(21)
    cols = torch.tensor([33, 1, 17, 49, 9, 55, 3, 20])
    gain = (torch.rand((n, 8), generator=g) * 4 - 2)
    W = torch.zeros((n, 8, 56)); W[:, torch.arange(8), cols] = gain
    W, gain, cols = W.cuda(), gain.cuda(), cols.cuda()
    env = SnakeVecEnv(num_envs=n, device=0)
    env.reset(as_torch=True)
    s0 = env.get_state().clone()
    first = None
    for r in range(reps):
        env.set_state(s0)
        ret = env.rollout_linear(W, T)
        st = env.get_state()
        if first is None:
            first = (ret.clone(), st.clone())
        else:
            assert torch.equal(ret, first[0]) and torch.equal(st, first[1]), r
    # the stepwise loop (the weights have one non-zero per row: the policy arithmetic is exact in both)
    env.set_state(s0)
    obs = env.observe()
    acc = torch.zeros(n, device="cuda")
    for t in range(T):
        obs, rew, done, _ = env.step(gain * obs[:, cols])
        acc += rew
    assert torch.equal(acc, first[0]) and torch.equal(env.get_state(), first[1])
    env.close()


def test_two_rollouts_at_once_on_one_device(torch):
    """Two handles, two streams, rollouts in flight together: the rollout kernel needs its whole grid resident (its lanes wait for
    environments other CTAs push), so it is launched cooperatively and the driver runs the two one after the other instead of
    interleaving half-resident grids (which would hang).  Results equal the same rollouts run alone."""
    from bullet_envs_b200 import SnakeVecEnv
    n, T = 60_000, 8
    g = torch.Generator().manual_seed(22)
    Ws = [(torch.randn((n, 8, 56), generator=g) * 0.05).cuda() for _ in range(2)]
    alone = []
    for W in Ws:
        e = SnakeVecEnv(num_envs=n, device=0); e.reset(as_torch=True)
        alone.append(e.rollout_linear(W, T).clone()); e.close()
    envs = [SnakeVecEnv(num_envs=n, device=0) for _ in range(2)]
    for e in envs:
        e.reset(as_torch=True)
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in range(2)]
    rets = []
    for e, W, s in zip(envs, Ws, streams):
        with torch.cuda.stream(s):
            rets.append(e.rollout_linear(W, T))
    torch.cuda.synchronize()
    for r, a in zip(rets, alone):
        assert torch.equal(r, a)
    for e in envs:
        e.close()
