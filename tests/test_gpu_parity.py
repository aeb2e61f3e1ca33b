"""CUDA path vs the CPU oracle, through the C ABI (ctypes -> libsnake_b200.so).  Needs a B200.

Tolerances (fp32 kernel vs fp64 oracle):
  * joint angles / rates follow the motor law and must agree to fp32 round-off (north_star: 1e-4
    relative after one step) in EVERY environment;
  * the base twist is the solution of a projected Gauss-Seidel over ~32 redundant contacts: the
    typical environment agrees to round-off, the tail (a different sweep count at the residual
    threshold, a contact flipping at the breaking distance) is bounded by percentiles;
  * integer outputs (tick counts, done flags) are compared exactly, on the fraction stated.
"""
import numpy as np
import pytest

from bullet_envs_b200 import default_params, gait_params
from oracle.oracle_py import Oracle
from scenarios import err_table, rollout_states, serpenoid_actions

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


def test_library_is_the_cuda_build(torch):
    from bullet_envs_b200 import _abi
    lib = _abi.load_library()
    assert b"sm_100a" in lib.snk_build_info()


@pytest.mark.parametrize("solver", [2, 0])
def test_reset_observation(torch, solver):
    p = default_params(motor_solver=solver)
    env = make_env(100, p); o = Oracle(100, p)
    assert np.array_equal(env.reset(as_torch=True).cpu().numpy().astype(np.float64), o.reset())
    assert np.array_equal(env.reset(), o.reset())  # numpy path (snk_reset_host)
    env.close()


@pytest.mark.parametrize("solver,n", [(1, 512), (0, 128)])
def test_one_tick_from_rollout_states(torch, solver, n):
    p = default_params(motor_solver=solver)
    o = Oracle(n, p)
    s, tg = rollout_states(o)
    env = make_env(n, p)
    env.set_state(s); o.set_state(s)
    assert np.array_equal(env.get_state().cpu().numpy(), s.astype(np.float32))  # import/export round trip
    env.tick(tg, 1); o.tick(tg.astype(np.float64), 1)
    g = env.get_state().cpu().numpy().astype(np.float64)
    t = err_table(o.get_state(), g)
    if solver == 1:  # prescribed joints: round-off
        assert t["q"][3] <= 1e-6 and t["qd"][3] <= 1e-5 * max(1.0, t["qd"][0]), (t["q"], t["qd"])
    else:            # motor rows relaxed by 50 (unconverged) sweeps together with the contacts: round-off for the
                     # typical environment, 1e-3 at the 99th percentile
        assert t["q"][1] <= 1e-6 and t["q"][2] <= 1e-3 and t["q"][3] <= 1e-2, t["q"]
        assert t["qd"][1] <= 1e-4 * max(1.0, t["qd"][0]) and t["qd"][2] <= 2e-2 * max(1.0, t["qd"][0]), t["qd"]
    for f in ("vel", "omega"):
        assert t[f][1] < 1e-4 * t[f][0] and t[f][2] < 3e-2 * t[f][0], (f, t[f])
    assert t["pos"][1] < 1e-6 and t["quat"][1] < 1e-5 and t["pos"][3] < 1e-3
    assert t["tau"][1] < 1e-3 * t["tau"][0] and t["fz"][1] < 1e-3 * t["fz"][0]
    c = env.counters()
    assert c["ticks"] == n
    env.close()


@pytest.mark.parametrize("solver", [2, 0])
def test_env_steps_vs_oracle(torch, solver):
    """config 2 of BASELINE.json at an oracle-sized sample: seeded U[-1,1] actions generated on the CPU as
    float32 and fed identically to both sides.  The first step starts from the common reset pose; before every
    later step the GPU batch is re-synchronised to the oracle's state (rounded to fp32), so each step is a
    clean one-env-step (~30 ticks) comparison; the free-running variant is test_free_running_statistics."""
    n, steps = (256, 5) if solver else (64, 3)
    p = default_params(motor_solver=solver)
    g = torch.Generator().manual_seed(0)
    acts = (torch.rand((steps, n, 8), generator=g) * 2 - 1).numpy()
    env = make_env(n, p); o = Oracle(n, p)
    env.reset(as_torch=True); o.reset()
    for t in range(steps):
        s32 = o.get_state().astype(np.float32)
        o.set_state(s32.astype(np.float64)); env.set_state(s32)
        obs, rew, done, infos = env.step(torch.from_numpy(acts[t]).cuda())
        tk = env.last_ticks.cpu().numpy()
        oo, orr, od, ot = o.step(acts[t].astype(np.float64), threads=8)
        og = obs.cpu().numpy().astype(np.float64); rg = rew.cpu().numpy().astype(np.float64); dg = done.cpu().numpy()
        # integer outputs: exact with the prescribed joints; the relaxed motor rows (solver 0) leave ~1e-3 rad of
        # unconverged joint error, which moves the 0.05 rad loop exit by one tick in a few environments
        frac = 0.99 if solver else 0.90
        assert (tk == ot).mean() >= frac, (t, (tk == ot).mean())
        assert (dg == od).mean() >= 0.97
        same = (tk == ot) & (dg == od)
        tol_q = 1e-5 if solver else 5e-3
        assert np.abs(og - oo)[same][:, :16].max() < tol_q, (t, np.abs(og - oo)[same][:, :16].max())
        assert np.median(np.abs(og - oo)[same][:, 48:51].max(1)) < 2e-3            # base position (contact sensitive)
        assert np.median(np.abs(rg - orr)[same]) < 5e-3                            # reward
        assert len(infos) == n and infos[0] == {}
    env.close()


def test_free_running_statistics(torch):
    """20 free-running env-steps of 512 environments (no re-synchronisation): trajectories of a contact-rich
    system separate exponentially, so the comparison is statistical -- tick counts (a function of the joints
    only) stay exact, batch means of reward, forward progress and episode ends agree."""
    n, steps = 512, 20
    g = torch.Generator().manual_seed(1)
    acts = (torch.rand((steps, n, 8), generator=g) * 2 - 1).numpy()
    p = default_params()
    env = make_env(n, p); o = Oracle(n, p)
    env.reset(as_torch=True); o.reset()
    tick_eq = []; rg_all = []; ro_all = []; dg_all = []; do_all = []; xg = xo = None
    for t in range(steps):
        obs, rew, done, _ = env.step(torch.from_numpy(acts[t]).cuda())
        oo, orr, od, ot = o.step(acts[t].astype(np.float64), threads=8)
        tick_eq.append((env.last_ticks.cpu().numpy() == ot).mean())
        rg_all.append(rew.cpu().numpy().astype(np.float64)); ro_all.append(orr); dg_all.append(done.cpu().numpy()); do_all.append(od)
    rg, ro, dg, do = (np.concatenate(x) for x in (rg_all, ro_all, dg_all, do_all))
    assert np.mean(tick_eq) > 0.97
    # episode ends: the |q9| > 0.5 rule depends on the (prescribed) joints only -> nearly identical counts
    assert abs(int(dg.sum()) - int(do.sum())) <= 0.1 * do.sum() + 5, (dg.sum(), do.sum())
    # the -10 "collision" term fires on |Fz of joint 0| > 10 N, i.e. on a base acceleration beyond ~0.2 m/s^2
    # (SnakeGymEnv.py:94): a rare hair-trigger event, compared as a count with a Poisson allowance
    eg, eo = int(((rg < -7.5) & ~dg).sum() + (rg < -12.5).sum()), int(((ro < -7.5) & ~do).sum() + (ro < -12.5).sum())
    assert abs(eg - eo) <= 4 * np.sqrt(max(eo, 1)) + 5, (eg, eo)
    # everything else (progress, drift, energy): batch mean of the event-free rewards
    mg, mo = rg[(rg > -4) & ~dg].mean(), ro[(ro > -4) & ~do].mean()
    assert abs(mg - mo) < 0.05 * abs(mo) + 2e-3, (mg, mo)
    env.close()


def test_golden_scenarios_on_the_gpu(torch, golden):
    """The reference-Python golden vectors (tests/golden, made by tests/golden/make_golden.py): tick counts, done
    flags and joint angles of every step; the contact-sensitive base pose over the first steps."""
    for name in ("const_half", "random", "clipped", "serpenoid", "terminate_q9", "lifted"):
        acts = golden[name + "/actions"]
        inj = golden[name + "/inject"]
        env = make_env(1)
        obs0 = env.reset()
        assert np.array_equal(obs0[0], golden[name + "/obs"][0])
        agree = 0
        for t, a in enumerate(acts):
            if inj[t].any():                                    # the generator lifted / threw the snake before this step
                s = env.get_state(); s[0, 2] += float(inj[t, 0]); s[0, 9] += float(inj[t, 1]); env.set_state(s)
            ob, r, d, _ = env.step(a[None, :])                  # numpy in -> numpy out (snk_step_host)
            tk = int(env.last_ticks[0])
            if tk != int(golden[name + "/ticks"][t]) or bool(d[0]) != bool(golden[name + "/done"][t]):
                break                                           # fp32 trajectories part ways after a discrete event
            agree += 1
            assert np.abs(ob[0, :16] - golden[name + "/obs"][t + 1][:16]).max() < 1e-5, (name, t)
            if t < 3:   # ~90 ticks of free-running fp32 vs fp64 contact dynamics: round-off grows about 3x per env-step (the same fp32
                        # code on the CPU, tests/hostemu, lands anywhere between 1e-4 and 1e-2 by the third step depending on how the
                        # arithmetic is associated), so this bounds the order of magnitude only
                assert np.abs(ob[0, 48:51] - golden[name + "/obs"][t + 1][48:51]).max() < (2e-3 if t == 0 else 2e-2), (name, t)
                assert abs(r[0] - golden[name + "/rew"][t]) < 2e-2, (name, t)
        assert agree >= min(len(acts), 8), (name, agree)
        if name == "terminate_q9":
            assert golden[name + "/done"][:agree].sum() >= 4    # the |q9| > 0.5 termination was really taken on the GPU
        if name == "lifted":
            assert agree == len(acts)                           # height break after tick 1, mid-loop break, zero-tick height termination
        env.close()


def test_gait_script_ticks_finite_motor_force(torch):
    """config 1 of BASELINE.json: raw ticks at dt = 0.01, g = -9.81, 4 N.m motors (snake_gait_test.py:50-53,96-104),
    serpenoid targets with t = tick * 0.01; the force limit makes the motor rows inequalities, so this runs
    the warp-per-env kernel with Bullet-order rows."""
    p = gait_params()
    n = 8
    env = make_env(n, p); o = Oracle(n, p)
    env.reset(as_torch=True); o.reset()
    nn = np.arange(16)
    qerr = []; free = make_env(n, p)
    free.reset(as_torch=True)
    for tick in range(120):
        tg = np.where(nn % 2 == 1, -(np.pi / 6) * np.sin(4 * nn + 2 * tick * 0.01), 0.0)[None, :].repeat(n, 0).astype(np.float32)
        if tick % 10 == 0:  # re-synchronise every 10 ticks: bounded per-segment error
            s32 = o.get_state().astype(np.float32)
            o.set_state(s32.astype(np.float64)); env.set_state(s32)
        env.tick(tg, 1); o.tick(tg.astype(np.float64), 1); free.tick(tg, 1)
        qerr.append(np.abs(env.get_state().cpu().numpy()[:, 13:29] - o.get_state()[:, 13:29]).max())
    assert max(qerr[:10]) < 1e-4 and max(qerr) < 2e-2, (max(qerr[:10]), max(qerr))
    # free running over the 120 ticks: the saturated 4 N.m motors let round-off grow; stated bound 0.2 rad
    assert np.abs(free.get_state().cpu().numpy()[:, 13:29] - o.get_state()[:, 13:29]).max() < 0.2
    env.close(); free.close()


def test_hundred_step_gait_rollout_bounds(torch):
    """north_star: 'within a stated bound over a 100-step gait rollout; episode rewards within 1%'.
    Stated bounds for the serpenoid gait driven through env.step (free running, no re-synchronisation):
    joint angles stay within 1e-4 rad of the oracle at every step where both sides ran the same number of
    ticks; the batch-mean 100-step return agrees within 1%; per-environment returns within 5% median."""
    n, steps = 64, 100
    rng = np.random.default_rng(7)
    acts = serpenoid_actions(steps, n, phase=rng.uniform(0, 2 * np.pi, n)).astype(np.float32)
    p = default_params()
    env = make_env(n, p); o = Oracle(n, p)
    env.reset(as_torch=True); o.reset()
    Rg = np.zeros(n); Ro = np.zeros(n); worst_q = 0.0; same_frac = []
    for t in range(steps):
        obs, rew, done, _ = env.step(torch.from_numpy(acts[t]).cuda())
        oo, orr, od, ot = o.step(acts[t].astype(np.float64), threads=8)
        tk = env.last_ticks.cpu().numpy()
        same = (tk == ot) & (done.cpu().numpy() == od)
        same_frac.append(same.mean())
        if same.any():
            worst_q = max(worst_q, float(np.abs(obs.cpu().numpy()[same][:, :16] - oo[same][:, :16]).max()))
        Rg += rew.cpu().numpy(); Ro += orr
    assert worst_q < 1e-4, worst_q
    assert np.mean(same_frac) > 0.9
    assert abs(Rg.mean() - Ro.mean()) <= 0.01 * abs(Ro.mean()) + 1e-3, (Rg.mean(), Ro.mean())
    assert np.median(np.abs(Rg - Ro) / (np.abs(Ro) + 1e-9)) < 0.05
    env.close()


def test_numpy_path_equals_device_path(torch):
    n = 97
    rng = np.random.default_rng(2)
    a = rng.uniform(-1, 1, (2, n, 8)).astype(np.float32)
    e1 = make_env(n); e2 = make_env(n); e3 = make_env(n, obs_dtype=np.float32, pinned_io=True)
    e1.reset(); e2.reset(); e3.reset()
    for t in range(2):
        o1, r1, d1, _ = e1.step(a[t])
        o2, r2, d2, _ = e2.step(torch.from_numpy(a[t]).cuda())
        o3, r3, d3, _ = e3.step(a[t])                           # page-locked result buffers, DMA without staging
        assert o3.dtype == np.float32 and d3.dtype == bool
        assert np.array_equal(o3, o2.cpu().numpy()) and np.array_equal(r3, r2.cpu().numpy()) and np.array_equal(d3, d2.cpu().numpy())
        assert o1.dtype == np.float64 and d1.dtype == bool
        assert np.array_equal(o1.astype(np.float32), o2.cpu().numpy()) and np.array_equal(r1.astype(np.float32), r2.cpu().numpy())
        assert np.array_equal(d1, d2.cpu().numpy())
    e1.close(); e2.close(); e3.close()


def test_numpy_path_large_batch_rows_widened_during_the_launch(torch):
    """snk_step_host_f64 with a batch of several waves: one launch, the host threads widen every row when its ticks word arrives in
    the mapped buffer (HandOut::flag_rows) -- the result must be the device path's, whatever order the rows finish in; twice, so
    that a stale ticks word of the previous call would be caught."""
    n = 150001
    g = torch.Generator().manual_seed(11)
    a = (torch.rand((3, n, 8), generator=g) * 2.4 - 1.2)
    e1 = make_env(n); e2 = make_env(n)
    e1.reset(); e2.reset()
    for t in range(3):
        o1, r1, d1, i1 = e1.step(a[t].numpy().astype(np.float64))
        o2, r2, d2, _ = e2.step(a[t].cuda()); torch.cuda.synchronize()
        assert o1.dtype == np.float64 and o1.shape == (n, 56)
        assert np.array_equal(o1.astype(np.float32), o2.cpu().numpy())
        assert np.array_equal(r1.astype(np.float32), r2.cpu().numpy()) and np.array_equal(d1, d2.cpu().numpy())
        assert np.array_equal(np.asarray(e1.last_ticks), e2.last_ticks.cpu().numpy())
    e1.close(); e2.close()


def test_host_path_call_waits_for_device_path_work_in_flight(torch):
    """A device-path step is asynchronous on the caller's stream; a numpy step issued right behind it runs on the library's own
    stream and must still see its result (event ordering inside the C ABI).  Large enough that the first kernel is still running."""
    n = 40000
    g = torch.Generator().manual_seed(5)
    a = (torch.rand((2, n, 8), generator=g) * 2 - 1)
    mixed = make_env(n, obs_dtype=np.float32); plain = make_env(n)
    mixed.reset(); plain.reset()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        mixed.step(a[0].cuda(non_blocking=True))                # enqueued on stream s, not waited for
    o_m, r_m, d_m, _ = mixed.step(a[1].numpy())                 # host path, the library's stream
    plain.step(a[0].cuda()); torch.cuda.synchronize()
    o_p, r_p, d_p, _ = plain.step(a[1].cuda()); torch.cuda.synchronize()
    assert np.array_equal(o_m, o_p.cpu().numpy()) and np.array_equal(r_m, r_p.cpu().numpy()) and np.array_equal(d_m, d_p.cpu().numpy())
    mixed.close(); plain.close()


def test_edge_cases(torch):
    # ragged sizes: 1 env, a partial last CTA, out-of-range actions are clipped like np.clip
    for n in (1, 31, 33):
        env = make_env(n); o = Oracle(n)
        env.reset(); o.reset()
        a = np.linspace(-3, 3, n * 8, dtype=np.float32).reshape(n, 8)
        ob, r, d, _ = env.step(a)
        oo, orr, od, ot = o.step(a.astype(np.float64))
        assert np.array_equal(np.asarray(env.last_ticks), ot)
        assert np.abs(ob[:, :16] - oo[:, :16]).max() < 1e-5
        env.close()
    # zero-tick step (snake.py:283-284): repeat an action that is already reached
    env = make_env(4); env.reset()
    a = np.full((4, 8), 0.5, np.float32)
    env.step(a); env.step(a)
    ob, r, d, _ = env.step(a)
    assert (np.asarray(env.last_ticks) == 0).all() and not d.any()
    energy = np.sum(ob[:, 16:32] * ob[:, 32:48] * 0.01, axis=1)
    assert np.allclose(r, -0.01 * np.abs(ob[:, 49]) - 0.1 * energy, atol=1e-6)
    env.close()
    # a NaN action is a no-op step, as in the reference (checkBound and checkFeedback compare false on NaN)
    env = make_env(64); env.reset()
    a = np.zeros((64, 8), np.float32); a[:, 1] = 0.8; a[5, 3] = np.nan
    ob, r, d, _ = env.step(a)
    tk = np.asarray(env.last_ticks)
    assert tk[5] == 0 and (tk[np.arange(64) != 5] > 0).all() and np.isfinite(ob).all()
    # a non-finite state poisons one environment only: it is force-reset, reported done and counted
    s = env.get_state().cpu().numpy(); s[7, 8] = np.nan
    env.set_state(s)
    a[:, 1] = -0.8; a[5, 3] = 0.0
    ob, r, d, _ = env.step(a)
    c = env.counters()
    assert np.isfinite(ob).all() and np.isfinite(r).all()
    assert d[7] and r[7] == -5.0 and c["nonfinite"] == 1 and not d[np.arange(64) != 7].any()
    env.close()
    # closed handle
    with pytest.raises(RuntimeError):
        env.step(a)
    # action / observation rows move as 16-byte vectors: a misaligned device pointer is refused instead of faulting in the kernel,
    # a misaligned host buffer is staged
    env = make_env(8); env.reset()
    flat = torch.zeros(8 * 8 + 1, device="cuda")
    with pytest.raises(RuntimeError, match="16-byte aligned"):
        env.step(flat[1:].view(8, 8))
    ob1, _, _, _ = env.step(np.zeros(8 * 8 + 1, np.float32)[1:].reshape(8, 8))      # numpy view at a 4-byte offset: host path stages it
    assert np.isfinite(ob1).all()
    env.close()


def test_single_env_facade_and_vec_surface(torch):
    """SnakeGymEnv(robot, args)-shaped facade (SnakeGymEnv.py:4-103) and the SubprocVecEnv surface
    (ppo/multiprocessing_env.py:31-153) that the reference's callers touch."""
    import types
    from bullet_envs_b200 import Snake, SnakeGymEnv, SnakeVecEnv
    args = types.SimpleNamespace(alpha=1, beta=0.01, gamma=0.1, mode="train", gaitSelection=1, scaling_factor=6)
    env = SnakeGymEnv(Snake(None, "snake/snake.urdf", args), args)
    assert env.observation_space.shape == (56,) and env.action_space.shape == (8,)
    assert env.robot.numMotors == 16 and env.alpha == 1 and env.mode == "train" and env._gaitSelection == 1
    ob = env.reset()
    assert ob.shape == (56,) and ob[54] == 1.0
    a = np.array([2.0, -3.0, 0.5, 0.1, 0.0, 0.2, -0.2, 0.9])
    ob, r, d, info = env.step(a)
    assert a[0] == 1.0 and a[1] == -1.0                         # checkBound clips the caller's array in place
    assert isinstance(r, float) and isinstance(d, bool) and info == {} and ob.shape == (56,)
    assert abs(ob[1] - np.pi / 6) < 0.06 and abs(ob[3] + np.pi / 6) < 0.06   # driven (odd) joints reach the clipped targets
    assert env.render().size == 0
    assert env.robot.calculateEnergy(ob) == pytest.approx(float(np.sum(ob[16:32] * ob[32:48] * 0.01)))
    soft = env.reset()
    assert np.abs(soft[32:48]).max() > 0 and np.allclose(soft[:32], 0)        # Q9: a soft reset keeps the last applied torques
    hard = env.reset(hardReset=True)                                          # snake.py:88-95: world rebuilt, nothing stale
    assert np.allclose(hard[:54], 0) and hard[54] == 1.0 and hard[55] == 0.0
    env.close()
    vec = SnakeVecEnv([None] * 16)                              # constructed from a list of env thunks, like SubprocVecEnv
    assert len(vec) == 16 and vec.num_envs == vec.nenvs == 16
    assert vec.observation_space.high[0] == pytest.approx(np.pi) and np.isinf(vec.observation_space.high[20]) and vec.action_space.high[0] == 1
    s = vec.reset()
    vec.step_async(np.zeros((16, 8))); s2, r2, d2, infos = vec.step_wait()
    assert s.shape == s2.shape == (16, 56) and s2.dtype == np.float64 and d2.dtype == bool and len(infos) == 16
    import torch as T
    assert T.FloatTensor(s2).shape == (16, 56) and (1 - d2).sum() == 16        # what ppo/train.py:114,134 do with the results
    vec.close(); vec.close()


def test_one_model_per_process_is_enforced(torch):
    """the thread-per-env kernel keeps the model tables in one __constant__ symbol: a second handle with a
    DIFFERENT model must be refused while the first is alive (and accepted afterwards), never silently mixed"""
    from bullet_envs_b200 import ImportRules, SnakeVecEnv, build_model
    other = build_model(rules=ImportRules(unit_mass_for_missing_inertial=False))
    e1 = make_env(8)
    e1b = make_env(8)                                           # same model: fine
    with pytest.raises(RuntimeError, match="one model per process"):
        SnakeVecEnv(num_envs=8, device=0, model=other)
    e1.close(); e1b.close()
    e2 = SnakeVecEnv(num_envs=8, device=0, model=other)         # all handles of the first model are gone
    o = Oracle(8, model=other)
    a = np.full((8, 8), 0.7, np.float32)
    e2.reset(); o.reset()
    ob, r, d, _ = e2.step(a)
    oo, orr, od, ot = o.step(a.astype(np.float64))
    assert np.array_equal(np.asarray(e2.last_ticks), ot) and np.abs(ob[:, :16] - oo[:, :16]).max() < 1e-5
    e2.close()


def test_free_flight_matches_independent_numpy_dynamics(torch, model):
    """Physics known-answer test of the CUDA tick that does not involve the oracle: far above the plane (no contacts)
    the base acceleration must satisfy the hybrid dynamics  M_bb a_b + M_bj qdd = F_b  with the joint accelerations the
    motor law prescribes, where M and F come from the dense world-frame Newton-Euler in tests/ref_dynamics_numpy.py."""
    import ref_dynamics_numpy as rd
    n = 16
    rng = np.random.default_rng(21)
    p = default_params()
    s = np.zeros((n, 64)); tg = rng.uniform(-0.5, 0.5, (n, 16)).astype(np.float32)
    for e in range(n):
        s[e, 0:3] = [0.3, -0.2, 5.0]
        q = rng.normal(size=4); s[e, 3:7] = q / np.linalg.norm(q)
        s[e, 7:10] = rng.normal(size=3) * 0.5; s[e, 10:13] = rng.normal(size=3)
        s[e, 13:29] = rng.uniform(-0.6, 0.6, 16); s[e, 29:45] = rng.normal(size=16) * 2.0
    s = s.astype(np.float32).astype(np.float64)
    env = make_env(n, p)
    env.set_state(s); env.tick(tg, 1)
    g = env.get_state().cpu().numpy().astype(np.float64)
    dt = p.dt
    for e in range(n):
        acc_free, Mw = rd.forward_dynamics(model, s[e, 0:3], s[e, 3:7], s[e, 7:10], s[e, 10:13], s[e, 13:29], s[e, 29:45],
                                           np.zeros(16), [0, 0, -9.8], 0.04, 0.04)
        F = Mw @ acc_free                                            # generalized force incl. all bias terms
        qd_new = 0.1 * (tg[e] - s[e, 13:29]) / dt                    # motor law
        qdd = (qd_new - s[e, 29:45]) / dt
        ab = np.linalg.solve(Mw[:6, :6], F[:6] - Mw[:6, 6:] @ qdd)   # [angular, linear] base acceleration, world axes
        got = np.concatenate([g[e, 10:13] - s[e, 10:13], g[e, 7:10] - s[e, 7:10]]) / dt
        assert np.abs(got - ab).max() < 2e-3 * max(1.0, np.abs(ab).max()), (e, got, ab)
        assert np.abs(g[e, 29:45] - qd_new).max() < 1e-3 and np.abs(g[e, 13:29] - (s[e, 13:29] + dt * qd_new)).max() < 1e-6
    env.close()


def test_rest_on_the_plane_and_gait_makes_progress(torch):
    """contact known-answers of the CUDA path: the reset pose is a resting contact (the chain neither sinks nor drifts),
    and the serpenoid gait on the anisotropic-friction skin moves the snake along its body axis like the oracle's."""
    n = 32
    env = make_env(n)
    env.reset(as_torch=True)
    env.tick(np.zeros((n, 16), np.float32), 120)                      # half a second at rest
    s = env.get_state().cpu().numpy()
    assert np.abs(s[:, 7:13]).max() < 2e-3 and np.abs(s[:, 2]).max() < 1.5e-3 and np.abs(s[:, 0:2]).max() < 1e-3
    assert np.abs(s[:, 13:45]).max() == 0.0 and np.abs(np.linalg.norm(s[:, 3:7], axis=1) - 1).max() < 1e-6
    env.reset(as_torch=True)
    o = Oracle(n); o.reset()
    acts = serpenoid_actions(40, n).astype(np.float32)
    for t in range(40):
        obs, _, _, _ = env.step(acts[t]); oo, _, _, _ = o.step(acts[t].astype(np.float64), threads=8)
    dx_gpu, dx_ref = float(np.mean(obs[:, 48])), float(np.mean(oo[:, 48]))
    assert abs(dx_ref) > 0.005 and np.sign(dx_gpu) == np.sign(dx_ref) and abs(dx_gpu - dx_ref) < 0.25 * abs(dx_ref), (dx_gpu, dx_ref)
    env.close()
