"""mode='test' info stream on the GPU (snk_step_trace) vs the oracle's step_trace and the golden vectors the
reference's own Python recorded (tests/golden/reference_python_test_mode.npz).  Needs a B200."""
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


@pytest.mark.parametrize("solver", [2, 0])
def test_trace_vs_oracle(torch, solver):
    """Every step starts from the oracle's state (rounded to fp32): per-tick observations and link positions of the env-steps
    whose tick count agrees are compared row by row; the traced step returns what the plain step returns."""
    n, steps = (96, 3) if solver else (40, 2)
    p = default_params(motor_solver=solver)
    rng = np.random.default_rng(3)
    env = make_env(n, p, mode="test"); plain = make_env(n, p); o = Oracle(n, p)
    env.reset(); plain.reset(); o.reset()
    for t in range(steps):
        act = rng.uniform(-1, 1, (n, 8)).astype(np.float32)
        s32 = o.get_state().astype(np.float32)
        o.set_state(s32.astype(np.float64)); env.set_state(s32); plain.set_state(s32)
        obs, rew, done, infos = env.step(act)                          # numpy in -> numpy out
        tk = np.asarray(env.last_ticks)
        pobs, prew, pdone, _ = plain.step(act)
        assert np.array_equal(obs, pobs) and np.array_equal(rew, prew) and np.array_equal(done, pdone)   # tracing changes nothing
        assert np.array_equal(tk, np.asarray(plain.last_ticks))
        oo, orr, od, ot, tobs, tlnk = o.step_trace(act.astype(np.float64))
        assert (tk == ot).mean() >= (0.99 if solver else 0.90)
        worst_q = worst_l = 0.0
        per_env_q = []
        for e in np.where(tk == ot)[0]:
            info = infos[e]
            assert info["frames"] == [] and len(info["internal_observations"]) == tk[e] == len(info["link_positions"])
            if tk[e] == 0:
                continue
            io = np.stack(info["internal_observations"]); lp = np.stack(info["link_positions"])
            assert io.shape == (tk[e], 56) and lp.shape == (tk[e], 51)
            per_env_q.append(np.abs(io[:, :16] - tobs[e, :tk[e], :16]).max())
            worst_q = max(worst_q, per_env_q[-1])
            worst_l = max(worst_l, np.median(np.abs(lp - tlnk[e, :tk[e]])))
            if not done[e]:
                assert np.array_equal(io[-1], obs[e])                   # last internal observation = the returned one
        # exact solver: the joints follow the motor law, round-off in every tick.  Relaxed motor rows (solver 0): 50 unconverged
        # sweeps amplify fp32 round-off in mid-step ticks: round-off for the typical environment only
        assert worst_q < (1e-5 if solver else 5e-2), worst_q
        assert np.median(per_env_q) < (1e-5 if solver else 2e-3), np.median(per_env_q)
        assert worst_l < 2e-3, worst_l                                  # link positions ride on the (contact sensitive) base pose
    env.close(); plain.close()


def test_trace_link_positions_are_forward_kinematics(torch, model):
    """link_positions of a tick = forward kinematics of the observation of the same tick (independent numpy FK from the model
    tables), for the GPU stream itself -- no oracle involved."""
    from scipy.spatial.transform import Rotation as Rot
    n = 8
    env = make_env(n, mode="test"); env.reset()
    act = np.random.default_rng(1).uniform(-1, 1, (n, 8)).astype(np.float32)
    _, _, _, infos = env.step(act)
    checked = 0
    for e in range(n):
        for io, lp in zip(infos[e]["internal_observations"], infos[e]["link_positions"]):
            R = Rot.from_quat(io[51:55]).as_matrix(); p = io[48:51].copy()
            pts = []
            for b in range(17):
                if b > 0:
                    j = b - 1
                    p = p + R @ np.asarray(model.joint_t[j])
                    Rq = Rot.from_rotvec(np.asarray(model.joint_axis[j]) * io[j]).as_matrix()
                    R = R @ np.asarray(model.joint_R0[j]).reshape(3, 3) @ Rq
                pts.append(p + R @ np.asarray(model.height_pt[b]))
            want = np.asarray(pts).T.reshape(-1)                        # [x0..x16 | y0..y16 | z0..z16]
            assert np.abs(lp - want).max() < 5e-6
            checked += 1
    assert checked > 20
    env.close()


def test_trace_vs_reference_python_golden(torch, golden_test_mode):
    """The stream the reference's own Python recorded (fake client on the oracle) vs the GPU, free running over 14 env-steps of the
    serpenoid scenario: joints to round-off wherever the tick counts agree, link positions within a centimetre or two."""
    g = golden_test_mode
    from bullet_envs_b200 import SnakeGymEnv
    env = SnakeGymEnv()
    env.mode = "test"
    env.reset()
    row, agree = 0, 0
    for t, a in enumerate(g["serpenoid/actions"]):
        ob, r, d, info = env.step(np.array(a))
        k = int(g["serpenoid/ticks"][t])
        if len(info["internal_observations"]) == k and k > 0:
            agree += 1
            io = np.stack(info["internal_observations"]); lp = np.stack(info["link_positions"])
            assert np.abs(io[:, :16] - g["serpenoid/internal_observations"][row:row + k, :16]).max() < 1e-4
            # free running: the fp32 base pose parts from the fp64 golden like any two runs of a contact-rich system (round-off grows
            # ~3x per env-step, see test_golden_scenarios_on_the_gpu; centimetres after a dozen env-steps), so the link positions are
            # compared over the first env-steps only and the later ones through the joints, which follow the motor law
            if t < 3:
                assert np.median(np.abs(lp - g["serpenoid/link_positions"][row:row + k])) < 2e-2
        assert bool(d) == bool(g["serpenoid/done"][t])
        row += k
    assert agree >= 12
    env.close()


def test_self_clearance_vs_oracle(torch):
    """snk_self_clearance: the device counter equals the oracle's bound on identical states."""
    from scenarios import rollout_states
    n = 256
    p = default_params()
    o = Oracle(n, p)
    s, _ = rollout_states(o)
    rng = np.random.default_rng(8)
    s[:32, 13:29] = rng.uniform(-1.2, 1.2, (32, 16))            # some poses far outside the reachable range too
    env = make_env(n, p)
    env.set_state(s); o.set_state(s.astype(np.float32).astype(np.float64))
    g = env.self_clearance().cpu().numpy()
    assert np.abs(g - o.self_clearance()).max() < 1e-4      # fp32 forward kinematics over a 1 m, 16-joint chain
    env.close()


def test_no_self_contact_is_missed_over_a_rollout(torch):
    """4096 environments x 20 env-steps of U[-1,1] actions: the closest pair of cylinders that Bullet would test under
    URDF_USE_SELF_COLLISION (snake.py:93) never comes within 15 mm, so leaving those pairs out of the kernels loses nothing."""
    n = 4096
    env = make_env(n)
    env.reset(as_torch=True)
    g = torch.Generator(device="cuda").manual_seed(0)
    worst = torch.full((), 1.0, device="cuda")
    for _ in range(20):
        env.step(torch.rand((n, 8), generator=g, device="cuda") * 2 - 1)
        worst = torch.minimum(worst, env.self_clearance().min())
    assert float(worst) > 0.015, float(worst)
    env.close()


def test_facade_raw_single_env_semantics_vs_reference_python(torch):
    """SnakeGymEnv (the facade) behaves like the reference class used without the vector wrapper (a3c/agent.py:99-125):
    on done it returns the terminal observation and the next reward uses the dead episode's x.  Golden: the reference's own Python
    stepping on after done (tests/golden/reference_python_raw_single_env.npz); free running, so contact-sensitive values are
    compared with tolerances and the joints (motor law) to round-off while the tick counts agree."""
    import os
    from bullet_envs_b200 import SnakeGymEnv
    g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reference_python_raw_single_env.npz"))
    env = SnakeGymEnv()
    wrap = SnakeGymEnv(raw_semantics=False)
    env.reset(); wrap.reset()
    acts, G_ob, G_r, G_d = g["clipped/actions"], g["clipped/obs"], g["clipped/rew"], g["clipped/done"]
    n_done = 0
    for t, a in enumerate(acts):
        ob, r, d, info = env.step(np.array(a))
        wob, wr, wd, _ = wrap.step(np.array(a))
        assert d == bool(G_d[t]) == wd
        assert np.abs(ob[:16] - G_ob[t + 1][:16]).max() < 1e-4
        assert abs(r - G_r[t]) < 0.05 and info == {}
        if d:
            n_done += 1
            assert np.abs(ob[:16]).max() > 0.4                 # terminal observation: the joints are where the episode died
            assert np.allclose(wob[:32], 0) and wob[54] == 1   # wrapper semantics: the post-reset observation
        elif t > 0 and G_d[t - 1]:
            assert abs((r - wr) + G_ob[t][48]) < 0.02          # Q8: the raw reward carries -alpha * x_terminal
    assert n_done >= 3
    env.close(); wrap.close()
