"""Task logic of the oracle vs the reference's own Python (golden vectors made by tests/golden/make_golden.py
from the unmodified SnakeGymEnv.py / snake.py) and vs the reference's documented quirks."""
import numpy as np
import pytest

from bullet_envs_b200 import default_params
from oracle.oracle_py import Oracle

SCENARIOS = ["const_half", "random", "clipped", "serpenoid", "terminate_q9", "lifted"]


def test_golden_meta(golden, model):
    assert list(golden["meta/motor_list"]) == model.motor_joint_indices
    assert int(golden["meta/obs_dim"]) == 56 and int(golden["meta/act_dim"]) == 8
    hi = golden["meta/obs_high"]
    assert np.allclose(hi[:16], np.pi) and np.isinf(hi[16:48]).all() and np.allclose(hi[48:], 1.0)


@pytest.mark.parametrize("solver", ["", "pgs/"])
@pytest.mark.parametrize("name", SCENARIOS)
def test_oracle_reproduces_reference_python(golden, model, name, solver):
    name = solver + name
    acts = golden[name + "/actions"]
    o = Oracle(1, default_params(motor_solver=0 if solver else 2), model)
    obs = o.reset()
    assert np.array_equal(obs[0], golden[name + "/obs"][0])
    inj = golden[name + "/inject"]
    for t, a in enumerate(acts):
        if inj[t].any():                                       # the state edit the generator applied to the fake simulator before this step
            s = o.get_state(); s[0, 2] += inj[t, 0]; s[0, 9] += inj[t, 1]; o.set_state(s)
        ob, r, d, tk = o.step(a[None, :])
        assert tk[0] == golden[name + "/ticks"][t], (name, t)
        assert bool(d[0]) == bool(golden[name + "/done"][t]), (name, t)
        assert np.allclose(ob[0], golden[name + "/obs"][t + 1], rtol=0, atol=1e-12), (name, t)
        assert r[0] == pytest.approx(golden[name + "/rew"][t], abs=1e-12), (name, t)
    # checkBound clipped in place exactly like np.clip
    assert np.array_equal(golden[name + "/clipped_actions"], np.clip(acts, -1, 1))


def test_golden_covers_the_branches(golden):
    assert golden["clipped/done"].sum() >= 1                 # -5 penalty + double reset path
    assert golden["terminate_q9/done"].sum() >= 10           # |q9| > 0.5 (SnakeGymEnv.py:100)
    assert np.abs(golden["terminate_q9/obs"][:, 9]).max() < 0.5   # ... and every returned observation is the post-reset / unterminated one
    # checkSnakeHeight (snake.py:237-245): break after the first tick, break in the middle of the loop, zero-tick termination
    lt, ld = golden["lifted/ticks"], golden["lifted/done"]
    assert lt[0] == 1 and ld[0] and 1 < lt[3] < 15 and ld[3] and lt[5] == 0 and ld[5] and ld.sum() == 3
    assert (golden["const_half/ticks"] == 0).any()           # Q5 zero-tick step
    assert golden["random/ticks"].max() <= 41                # `counter > 40` cap


def test_tick_cap(model):
    """`if self.counter > 40: break` (snake.py:303): unreachable with kp=0.1 and |a|<=1 (0.9^41 * 2.96 < 0.05),
    so exercise the cap through the parameter."""
    o = Oracle(1, default_params(max_ticks=5), model); o.reset()
    _, _, _, tk = o.step(np.ones((1, 8)))
    assert tk[0] == 5


def test_zero_tick_step_and_reward_terms(model):
    o = Oracle(1, default_params(), model); o.reset()
    a = np.full((1, 8), 0.5)
    o.step(a)
    ob, r, d, tk = o.step(a)                                  # already within 0.05: no tick (snake.py:283-284)
    assert tk[0] == 0 and not d[0]
    energy = np.sum(ob[0, 16:32] * ob[0, 32:48] * 0.01)
    assert r[0] == pytest.approx(-0.01 * abs(ob[0, 49]) - 0.1 * energy, abs=1e-15)


def test_done_returns_post_reset_obs_with_stale_torques(model):
    rng = np.random.default_rng(11)
    o = Oracle(4, default_params(), model); o.reset()
    seen = False
    for _ in range(60):
        before = o.get_state()
        ob, r, d, tk = o.step(rng.uniform(-1, 1, (4, 8)) * 3)     # clipped to +-1 => |q9| can exceed 0.5
        for e in np.where(d)[0]:
            seen = True
            assert np.allclose(ob[e, 0:32], 0) and np.allclose(ob[e, 48:51], 0) and np.allclose(ob[e, 51:55], [0, 0, 0, 1])
            assert np.abs(ob[e, 32:48]).max() > 0                  # Q9: stale applied torques survive the reset
            assert r[e] < -4.0                                      # -5 penalty added
            assert o.get_state()[e, 63] == 0                        # episode length restarted
    assert seen
    z = Oracle(1, default_params(stale_obs_on_reset=0), model); z.reset()
    z.step(np.ones((1, 8))); assert np.allclose(z.reset()[0, 32:48], 0)


def test_gait_selection_maps_actions(model):
    for gait, joints in ((0, range(0, 16, 2)), (1, range(1, 16, 2)), (2, range(16))):
        o = Oracle(1, default_params(gait_selection=gait), model); o.reset()
        assert o.act_dim == len(list(joints))
        ob, *_ = o.step(np.full((1, o.act_dim), 0.6))
        moved = np.abs(ob[0, :16]) > 0.1
        assert set(np.where(moved)[0]) == set(joints)


def test_masked_reset_and_counters(model):
    o = Oracle(3, default_params(), model); o.reset()
    o.step(np.full((3, 8), 0.7))
    obs = o.reset(mask=[0, 1, 0])
    assert np.allclose(obs[1, :32], 0) and np.abs(obs[0, :16]).max() > 0.1 and np.abs(obs[2, :16]).max() > 0.1
    c = o.counters()
    assert c["ticks"] > 0 and c["pgs_iterations"] >= c["ticks"]
