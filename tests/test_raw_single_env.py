"""Raw single-env semantics: bullet_envs_b200.gym_env.RawSingleEnv turns a batch of one with vector-wrapper
semantics into the unwrapped SnakeGymEnv.step behaviour (terminal observation on done, the dead episode's x as x_prev of the next
reward).  Here the batch of one is the CPU oracle, and the expected values come from the reference's own Python run WITHOUT the
wrapper (tests/golden/reference_python_raw_single_env.npz, made by tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from bullet_envs_b200 import default_params
from bullet_envs_b200.gym_env import RawSingleEnv
from oracle.oracle_py import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBatchOfOne:
    def __init__(self, model):
        self.o = Oracle(1, default_params(), model)
        self.o.reset()

    def observe(self):
        return self.o.observe()

    def step_traced(self, a):
        obs, r, d, tk, tobs, _ = self.o.step_trace(a)
        return obs, r, d, (tobs[0, tk[0] - 1] if tk[0] > 0 else None)


@pytest.mark.parametrize("name", ["serpenoid", "clipped"])
def test_raw_semantics_reproduce_the_unwrapped_reference(model, name):
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_python_raw_single_env.npz"))
    env = RawSingleEnv(OracleBatchOfOne(model), alpha=1.0)
    env.reset()
    assert g[name + "/done"].sum() >= 3
    for t, a in enumerate(g[name + "/actions"]):
        ob, r, d = env.step(np.asarray(a)[None, :])
        assert d == bool(g[name + "/done"][t]), (name, t)
        assert np.allclose(ob, g[name + "/obs"][t + 1], rtol=0, atol=1e-12), (name, t)
        assert r == pytest.approx(g[name + "/rew"][t], abs=1e-12), (name, t)


def test_raw_and_wrapper_semantics_differ_only_around_done(model, golden):
    """Same scenario through the wrapper (golden of tests/test_oracle_task_logic.py) and raw: rewards agree except on the step
    after a done, where the raw reward carries the dead episode's x (Q8)."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_python_raw_single_env.npz"))
    rw, rr, dn = golden["serpenoid/rew"], g["serpenoid/rew"], golden["serpenoid/done"]
    assert np.array_equal(dn, g["serpenoid/done"])
    after = np.zeros_like(dn); after[1:] = dn[:-1]
    assert np.allclose(rw[~after], rr[~after], atol=1e-12)
    x_terminal = g["serpenoid/obs"][1:][dn, 48]                      # terminal observations are only visible in the raw run
    assert np.allclose(rr[after] - rw[after], -x_terminal[:after.sum()], atol=1e-12)
