"""mode='test' info stream: the oracle's step_trace vs what the reference's own Python records in
``info['internal_observations']`` / ``info['link_positions']`` (snake.py:275-278,292-293; SnakeGymEnv.py:43-44), golden vectors
made by tests/golden/make_golden.py from the unmodified SnakeGymEnv.py / snake.py."""
import numpy as np
import pytest

from bullet_envs_b200 import default_params
from oracle.oracle_py import Oracle


@pytest.mark.parametrize("name", ["serpenoid", "terminate_q9"])
def test_step_trace_reproduces_reference_python(golden_test_mode, model, name):
    g = golden_test_mode
    acts, ticks = g[name + "/actions"], g[name + "/ticks"]
    io, lp = g[name + "/internal_observations"], g[name + "/link_positions"]
    assert io.shape == (ticks.sum(), 56) and lp.shape == (ticks.sum(), 51)
    o = Oracle(1, default_params(), model)
    o.reset()
    row = 0
    for t, a in enumerate(acts):
        ob, r, d, tk, tobs, tlnk = o.step_trace(a[None, :])
        assert tk[0] == ticks[t]
        k = int(tk[0])
        assert np.allclose(tobs[0, :k], io[row:row + k], rtol=0, atol=1e-12), (name, t)
        assert np.allclose(tlnk[0, :k], lp[row:row + k], rtol=0, atol=1e-12), (name, t)
        assert np.isnan(tobs[0, k:]).all() and np.isnan(tlnk[0, k:]).all()      # rows past the last tick are not written
        assert np.allclose(ob[0], g[name + "/obs"][t], rtol=0, atol=1e-12)
        assert r[0] == pytest.approx(g[name + "/rew"][t], abs=1e-12) and bool(d[0]) == bool(g[name + "/done"][t])
        row += k
    assert row == ticks.sum()


def test_step_trace_is_step(model):
    """The traced step changes nothing: same outputs and state as the plain step; the last traced row of an env-step that
    did not end an episode is the returned observation."""
    rng = np.random.default_rng(5)
    a, b = Oracle(6, default_params(), model), Oracle(6, default_params(), model)
    a.reset(); b.reset()
    for _ in range(4):
        act = rng.uniform(-1, 1, (6, 8))
        o1, r1, d1, t1 = a.step(act)
        o2, r2, d2, t2, tobs, tlnk = b.step_trace(act)
        assert np.array_equal(o1, o2) and np.array_equal(r1, r2) and np.array_equal(d1, d2) and np.array_equal(t1, t2)
        assert np.array_equal(a.get_state(), b.get_state())
        for e in range(6):
            if t2[e] > 0 and not d2[e]:
                assert np.array_equal(tobs[e, t2[e] - 1], o2[e])
                # link 0 (`base`) sits at the base origin + R0 * hpt[0]; its z stays near the cylinder radius on the ground
                assert 0.0 < tlnk[e, t2[e] - 1, 34] < 0.1
