"""SURVEY.md 8f rank 2 in the oracle (test infrastructure only; the kernels implement deviations D1 / D3): Bullet-style persistent
contact manifolds -- support vertex of the 32-gon hull, 4 cached points per cylinder, Bullet's refresh / breaking rule, optional
warm start x 0.1 -- and joint-limit rows, each behind a switch, measured against the default contact model."""
import numpy as np
import pytest

from bullet_envs_b200 import default_params
from oracle.oracle_py import Oracle


def play(golden, name, manifold, warm=0.0):
    o = Oracle(1, default_params())
    if manifold:
        o.set_manifold(True, warm)
    o.reset()
    ret, ticks, dones = 0.0, [], 0
    for a in golden[name + "/actions"]:
        ob, r, d, tk = o.step(a[None, :])
        ret += float(r[0]); ticks.append(int(tk[0])); dones += int(d[0])
    pts = o.row_stats()[0]
    return ret, np.array(ticks), dones, pts


@pytest.mark.parametrize("name", ["random", "serpenoid", "clipped", "terminate_q9"])
def test_manifold_switch_on_the_golden_scenarios(golden, name):
    """The task's integer outputs do not depend on the contact model (the joints are prescribed by the motor law): tick counts and
    episode ends are identical; the return moves by a few per cent (measured: random -4.83 -> -4.72, serpenoid -14.80 -> -14.68)."""
    r0, t0, d0, _ = play(golden, name, False)
    assert r0 == pytest.approx(float(golden[name + "/rew"].sum()), abs=1e-9)           # switch off = the pinned default
    for warm in (0.0, 0.1):
        r1, t1, d1, pts = play(golden, name, True, warm)
        assert np.array_equal(t0, t1) and d0 == d1
        assert abs(r1 - r0) <= 0.05 * abs(r0) + 0.02, (name, warm, r0, r1)
        assert 20 <= pts <= 4 * 32                                                     # ~1 point per cylinder while moving, up to 4 at rest


def test_manifold_accumulates_points_at_rest_and_breaks_them_when_sliding():
    o = Oracle(1, default_params()); o.set_manifold(True, 0.0); o.reset()
    o.tick(np.zeros((1, 16)), 1)
    assert o.row_stats()[0] == 32                                                      # first tick: one support vertex per cylinder
    o.tick(np.zeros((1, 16)), 150)
    assert o.row_stats()[0] > 45                                                       # the caches fill up while the chain settles on its facets
    s = o.get_state()
    assert np.isfinite(s).all() and abs(s[0, 2]) < 3e-3                                # ... on the plane (margin 1 mm)
    o2 = Oracle(1, default_params()); o2.set_manifold(True, 0.0); o2.reset()
    tg = np.zeros((1, 16)); tg[0, 1::2] = 0.5
    o2.tick(tg, 60)
    assert o2.row_stats()[0] < 45                                                      # sliding contacts drift off their anchors (0.84 mm) and are dropped


def test_manifold_batch_statistics_against_the_default_contact_model():
    n, steps = 128, 6
    rng = np.random.default_rng(0)
    acts = rng.uniform(-1, 1, (steps, n, 8))
    out = {}
    for label, on in (("d1", False), ("manifold", True)):
        o = Oracle(n, default_params())
        if on:
            o.set_manifold(True, 0.1)
        o.reset()
        R = np.zeros(n); T = []
        for t in range(steps):
            ob, r, d, tk = o.step(acts[t], threads=8); R += r; T.append(tk)
        out[label] = (R, np.array(T), ob)
    assert (out["d1"][1] == out["manifold"][1]).mean() > 0.99
    assert np.median(np.abs(out["d1"][2][:, 48:51] - out["manifold"][2][:, 48:51]).max(1)) < 0.05   # base within centimetres after 6 env-steps
    ev = lambda R: R[R > -4]
    assert abs(ev(out["d1"][0]).mean() - ev(out["manifold"][0]).mean()) < 0.05


@pytest.mark.parametrize("name", ["random", "clipped"])
def test_joint_limit_rows_never_activate(golden, name):
    """btMultiBodyJointLimitConstraint rows at +-1.57 rad in the Bullet-order tick: with |target| <= pi/6 they are speculative rows with a
    1 rad gap, their impulses stay exactly zero and every output is bit-identical with the rows switched off (deviation D3 is exact)."""
    outs = []
    for on in (False, True):
        o = Oracle(1, default_params(motor_solver=0))
        if on:
            o.set_joint_limits(True)
        o.reset()
        rec = []
        for a in golden[name + "/actions"][:6]:
            rec.append(np.concatenate([o.step(a[None, :])[0][0], o.get_state()[0]]))
        outs.append(np.array(rec))
        if on:
            assert o.row_stats()[1] == 0
    assert np.array_equal(outs[0], outs[1])
    # and the rows do act when a limit is within reach: a limit at 0.3 rad stops a joint commanded to 0.5
    o = Oracle(1, default_params(motor_solver=0)); o.set_joint_limits(True, 0.3); o.reset()
    a = np.zeros((1, 8)); a[0, 2] = 1.0
    for _ in range(3):
        ob, _, _, _ = o.step(a)
    assert o.row_stats()[1] > 0 and ob[0, 5] < 0.4      # (the unlimited-force motor row pushes against the limit row: the sweeps settle in between)
