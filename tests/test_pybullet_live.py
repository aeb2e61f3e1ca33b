"""Import-guarded comparison with a live PyBullet.  PyBullet is not installable in the
authoring image or on the GPU box, so this module is SKIPPED there; it activates wherever `pybullet` and
`pybullet_data` import, and is the first thing to run when they do: it is what would turn the oracle's
"parity unpinned" into a pinned statement.

What it checks (CPU oracle vs PyBullet DIRECT, same URDF, same settings as snake.py:86-107):
  * the motor law: after one stepSimulation with POSITION_CONTROL targets and force = inf the joints have
    closed 10 % of their error (the closed form of SURVEY A.4) -- independent of the contact model;
  * env.step tick counts for seeded actions (a function of the joints only);
  * base pose after one env-step within a loose tolerance (contact model deviations D1-D3).
"""
import os

import numpy as np
import pytest

pybullet = pytest.importorskip("pybullet")
pybullet_data = pytest.importorskip("pybullet_data")

from bullet_envs_b200 import default_params  # noqa: E402
from bullet_envs_b200.urdf_model import DEFAULT_URDF, build_model  # noqa: E402
from oracle.oracle_py import Oracle  # noqa: E402


@pytest.fixture()
def world():
    p = pybullet
    cid = p.connect(p.DIRECT)
    p.setAdditionalSearchPath(pybullet_data.getDataPath())
    p.resetSimulation()
    p.setGravity(0, 0, -9.8)                                     # snake.py:8,91
    p.loadURDF("plane.urdf")                                     # snake.py:92
    snake = p.loadURDF(DEFAULT_URDF, [0, 0, 0], useFixedBase=0, flags=p.URDF_USE_SELF_COLLISION)  # snake.py:93
    for j in range(-1, p.getNumJoints(snake)):                   # snake.py:103-107
        p.changeDynamics(snake, j, lateralFriction=2, anisotropicFriction=[1, 0.1, 0.01])
    yield p, snake
    p.disconnect(cid)


def test_motor_law_and_tick_counts(world):
    p, snake = world
    model = build_model()
    motors = model.motor_joint_indices
    rng = np.random.default_rng(0)
    o = Oracle(1, default_params()); o.reset()
    for step in range(5):
        a = rng.uniform(-1, 1, 8)
        tgt = np.zeros(16); tgt[1::2] = a * np.pi / 6
        ticks = 0
        while True:                                              # snake.py:284-304 without the sleep
            q = np.array([p.getJointState(snake, j)[0] for j in motors])
            if not np.linalg.norm(tgt - q) > 0.05 or ticks > 40:
                break
            p.setJointMotorControlArray(snake, motors, p.POSITION_CONTROL, targetPositions=list(tgt), forces=[np.inf] * 16)
            p.stepSimulation()
            q1 = np.array([p.getJointState(snake, j)[0] for j in motors])
            assert np.allclose(q1 - q, 0.1 * (tgt - q), atol=2e-3)   # closed form, SURVEY A.4
            ticks += 1
        _, _, _, ot = o.step(a[None, :])
        assert abs(int(ot[0]) - ticks) <= 1
    pos, _ = p.getBasePositionAndOrientation(snake)
    assert np.abs(np.array(pos) - o.observe()[0, 48:51]).max() < 0.05


def test_which_variant_matches_pybullet(world, capsys):
    """The two open questions of DESIGN.md section 4, answered wherever PyBullet imports: runs the same 10 env-steps in PyBullet and in
    the three variants of the oracle that have a CUDA twin -- one point per cylinder with the motor rows eliminated (the default), the same
    with persistent manifolds + warm start 0.1 (snk_set_manifold), and Bullet-order motor rows (motor_solver = 0) -- and prints the base-position and
    reward-relevant distances -- the combination with the smallest one is the configuration to ship as the default."""
    p, snake = world
    model = build_model()
    motors = model.motor_joint_indices
    rng = np.random.default_rng(1)
    acts = rng.uniform(-1, 1, (10, 8))
    ref = []
    for a in acts:
        tgt = np.zeros(16); tgt[1::2] = a * np.pi / 6
        ticks = 0
        while True:
            q = np.array([p.getJointState(snake, j)[0] for j in motors])
            if not np.linalg.norm(tgt - q) > 0.05 or ticks > 40:
                break
            p.setJointMotorControlArray(snake, motors, p.POSITION_CONTROL, targetPositions=list(tgt), forces=[np.inf] * 16)
            p.stepSimulation()
            ticks += 1
        pos, quat = p.getBasePositionAndOrientation(snake)
        ref.append((ticks, np.array(pos), np.array(quat)))
    rows = []
    for solver in (2, 0):
        for man in (False, True):
            if man and solver == 0:
                continue  # the oracle's manifolds live in the exact tick
            o = Oracle(1, default_params(motor_solver=solver))
            if man:
                o.set_manifold(True, 0.1)
            o.reset()
            dt, dp = 0, 0.0
            for a, (ticks, pos, quat) in zip(acts, ref):
                ob, _, _, ot = o.step(a[None, :])
                dt += abs(int(ot[0]) - ticks)
                dp = max(dp, float(np.abs(ob[0, 48:51] - pos).max()))
            rows.append(("motor rows %s, %s" % ("eliminated" if solver else "Bullet order", "manifolds + warm 0.1" if man else "one point per cylinder"), dt, dp))
    with capsys.disabled():
        print("\nvariant                                              sum |tick diff|   max base position diff over 10 env-steps [m]")
        for name, dt, dp in rows:
            print("%-55s %6d %12.4f" % (name, dt, dp))
    assert min(r[2] for r in rows) < 0.1
