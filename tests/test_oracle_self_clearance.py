"""Self-collision clearance counter: the oracle's lower bound vs brute-force sampled cylinder
distances, and the claim it exists to prove -- at the joint angles the task reaches, no pair of cylinders that Bullet would test
under URDF_USE_SELF_COLLISION (snake.py:93) comes near contact."""
import numpy as np

from bullet_envs_b200 import default_params
from oracle.oracle_py import Oracle


def cylinders_world(o, e):
    m = o.model
    Rw, pw, _ = o.kinematics(e)
    c = np.array([pw[m.cyl_body[k]] + Rw[m.cyl_body[k]] @ m.cyl_center[k] for k in range(32)])
    a = np.array([Rw[m.cyl_body[k]] @ m.cyl_axis[k] for k in range(32)])
    return c, a


def surface_points(c, a, h, r, n_ang=40, n_len=7, n_rad=3):
    """Points on the surface of a finite cylinder: mantle + both caps."""
    u = np.cross(a, [1.0, 0.3, 0.2]); u /= np.linalg.norm(u)
    v = np.cross(a, u)
    th = np.linspace(0, 2 * np.pi, n_ang, endpoint=False)
    ring = np.cos(th)[:, None] * u + np.sin(th)[:, None] * v
    pts = [c + a * z + r * ring for z in np.linspace(-h, h, n_len)]
    for z in (-h, h):
        pts += [c + a * z + rr * ring for rr in np.linspace(0, r, n_rad, endpoint=False)]
    return np.concatenate(pts)


def sampled_min_distance(o, e):
    m = o.model
    c, a = cylinders_world(o, e)
    P = [surface_points(c[k], a[k], m.cyl_halflen[k], m.cyl_radius[k]) for k in range(32)]
    best = np.inf
    for i in range(30):
        for j in range(i + 2, min(i + 5, 32)):       # only near neighbours can be the minimum on an open chain pose
            d = np.linalg.norm(P[i][:, None, :] - P[j][None, :, :], axis=2).min()
            best = min(best, d)
    return best


def test_rest_pose_clearance_is_the_axial_gap(model):
    o = Oracle(1, default_params(), model); o.reset()
    # INPUT_k ends 0.0348 above its frame, INPUT_{k+1} starts 0.0366 + 0.0273 + 0.0018 above it (snake.urdf:806-811,833-836,874-878)
    assert abs(o.self_clearance()[0] - (0.0366 + 0.0273 + 0.0018 - 0.0348)) < 1e-9


def test_bound_never_exceeds_the_sampled_distance(model):
    rng = np.random.default_rng(4)
    n = 4
    o = Oracle(n, default_params(), model); o.reset()
    s = o.get_state()
    s[:, 13:29] = rng.uniform(-0.9, 0.9, (n, 16))      # well beyond the reachable +-pi/6, so that some bounds go near zero
    s[0, 13:29] = 0.0
    o.set_state(s)
    clr = o.self_clearance()
    for e in range(n):
        d = sampled_min_distance(o, e)
        assert clr[e] <= d + 1e-9, (e, clr[e], d)       # a lower bound ...
        assert clr[e] >= d - 0.012, (e, clr[e], d)      # ... and not a vacuous one (sampling itself over-estimates by ~2 mm)


def test_reachable_joint_range_keeps_every_pair_apart(model):
    """|q| <= pi/6 on every joint (the action bound times SCALING_FACTOR, snake.py:41,223-225): the closest non-consecutive
    cylinders stay >= 17 mm apart, so Bullet's self-collision pairs never produce a contact and omitting them (D3) is exact."""
    rng = np.random.default_rng(0)
    n = 256
    o = Oracle(n, default_params(), model); o.reset()
    s = o.get_state()
    q = rng.uniform(-np.pi / 6, np.pi / 6, (n, 16))
    q[0] = np.pi / 6; q[1] = -np.pi / 6; q[2] = np.where(np.arange(16) % 2 == 1, np.pi / 6, 0.0)     # the extreme curls
    q[3] = np.where(np.arange(16) % 2 == 1, np.pi / 6, -np.pi / 6)
    s[:, 13:29] = q
    o.set_state(s)
    assert o.self_clearance().min() > 0.017
    # ... and the bound keeps shrinking as one joint bends on towards its +-1.57 limit, where the rims would meet
    for e, ang in enumerate((0.3, 0.6, 0.9, 1.2, 1.5)):
        s[e, 13:29] = 0.0; s[e, 13 + 5] = ang
    o.set_state(s)
    c = o.self_clearance()[:5]
    assert (np.diff(c) < 0).all() and c[3] < 0.006 and c[4] < 0.004


def test_clearance_over_a_random_rollout(model):
    rng = np.random.default_rng(2)
    n = 48
    o = Oracle(n, default_params(), model); o.reset()
    worst = np.inf
    for _ in range(6):
        o.step(rng.uniform(-1, 1, (n, 8)), threads=8)
        worst = min(worst, o.self_clearance().min())
    assert worst > 0.017
