"""URDF -> merged rigid-body tables for the snake chain (host side, numpy only).

Replaces what ``pybullet.loadURDF(snake.urdf, useFixedBase=0, flags=URDF_USE_SELF_COLLISION)``
does for the reference (``snake.py:93``): parse links/joints, number them depth-first the way
PyBullet does (so ``motorList = arange(3, numJoints, 3)``, ``snake.py:78-81``, still names the
revolute joints), apply Bullet's import rules for mass/inertia, then merge every fixed joint so
the dynamics see 17 rigid bodies in a serial chain with 16 revolute joints.

Bullet's import rules are recalled semantics and therefore each one is
a named switch of :class:`ImportRules`; the defaults are the PyBullet behaviour.
"""
from __future__ import annotations

import ctypes
import math
import os
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field

import numpy as np

NB, NJ, NC, NDOF, OBS_DIM, STATE_STRIDE = 17, 16, 32, 22, 56, 64

DEFAULT_URDF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets", "snake_chain.urdf")


@dataclass
class ImportRules:
    """Switches for the URDF import semantics of PyBullet."""
    unit_mass_for_missing_inertial: bool = True   # "No inertial data for link, using mass=1"
    inertia_from_collision_aabb: bool = True      # no URDF_USE_INERTIA_FROM_FILE flag
    collision_margin: float = 0.001               # gUrdfDefaultCollisionMargin
    hull_segments: int = 32                       # cylinder -> 32-gon hull (AABB only)
    contact_breaking_factor: float = 0.02         # gContactBreakingThreshold scale


def rpy_to_mat(r, p, y):
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


def _origin(elem):
    R, t = np.eye(3), np.zeros(3)
    if elem is not None:
        o = elem.find("origin")
        if o is not None:
            t = np.array([float(v) for v in o.get("xyz", "0 0 0").split()])
            R = rpy_to_mat(*[float(v) for v in o.get("rpy", "0 0 0").split()])
    return R, t


@dataclass
class UrdfLink:
    name: str
    has_inertial: bool
    mass: float
    com_R: np.ndarray
    com_t: np.ndarray
    inertia_file: np.ndarray
    collisions: list  # (R, t, radius, length)


@dataclass
class UrdfJoint:
    name: str
    jtype: str
    parent: str
    child: str
    R: np.ndarray
    t: np.ndarray
    axis: np.ndarray
    damping: float
    lower: float
    upper: float
    effort: float


def parse_urdf(path):
    """Return (links: dict name->UrdfLink, joints: list[UrdfJoint] in file order)."""
    root = ET.parse(path).getroot()
    links, joints = {}, []
    for le in root.findall("link"):
        ie = le.find("inertial")
        if ie is not None:
            cR, ct = _origin(ie)
            m = float(ie.find("mass").get("value"))
            ii = ie.find("inertia")
            g = lambda k: float(ii.get(k, "0"))
            I = np.array([[g("ixx"), g("ixy"), g("ixz")], [g("ixy"), g("iyy"), g("iyz")], [g("ixz"), g("iyz"), g("izz")]])
        else:
            cR, ct, m, I = np.eye(3), np.zeros(3), 0.0, np.zeros((3, 3))
        cols = []
        for ce in le.findall("collision"):
            R, t = _origin(ce)
            cyl = ce.find("geometry").find("cylinder")
            if cyl is None:
                raise ValueError("only <cylinder> collision geometry is supported (link %s)" % le.get("name"))
            cols.append((R, t, float(cyl.get("radius")), float(cyl.get("length"))))
        links[le.get("name")] = UrdfLink(le.get("name"), ie is not None, m, cR, ct, I, cols)
    for je in root.findall("joint"):
        if je.find("parent") is None:
            continue  # <joint> inside <transmission>
        R, t = _origin(je)
        ax = je.find("axis")
        axis = np.array([float(v) for v in ax.get("xyz").split()]) if ax is not None else np.array([1.0, 0, 0])
        dyn, lim = je.find("dynamics"), je.find("limit")
        joints.append(UrdfJoint(je.get("name"), je.get("type"), je.find("parent").get("link"),
                                je.find("child").get("link"), R, t, axis,
                                float(dyn.get("damping", "0")) if dyn is not None else 0.0,
                                float(lim.get("lower", "0")) if lim is not None else 0.0,
                                float(lim.get("upper", "0")) if lim is not None else 0.0,
                                float(lim.get("effort", "0")) if lim is not None else 0.0))
    return links, joints


def dfs_order(links, joints):
    """PyBullet numbering: joint/link index i = i-th joint met in a depth-first preorder walk
    from the root link, children in URDF file order.  Returns (root_name, [joint, ...])."""
    children = {n: [] for n in links}
    has_parent = set()
    for j in joints:
        children[j.parent].append(j)
        has_parent.add(j.child)
    roots = [n for n in links if n not in has_parent]
    if len(roots) != 1:
        raise ValueError("expected exactly one root link, got %r" % roots)
    order = []

    def walk(name):
        for j in children[name]:
            order.append(j)
            walk(j.child)

    import sys
    sys.setrecursionlimit(max(10000, sys.getrecursionlimit()))
    walk(roots[0])
    return roots[0], order


def _hull_aabb_inertia(mass, cols, com_R, com_t, rules):
    """btCompoundShape::calculateLocalInertia: box inertia of the compound's AABB, taken in the
    inertial frame, hull vertices + margin, AABB centre offset ignored (SURVEY A.1 [M])."""
    if not cols:
        return np.zeros((3, 3))
    lo, hi = np.full(3, np.inf), np.full(3, -np.inf)
    n = rules.hull_segments
    for (R, t, rad, length) in cols:
        ang = 2 * math.pi * np.arange(n) / n
        v = np.stack([rad * np.sin(ang), rad * np.cos(ang), np.full(n, 0.5 * length)], 1)
        v = np.concatenate([v, v * np.array([1, 1, -1.0])], 0)
        vw = (com_R.T @ ((R @ v.T).T + t - com_t).T).T  # into the inertial frame
        lo = np.minimum(lo, vw.min(0) - rules.collision_margin)
        hi = np.maximum(hi, vw.max(0) + rules.collision_margin)
    lx, ly, lz = hi - lo
    return np.diag([mass / 12.0 * (ly * ly + lz * lz), mass / 12.0 * (lx * lx + lz * lz), mass / 12.0 * (lx * lx + ly * ly)])


@dataclass
class SnakeModel:
    joint_R0: np.ndarray
    joint_t: np.ndarray
    joint_axis: np.ndarray
    joint_damping: np.ndarray
    body_mass: np.ndarray
    body_com: np.ndarray
    body_inertia: np.ndarray
    cyl_center: np.ndarray
    cyl_axis: np.ndarray
    cyl_fric_R: np.ndarray
    cyl_radius: np.ndarray
    cyl_halflen: np.ndarray
    cyl_end: np.ndarray
    cyl_margin: np.ndarray
    cyl_break: np.ndarray
    height_pt: np.ndarray
    fz_axis: np.ndarray
    root_mass: float
    cyl_body: np.ndarray
    height_body: np.ndarray
    # bookkeeping (not part of the C struct)
    motor_joint_indices: list = field(default_factory=list)  # PyBullet joint indices of the motors
    num_urdf_links: int = 0
    num_urdf_joints: int = 0
    link_names: list = field(default_factory=list)
    link_body: list = field(default_factory=list)      # merged body of each URDF link (DFS order, root first)
    link_pose_in_body: list = field(default_factory=list)
    joint_limits: np.ndarray = None

    def to_ctypes(self):
        m = CModel()
        for name, _ in CModel._fields_:
            v = getattr(self, name)
            dst = getattr(m, name)
            if isinstance(v, float):
                setattr(m, name, v)
            else:
                flat = np.ascontiguousarray(v).reshape(-1)
                flat = np.ascontiguousarray(flat, dtype=np.int32 if flat.dtype.kind in "iu" else np.float64)
                if flat.nbytes != ctypes.sizeof(dst):
                    raise ValueError("model field %s has %d bytes, ABI expects %d" % (name, flat.nbytes, ctypes.sizeof(dst)))
                ctypes.memmove(ctypes.addressof(dst), flat.ctypes.data, flat.nbytes)
        return m


class CModel(ctypes.Structure):
    """ctypes mirror of ``snk_model`` (include/snake_b200.h)."""
    _fields_ = [
        ("joint_R0", ctypes.c_double * (NJ * 9)), ("joint_t", ctypes.c_double * (NJ * 3)),
        ("joint_axis", ctypes.c_double * (NJ * 3)), ("joint_damping", ctypes.c_double * NJ),
        ("body_mass", ctypes.c_double * NB), ("body_com", ctypes.c_double * (NB * 3)),
        ("body_inertia", ctypes.c_double * (NB * 9)),
        ("cyl_center", ctypes.c_double * (NC * 3)), ("cyl_axis", ctypes.c_double * (NC * 3)),
        ("cyl_fric_R", ctypes.c_double * (NC * 9)), ("cyl_radius", ctypes.c_double * NC),
        ("cyl_halflen", ctypes.c_double * NC), ("cyl_end", ctypes.c_double * NC),
        ("cyl_margin", ctypes.c_double * NC), ("cyl_break", ctypes.c_double * NC),
        ("height_pt", ctypes.c_double * (NB * 3)), ("fz_axis", ctypes.c_double * 3),
        ("root_mass", ctypes.c_double),
        ("cyl_body", ctypes.c_int32 * NC), ("height_body", ctypes.c_int32 * NB),
    ]


def build_model(urdf_path=None, rules: ImportRules | None = None) -> SnakeModel:
    """Parse ``urdf_path`` and return the merged tables the simulator consumes."""
    rules = rules or ImportRules()
    urdf_path = urdf_path or DEFAULT_URDF
    links, joints = parse_urdf(urdf_path)
    root, order = dfs_order(links, joints)

    # --- per-URDF-link mass properties under Bullet's import rules -------------------------
    def link_props(L):
        if L.has_inertial:
            m, cR, ct = L.mass, L.com_R, L.com_t
        elif rules.unit_mass_for_missing_inertial and L.name != "world":
            m, cR, ct = 1.0, np.eye(3), np.zeros(3)
        else:
            m, cR, ct = 0.0, np.eye(3), np.zeros(3)
        if rules.inertia_from_collision_aabb:
            I = _hull_aabb_inertia(m, L.collisions, cR, ct, rules)
        else:
            I = L.inertia_file if L.has_inertial else (np.eye(3) if m > 0 else np.zeros((3, 3)))
        return m, cR, ct, I

    # --- walk the tree, merging fixed joints --------------------------------------------
    body_of = {root: 0}
    pose_in_body = {root: (np.eye(3), np.zeros(3))}
    body_parent, jR0, jt, jaxis, jdamp, jlim = [-1], [], [], [], [], []
    motor_idx = []
    link_seq = [root]
    for idx, j in enumerate(order):
        pR, pt = pose_in_body[j.parent]
        R, t = pR @ j.R, pt + pR @ j.t
        if j.jtype == "fixed":
            body_of[j.child] = body_of[j.parent]
            pose_in_body[j.child] = (R, t)
        elif j.jtype in ("revolute", "continuous"):
            b = len(body_parent)
            if body_of[j.parent] != b - 1:
                raise ValueError("the simulator supports a serial chain only (joint %s branches)" % j.name)
            body_parent.append(body_of[j.parent])
            body_of[j.child] = b
            pose_in_body[j.child] = (np.eye(3), np.zeros(3))
            jR0.append(R); jt.append(t); jaxis.append(j.axis / np.linalg.norm(j.axis))
            jdamp.append(j.damping); jlim.append((j.lower, j.upper, j.effort))
            motor_idx.append(idx)
        else:
            raise ValueError("unsupported joint type %r" % j.jtype)
        link_seq.append(j.child)
    nb = len(body_parent)
    if nb != NB:
        raise ValueError("kernels are compiled for %d bodies, URDF merges to %d" % (NB, nb))

    # --- accumulate mass properties per merged body ----------------------------------------
    mass = np.zeros(nb); first = np.zeros((nb, 3)); Io = np.zeros((nb, 3, 3))
    cyl = []
    for name in link_seq:
        L = links[name]
        b = body_of[name]
        R, t = pose_in_body[name]
        m, cR, ct, I = link_props(L)
        c = t + R @ ct                      # link COM in body frame
        Rc = R @ cR
        mass[b] += m
        first[b] += m * c
        Io[b] += Rc @ I @ Rc.T + m * ((c @ c) * np.eye(3) - np.outer(c, c))
        for (cRot, ctr, rad, length) in L.collisions:
            cyl.append(dict(body=b, center=t + R @ ctr, axis=(R @ cRot)[:, 2], fricR=Rc, radius=rad,
                            halflen=0.5 * length))
    com = first / mass[:, None]
    Ic = np.stack([Io[b] - mass[b] * ((com[b] @ com[b]) * np.eye(3) - np.outer(com[b], com[b])) for b in range(nb)])
    if len(cyl) != NC:
        raise ValueError("kernels are compiled for %d collision cylinders, URDF has %d" % (NC, len(cyl)))

    # --- contact rim selection: the end of each cylinder that is an extreme of its body -----
    jt_arr = np.array(jt)
    for i, c in enumerate(cyl):
        others = [o for k, o in enumerate(cyl) if o["body"] == c["body"] and k != i]
        if others:
            ref = np.mean([o["center"] for o in others], 0)
        elif c["body"] + 1 < nb:
            ref = jt_arr[c["body"]]          # origin of the joint to the next body, this body's frame
        else:
            ref = np.zeros(3)                # last body: its own joint sits at the frame origin
        s = float(np.dot(c["center"] - ref, c["axis"]))
        c["end"] = 1.0 if s >= 0 else -1.0
        ext = np.array([c["radius"], c["radius"], c["halflen"]]) + rules.collision_margin
        c["break"] = rules.contact_breaking_factor * float(np.linalg.norm(ext))

    # --- checkSnakeHeight points: COM of URDF links 0,3,6,...,48 (snake.py:237-245) ---------
    hb, hp = [], []
    for li in range(0, len(order), 3):
        name = order[li].child
        L = links[name]
        R, t = pose_in_body[name]
        _, cR, ct, _ = link_props(L)
        hb.append(body_of[name]); hp.append(t + R @ ct)
    if len(hb) != NB:
        raise ValueError("expected %d height points, got %d" % (NB, len(hb)))

    link0 = order[0].child
    R0, _ = pose_in_body[link0]
    root_mass = link_props(links[root])[0]

    return SnakeModel(
        joint_R0=np.array(jR0).reshape(NJ, 9), joint_t=jt_arr, joint_axis=np.array(jaxis),
        joint_damping=np.array(jdamp), body_mass=mass, body_com=com, body_inertia=Ic.reshape(nb, 9),
        cyl_center=np.array([c["center"] for c in cyl]), cyl_axis=np.array([c["axis"] for c in cyl]),
        cyl_fric_R=np.array([c["fricR"].reshape(9) for c in cyl]),
        cyl_radius=np.array([c["radius"] for c in cyl]), cyl_halflen=np.array([c["halflen"] for c in cyl]),
        cyl_end=np.array([c["end"] for c in cyl]),
        cyl_margin=np.full(NC, rules.collision_margin), cyl_break=np.array([c["break"] for c in cyl]),
        height_pt=np.array(hp), fz_axis=R0[:, 2].copy(), root_mass=float(root_mass),
        cyl_body=np.array([c["body"] for c in cyl], dtype=np.int32), height_body=np.array(hb, dtype=np.int32),
        motor_joint_indices=motor_idx, num_urdf_links=len(links), num_urdf_joints=len(order),
        link_names=link_seq, link_body=[body_of[n] for n in link_seq],
        link_pose_in_body=[pose_in_body[n] for n in link_seq], joint_limits=np.array(jlim))
