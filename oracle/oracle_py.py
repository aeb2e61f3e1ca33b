"""ctypes front-end of the CPU oracle (oracle/snake_oracle.c).  TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; nothing under bullet_envs_b200/ imports this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from bullet_envs_b200._abi import CParams, default_params  # noqa: E402  (ABI structs only)
from bullet_envs_b200.urdf_model import CModel, build_model, NJ, OBS_DIM, STATE_STRIDE  # noqa: E402

LIB = os.path.join(_HERE, "_ref", "libsnake_oracle.so")


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(_HERE, "snake_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return LIB


_libs = {}


def _load():
    if "lib" not in _libs:
        if not os.path.exists(LIB):
            build()
        _libs["lib"] = ctypes.CDLL(LIB)
    return _libs["lib"]


class Oracle:
    """N independent snake environments stepped on the CPU in fp64 (or fp32 with ``f32=True``)."""

    def __init__(self, n_envs, params: CParams | None = None, model=None, f32=False):
        lib = _load()
        self.pfx = "snk_cpu32_" if f32 else "snk_cpu_"
        self.dtype = np.float32 if f32 else np.float64
        self.n = int(n_envs)
        self.params = params or default_params()
        self.model = model or build_model()
        self._cm = self.model.to_ctypes()
        self._h = ctypes.c_void_p()
        f = self._fn("create")
        f.argtypes = [ctypes.POINTER(CModel), ctypes.POINTER(CParams), ctypes.c_int64, ctypes.POINTER(ctypes.c_void_p)]
        rc = f(ctypes.byref(self._cm), ctypes.byref(self.params), self.n, ctypes.byref(self._h))
        if rc != 0:
            raise RuntimeError("oracle create failed: %d" % rc)
        self.act_dim = self._fn("action_dim")(self._h)
        for name in ("reset", "step", "step_range", "tick", "observe", "get_state", "set_state", "counters", "kinematics", "destroy"):
            getattr(lib, self.pfx + name).restype = ctypes.c_int

    def _fn(self, name):
        return getattr(_load(), self.pfx + name)

    @staticmethod
    def _p(a):
        return None if a is None else ctypes.c_void_p(a.ctypes.data)

    def close(self):
        if self._h:
            self._fn("destroy")(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self, mask=None):
        obs = np.empty((self.n, OBS_DIM), self.dtype)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        self._fn("reset")(self._h, self._p(m), self._p(obs))
        return obs

    def step(self, actions, threads=1):
        a = np.ascontiguousarray(actions, self.dtype).reshape(self.n, self.act_dim)
        obs = np.empty((self.n, OBS_DIM), self.dtype)
        rew = np.empty(self.n, self.dtype)
        done = np.empty(self.n, np.uint8)
        ticks = np.empty(self.n, np.int32)
        if threads <= 1:
            self._fn("step")(self._h, self._p(a), self._p(obs), self._p(rew), self._p(done), self._p(ticks))
        else:
            f = self._fn("step_range")
            bounds = np.linspace(0, self.n, threads + 1).astype(np.int64)
            with ThreadPoolExecutor(threads) as ex:
                list(ex.map(lambda i: f(self._h, ctypes.c_int64(int(bounds[i])), ctypes.c_int64(int(bounds[i + 1])),
                                        self._p(a), self._p(obs), self._p(rew), self._p(done), self._p(ticks)),
                            range(threads)))
        return obs, rew, done.astype(bool), ticks

    def step_trace(self, actions):
        """Twin of ``snk_step_trace`` (mode='test' info stream, snake.py:275-278,292-293): ``step`` plus the per-tick
        observations [n, max_ticks, 56] and link positions [n, max_ticks, 51]; rows >= ticks[e] are NaN."""
        a = np.ascontiguousarray(actions, self.dtype).reshape(self.n, self.act_dim)
        mt = int(self.params.max_ticks)
        obs = np.empty((self.n, OBS_DIM), self.dtype)
        rew = np.empty(self.n, self.dtype)
        done = np.empty(self.n, np.uint8)
        ticks = np.empty(self.n, np.int32)
        tobs = np.full((self.n, mt, OBS_DIM), np.nan, self.dtype)
        tlnk = np.full((self.n, mt, 51), np.nan, self.dtype)
        f = self._fn("step_trace"); f.restype = ctypes.c_int
        f(self._h, self._p(a), self._p(obs), self._p(rew), self._p(done), self._p(ticks), self._p(tobs), self._p(tlnk))
        return obs, rew, done.astype(bool), ticks, tobs, tlnk

    def rollout_linear(self, weights, n_steps, mean=None, inv_std=None, noise=None, trace=False):
        w = np.ascontiguousarray(weights, self.dtype).reshape(self.n, self.act_dim, OBS_DIM)
        m = None if mean is None else np.ascontiguousarray(mean, self.dtype)
        s = None if inv_std is None else np.ascontiguousarray(inv_std, self.dtype)
        z = None if noise is None else np.ascontiguousarray(noise, self.dtype).reshape(n_steps, self.n, OBS_DIM)
        ret = np.empty(self.n, self.dtype)
        tr = np.empty((n_steps, self.n, OBS_DIM), self.dtype) if trace else None
        f = self._fn("rollout_linear"); f.restype = ctypes.c_int
        f(self._h, self._p(w), self._p(m), self._p(s), self._p(z), ctypes.c_int32(n_steps), self._p(ret), self._p(tr))
        return (ret, tr) if trace else ret

    def tick(self, targets, n_ticks=1):
        t = np.ascontiguousarray(targets, self.dtype).reshape(self.n, NJ)
        iters = np.empty(self.n, np.int32)
        self._fn("tick")(self._h, self._p(t), ctypes.c_int32(n_ticks), self._p(iters))
        return iters

    def observe(self):
        obs = np.empty((self.n, OBS_DIM), self.dtype)
        self._fn("observe")(self._h, self._p(obs))
        return obs

    def get_state(self):
        s = np.empty((self.n, STATE_STRIDE), self.dtype)
        self._fn("get_state")(self._h, self._p(s))
        return s

    def set_state(self, s):
        s = np.ascontiguousarray(s, self.dtype).reshape(self.n, STATE_STRIDE)
        self._fn("set_state")(self._h, self._p(s))

    def counters(self, clear=False):
        out = (ctypes.c_int64 * 4)()
        self._fn("counters")(self._h, out, ctypes.c_int(int(clear)))
        return dict(ticks=out[0], pgs_iterations=out[1], dones=out[2], nonfinite=out[3])

    def set_contact_points(self, cpp):
        """Oracle-only sensitivity switch for deviation D1: contact points per cylinder in the exact tick (1 = default, 2 = both rims)."""
        f = self._fn("set_contact_points"); f.restype = ctypes.c_int
        if f(self._h, ctypes.c_int(int(cpp))) != 0:
            raise ValueError("contact points per cylinder must be 1 or 2")

    def set_manifold(self, on=True, warm=0.0):
        """Oracle-only switch for SURVEY 8f rank 2: Bullet-style persistent contact manifolds (support vertex of the 32-gon hull, 4 cached
        points per cylinder, Bullet's refresh / breaking rule) in the exact tick; ``warm`` = warm-start factor of the cached normal
        impulses (Bullet's warmstartingFactor is 0.1; 0 = off).  Clears the caches."""
        f = self._fn("set_manifold"); f.restype = ctypes.c_int; f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double]
        if f(self._h, int(bool(on)), float(warm)) != 0:
            raise RuntimeError("set_manifold failed")

    def set_joint_limits(self, on=True, angle=1.57):
        """Oracle-only switch: joint-limit rows (snake.urdf:833-840, +-1.57 rad) in the Bullet-order tick (motor_solver = 0)."""
        f = self._fn("set_joint_limits"); f.restype = ctypes.c_int; f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double]
        f(self._h, int(bool(on)), float(angle))

    def row_stats(self):
        """(mean cached contact points per tick, joint-limit row activations) of the optional rows."""
        out = (ctypes.c_double * 2)()
        f = self._fn("row_stats"); f.restype = ctypes.c_int
        f(self._h, out)
        return float(out[0]), int(out[1])

    def self_clearance(self):
        """Twin of ``snk_self_clearance``: lower bound [n] of the smallest distance between non-consecutive cylinders."""
        out = np.empty(self.n, self.dtype)
        f = self._fn("self_clearance"); f.restype = ctypes.c_int
        f(self._h, self._p(out))
        return out

    def kinematics(self, e=0):
        Rw = np.empty((17, 3, 3), self.dtype)
        pw = np.empty((17, 3), self.dtype)
        h = np.empty(1, self.dtype)
        self._fn("kinematics")(self._h, ctypes.c_int64(e), self._p(Rw), self._p(pw), self._p(h))
        return Rw, pw, float(h[0])
