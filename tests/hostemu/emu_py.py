"""ctypes front-end of tests/hostemu (CPU run of the CUDA kernel core; development check only)."""
import ctypes
import os
import subprocess

import numpy as np

from bullet_envs_b200._abi import CParams, default_params
from bullet_envs_b200.urdf_model import CModel, build_model

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_build", "libhostemu.so")


def build():
    deps = [os.path.join(_HERE, "hostemu.cpp")] + [os.path.join(_HERE, "..", "..", "bullet_envs_b200", "csrc", f)
                                                    for f in ("snake_exact_core.cuh", "snake_host.h", "snake_step.cuh")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.check_call([os.path.join(_HERE, "build.sh")])
    return LIB


class Emu:
    def __init__(self, n, params=None, model=None):
        self.lib = ctypes.CDLL(build())
        self.n = n
        self.params = params or default_params(motor_solver=1)
        self.model = model or build_model()
        self._cm = self.model.to_ctypes()
        self._h = ctypes.c_void_p()
        self.lib.emu_create.argtypes = [ctypes.POINTER(CModel), ctypes.POINTER(CParams), ctypes.c_int64, ctypes.POINTER(ctypes.c_void_p)]
        rc = self.lib.emu_create(ctypes.byref(self._cm), ctypes.byref(self.params), n, ctypes.byref(self._h))
        assert rc == 0
        self.act_dim = 8 if self.params.gait_selection in (0, 1) else 16

    @staticmethod
    def _p(a):
        return None if a is None else ctypes.c_void_p(a.ctypes.data)

    def set_state(self, s):
        s = np.ascontiguousarray(s, np.float32).reshape(self.n, 64)
        self.lib.emu_set_state(self._h, self._p(s))

    def get_state(self):
        s = np.empty((self.n, 64), np.float32)
        self.lib.emu_get_state(self._h, self._p(s))
        return s

    def tick(self, targets, n_ticks=1):
        t = np.ascontiguousarray(targets, np.float32).reshape(self.n, 16)
        it = np.empty(self.n, np.int32); nc = np.empty(self.n, np.int32); hh = np.empty(self.n, np.float32)
        self.lib.emu_tick(self._h, self._p(t), ctypes.c_int(n_ticks), self._p(it), self._p(nc), self._p(hh))
        return it, nc, hh

    def step(self, actions):
        a = np.ascontiguousarray(actions, np.float32).reshape(self.n, self.act_dim)
        obs = np.empty((self.n, 56), np.float32); rew = np.empty(self.n, np.float32)
        done = np.empty(self.n, np.uint8); ticks = np.empty(self.n, np.int32)
        self.lib.emu_step(self._h, self._p(a), self._p(obs), self._p(rew), self._p(done), self._p(ticks))
        return obs, rew, done.astype(bool), ticks

    def link_positions(self):
        out = np.empty((self.n, 51), np.float32)
        self.lib.emu_link_positions(self._h, self._p(out))
        return out

    def counters(self, clear=False):
        out = (ctypes.c_int64 * 4)()
        self.lib.emu_counters(self._h, out, ctypes.c_int(int(clear)))
        return dict(ticks=out[0], pgs_iterations=out[1], dones=out[2], nonfinite=out[3])
