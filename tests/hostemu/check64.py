"""fp64 run of the kernel core vs the fp64 oracle: checks the world-frame / centre-of-mass formulation
of the CUDA kernel against the oracle's body-frame formulation to round-off (development check)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np, ctypes
import hostemu.emu_py as ep
from oracle.oracle_py import Oracle
from bullet_envs_b200 import default_params


class Emu64(ep.Emu):
    def __init__(self, n, params):
        super().__init__(n, params)
        self.lib = ctypes.CDLL(ep.LIB.replace("libhostemu.so", "libhostemu64.so"))
        self._h = ctypes.c_void_p()
        self.lib.emu_create.argtypes = [ctypes.POINTER(ep.CModel), ctypes.POINTER(ep.CParams), ctypes.c_int64, ctypes.POINTER(ctypes.c_void_p)]
        assert self.lib.emu_create(ctypes.byref(self._cm), ctypes.byref(self.params), n, ctypes.byref(self._h)) == 0

    def set_state(self, s):
        s = np.ascontiguousarray(s, np.float64).reshape(self.n, 64); self.lib.emu_set_state(self._h, self._p(s))

    def get_state(self):
        s = np.empty((self.n, 64), np.float64); self.lib.emu_get_state(self._h, self._p(s)); return s

    def tick(self, targets, n_ticks=1):
        t = np.ascontiguousarray(targets, np.float64).reshape(self.n, 16)
        it = np.empty(self.n, np.int32); nc = np.empty(self.n, np.int32); hh = np.empty(self.n, np.float64)
        self.lib.emu_tick(self._h, self._p(t), ctypes.c_int(n_ticks), self._p(it), self._p(nc), self._p(hh)); return it, nc, hh

    def step(self, actions):
        a = np.ascontiguousarray(actions, np.float64).reshape(self.n, self.act_dim)
        obs = np.empty((self.n, 56)); rew = np.empty(self.n); done = np.empty(self.n, np.uint8); ticks = np.empty(self.n, np.int32)
        self.lib.emu_step(self._h, self._p(a), self._p(obs), self._p(rew), self._p(done), self._p(ticks)); return obs, rew, done.astype(bool), ticks


if __name__ == "__main__":
    n = 256
    p = default_params(motor_solver=1)
    o = Oracle(n, p); e = Emu64(n, p)
    rng = np.random.default_rng(1)
    s = np.zeros((n, 64)); s[:, 6] = 1
    for i in range(n):
        q = np.array([0, 0, 0, 1.0]) + rng.normal(size=4) * 0.02; s[i, 3:7] = q / np.linalg.norm(q)
    s[:, 0:3] = rng.normal(size=(n, 3)) * 0.1; s[:, 2] = rng.uniform(-0.002, 0.004, n)
    s[:, 7:10] = rng.normal(size=(n, 3)) * 0.3; s[:, 10:13] = rng.normal(size=(n, 3)) * 1.0
    s[:, 13:29] = rng.uniform(-0.5, 0.5, (n, 16)); s[:, 29:45] = rng.normal(size=(n, 16)) * 3
    tg = rng.uniform(-0.5, 0.5, (n, 16))
    o.set_state(s); e.set_state(s)
    it = o.tick(tg, 1); ite, nc, hh = e.tick(tg, 1)
    a = o.get_state(); b = e.get_state()
    print("iters equal", (it == ite).mean(), "max abs state diff", np.abs(a - b).max(), "rel tau", np.abs(a[:, 45:61] - b[:, 45:61]).max() / np.abs(a[:, 45:61]).max())
    o.reset(); e2 = Emu64(n, p)
    for t in range(5):
        A = rng.uniform(-1, 1, (n, 8))
        oo, orr, od, ot = o.step(A); eo, er, ed, et = e2.step(A)
        print(t, "ticks eq", (ot == et).mean(), "done eq", (od == ed).mean(), "obs diff", np.abs(oo - eo).max(), "rew diff", np.abs(orr - er).max())
    d = np.abs(a - b)
    i, k = np.unravel_index(d.argmax(), d.shape)
    print("worst env", i, "slot", k, "contacts", nc[i], "iters", it[i], a[i, k], b[i, k])
    print("per-slot max", np.round(np.log10(d.max(0) + 1e-30), 1))
    print("envs with diff>1e-9:", np.where(d.max(1) > 1e-9)[0], nc[d.max(1) > 1e-9], it[d.max(1) > 1e-9])
