#!/usr/bin/env python3
"""How much of the solver time is lost to lock-step sweeps, and could regrouping lanes recover it?  

The env-step kernel runs one environment per lane; a warp sweeps the projected Gauss-Seidel until its SLOWEST lane has
converged (<= 50 sweeps).  This script records per-tick sweep counts of the oracle's exact tick (the kernel's CPU twin) for a
batch of environments under the bench workload (U[-1,1] actions), then replays three assignments of environments to the 6 warps
of a CTA, charging every warp-tick  c0 + c1 * max(sweeps of its lanes)  with the solver at 85 % of a 50-sweep tick:
  base          lanes keep their environment (what the kernel does);
  perfect-sort  before every tick the 192 environments of the CTA are sorted by the sweep count they are ABOUT to need (oracle
                knowledge: an upper bound on any regrouping scheme);
  pred-sort     sorted by the sweep count of their previous tick (the best predictor available before the solve).
Test infrastructure (uses oracle/): python tests/sweep_regroup_sim.py [n_envs] [env_steps]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.oracle_py import Oracle  # noqa: E402


def traces(n, steps, seed=0):
    o = Oracle(n)
    rng = np.random.default_rng(seed)
    o.reset()
    sf = np.pi / 6
    out = []
    for _ in range(steps):
        a = rng.uniform(-1, 1, (n, 8))
        tg = np.zeros((n, 16)); tg[:, 1::2] = a * sf
        tr = -np.ones((41, n), np.int32)
        for t in range(41):                                   # the tick loop of snake.py:284-304 (no height break / resets)
            st = o.get_state()
            need = np.sqrt(((tg - st[:, 13:29]) ** 2).sum(1)) > 0.05
            if not need.any():
                break
            it = o.tick(tg, 1)
            st2 = o.get_state(); st2[~need] = st[~need]; o.set_state(st2)
            tr[t, need] = it[need]
        out.append(tr)
    return np.array(out)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1536
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    T = traces(n, steps)[1:]                                   # drop the step that leaves the rest pose
    seqs = [np.concatenate([T[k, :, e][T[k, :, e] >= 0] for k in range(T.shape[0])]) for e in range(n)]
    L = min(len(s) for s in seqs)
    A = np.array([s[:L] for s in seqs])
    print("environments %d, ticks each %d, sweeps per tick: lane mean %.1f, hit the 50-sweep cap %.0f %%, tick-to-tick correlation %.2f"
          % (n, L, A.mean(), 100 * (A >= 50).mean(), np.corrcoef(A[:, :-1].ravel(), A[:, 1:].ravel())[0, 1]))
    c0, c1 = 0.15, 0.017
    cost = lambda m: c0 + c1 * m
    base = perf = pred = ideal = 0.0
    for g in range(n // 192):
        X = A[g * 192:(g + 1) * 192]
        base += cost(X.reshape(6, 32, L).max(1)).sum()
        ideal += cost(X.mean(0)).sum() * 6
        for t in range(L):
            x = X[:, t]
            perf += cost(np.sort(x).reshape(6, 32).max(1)).sum()
            p = X[:, t - 1] if t > 0 else np.full(192, 30)
            pred += cost(x[np.argsort(p, kind="stable")].reshape(6, 32).max(1)).sum()
    print("warp-level sweeps per tick (base): %.1f" % np.mean([A[g * 192:(g + 1) * 192].reshape(6, 32, L).max(1).mean() for g in range(n // 192)]))
    print("relative time: base 1.000 | perfect-sort %.3f | pred-sort %.3f | no lock-step loss at all %.3f" % (perf / base, pred / base, ideal / base))


if __name__ == "__main__":
    main()
