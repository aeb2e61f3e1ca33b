"""First-contact GPU debug: one raw tick and short rollouts vs the oracle; prints error tables."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import __graft_entry__ as ge
ge.build()
from bullet_envs_b200 import SnakeVecEnv, default_params
from oracle.oracle_py import Oracle

np.set_printoptions(precision=6, suppress=True, linewidth=180)
n = 256
rng = np.random.default_rng(0)

def rand_states(n, lifted=False):
    s = np.zeros((n, 64))
    s[:, 6] = 1
    s[:, 13:29] = rng.uniform(-0.5, 0.5, (n, 16))
    s[:, 29:45] = rng.normal(size=(n, 16)) * 0.5
    s[:, 7:10] = rng.normal(size=(n, 3)) * 0.05
    s[:, 10:13] = rng.normal(size=(n, 3)) * 0.1
    yaw = rng.uniform(-np.pi, np.pi, n)
    s[:, 5] = np.sin(yaw / 2); s[:, 6] = np.cos(yaw / 2)
    s[:, 0:2] = rng.uniform(-1, 1, (n, 2))
    s[:, 2] = 0.3 if lifted else 0.0005
    return s

print(torch.cuda.get_device_name(0))
for lifted in (True, False):
    env = SnakeVecEnv(num_envs=n, device=0)
    o64 = Oracle(n); o32 = Oracle(n, f32=True)
    s0 = rand_states(n, lifted)
    tg = rng.uniform(-0.5, 0.5, (n, 16))
    env.set_state(s0); o64.set_state(s0); o32.set_state(s0)
    env.tick(tg, 1); it64 = o64.tick(tg, 1); it32 = o32.tick(tg, 1)
    g = env.get_state().cpu().numpy().astype(np.float64)
    a = o64.get_state(); b = o32.get_state().astype(np.float64)
    print("=== one tick, lifted=%s; gpu counters %s; oracle iters mean %.1f" % (lifted, env.counters(), it64.mean()))
    for name, sl in (("pos", slice(0, 3)), ("quat", slice(3, 7)), ("vel", slice(7, 10)), ("omega", slice(10, 13)), ("q", slice(13, 29)),
                     ("qd", slice(29, 45)), ("tau", slice(45, 61)), ("fz", slice(61, 62))):
        sc = np.abs(a[:, sl]).max() + 1e-12
        print("  %-6s scale %10.4g | gpu-o64 max %10.3g | gpu-o32 max %10.3g | o32-o64 max %10.3g" % (
            name, sc, np.abs(g[:, sl] - a[:, sl]).max(), np.abs(g[:, sl] - b[:, sl]).max(), np.abs(b[:, sl] - a[:, sl]).max()))
    env.close()

# rollouts through the task logic
n = 512
env = SnakeVecEnv(num_envs=n, device=0)
o64 = Oracle(n); o32 = Oracle(n, f32=True)
T = 30
acts = rng.uniform(-1, 1, (T, n, 8)).astype(np.float32)
og = env.reset(as_torch=True).cpu().numpy(); oa = o64.reset(); ob = o32.reset()
print("reset obs diff", np.abs(og - oa).max())
Rg = np.zeros(n); Ra = np.zeros(n); Rb = np.zeros(n)
for t in range(T):
    obs, rew, done, _ = env.step(torch.from_numpy(acts[t]).cuda())
    tk = env.last_ticks.cpu().numpy()
    xa, ra, da, ta = o64.step(acts[t], threads=8); xb, rb, db, tb = o32.step(acts[t], threads=8)
    xg = obs.cpu().numpy().astype(np.float64); rg = rew.cpu().numpy().astype(np.float64)
    Rg += rg; Ra += ra; Rb += rb
    same = (tk == ta)
    print("t=%2d ticks gpu %.2f o64 %.2f | same-ticks %.3f (o32 vs o64 %.3f) | done gpu %d o64 %d | q err max %.3g (o32: %.3g) | base xyz err %.3g (o32 %.3g) | rew err %.3g (o32 %.3g) | tau err %.3g" % (
        t, tk.mean(), ta.mean(), same.mean(), (tb == ta).mean(), done.sum().item(), da.sum(),
        np.abs(xg[same][:, :16] - xa[same][:, :16]).max(), np.abs(xb[:, :16] - xa[:, :16]).max(),
        np.abs(xg[same][:, 48:51] - xa[same][:, 48:51]).max(), np.abs(xb[:, 48:51] - xa[:, 48:51]).max(),
        np.abs(rg[same] - ra[same]).max(), np.abs(rb - ra).max(), np.abs(xg[same][:, 32:48] - xa[same][:, 32:48]).max()))
print("episode-return rel err gpu vs o64: median %.3g max %.3g ; o32 vs o64 median %.3g max %.3g" % (
    np.median(np.abs(Rg - Ra) / (np.abs(Ra) + 1e-9)), np.max(np.abs(Rg - Ra) / (np.abs(Ra) + 1e-9)),
    np.median(np.abs(Rb - Ra) / (np.abs(Ra) + 1e-9)), np.max(np.abs(Rb - Ra) / (np.abs(Ra) + 1e-9))))
env.close()

# timing
for n in (4096, 16384):
    env = SnakeVecEnv(num_envs=n, device=0)
    env.reset(as_torch=True)
    g = torch.Generator(device="cuda").manual_seed(0)
    acts = torch.rand((12, n, 8), device="cuda", generator=g) * 2 - 1
    for t in range(2):
        env.step(acts[t])
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    tot_ticks = 0
    e0.record()
    for t in range(2, 12):
        env.step(acts[t])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    c = env.counters()
    print("N=%d: %.3f ms/step -> %.3g env-steps/s ; last step ticks/env %.2f iters/tick %.2f -> %.3g ticks/s" % (
        n, ms, n / ms * 1e3, c["ticks"] / n, c["pgs_iterations"] / max(1, c["ticks"]), c["ticks"] / ms * 1e3))
    env.close()
