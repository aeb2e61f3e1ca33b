"""Event-driven model of the step kernel's hand-out (no GPU needed): how close is the launch time to total work / lanes?

The persistent grid has L = 148 SMs x 8 warps x 32 lanes.  An env-step costs its predicted tick count in warp iterations (a lane
ticks whenever its warp ticks; a zero-tick step costs one iteration), a lane that finishes takes the next environment of its pool at
once, and a warp whose partner on the scheduler has run out of work iterates faster (`lone`, measured: one warp alone on a
scheduler needs about 0.68 of the time per iteration).  Policies:

  lpt   one pool, longest first (SNK_EXACT_BALANCE=0)
  two   the launcher's two pools: with n = k L + r, the first ceil(r / 32) warps draw k + 1 env-steps each from the SHORTEST
        (k + 1) r environments, the others k from the longest ("Balanced last wave")

  split the r shortest env-steps are cut in two at a tick boundary (SplitCta in csrc/snake_exact.cu): r lanes, dealt out evenly over
        the CTAs and warps, start with a first part of `phi` of the predicted ticks and park it; when the whole env-steps have run
        out, the free lanes of the same CTA finish the parked ones (phi = 0.6 for r <= L / 2, else r / L + 0.1)

and a reference figure for the integrality of whole env-steps: r lanes must run k + 1 env-steps, and the (k + 1) r SHORTEST jobs spread
evenly over them already cost `ref` iterations each -- with the narrow tick distribution of this workload (30.1 +- 2.2) the loss against
n T / L is a property of the batch size, not of the policy (the dynamic pools get slightly below `ref` because warps switch pools and a
lone warp iterates faster).  Used for DESIGN.md section 6 (why 131 072 environments per GPU run at 91 % of the 2^20 rate).

    python tools/handout_sim.py 65536,100000,131072,200000,262144
"""
import heapq
import sys

import numpy as np

SM, WPS = 148, 8


def ticks_sample(n, rng):
    """tick counts of U[-1,1]^8 actions following U[-1,1]^8 actions: ceil(log(0.05 / err) / log(0.9)), err = pi/6 |a1 - a0|"""
    a0 = rng.uniform(-1, 1, (n, 8)); a1 = rng.uniform(-1, 1, (n, 8))
    err = np.pi / 6 * np.linalg.norm(a1 - 0.97 * a0, axis=1)
    k = np.where(err > 0.05, np.clip(np.ceil(np.log(0.05 / np.maximum(err, 1e-9)) / np.log(0.9)), 1, 41), 0)
    return k.astype(int)


def run(long_list, short_list, short_warps, lone=0.68):
    W = SM * WPS
    lists, cur = [list(long_list), list(short_list)], [0, 0]
    rem = np.zeros((W, 32), int); have = np.zeros((W, 32), bool)
    pool = np.array([1 if gw < short_warps else 0 for gw in range(W)])
    active = np.ones(W, bool); endt = np.zeros(W)

    def partner(gw):  # warps w and w + 4 of an SM share a scheduler; gw = warp * SM + sm (warp-major over the grid)
        w, sm = divmod(gw, SM)
        return ((w + 4) % WPS) * SM + sm

    h = [(0.0, gw) for gw in range(W)]
    heapq.heapify(h)
    while h:
        t, gw = heapq.heappop(h)
        for lane in range(32):
            if have[gw, lane]:
                continue
            p = pool[gw]
            for _ in range(2):
                if cur[p] < len(lists[p]):
                    rem[gw, lane] = lists[p][cur[p]]; cur[p] += 1; have[gw, lane] = True
                    break
                p ^= 1
            if have[gw, lane]:
                pool[gw] = p
        if not have[gw].any():
            active[gw] = False; endt[gw] = t
            continue
        dt = 1.0 if active[partner(gw)] else lone
        rem[gw][have[gw]] -= 1
        have[gw][have[gw] & (rem[gw] <= 0)] = False
        heapq.heappush(h, (t + dt, gw))
    return endt.max()


def run_split(sc, S, phi, lone=0.68):
    """sc: tick counts in descending order; the last S (shortest) are split; a-part = round(phi * ticks) >= 1."""
    W = SM * WPS
    n = len(sc)
    whole, cur = list(sc[:n - S]), 0
    split = sc[n - S:]
    rem = np.zeros((W, 32), int); have = np.zeros((W, 32), bool); bpend = np.zeros((W, 32), int)
    parked = [[] for _ in range(SM)]; taken = [0] * SM
    idx = 0
    for sm in range(SM):
        for j in range(S // SM + (1 if sm < S % SM else 0)):  # entry j of the CTA: warp j % 8, lane j // 8
            gw, lane = (j % WPS) * SM + sm, j // WPS
            t = int(split[idx]); idx += 1
            a = max(1, int(round(t * phi)))
            rem[gw, lane] = min(a, t); have[gw, lane] = True; bpend[gw, lane] = max(t - a, 0)
    active = np.ones(W, bool); endt = np.zeros(W)

    def partner(gw):
        w, sm = divmod(gw, SM)
        return ((w + 4) % WPS) * SM + sm

    h = [(0.0, gw) for gw in range(W)]
    heapq.heapify(h)
    while h:
        t, gw = heapq.heappop(h)
        sm = gw % SM
        for lane in range(32):
            if have[gw, lane]:
                continue
            if cur < len(whole):
                rem[gw, lane] = whole[cur]; cur += 1; have[gw, lane] = True
            elif taken[sm] < len(parked[sm]):
                rem[gw, lane] = parked[sm][taken[sm]]; taken[sm] += 1; have[gw, lane] = True
        if not have[gw].any():
            active[gw] = False; endt[gw] = t
            continue
        dt = 1.0 if active[partner(gw)] else lone
        rem[gw][have[gw]] -= 1
        fin = have[gw] & (rem[gw] <= 0)
        for lane in np.nonzero(fin)[0]:
            if bpend[gw, lane] > 0:
                parked[sm].append(bpend[gw, lane]); bpend[gw, lane] = 0
        have[gw][fin] = False
        heapq.heappush(h, (t + dt, gw))
    assert all(taken[s] == len(parked[s]) for s in range(SM))
    return endt.max()


def main():
    L = SM * WPS * 32
    sizes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "65536,100000,131072,200000,262144").split(",")]
    print("%8s %7s %5s | %17s | %17s | %17s | %s" % ("envs", "ideal", "k+r/L", "longest first", "two pools", "split", "k + 1 shortest jobs on r lanes"))
    for n in sizes:
        rng = np.random.default_rng(0)
        sc = np.sort(np.maximum(ticks_sample(n, rng), 1))[::-1]
        T = sc.sum() / L
        k, r = divmod(n, L)
        sw = (r + 31) // 32
        n_short = min(sw * 32 * (k + 1), n)
        lpt = run(sc, [], 0)
        two = run(sc[:n - n_short], sc[n - n_short:], sw) if k >= 1 and r > 0 else lpt
        # r lanes carry k + 1 jobs: at best the (k + 1) r shortest, evenly spread
        bound = max(T, sc[n - n_short:].sum() / max(sw * 32, 1)) if r > 0 else T
        spl = run_split(sc, r, 0.6 if 2 * r <= L else min(0.9, r / L + 0.1)) if k >= 1 and 16 * r >= L else lpt
        print("%8d %7.1f %5.2f | %7.1f (-%4.1f %%) | %7.1f (-%4.1f %%) | %7.1f (-%4.1f %%) | %7.1f (-%4.1f %%)" % (
            n, T, n / L, lpt, 100 * (1 - T / lpt), two, 100 * (1 - T / two), spl, 100 * (1 - T / spl), bound, 100 * (1 - T / bound)))


if __name__ == "__main__":
    main()
