"""The hand-out model behind DESIGN.md section 6 (tools/handout_sim.py, no GPU): at the per-GPU batch of the 8-GPU run (131 072
environments = 3.46 waves of the 37 888-lane grid) whole env-steps cannot get below the integrality bound, the two pools get close to
it, and cutting the r shortest env-steps in two (the split hand-out of csrc/snake_exact.cu) goes below it; a conservation check of the
split model (every tick of every environment is run exactly once)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_split_hand_out_beats_the_whole_env_step_policies_in_the_model():
    import handout_sim as hs
    L = hs.SM * hs.WPS * 32
    n = 131072
    sc = np.sort(np.maximum(hs.ticks_sample(n, np.random.default_rng(0)), 1))[::-1]
    ideal = sc.sum() / L
    k, r = divmod(n, L)
    sw = (r + 31) // 32
    n_short = min(sw * 32 * (k + 1), n)
    lpt = hs.run(sc, [], 0)
    two = hs.run(sc[:n - n_short], sc[n - n_short:], sw)
    spl = hs.run_split(sc, r, 0.6)
    assert k == 3 and 2 * r <= L
    assert ideal < spl < two < lpt
    # the figures DESIGN.md quotes: 104.3 ideal, 119.7 longest first, 110.2 two pools, 107.0 split
    assert abs(ideal - 104.3) < 0.1 and abs(lpt - 119.7) < 0.6 and abs(two - 110.2) < 0.6 and abs(spl - 107.0) < 0.6
    # r > L / 2: several second parts per lane
    n2 = 100000
    sc2 = np.sort(np.maximum(hs.ticks_sample(n2, np.random.default_rng(0)), 1))[::-1]
    k2, r2 = divmod(n2, L)
    assert 2 * r2 > L
    assert hs.run_split(sc2, r2, min(0.9, r2 / L + 0.1)) < hs.run(sc2, [], 0) - 5
