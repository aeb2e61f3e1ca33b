#!/usr/bin/env python3
"""Static screening of operation-order variants of the solver rows (no GPU needed).

Compiles only the benchmarked kernel (snake_exact.cu with -DSNK_SCREEN) for every combination of the SNK_VAR_* macros of
snake_exact_core.cuh, disassembles it and counts, in the solver's sweep loop, the register-bank conflicts
(tools/sass_bank_conflicts.py: two source registers of one fp32 instruction in the same bank = one lost issue cycle) and the
instructions.  Prints the combinations sorted by conflicts; the best few are then measured on the GPU (profiles/README.md).

    python tools/screen_variants.py [-j 8] [--extra "-DSNK_UNROLL_F=2"]
"""
import argparse
import itertools
import os
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import sass_bank_conflicts as sbc  # noqa: E402

SRC = os.path.join(ROOT, "bullet_envs_b200", "csrc", "snake_exact.cu")
AXES = {"G": 3, "DW": 3, "F": 2, "T": 2, "U": 2, "NDW": 2}


def measure(combo, extra, td):
    tag = "_".join("%s%d" % kv for kv in combo.items())
    cubin = os.path.join(td, tag + ".cubin")
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-DSNK_SCREEN", "-cubin", "-o", cubin, SRC]
    cmd += ["-DSNK_VAR_%s=%d" % kv for kv in combo.items()] + extra
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if r.returncode != 0:
        return tag, None
    sass = os.path.join(td, tag + ".sass")
    with open(sass, "w") as f:
        subprocess.check_call(["cuobjdump", "-sass", cubin], stdout=f)
    k = sbc.kernels(sass)
    name = [n for n in k if "snk_hyb_step_kernel" in n][0]
    sweep = None
    for body in sbc.loops(sbc.ins_of(k[name])):
        n_ld4 = sum("LDTM.x4" in t for _, t in body)
        if n_ld4 >= 32 and len(body) < 1500:               # the sweep loop: 32 unrolled normal rows + the friction loop
            if sweep is None or len(body) < len(sweep):
                sweep = body
    lo, hi = sweep[0][0], sweep[-1][0]
    inner = [b for b in sbc.loops(sbc.ins_of(k[name])) if lo < b[0][0] and b[-1][0] < hi and any("LDTM.x8" in t for _, t in b)]
    inner = min(inner, key=len)                            # the rolled friction loop (2 pairs per iteration, 15 iterations per sweep)
    conf, reuse, nfp = sbc.metric(sweep)
    ci, _, _ = sbc.metric(inner)
    per_sweep = (conf - ci) + 15 * ci                      # executed conflict cycles per sweep: unrolled normal rows + peeled pairs once, the loop 15 x
    n_exec = (len(sweep) - len(inner)) + 15 * len(inner)
    os.remove(cubin); os.remove(sass)
    return tag, (per_sweep, n_exec, conf, ci)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-j", type=int, default=os.cpu_count() or 4)
    ap.add_argument("--extra", default="")
    a = ap.parse_args()
    combos = [dict(zip(AXES, v)) for v in itertools.product(*[range(n) for n in AXES.values()])]
    with tempfile.TemporaryDirectory() as td, ThreadPoolExecutor(a.j) as ex:
        res = list(ex.map(lambda c: measure(c, a.extra.split(), td), combos))
    res = [r for r in res if r[1]]
    res.sort(key=lambda r: (r[1][0], r[1][1]))
    for tag, (per_sweep, n_exec, conf, ci) in res:
        print("%-28s conflict cycles per contact and sweep %6.2f   instructions per contact and sweep %6.2f   (static: %d in the sweep body, %d in the friction loop)"
              % (tag, per_sweep / 32.0, n_exec / 32.0, conf, ci))


if __name__ == "__main__":
    main()
