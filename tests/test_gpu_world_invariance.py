""""action RNG keyed by global env id so results are invariant to W; W=1 vs W=2 bit-identical returns".
The same small workload (stepwise env-steps + a fused ARS rollout + the all-gather of the returns) is run as one process and
as two torchrun ranks; the gathered return vectors must be equal bit for bit."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_returns_do_not_depend_on_the_number_of_ranks(tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (there is no CPU fallback)")
    worker = os.path.join(ROOT, "tests", "w_invariance_worker.py")
    total, T = 3001, 6                                           # ragged shards: 1501 + 1500
    outs = []
    for world in (1, 2):
        out = str(tmp_path / ("w%d.npy" % world))
        if world == 1:
            cmd = [sys.executable, worker, out, str(total), str(T)]
        else:
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                   "--master-port", "29731", worker, out, str(total), str(T)]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, cwd=ROOT)
        assert r.returncode == 0, r.stderr[-3000:]
        outs.append(np.load(out))
    assert outs[0].shape == (total, 2) and np.isfinite(outs[0]).all() and np.abs(outs[0][:, 1]).max() > 0
    assert np.array_equal(outs[0], outs[1])
