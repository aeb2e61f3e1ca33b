"""The C-ABI library loads and exports every symbol include/*.h declares (no compute without a GPU)."""
import ctypes
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names += re.findall(r"\b(snk_[a-z_0-9]+)\s*\(", src)
    return sorted(set(names))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from bullet_envs_b200 import _abi
    return _abi.load_library()


def test_header_declares_the_boundary():
    fns = declared_functions()
    for must in ("snk_create", "snk_destroy", "snk_reset", "snk_step", "snk_step_host", "snk_tick", "snk_get_state", "snk_set_state",
                 "snk_last_error"):
        assert must in fns


def test_library_exports_every_declared_symbol(lib):
    for name in declared_functions():
        assert hasattr(lib, name), "libsnake_b200.so does not export %s" % name


def test_struct_sizes_match_the_header(lib):
    from bullet_envs_b200._abi import CParams
    from bullet_envs_b200.urdf_model import CModel
    # snk_default_params fills the same bytes as the Python defaults
    from bullet_envs_b200 import default_params
    c = CParams()
    assert lib.snk_default_params(ctypes.byref(c)) == 0
    p = default_params()
    assert bytes(c) == bytes(p)
    assert ctypes.sizeof(CModel) == 8 * (16 * 9 + 16 * 3 + 16 * 3 + 16 + 17 + 17 * 3 + 17 * 9 + 32 * 3 + 32 * 3 + 32 * 9 + 32 * 5 + 17 * 3 + 3 + 1) + 4 * (32 + 17) + 4


def test_argument_errors_do_not_need_a_gpu(lib):
    assert lib.snk_default_params(None) < 0
    assert b"null" in lib.snk_last_error()
    h = ctypes.c_void_p()
    assert lib.snk_create(None, None, 4, 0, ctypes.byref(h)) < 0
    assert lib.snk_step(None, None, None, None, None, None, None) < 0
    assert lib.snk_step_trace(None, None, None, None, None, None, None, None, None) < 0 and lib.snk_self_clearance(None, None, None) < 0
    assert lib.snk_step_host_f64(None, None, None, None, None, None) < 0
    assert b"sm_100a" in lib.snk_build_info()


def test_no_cpu_fallback_in_the_product(lib, model):
    """Without a CUDA device snk_create must fail loudly (and the package never imports oracle/)."""
    import torch
    from bullet_envs_b200 import default_params
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    cm, p = model.to_ctypes(), default_params()
    rc = lib.snk_create(ctypes.byref(cm), ctypes.byref(p), 4, 0, ctypes.byref(h))
    assert rc < 0 and b"no CPU fallback" in lib.snk_last_error()
    import subprocess
    out = subprocess.run(["grep", "-rlE", r"oracle_py|from oracle|import oracle|libsnake_oracle", os.path.join(ROOT, "bullet_envs_b200")],
                         stdout=subprocess.PIPE, text=True).stdout.strip()
    assert out == "", "product package references the oracle: %s" % out
