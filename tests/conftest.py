import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def model():
    from bullet_envs_b200 import build_model
    return build_model()


@pytest.fixture(scope="session", autouse=True)
def _built_oracle():
    from oracle import oracle_py
    oracle_py.build()


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_python_task_logic.npz"))


@pytest.fixture(scope="session")
def golden_test_mode():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_python_test_mode.npz"))
