import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_binding
    return oracle_binding.Oracle()


@pytest.fixture(scope="session")
def asb():
    import arrowspace_b200
    return arrowspace_b200


@pytest.fixture(scope="session")
def ctx(asb):
    """The CUDA context: creation FAILS (never falls back) when no B200 is present."""
    return asb.Context(0)


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    g = ROOT / "tests" / "golden"
    return {"proteins": np.load(g / "proteins_64x24.npy"), "quora": np.load(g / "quora_15x384.npy")}
