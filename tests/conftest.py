import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree shared library (built on demand; nvcc cross-compiles without a GPU)."""
    from recguru_b200 import build
    return build.build()


def load_golden(name):
    g = np.load(GOLDEN / name)
    return {k: g[k] for k in g.files}


@pytest.fixture(scope="session")
def golden():
    return load_golden


def t(a, dtype=None):
    a = np.asarray(a)
    x = torch.from_numpy(a.copy() if a.ndim else a.reshape(()).copy())
    return x if dtype is None else x.to(dtype)
