import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: torch.from_numpy(z[k]) for k in z.files}


@pytest.fixture(scope="session")
def golden_ops():
    return load_golden("gla_ops.npz")


@pytest.fixture(scope="session")
def golden_layer():
    return load_golden("gla_layer_cfg1.npz")


@pytest.fixture(scope="session")
def golden_model():
    return load_golden("lina_tiny.npz")


@pytest.fixture(scope="session")
def golden_codec():
    return load_golden("codec_small.npz")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Every test session works against a freshly built in-tree liblina_b200.so when nvcc is present."""
    import shutil
    if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
        from lina_speech_b200 import _build
        _build.build()
    yield
