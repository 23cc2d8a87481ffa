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


@pytest.fixture(scope="session")
def golden_triton():
    """Outputs of the reference's OWN Triton kernels (fla fused_chunk_gla / chunk_gla, bf16) run on a B200 by
    profiles/triton_reference_bench.py; inputs are rebuilt from the seed by :func:`triton_golden_inputs`."""
    with np.load(os.path.join(GOLDEN, "gla_triton_bf16.npz")) as z:
        out = {}
        for k in z.files:
            if k.endswith("_bf16bits"):
                out[k[:-9]] = torch.from_numpy(z[k]).view(torch.bfloat16)
            else:
                out[k] = torch.from_numpy(z[k])
        return out


def triton_golden_inputs(shape_row):
    """Same generator calls as profiles/triton_reference_bench.py:inputs (CPU generator: identical on every machine)."""
    import torch.nn.functional as F
    B, H, T, K, V, gates, seed = (int(x) for x in shape_row)
    g = torch.Generator().manual_seed(seed)
    q, k = (torch.randn(B, H, T, K, generator=g).bfloat16() for _ in range(2))
    v = torch.randn(B, H, T, V, generator=g).bfloat16()
    x = torch.randn(B, H, T, K, generator=g)
    gk = (F.logsigmoid(x).clamp_min(-5) if gates else F.logsigmoid(x) / 16).bfloat16()
    return q, k, v, gk


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Every test session works against a freshly built in-tree liblina_b200.so when nvcc is present."""
    import shutil
    if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
        from lina_speech_b200 import _build
        _build.build()
    yield
