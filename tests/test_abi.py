"""CPU: the C-ABI library builds, loads and exports every symbol include/lina_b200.h declares; the ctypes
prototype table covers exactly those symbols; ops fail loudly without a GPU (no CPU / oracle fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "lina_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lina_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from lina_speech_b200 import _lib
    names = _declared()
    assert len(names) >= 25
    assert os.path.exists(_lib.LIB_PATH), "liblina_b200.so must be built in-tree"
    raw = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(raw, n)]
    assert not missing, f"declared in include/lina_b200.h but not exported: {missing}"


def test_debug_library_is_separate_from_the_product():
    """the bring-up probes live in liblina_b200_debug.so / include/lina_b200_debug.h, not in the product library"""
    from lina_speech_b200 import _lib
    src = open(os.path.join(ROOT, "include", "lina_b200_debug.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = sorted(set(re.findall(r"\b(lina_[a-z0-9_]+)\s*\(", src)))
    assert names == sorted(_lib.DEBUG_PROTOTYPES)
    dbg = ctypes.CDLL(_lib.DEBUG_LIB_PATH)
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(dbg, n), n
        # each library has its own copy of the A/B variant table (and its setter); everything else is debug-only
        assert n == "lina_debug_set_variant" or not hasattr(raw, n), n


def test_ctypes_prototypes_match_header():
    from lina_speech_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == _declared()
    lib = _lib.lib()
    assert lib.lina_abi_version() >= 1
    assert lib.lina_last_error_string() is not None


def test_argument_validation_happens_before_any_cuda_call():
    """error codes + messages travel through the C ABI without a device (no compute is launched)."""
    from lina_speech_b200 import _lib
    lib = _lib.lib()
    rc = lib.lina_gla_recurrent_fwd(None, None, None, None, None, 0, None, None, 1, 1, 8, 512, 64, _lib.F32, 1.0, None)
    assert rc == -2 and b"K=512" in lib.lina_last_error_string()          # LINA_ERR_UNSUPPORTED
    rc = lib.lina_gla_recurrent_fwd(None, None, None, None, None, 0, None, None, 1, 1, 8, 64, 64, _lib.F32, 1.0, None)
    assert rc == -1                                                        # LINA_ERR_BAD_ARG: null pointers
    rc = lib.lina_rmsnorm_swishgate_fwd(None, None, None, None, None, 4, 8, 1e-5, 7, None)
    assert rc == -1
    assert lib.lina_gla_chunk_fwd_uses_tensor_cores(32, 4, 2048, 256, 512, _lib.BF16) == 1
    assert lib.lina_gla_chunk_fwd_uses_tensor_cores(32, 4, 2048, 256, 512, _lib.F32) == 0
    assert lib.lina_gla_chunk_fwd_uses_tensor_cores(1, 4, 128, 48, 128, _lib.BF16) == 0
    assert lib.lina_gla_recurrent_bwd_workspace_bytes(2, 3, 5, 7, 11) == (2 * 2 * 3 * 5 * 7 + 2 * 3 * 7) * 4


def test_ops_refuse_cpu_tensors():
    from lina_speech_b200.fla_api import fused_chunk_gla, chunk_gla, ShortConvolution, FusedRMSNormSwishGate
    x = torch.randn(1, 2, 8, 16)
    for fn in (fused_chunk_gla, chunk_gla):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            fn(x, x, x, x)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ShortConvolution(16, 4)(torch.randn(1, 8, 16))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        FusedRMSNormSwishGate(16)(torch.randn(4, 16), torch.randn(4, 16))


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under lina_speech_b200/ may reference it."""
    pkg = os.path.join(ROOT, "lina_speech_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports oracle"


def test_new_entry_points_validate_arguments_without_a_device():
    """the entries added for the pre-gated path, the tensor-core backward glue, the sampler and the autocast LayerNorm
    reject bad arguments before touching CUDA (error code + message through the C ABI)."""
    from lina_speech_b200 import _lib
    lib = _lib.lib()
    err = lambda: lib.lina_last_error_string().decode()
    assert lib.lina_gla_prefill_prep_gated(None, 0, None, 0, None, 0, None, None, None, None, 0, None, None, None, None, None,
                                           None, None, 0, 1, 8, 4, 256, 512, 4, 16.0, 0.0625, None, None) == -1
    assert "null pointer" in err()
    assert lib.lina_gla_chunk_fwd_pregated_bthd(None, None, None, None, None, 0, None, None, 1, 4, 128, 256, 512, None) == -1
    assert lib.lina_gla_chunk_fwd_pregated(None, None, None, None, None, 0, None, None, 1, 4, 128, 256, 512, 0, 0, 0, 0, 0, 0,
                                           None) == -1
    assert lib.lina_gla_bwd_prep(None, None, None, None, None, None, None, None, 1, 4, 128, 256, 0, 0.0625, None) == -1
    assert lib.lina_gla_bwd_post(None, None, None, None, None, None, None, None, None, None, None, 1, 4, 128, 256, 0, 0.0625,
                                 None) == -1
    assert lib.lina_topk_sample(None, 0, 1, 10, 1, 1.0, None, None, _lib.F32, None) == -1
    assert lib.lina_layernorm_f32in_fwd(None, None, None, None, None, None, 4, 8, 1e-5, _lib.BF16, None) == -1
    assert lib.lina_cross_entropy_rows(None, 0, None, None, None, None, 4, 8, 1, _lib.BF16, None) == -1
    assert lib.lina_debug_set_variant(99, 1) == -1 and lib.lina_debug_set_variant(0, 0) == 0


def test_gla_time_cut_plan_sizes():
    """lina_gla_chunk_fwd_pregated_ws_bytes: a workspace is asked for exactly when the last wave of CTA-pair tiles is at most half
    full (and a full wave precedes it, and T has >= 4 chunks); its size is the flag block + one fp32 [K <= 256, 128] state slice
    per CTA of the cut tiles.  Without a device the plan assumes 148 SMs (B200)."""
    from lina_speech_b200 import _lib
    lib = _lib.lib()
    f = lib.lina_gla_chunk_fwd_pregated_ws_bytes
    # bench shape: 32 * 4 * 2 = 256 pair tiles on 74 slots -> 3 waves + 34 tiles: cut those 34
    assert f(32, 4, 2048, 256, 512) == 512 + 34 * 2 * 256 * 128 * 4
    assert f(8, 4, 4096, 256, 512) == 0            # 64 tiles < 74 slots: one partial wave, nothing to gain
    assert f(2, 4, 1024, 256, 512) == 0
    assert f(37, 4, 2048, 256, 512) == 0           # 296 tiles = 4 full waves exactly
    assert f(32, 4, 128, 256, 512) == 0            # two chunks only: too short to cut
    assert f(32, 4, 2048, 64, 512) == 0            # K = 64 runs on the one-CTA-per-tile kernel
    assert f(32, 4, 2048, 256, 384) == 0           # V / 128 odd: no CTA pairs
    assert f(32, 4, 2048, 48, 512) == 0            # outside the tensor-core envelope
