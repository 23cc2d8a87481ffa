"""ctypes binding of liblina_b200.so (the C ABI declared in include/lina_b200.h).

There is no fallback: if the library is missing, ``lib()`` raises; every op
requires CUDA tensors and raises otherwise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "liblina_b200.so")

F32, BF16, F16 = 0, 1, 2
_DT = {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16}

_p, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t

# name -> (restype, argtypes).  Kept in lock-step with include/lina_b200.h (tests/test_abi.py checks it).
PROTOTYPES = {
    "lina_abi_version": (_i, []),
    "lina_last_error_string": (C.c_char_p, []),
    "lina_gla_recurrent_fwd": (_i, [_p] * 5 + [_i, _p, _p] + [_i] * 6 + [_f, _p]),
    "lina_rwkv6_recurrent_fwd": (_i, [_p] * 6 + [_i, _p, _p] + [_i] * 6 + [_f, _p]),
    "lina_gla_recurrent_bwd_workspace_bytes": (_sz, [_i] * 5),
    "lina_gla_recurrent_bwd": (_i, [_p] * 5 + [_i] + [_p] * 8 + [_i] * 6 + [_f, _p]),
    "lina_gla_chunk_fwd_workspace_bytes": (_sz, [_i] * 6),
    "lina_gla_chunk_fwd": (_i, [_p] * 5 + [_i, _p, _p, _p] + [_i] * 6 + [_f, _p]),
    "lina_gla_chunk_fwd_bthd": (_i, [_p] * 5 + [_i, _p, _p, _p] + [_i] * 6 + [_f, _p]),
    "lina_gla_chunk_fwd_uses_tensor_cores": (_i, [_i] * 6),
    "lina_gla_step_workspace_bytes": (_sz, [_i] * 4),
    "lina_gla_step": (_i, [_p] * 15 + [_i] * 7 + [_f] * 3 + [_p]),
    "lina_gla_step_ld": (_i, [_p] * 15 + [_i] * 7 + [_f] * 3 + [_i, _i, _p]),
    "lina_gla_step_lr": (_i, [_p] * 4 + [_i, _p, _p, _i] + [_p] * 11 + [_i] * 7 + [_f] * 3 + [_i, _p]),
    "lina_gla_prefill_prep": (_i, [_p] * 3 + [C.c_longlong] + [_p] * 4 + [C.c_longlong] + [_p] * 7 + [_i] * 6 + [_f, _f, _i, _i, _p]),
    "lina_lowrank_linear": (_i, [_p, C.c_longlong, _p, _p, _p, C.c_longlong, _i, _i, _i, _i, _p]),
    "lina_gla_prefill_prep_gated": (_i, [_p, C.c_longlong] * 3 + [_p] * 4 + [C.c_longlong] + [_p] * 7 + [_i] * 7 + [_f, _f, _p, _p]),
    "lina_gla_chunk_fwd_pregated_bthd": (_i, [_p] * 5 + [_i, _p, _p] + [_i] * 5 + [_p]),
    "lina_gla_chunk_fwd_pregated_ws_bytes": (_sz, [_i] * 5),
    "lina_gla_chunk_fwd_pregated_bthd_ws": (_i, [_p] * 5 + [_i, _p, _p, _p, _sz] + [_i] * 5 + [_p]),
    "lina_gla_chunk_fwd_pregated": (_i, [_p] * 5 + [_i, _p, _p] + [_i] * 11 + [_p]),
    "lina_gla_bwd_prep": (_i, [_p] * 8 + [_i] * 5 + [_f, _p]),
    "lina_time_reverse_pad2": (_i, [_p] * 4 + [C.c_longlong, _i, _i, _i, _p]),
    "lina_gla_bwd_post": (_i, [_p] * 11 + [_i] * 5 + [_f, _p]),
    "lina_gla_bwd_dgk_finish": (_i, [_p] * 3 + [_i] * 5 + [_p]),
    "lina_short_conv_fwd": (_i, [_p] * 4 + [_i] * 7 + [_p]),
    "lina_short_conv_bwd": (_i, [_p] * 5 + [_i] * 6 + [_p]),
    "lina_short_conv_update": (_i, [_p, _p, _i, _p, _p] + [_i] * 5 + [_p]),
    "lina_rmsnorm_swishgate_fwd": (_i, [_p] * 5 + [_i, _i, _f, _i, _p]),
    "lina_rmsnorm_swishgate_fwd_ld": (_i, [_p] * 5 + [_i, _i, _f, _i, C.c_longlong, _i, _p]),
    "lina_rmsnorm_swishgate_bwd": (_i, [_p] * 8 + [_i, _i, _i, _p]),
    "lina_gate_logsigmoid": (_i, [_p, _p, C.c_longlong, _f, _f, _i, _i, _p]),
    "lina_swiglu_act": (_i, [_p, _p, _i, _i, _i, _p]),
    "lina_add_layernorm": (_i, [_p] * 6 + [_i, _i, _f, _i, _p]),
    "lina_layernorm_f32in_fwd": (_i, [_p] * 6 + [_i, _i, _f, _i, _p]),
    "lina_layernorm_f32in_bwd": (_i, [_p] * 8 + [_i, _i, _i, _p]),
    "lina_cross_entropy_rows": (_i, [_p, C.c_longlong, _p, _p, _p, _p, _i, _i, C.c_longlong, _i, _p]),
    "lina_topk_sample": (_i, [_p, C.c_longlong, _i, _i, _i, _f, _p, _p, _i, _p]),
    "lina_codec_codes_to_features": (_i, [_p] * 3 + [_i] * 5 + [_p]),
    "lina_codec_groupnorm_swish": (_i, [_p] * 5 + [_i] * 4 + [_f, _i, _p]),
    "lina_codec_dwconv_adaln": (_i, [_p] * 6 + [_i] * 3 + [_f, _p]),
    "lina_codec_scale_residual_t": (_i, [_p] * 4 + [_i] * 3 + [_p]),
    "lina_codec_layernorm_t": (_i, [_p] * 4 + [_i] * 3 + [_f, _p]),
    "lina_codec_dwconv_adaln_workspace_bytes": (_sz, [_i] * 3),
    "lina_codec_dwconv_adaln_ws": (_i, [_p] * 7 + [_i] * 3 + [_f, _p]),
    "lina_codec_layernorm_t_ws": (_i, [_p] * 5 + [_i] * 3 + [_f, _p]),
    "lina_codec_istft_workspace_bytes": (_sz, [_i] * 3),
    "lina_codec_istft_head": (_i, [_p] * 4 + [_i] * 4 + [_p]),
    "lina_codec_istft_head_ld": (_i, [_p, C.c_longlong] + [_p] * 3 + [_i] * 4 + [_p]),
    "lina_codec_cl_gather": (_i, [_p] * 4 + [_i] * 6 + [_p]),
    "lina_codec_cl_gn_partials_bytes": (_sz, [_i] * 3),
    "lina_codec_cl_gn_partials": (_i, [_p, _p] + [_i] * 4 + [_p]),
    "lina_codec_cl_rows": (_i, [_p] * 6 + [_i, _f, _i, _p, _p, _f, _p, _p] + [_i] * 4 + [_p]),
    "lina_cross_att_step": (_i, [_p, C.c_longlong, _p, _p, _f, _p, C.c_longlong, _p, C.c_longlong, _p, C.c_longlong, _p, C.c_longlong,
                                 _i, _i, _i, _f, _i, _p]),
    "lina_skinny_linear_max_rows": (_i, []),
    "lina_skinny_linear": (_i, [_p, C.c_longlong, _p, C.c_longlong, _p, _p, _f, _p, _p, C.c_longlong, _p, _p, C.c_longlong, _i, _i, _i, _i, _p]),
    "lina_codec_cl_softmax": (_i, [_p, _p, _i, C.c_longlong, _i, C.c_longlong, C.c_longlong, _p]),
    "lina_debug_set_variant": (_i, [_i, _i]),
}

# liblina_b200_debug.so (include/lina_b200_debug.h): bring-up probes, not part of the product library
DEBUG_LIB_PATH = os.path.join(_HERE, "lib", "liblina_b200_debug.so")
DEBUG_PROTOTYPES = {
    "lina_debug_umma_probe": (_i, [_p] * 3 + [_i] * 5 + [_p]),
    "lina_debug_umma_probe_m": (_i, [_p] * 3 + [_i] * 3 + [_p]),
    "lina_debug_umma_probe_sw128": (_i, [_p] * 4 + [_i] * 5 + [_p]),
    "lina_debug_umma_timing": (_i, [_p] + [_i] * 6 + [_p]),
    "lina_debug_set_variant": (_i, [_i, _i]),
    "lina_debug_gla_chunk_trace": (_i, [_p] * 5 + [_i] * 5 + [_f, _p, _p]),
    "lina_debug_gla_pregated_trace": (_i, [_p] * 5 + [_i] * 5 + [_p, _p]),
}


class GemmArgs(C.Structure):
    """``lina_gemm_args`` of include/lina_b200.h (field order and types are the header's)."""
    _fields_ = [("a", _p * 3), ("lda", C.c_longlong), ("a_batch_stride", C.c_longlong), ("a_parts", _i),
                ("b", _p * 3), ("ldb", C.c_longlong), ("b_batch_stride", C.c_longlong), ("b_parts", _i), ("b_batched", _i),
                ("n_terms", _i), ("term_a", _i * 6), ("term_b", _i * 6),
                ("NB", _i), ("L", _i), ("N", _i), ("K", _i), ("taps", _i), ("pad", _i),
                ("alpha", _f), ("bias", _p), ("gamma", _p), ("residual", _p), ("ld_res", C.c_longlong), ("act", _i),
                ("out_f32", _p), ("ld_out", C.c_longlong), ("out_split", _p * 3), ("ld_split", C.c_longlong),
                ("out_parts", _i), ("b_mn", _i), ("span", _i)]


PROTOTYPES["lina_gemm_bf16_terms"] = (_i, [C.POINTER(GemmArgs), _p])

_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """Load the CUDA library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"liblina_b200.so not found at {LIB_PATH}: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a). There is no CPU or PyTorch fallback for these ops.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


_debug_lib: Optional[C.CDLL] = None


def debug_lib() -> C.CDLL:
    """The bring-up probe library (tests/test_umma_probe_gpu.py, profiles/probe_m64.py, profiles/umma_timing.py)."""
    global _debug_lib
    if _debug_lib is None:
        if not os.path.exists(DEBUG_LIB_PATH):
            raise RuntimeError(f"liblina_b200_debug.so not found at {DEBUG_LIB_PATH}: run __graft_entry__.build()")
        l = C.CDLL(DEBUG_LIB_PATH)
        for name, (res, args) in DEBUG_PROTOTYPES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        fn = l.lina_last_error_string
        fn.restype, fn.argtypes = C.c_char_p, []
        _debug_lib = l
    return _debug_lib


def check(rc: int, what: str, l: Optional[C.CDLL] = None) -> None:
    if rc != 0:
        msg = (l or lib()).lina_last_error_string().decode(errors="replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def dt(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported dtype {t.dtype} (float32 / bfloat16 / float16 only)") from None


def ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def stream(t: torch.Tensor):
    return torch.cuda.current_stream(t.device).cuda_stream


def require_cuda(*ts: Optional[torch.Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("lina_speech_b200 ops run on CUDA (B200) tensors only; got a tensor on "
                               f"{t.device}. There is no CPU fallback.")


_launches = 0


def count_launches(n: int) -> None:
    """Book-keeping for bench.py's gpu_launches field: kernels of OURS enqueued so far."""
    global _launches
    _launches += n


def launches() -> int:
    return _launches
