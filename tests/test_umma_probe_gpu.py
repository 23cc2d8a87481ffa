"""GPU: pins the tcgen05 descriptor conventions of csrc/sm100.cuh (smem K-major / MN-major operands in the
core-matrix-interleave layout, A operand from TMEM, M=128 accumulator layout) on real hardware."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(N, KD, a_mode, b_mode, swap):
    from lina_speech_b200 import _lib as L
    torch.manual_seed(N + KD + a_mode * 7 + b_mode * 3)
    A = torch.randn(128, KD, device="cuda")
    B = torch.randn(N, KD, device="cuda")
    D = torch.zeros(128, N, device="cuda")
    rc = L.debug_lib().lina_debug_umma_probe(L.ptr(A), L.ptr(B), L.ptr(D), N, KD, a_mode, b_mode, swap, L.stream(A))
    L.check(rc, "lina_debug_umma_probe")
    torch.cuda.synchronize()
    ref = A.bfloat16().float() @ B.bfloat16().float().t()
    return (D - ref).abs().max().item(), ref.abs().max().item()


@pytest.mark.parametrize("a_mode", [0, 1, 2])
@pytest.mark.parametrize("b_mode", [0, 1])
@pytest.mark.parametrize("N,KD", [(64, 64), (256, 64), (64, 128)])
def test_descriptor_conventions(a_mode, b_mode, N, KD):
    # (the swapped convention was tried once on hardware: wrong results at N=64, out-of-bounds smem reads at N=256)
    err, mag = _run(N, KD, a_mode, b_mode, 0)
    print(f"a_mode={a_mode} b_mode={b_mode} N={N} KD={KD}: err={err:.3e} |ref|max={mag:.1f}")
    assert err < 1e-3 * mag, f"convention in sm100.cuh is wrong for a_mode={a_mode} b_mode={b_mode}"


def _run_sw(N, KD, a_mode, b_mode, use_tma):
    from lina_speech_b200 import _lib as L
    torch.manual_seed(N + KD + a_mode * 7 + b_mode * 3 + use_tma)
    A = torch.randn(128, KD, device="cuda")
    B = torch.randn(N, KD, device="cuda")
    D = torch.zeros(128, N, device="cuda")
    Ab = A.bfloat16().contiguous()
    rc = L.debug_lib().lina_debug_umma_probe_sw128(L.ptr(A), L.ptr(B), L.ptr(D), L.ptr(Ab), N, KD, a_mode, b_mode, use_tma,
                                             L.stream(A))
    L.check(rc, "lina_debug_umma_probe_sw128")
    torch.cuda.synchronize()
    ref = A.bfloat16().float() @ B.bfloat16().float().t()
    return (D - ref).abs().max().item(), ref.abs().max().item()


@pytest.mark.parametrize("a_mode,b_mode,use_tma", [(0, 0, 0), (0, 1, 0), (1, 0, 0), (1, 1, 0), (0, 0, 1), (0, 1, 1)])
@pytest.mark.parametrize("N,KD", [(64, 64), (256, 128)])
def test_sw128_descriptor_conventions_and_tma(a_mode, b_mode, use_tma, N, KD):
    err, mag = _run_sw(N, KD, a_mode, b_mode, use_tma)
    print(f"sw128 a_mode={a_mode} b_mode={b_mode} tma={use_tma} N={N} KD={KD}: err={err:.3e} |ref|max={mag:.1f}")
    assert err < 1e-3 * mag
