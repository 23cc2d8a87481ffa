"""lina_skinny_linear (csrc/skinny_linear.cu): the decode step's weight-streaming linears with fused add + LayerNorm prologue and
SwiGLU epilogue, against the unfused sequence in torch (same bf16 roundings, fp32 accumulation)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"
BF = torch.bfloat16


def _call(x, W, *, delta=None, norm=None, bias=None, N=None, pair=0):
    from lina_speech_b200 import _lib as L
    M, K = x.shape
    N = N if N is not None else W.shape[0]
    out = torch.empty(M, N, dtype=BF, device=DEV)
    s_out = torch.empty_like(x) if delta is not None else None
    rc = L.lib().lina_skinny_linear(L.ptr(x), x.stride(0), L.ptr(delta), delta.stride(0) if delta is not None else 0,
                                    L.ptr(norm[0]) if norm else None,
                                    L.ptr(norm[1]) if norm else None, norm[2] if norm else 0.0, L.ptr(s_out), L.ptr(W), W.stride(0),
                                    L.ptr(bias), L.ptr(out), out.stride(0), M, N, K, pair, L.stream(x))
    L.check(rc, "lina_skinny_linear")
    return out, s_out


def _close_bf16(got, ref, what):
    """within one bf16 ulp of the fp32-accumulated reference (+ a small absolute term for values near zero)"""
    got, ref = got.float().cpu(), ref.float().cpu()
    tol = ref.abs() * 2 ** -7 + 2e-3 * ref.abs().max()
    bad = (got - ref).abs() > tol
    assert not bool(bad.any()), f"{what}: {int(bad.sum())} elements off, max diff {(got - ref).abs().max():.3e}"


@pytest.mark.parametrize("M", [1, 5, 32])
@pytest.mark.parametrize("N,K", [(6160, 1024), (1024, 2048), (1024, 1368), (2736, 1024), (40, 64)])
def test_plain_linear(M, N, K):
    torch.manual_seed(M + N)
    x = torch.randn(M, K).to(BF).to(DEV)
    W = (torch.randn(N, K) / K ** 0.5).to(BF).to(DEV)
    b = torch.randn(N).to(BF).to(DEV)
    out, _ = _call(x, W, bias=b)
    ref = (x.float() @ W.float().t() + b.float())
    _close_bf16(out, ref, f"plain {M}x{N}x{K}")
    out2, _ = _call(x, W)
    _close_bf16(out2, x.float() @ W.float().t(), "plain, no bias")


@pytest.mark.parametrize("M", [2, 32])
@pytest.mark.parametrize("with_delta", [False, True])
def test_layernorm_prologue_matches_add_layernorm_then_linear(M, with_delta):
    from lina_speech_b200.model.base_blocks import add_layernorm
    torch.manual_seed(M)
    K, N = 1024, 6160
    ln = torch.nn.LayerNorm(K).to(DEV).to(BF)
    with torch.no_grad():
        ln.weight.uniform_(0.5, 1.5)
        ln.bias.normal_(0, 0.2)
    x = torch.randn(M, K).to(BF).to(DEV)
    d = torch.randn(M, K).to(BF).to(DEV) if with_delta else None
    W = (torch.randn(N, K) / K ** 0.5).to(BF).to(DEV)
    s_ref, h_ref = add_layernorm(d, x, ln)                       # the unfused kernel (same rounding points)
    out, s = _call(x, W, delta=d, norm=(ln.weight, ln.bias, float(ln.eps)))
    if with_delta:
        assert torch.equal(s, s_ref)
    _close_bf16(out, h_ref.float() @ W.float().t(), f"LN prologue M={M} delta={with_delta}")


@pytest.mark.parametrize("M", [3, 32])
def test_swiglu_epilogue(M):
    torch.manual_seed(7)
    K, hp = 1024, 1368
    x = torch.randn(M, K).to(BF).to(DEV)
    W = (torch.randn(2 * hp, K) / K ** 0.5).to(BF).to(DEV)
    b = (0.1 * torch.randn(2 * hp)).to(BF).to(DEV)
    out, _ = _call(x, W, bias=b, N=hp, pair=hp)
    h = (x.float() @ W.float().t() + b.float()).to(BF).float()   # the linear's bf16 output, as lina_swiglu_act reads it
    ref = F.silu(h[:, :hp]) * h[:, hp:]
    _close_bf16(out, ref, f"swiglu M={M}")


def test_block_step_equals_the_unfused_step():
    """MixingBlock single-token step through the skinny path == the same block with LINA_SKINNY_STEP off (library GEMMs +
    add_layernorm + swiglu_act), to bf16 rounding; the caches end identical to rounding as well."""
    import lina_speech_b200.model as M
    import lina_speech_b200.model.base_blocks as BB
    torch.manual_seed(0)
    rnn = M.AttentiveGLA(1024, 1, 4, blind=True, use_short_conv=True, pos_type="convolutional").to(DEV).to(BF).eval()
    B = 4
    x = torch.randn(B, 6, 1024).to(BF).to(DEV)
    ctx = torch.randn(B, 9, 1024).to(BF).to(DEV)
    outs = {}
    saved = BB.SKINNY_STEP
    try:
        for name, flag in (("skinny", True), ("unfused", False)):
            BB.SKINNY_STEP = flag
            cache = rnn.init_state(batch_size=B)
            ys = []
            with torch.inference_mode():
                for t in range(x.shape[1]):
                    y, _, _ = rnn.step(x[:, t:t + 1], ctx, t, cache)
                    ys.append(y)
            outs[name] = (torch.cat(ys, 1).float().cpu(), [s[-1].float().cpu() for s in cache.states])
    finally:
        BB.SKINNY_STEP = saved
    a, b = outs["skinny"], outs["unfused"]
    scale = b[0].abs().max().item()
    assert (a[0] - b[0]).abs().max().item() <= 3e-2 * scale
    for sa, sb in zip(a[1], b[1]):
        assert (sa - sb).abs().max().item() <= 3e-2 * max(sb.abs().max().item(), 1e-3)
