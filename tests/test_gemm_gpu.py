"""lina_gemm_bf16_terms (csrc/gemm_sm100.cu) against fp64 torch: plain bf16, the 3-term / 6-term split products that carry
fp32 tensors on bf16 tensor cores, convolution taps (vs F.conv1d), batched B operands, ragged M / N / K, epilogues."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel(got, ref):
    return ((got.double().cpu() - ref.double().cpu()).abs().max() / ref.double().abs().max()).item()


@pytest.mark.parametrize("NB,Ln,N,K", [(1, 128, 256, 64), (2, 300, 768, 768), (3, 750, 1282, 200), (1, 77, 40, 2304)])
def test_plain_bf16_product(NB, Ln, N, K):
    from lina_speech_b200.codec import gemm as G
    torch.manual_seed(NB + Ln)
    Kp = (K + 7) // 8 * 8
    a = torch.zeros(NB, Ln, Kp).bfloat16(); a[..., :K] = torch.randn(NB, Ln, K).bfloat16()
    b = torch.zeros(N, Kp).bfloat16(); b[:, :K] = torch.randn(N, K).bfloat16()
    ref = a[..., :K].double() @ b[:, :K].double().t()
    out, _ = G.gemm_terms((a.to(DEV),), (b.to(DEV),), NB=NB, Ln=Ln, N=N, K=K)
    assert out.shape == (NB, Ln, N)
    assert _rel(out, ref) < 5e-6                      # exact products, fp32 accumulation over up to 2304 terms


@pytest.mark.parametrize("parts,tol", [(2, 3e-5), (3, 1.5e-6)])
@pytest.mark.parametrize("NB,Ln,N,K", [(2, 750, 768, 768), (1, 200, 2304, 768), (2, 130, 768, 2304)])
def test_split_products_carry_fp32_tensors(parts, tol, NB, Ln, N, K):
    from lina_speech_b200.codec import gemm as G
    torch.manual_seed(7)
    x, w = torch.randn(NB, Ln, K), torch.randn(N, K) / K ** 0.5
    bias, gamma, res = torch.randn(N), torch.rand(N) + 0.5, torch.randn(NB, Ln, N)
    ref = F.gelu(x.double() @ w.double().t() + bias.double()) * gamma.double() + res.double()
    out, sp = G.gemm_terms(G.split(x.to(DEV), parts), G.split(w.to(DEV), parts), NB=NB, Ln=Ln, N=N, K=K, bias=bias.to(DEV),
                           gamma=gamma.to(DEV), residual=res.to(DEV), act="gelu", out_parts=parts)
    assert _rel(out, ref) < tol
    rec = sum(p.float() for p in sp)[..., :N]         # the split written by the epilogue re-assembles the fp32 result
    assert _rel(rec, out) < (3e-5 if parts == 2 else 2e-7)
    fp32 = F.gelu(x.to(DEV) @ w.to(DEV).t() + bias.to(DEV)) * gamma.to(DEV) + res.to(DEV)
    print(f"parts={parts} {NB}x{Ln}x{N}x{K}: rel err {_rel(out, ref):.2e} (torch fp32 matmul: {_rel(fp32, ref):.2e})")


@pytest.mark.parametrize("taps,Ci,Co,Ln", [(3, 768, 768, 750), (7, 512, 768, 333), (1, 64, 40, 50)])
def test_conv_taps_equal_conv1d(taps, Ci, Co, Ln):
    """nn.Conv1d(Ci, Co, k, padding=k//2) on [B, Ci, L] == the tap-term GEMM on the channels-last [B, L, Ci] tensor with the
    weight laid out [Co][tap][Ci]; rows outside the sequence are TMA zero fill (per batch)."""
    from lina_speech_b200.codec import gemm as G
    torch.manual_seed(taps)
    B = 3
    x = torch.randn(B, Ci, Ln)
    conv = torch.nn.Conv1d(Ci, Co, taps, padding=taps // 2)
    ref = conv.double()(x.double()).transpose(1, 2)                                    # [B, L, Co]
    w = conv.weight.detach().float().permute(0, 2, 1).reshape(Co, taps * Ci)
    xcl = x.transpose(1, 2).contiguous()
    out, _ = G.gemm_terms(G.split(xcl.to(DEV), 3), G.split(w.to(DEV), 3), NB=B, Ln=Ln, N=Co, K=Ci, taps=taps, pad=taps // 2,
                          bias=conv.bias.detach().float().to(DEV))
    print(f"conv taps={taps}: rel err {_rel(out, ref):.2e}")
    assert _rel(out, ref) < 1.5e-6


def test_batched_b_operand_and_alpha():
    """the attention block's products: S = alpha q k^T per batch (both operands activations), then P V with V^T as B."""
    from lina_speech_b200.codec import gemm as G
    torch.manual_seed(1)
    B, Ln, Cc = 3, 750, 768
    qkv = torch.randn(B, Ln, 3 * Cc)
    parts = G.split(qkv.to(DEV), 2)
    q = tuple(p[..., :Cc] for p in parts)
    k = tuple(p[..., Cc:2 * Cc] for p in parts)
    ref = (qkv[..., :Cc].double() @ qkv[..., Cc:2 * Cc].double().transpose(1, 2)) * Cc ** -0.5
    S, _ = G.gemm_terms(q, k, NB=B, Ln=Ln, N=Ln, K=Cc, b_batched=True, alpha=Cc ** -0.5, lda=3 * Cc, ldb=3 * Cc,
                        a_batch_stride=Ln * 3 * Cc, b_batch_stride=Ln * 3 * Cc)
    assert _rel(S, ref) < 3e-5
    # P V: K = Ln = 750 is not a multiple of 64 (TMA zero fill past K), row stride padded to 752
    Lp = 752
    P = torch.softmax(ref.float(), dim=-1)
    Pp = torch.zeros(B, Ln, Lp); Pp[..., :Ln] = P
    Vt = torch.zeros(B, Cc, Lp); Vt[..., :Ln] = qkv[..., 2 * Cc:].transpose(1, 2)
    ref2 = P.double() @ qkv[..., 2 * Cc:].double()
    O, _ = G.gemm_terms(G.split(Pp.to(DEV), 2), G.split(Vt.to(DEV), 2), NB=B, Ln=Ln, N=Cc, K=Ln, b_batched=True)
    assert _rel(O, ref2) < 3e-5
    # the same product with V read where the qkv GEMM left it: [B, L (= K), 3C] rows, N contiguous ("MN-major" B operand)
    v = tuple(p[..., 2 * Cc:] for p in parts)
    O2, sp = G.gemm_terms(G.split(Pp.to(DEV), 2), v, NB=B, Ln=Ln, N=Cc, K=Ln, b_batched=True, b_mn=True, ldb=3 * Cc,
                          b_batch_stride=Ln * 3 * Cc, out_parts=2)
    assert _rel(O2, ref2) < 3e-5
    assert _rel(sum(p.float() for p in sp), O2) < 3e-5


def test_promotion_span_controls_the_accumulation_error():
    """The tensor core's fp32 accumulator truncates after every K = 16 step; partial sums are promoted to fp32 registers every
    ``span`` K blocks.  With one span for the whole K loop the error is the raw tensor-core one (~1e-5 at K = 3 x 2304), with
    the default span it is that of an fp32 SIMT GEMM."""
    from lina_speech_b200.codec import gemm as G
    torch.manual_seed(3)
    NB, Ln, N, K = 1, 256, 512, 2304
    x, w = torch.randn(NB, Ln, K), torch.randn(N, K) / K ** 0.5
    ref = x.double() @ w.double().t()
    xs, ws = G.split(x.to(DEV), 3), G.split(w.to(DEV), 3)
    errs = {}
    for span in (1, 2, 3, 4, 6, 8, 12, 1000):
        out, _ = G.gemm_terms(xs, ws, NB=NB, Ln=Ln, N=N, K=K, span=span)
        errs[span] = _rel(out, ref)
    fp32 = _rel(x.to(DEV) @ w.to(DEV).t(), ref)
    print(f"relative error by promotion span: {errs}; torch fp32 matmul {fp32:.2e}")
    assert errs[2] < 1.5e-6 and errs[1] < 1.5e-6
    assert errs[1000] > 3 * errs[2]


def test_bad_arguments_are_refused():
    from lina_speech_b200.codec import gemm as G
    a = torch.zeros(1, 8, 12, dtype=torch.bfloat16, device=DEV)          # row stride 12: not a multiple of 8
    b = torch.zeros(8, 12, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(RuntimeError, match="multiples of 8"):
        G.gemm_terms((a,), (b,), NB=1, Ln=8, N=8, K=12)
