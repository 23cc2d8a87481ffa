"""GPU, BASELINE-size shapes (head dims K=256, V=512, T=2048, bf16): size-independent properties of the hot path --
the oracle is too slow here, so parity rests on identities the function must satisfy."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"
B, H, T, K, V = 4, 4, 2048, 256, 512


def _inputs(seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    q, k = (torch.randn(B, H, T, K, device=DEV, generator=g).bfloat16() for _ in range(2))
    v = torch.randn(B, H, T, V, device=DEV, generator=g).bfloat16()
    gk = (F.logsigmoid(torch.randn(B, H, T, K, device=DEV, generator=g)) / 16).bfloat16()
    return q, k, v, gk


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def test_tensor_core_kernel_agrees_with_the_recurrence_kernel():
    """two independent CUDA implementations (tcgen05 chunk form vs CUDA-core recurrence) of the same function."""
    from lina_speech_b200.fla_api import fused_chunk_gla, fused_recurrent_gla
    q, k, v, gk = _inputs()
    o1, h1 = fused_chunk_gla(q, k, v, gk, output_final_state=True)
    o2, h2 = fused_recurrent_gla(q, k, v, gk, output_final_state=True)
    assert torch.isfinite(o1).all() and torch.isfinite(h1).all()
    assert _rel(o1, o2) < 1e-2 and _rel(h1, h2) < 5e-3
    assert (o1.float() - o2.float()).abs().max() <= 3e-2 * o2.float().abs().max()


def test_prefix_continuation_at_full_length():
    """state hand-off: [0,1000) then [1000,2048) with the carried state == one shot (prefill -> decode contract)."""
    from lina_speech_b200.fla_api import chunk_gla
    q, k, v, gk = _inputs(1)
    o, ht = chunk_gla(q, k, v, gk, output_final_state=True)
    o1, h1 = chunk_gla(q[:, :, :1000], k[:, :, :1000], v[:, :, :1000], gk[:, :, :1000], output_final_state=True)
    o2, h2 = chunk_gla(q[:, :, 1000:], k[:, :, 1000:], v[:, :, 1000:], gk[:, :, 1000:], initial_state=h1,
                       output_final_state=True)
    assert torch.equal(o1, o[:, :, :1000].contiguous())               # identical chunks before the cut (1000 % 64 != 0 only affects the tail item)
    assert _rel(torch.cat([o1, o2], 2), o) < 5e-3 and _rel(h2, ht) < 5e-3      # the cut shifts the chunk grid: bf16 operand rounding differs


def test_linearity_in_v_and_in_the_initial_state():
    """o is linear in (v, S0) for fixed q, k, gk:  f(v1+v2, S1+S2) = f(v1,S1) + f(v2,S2)."""
    from lina_speech_b200.fla_api import fused_chunk_gla
    q, k, v1, gk = _inputs(2)
    v2 = torch.randn_like(v1)
    s1, s2 = torch.randn(B, H, K, V, device=DEV), torch.randn(B, H, K, V, device=DEV)
    oa, ha = fused_chunk_gla(q, k, v1, gk, initial_state=s1, output_final_state=True)
    ob, hb = fused_chunk_gla(q, k, v2, gk, initial_state=s2, output_final_state=True)
    oc, hc = fused_chunk_gla(q, k, (v1.float() + v2.float()).bfloat16(), gk, initial_state=s1 + s2, output_final_state=True)
    assert _rel(oa.float() + ob.float(), oc) < 1e-2 and _rel(ha + hb, hc) < 5e-3


def test_step_loop_equals_chunk_forward_at_flagship_dims():
    """64 decode steps of the fused step kernel (bf16 cache) == the chunk kernel over the same 64 tokens, layer level."""
    from lina_speech_b200.model import GatedLinearAttention
    from lina_speech_b200.fla_api import Cache
    torch.manual_seed(0)
    layer = GatedLinearAttention(hidden_size=1024, num_heads=4, use_short_conv=True, layer_idx=0).to(DEV).bfloat16().eval()
    x = torch.randn(2, 64, 1024, device=DEV).bfloat16()
    with torch.inference_mode():
        y = layer(x)
        cache = Cache()
        cache.update(layer.init_state(2), 0, offset=0)
        ys = torch.cat([layer(x[:, t:t + 1], past_key_values=cache, use_cache=True) for t in range(64)], 1)
    assert _rel(ys, y) < 3e-2


def test_decode_state_kernel_bytes_are_what_the_roofline_assumes():
    """the cache really is bf16 [B,4,256,512] per block and the step updates it in place (no reallocation)."""
    import lina_speech_b200.model as m
    rnn = m.AttentiveGLA(1024, 1, 4, blind=True, use_short_conv=True, pos_type="convolutional").to(DEV).bfloat16().eval()
    cache = rnn.init_state(batch_size=8)
    S = cache[0][-1]
    assert S.dtype == torch.bfloat16 and tuple(S.shape) == (8, 4, 256, 512)
    ptrs = [t.data_ptr() for st in cache.states for t in st]
    with torch.inference_mode():
        y, att, _ = rnn.step(torch.randn(8, 1, 1024, device=DEV).bfloat16(), torch.randn(8, 16, 1024, device=DEV).bfloat16(), 0, cache)
    assert ptrs == [t.data_ptr() for st in cache.states for t in st] and S.abs().sum() > 0 and torch.isfinite(y).all()
