"""CPU: the oracle (oracle/*.py) against the golden fixtures generated from the reference
(tests/golden/make_golden.py), plus internal consistency of its restatements."""
import pytest
import math

import torch
import torch.nn.functional as F

from oracle import gla_oracle as GO, lina_oracle as LO, codec_oracle as CO


def _case(g, ci):
    p = f"c{ci}_"
    return {k[len(p):]: v for k, v in g.items() if k.startswith(p)}


def test_recurrence_matches_reference_goldens(golden_ops):
    for ci in range(int(golden_ops["n_cases"])):
        c = _case(golden_ops, ci)
        o, ht = GO.recurrent_gla(c["q"], c["k"], c["v"], c["gk"], initial_state=c.get("h0"))
        assert torch.allclose(o, c["o"], atol=1e-5, rtol=1e-5), ci
        assert torch.allclose(ht, c["ht"], atol=1e-4, rtol=1e-5), ci


def test_backward_identities_match_reference_autograd(golden_ops):
    for ci in range(int(golden_ops["n_cases"])):
        c = _case(golden_ops, ci)
        dq, dk, dv, dgk, dh0 = GO.recurrent_gla_bwd(c["q"], c["k"], c["v"], c["gk"], c.get("h0"), c["do"], c["dht"])
        for name, got in (("dq", dq), ("dk", dk), ("dv", dv), ("dgk", dgk)):
            ref = c[name]
            assert torch.allclose(got.float(), ref, atol=2e-3 * max(1.0, ref.abs().max().item()), rtol=1e-4), (ci, name)
        if "dh0" in c:
            assert torch.allclose(dh0.float(), c["dh0"], atol=1e-3, rtol=1e-4), ci


@pytest.mark.parametrize("chunk", [16, 64])
def test_chunk_form_equals_recurrence(golden_ops, chunk):
    for ci in (1, 3, 4, 6):
        c = _case(golden_ops, ci)
        if ci == 6 and chunk == 64:
            continue      # two -20 resets inside one 64-chunk: exp(-G) overflows a single-pivot chunk by design
        o, ht = GO.chunk_gla(c["q"].double(), c["k"].double(), c["v"].double(), c["gk"].double(),
                             initial_state=c.get("h0"), chunk=chunk, acc_dtype=torch.float64)
        assert torch.allclose(o.float(), c["o"], atol=1e-4, rtol=1e-4), ci
        assert torch.allclose(ht, c["ht"], atol=1e-4, rtol=1e-4), ci


def test_continuation_equals_one_shot():
    torch.manual_seed(0)
    B, H, T, K, V = 2, 2, 40, 16, 32
    q, k, v = torch.randn(B, H, T, K), torch.randn(B, H, T, K), torch.randn(B, H, T, V)
    gk = F.logsigmoid(torch.randn(B, H, T, K)) / 16
    o, ht = GO.recurrent_gla(q, k, v, gk)
    o1, h1 = GO.recurrent_gla(q[:, :, :23], k[:, :, :23], v[:, :, :23], gk[:, :, :23])
    o2, h2 = GO.recurrent_gla(q[:, :, 23:], k[:, :, 23:], v[:, :, 23:], gk[:, :, 23:], initial_state=h1)
    assert torch.allclose(torch.cat([o1, o2], 2), o, atol=1e-5)
    assert torch.allclose(h2, ht, atol=1e-5)


def test_short_conv_step_equals_prefill():
    torch.manual_seed(0)
    B, Ln, D, W = 2, 11, 8, 4
    x, w = torch.randn(B, Ln, D), torch.randn(D, W)
    y = GO.short_conv_prefill(x, w)
    cache = torch.zeros(B, D, W)
    ys = torch.cat([GO.short_conv_step(x[:, t:t + 1], cache, w) for t in range(Ln)], 1)
    assert torch.allclose(y, ys, atol=1e-6)
    cache2 = torch.zeros(B, D, W)
    GO.short_conv_prefill(x, w, cache2)
    assert torch.equal(cache, cache2)
    cache3 = torch.ones(B, D, W)
    GO.short_conv_prefill(x[:, :2], w, cache3)          # L < W: zero left-padded
    assert torch.equal(cache3[:, :, :2], torch.zeros(B, D, 2)) and torch.equal(cache3[:, :, 2:], x[:, :2].transpose(1, 2))


def test_layer_cfg1_matches_reference(golden_layer):
    g = golden_layer
    for p, sc in (("nosc_", False), ("sc_", True)):
        sd = {"l." + k[len(p) + 2:]: v for k, v in g.items() if k.startswith(p + "w.")}
        x = g[p + "x"]
        y = LO.gla_layer(sd, "l", x, 4, None, use_short_conv=sc)
        assert torch.allclose(y, g[p + "y"], atol=2e-5), p
        st = LO.init_state({"d_model": 256, "heads": 4, "n_layer": 0}, 1)[0]
        st = st if sc else st[-1:]
        yp = LO.gla_layer(sd, "l", x[:, :100], 4, st, use_short_conv=sc)
        ys = torch.cat([LO.gla_layer(sd, "l", x[:, t:t + 1], 4, st, use_short_conv=sc) for t in range(100, 128)], 1)
        assert torch.allclose(yp, g[p + "y_pre"], atol=2e-5) and torch.allclose(ys, g[p + "y_steps"], atol=2e-5)
        for i, s in enumerate(st):
            assert torch.allclose(s, g[p + f"state{i}"], atol=1e-4)


CFG_TINY = {"d_model": 64, "n_layer": 2, "heads": 2, "txt_heads": 2, "pos_type": "convolutional"}


def test_tiny_model_matches_reference(golden_model):
    g = golden_model
    sd = {k[2:]: v for k, v in g.items() if k.startswith("w.")}
    logits, loss, att = LO.lina_forward(sd, CFG_TINY, g["x"], g["y"], g["enc_mask"], g["ca_mask"], g["y_mask"])
    assert torch.allclose(logits, g["logits"], atol=1e-4)
    assert torch.allclose(loss, g["loss"], atol=1e-5)
    assert torch.allclose(att, g["att"], atol=1e-5)
    qs, atts, step_logits = LO.lina_generate_greedy(sd, CFG_TINY, g["xt"], 3, g["prompt"], 24)
    assert torch.equal(qs, g["qs"])                       # bit-exact greedy token ids
    assert torch.allclose(atts, g["atts"], atol=1e-5)


def test_codec_matches_reference(golden_codec):
    g = golden_codec
    sd = {k[2:]: v for k, v in g.items() if k.startswith("w.")}
    for Ln in (1, 7, 40):
        codes, bw = g[f"L{Ln}_codes"], g[f"L{Ln}_bw"]
        wav = CO.decode(sd, CO.codes_to_features(sd, codes), bw)
        assert wav.shape == (codes.shape[1], 320 * Ln)
        assert torch.allclose(wav, g[f"L{Ln}_wav"], atol=1e-5, rtol=1e-5)


def test_rwkv6_recurrence_matches_reference(request):
    from conftest import load_golden
    g = load_golden("rwkv6_ops.npz")
    for ci in range(int(g["n_cases"])):
        c = _case(g, ci)
        o, ht = GO.recurrent_rwkv6(c["r"], c["k"], c["v"], c["w"], c["u"], initial_state=c.get("h0"))
        assert torch.allclose(o, c["o"], atol=1e-5, rtol=1e-5) and torch.allclose(ht, c["ht"], atol=1e-4, rtol=1e-5)


def test_random_state_dict_has_the_host_classes_keys_and_shapes():
    """bench.py's CPU arm builds its model from oracle.lina_oracle.random_state_dict alone; it must describe the same
    architecture as the host mirror of LinaModel (model/modeling_lina.py:24-58) and run through lina_forward."""
    import torch
    from oracle import lina_oracle as LO
    import lina_speech_b200.model as m
    d, nl, h = 64, 2, 2
    rnn = m.AttentiveGLA(d, nl, h, blind=True, use_short_conv=True, pos_type="convolutional")
    lm = m.LinaModel(rnn, d, 1, 4096, 3, 3, 256, txt_encoder=m.TextEncoder(d, 2, n_layers=1, dropout=0.0, rotary=False))
    cfg = {"d_model": d, "n_layer": nl, "heads": h, "txt_heads": 2, "txt_layers": 1, "pos_type": "convolutional"}
    sd = LO.random_state_dict(cfg, 4096, 3, 256)
    assert {k: tuple(v.shape) for k, v in sd.items()} == {k: tuple(v.shape) for k, v in lm.state_dict().items()}
    x = torch.randint(3, 256, (2, 9))
    y = torch.randint(3, 4099, (2, 12, 1))
    em = torch.ones(2, 9, 9, dtype=torch.bool)
    cm = torch.ones(2, 12, 9, dtype=torch.bool)
    logits, loss = LO.lina_forward(sd, cfg, x, y, em, cm)[:2]
    assert torch.isfinite(logits).all() and torch.isfinite(loss)


def test_reference_rounding_emulation_reproduces_the_references_triton_kernels(golden_triton):
    """oracle.gla_oracle.fused_chunk_gla_as_reference_rounds restates every rounding of the reference's default op
    (FLA/fla/ops/gla/chunk_fuse.py:302-399).  The fixture holds what the reference's Triton kernels actually returned on a
    B200 for seeded bf16 inputs: the restatement must reproduce it -- bit for bit on (almost) every element, the rest (fp32
    summation order inside tl.dot flipping a rounding) within one bf16 ulp at the output scale, final state to fp32 rounding."""
    from conftest import triton_golden_inputs
    from oracle import gla_oracle as GO
    for tag in "abc":
        q, k, v, gk = triton_golden_inputs(golden_triton[f"{tag}_shape"])
        emu, emu_h = GO.fused_chunk_gla_as_reference_rounds(q, k, v, gk)
        ref = golden_triton[f"{tag}_fused_chunk_gla_o"]
        assert emu.dtype == ref.dtype == torch.bfloat16 and emu.shape == ref.shape
        same = (emu.view(torch.int16) == ref.view(torch.int16)).float().mean().item()
        assert same >= 0.995, f"{tag}: only {same:.4f} of the elements are bit-identical"
        d = (emu.float() - ref.float()).abs().max().item()
        ulp_top = 2.0 ** (math.floor(math.log2(ref.float().abs().max().item())) - 7)   # inter + intra are rounded separately:
        assert d <= ulp_top, f"{tag}: max difference {d:.3e} exceeds one bf16 ulp at the output scale ({ulp_top:.3e})"
        if f"{tag}_fused_chunk_gla_ht" in golden_triton:
            hd = (emu_h - golden_triton[f"{tag}_fused_chunk_gla_ht"]).abs().max().item()
            assert hd <= 1e-5, f"{tag}: final state differs by {hd:.2e}"
