#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/*.npz from the REFERENCE itself.

Runs only in the build container, where /root/reference is mounted (it does not
exist on the GPU box, and nothing under tests/ reads it at test time).  The
reference's own Python is imported unmodified, with the four shims SURVEY.md
section 8c lists for running it on a CPU without Triton / causal-conv1d /
rotary-embedding-torch:

  1. ``rotary_embedding_torch`` stub module (imported, never called: rotary=False);
  2. ``causal_conv1d`` stub whose two functions are the in-tree torch branch of
     FLA/fla/modules/convolution.py:175-178,197-204;
  3. ``model.gla.{fused_recurrent,fused_chunk,chunk}_gla`` rebound to the
     reference's own ``naive_recurrent_gla`` (FLA/fla/ops/gla/naive.py:13-44);
  4. ``FusedRMSNormSwishGate.forward`` = ``rms_norm_ref(...) * o * sigmoid(o)``
     (FLA/fla/modules/fused_norm_gate.py:41-55,134-136).

While generating, every fixture is also checked against oracle/ (the CPU
restatement), which is what pins the oracle to the reference.

    python tests/golden/make_golden.py
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "3rdparty"))
sys.path.insert(0, os.path.join(REF, "3rdparty", "flash-linear-attention"))

# ---- shim 1
rot = types.ModuleType("rotary_embedding_torch")
rot.RotaryEmbedding = type("RotaryEmbedding", (), {"__init__": lambda self, *a, **k: None})
rot.apply_rotary_emb = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("rotary stub"))
sys.modules["rotary_embedding_torch"] = rot

# ---- shim 2
cc = types.ModuleType("causal_conv1d")


def _cc_fn(x, weight, bias=None, activation=None, **kw):
    W = weight.shape[-1]
    y = F.conv1d(F.pad(x, (W - 1, 0)), weight.unsqueeze(1), bias, groups=x.shape[1])
    return F.silu(y) if activation in ("silu", "swish") else y


def _cc_update(x, conv_state, weight, bias=None, activation=None, **kw):
    conv_state.copy_(torch.roll(conv_state, shifts=-1, dims=-1))
    conv_state[:, :, -1] = x
    y = torch.sum(conv_state * weight, dim=-1)
    if bias is not None:
        y = y + bias
    return F.silu(y) if activation in ("silu", "swish") else y


cc.causal_conv1d_fn, cc.causal_conv1d_update = _cc_fn, _cc_update
sys.modules["causal_conv1d"] = cc

import fla  # noqa: E402  (vendored 0.1 tree, not the site-packages one)

assert fla.__file__.startswith(REF), fla.__file__
from fla.ops.gla.naive import naive_recurrent_gla  # noqa: E402
from fla.modules import fused_norm_gate as fng  # noqa: E402
import model.gla as ref_gla  # noqa: E402
from model.modeling_lina import LinaModel  # noqa: E402
from model.encoder import TextEncoder  # noqa: E402


# ---- shim 3
def _naive(q, k, v, gk, scale=None, initial_state=None, output_final_state=False, **kw):
    return naive_recurrent_gla(q, k, v, gk, initial_state=initial_state, output_final_state=output_final_state)


ref_gla.fused_recurrent_gla = ref_gla.fused_chunk_gla = ref_gla.chunk_gla = _naive


# ---- shim 4
def _ng_forward(self, x, o, residual=None, prenorm=False, residual_in_fp32=False):
    return fng.rms_norm_ref(x, self.weight, None, eps=self.eps, upcast=True) * o * torch.sigmoid(o)


fng.FusedRMSNormSwishGate.forward = _ng_forward

from oracle import gla_oracle as GO, lina_oracle as LO, codec_oracle as CO  # noqa: E402


def save(name, **arrs):
    out = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrs.items()}
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **out)
    print(f"  wrote {name}: {os.path.getsize(path) / 1e6:.2f} MB")


def close(a, b, tol, what):
    err = (a.double() - b.double()).abs().max().item()
    assert err <= tol, f"oracle != reference for {what}: {err} > {tol}"
    return err


# ----------------------------------------------------------------------------
# 1. op level: FLA/tests/ops/test_gla.py:58-102 procedure (recurrent vs naive, fp32, h0, all grads)
# ----------------------------------------------------------------------------
def gen_ops():
    cases = [  # B, H, T, K, V, gate kind, use_h0
        (2, 2, 1, 32, 64, "model", True),
        (2, 2, 17, 32, 64, "fla", True),
        (1, 2, 64, 64, 64, "fla", False),
        (1, 2, 100, 64, 128, "model", True),
        (1, 4, 128, 64, 128, "model", False),     # BASELINE config 1 op shape (d256/h4 -> K64, V128)
        (1, 1, 66, 256, 512, "model", False),     # flagship head dims
        (1, 2, 48, 32, 32, "reset", True),        # rows of -20 (reset_val, model/gla.py:182-184)
    ]
    out = {}
    for ci, (B, H, T, K, V, kind, use_h0) in enumerate(cases):
        torch.manual_seed(42 + ci)
        q, k, v = torch.randn(B, H, T, K), torch.randn(B, H, T, K), torch.randn(B, H, T, V)
        raw = torch.randn(B, H, T, K)
        if kind == "model":
            gk = F.logsigmoid(raw) / 16
        elif kind == "fla":
            gk = F.logsigmoid(raw).clamp_min(-3)
        else:
            gk = F.logsigmoid(raw) / 16
            gk[:, :, T // 3] = -20.0
            gk[:, :, T // 3 + 1] = -20.0
        h0 = torch.randn(B, H, K, V) if use_h0 else None
        do, dht = torch.randn(B, H, T, V), torch.randn(B, H, K, V)
        leaves = [t.clone().requires_grad_(True) for t in (q, k, v, gk)]
        h0l = h0.clone().requires_grad_(True) if use_h0 else None
        o, ht = naive_recurrent_gla(*leaves, initial_state=h0l, output_final_state=True)
        ((o * do).sum() + (ht * dht).sum()).backward()
        grads = [t.grad for t in leaves] + ([h0l.grad] if use_h0 else [])
        # pin the oracle
        o2, ht2 = GO.recurrent_gla(q, k, v, gk, initial_state=h0)
        close(o, o2, 1e-5, f"ops[{ci}].o"); close(ht, ht2, 1e-4, f"ops[{ci}].ht")
        g2 = GO.recurrent_gla_bwd(q, k, v, gk, h0, do, dht)
        for n, a, b in zip(("dq", "dk", "dv", "dgk", "dh0"), grads, g2):
            close(a, b.float(), 2e-3 * max(1.0, a.abs().max().item()), f"ops[{ci}].{n}")
        o3, ht3 = GO.chunk_gla(q, k, v, gk, initial_state=h0, chunk=16 if kind != "model" else 64)
        close(o, o3, 2e-3, f"ops[{ci}].chunk.o"); close(ht, ht3, 2e-3 * max(1, ht.abs().max().item()), f"ops[{ci}].chunk.ht")
        p = f"c{ci}_"
        out.update({p + "q": q, p + "k": k, p + "v": v, p + "gk": gk, p + "do": do, p + "dht": dht,
                    p + "o": o, p + "ht": ht, p + "dq": grads[0], p + "dk": grads[1], p + "dv": grads[2],
                    p + "dgk": grads[3]})
        if use_h0:
            out.update({p + "h0": h0, p + "dh0": grads[4]})
    out["n_cases"] = len(cases)
    save("gla_ops.npz", **out)


# ----------------------------------------------------------------------------
# 2. layer level: BASELINE config 1 -- GatedLinearAttention d256 h4 T128 B1, CPU reference (fla off)
# ----------------------------------------------------------------------------
def gen_layer():
    out = {}
    for sc in (False, True):
        torch.manual_seed(0)
        layer = ref_gla.GatedLinearAttention(hidden_size=256, num_heads=4, expand_k=1.0, expand_v=2.0,
                                             use_short_conv=sc, layer_idx=0).eval()
        with torch.no_grad():
            layer.g_norm_swish_gate.weight.uniform_(0.5, 1.5)
        x = torch.randn(1, 128, 256)
        with torch.no_grad():
            y = layer(x)
            # prefill 100 tokens through the cache, then 28 single-token steps
            cache = ref_gla.Cache()
            cache.update(layer.init_state(1), 0, offset=0)
            y_pre = layer(x[:, :100], past_key_values=cache, use_cache=True)
            y_steps = torch.cat([layer(x[:, t:t + 1], past_key_values=cache, use_cache=True) for t in range(100, 128)], 1)
            final_state = [s.clone() for s in cache.states[0]]
        sd = {"l." + k: v for k, v in layer.state_dict().items()}
        st = LO.init_state({"d_model": 256, "heads": 4, "n_layer": 0}, 1)[0]
        if not sc:
            st = st[-1:]
        y2 = LO.gla_layer(sd, "l", x, 4, None, use_short_conv=sc)
        e = close(y, y2, 2e-5, f"layer(sc={sc})")
        y2p = LO.gla_layer(sd, "l", x[:, :100], 4, st, use_short_conv=sc)
        y2s = torch.cat([LO.gla_layer(sd, "l", x[:, t:t + 1], 4, st, use_short_conv=sc) for t in range(100, 128)], 1)
        close(y_pre, y2p, 2e-5, "layer prefill"); close(y_steps, y2s, 2e-5, "layer steps")
        for a, b in zip(final_state, st):
            close(a, b, 1e-4, "layer state")
        print(f"  layer sc={sc}: oracle vs reference max err {e:.2e}; prefill+step vs one-shot "
              f"{(torch.cat([y_pre, y_steps], 1) - y).abs().max():.2e}")
        p = "sc_" if sc else "nosc_"
        out.update({p + "x": x, p + "y": y, p + "y_pre": y_pre, p + "y_steps": y_steps})
        out.update({p + "state%d" % i: s for i, s in enumerate(final_state)})
        out.update({p + "w." + k: v for k, v in layer.state_dict().items()})
    save("gla_layer_cfg1.npz", **out)


# ----------------------------------------------------------------------------
# 3. model level: tiny LinaModel forward (teacher forced) + generate_batch greedy
# ----------------------------------------------------------------------------
def gen_model():
    torch.manual_seed(1)
    d, N, Hh = 64, 2, 2
    rnn = ref_gla.AttentiveGLA(d_model=d, n_layer=N, heads=Hh, blind=True, use_short_conv=True,
                               pos_type="convolutional")
    m = LinaModel(rnn, d_model=d, n_quant=1, n_codebook=64, n_special_token_in=3, n_special_token_out=3,
                  n_txt_vocab=32, txt_encoder=TextEncoder(d, 2, n_layers=1, dropout=0.0, rotary=False)).eval()
    cfg = {"d_model": d, "n_layer": N, "heads": Hh, "txt_heads": 2, "pos_type": "convolutional"}
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    B, Tx, T = 2, 9, 21
    x = torch.randint(3, 32, (B, Tx))
    y = torch.randint(3, 67, (B, T, 1))
    y[:, 0] = 1
    xlen, ylen = torch.tensor([9, 7]), torch.tensor([21, 16])
    xm = torch.arange(Tx)[None] < xlen[:, None]
    ym = torch.arange(T)[None] < ylen[:, None]
    enc_mask = xm.unsqueeze(1) & xm.unsqueeze(2)
    ca_mask = xm.unsqueeze(1) & ym.unsqueeze(2)
    ca_mask[:, :, 0] = True
    with torch.no_grad():
        logits, loss, att, _, _ = m(x.clone(), y, enc_mask, ca_mask, logits_mask=ym)
        l2, loss2, att2 = LO.lina_forward(sd, cfg, x, y, enc_mask, ca_mask, ym)
    close(logits, l2, 1e-4, "model.logits"); close(loss, loss2, 1e-5, "model.loss"); close(att, att2, 1e-5, "model.att")
    # forward with an initial state (initial-state tuning entry point, model/gla.py:315-325)
    with torch.no_grad():
        params = rnn.get_init_state_tuning_params(lora=1)
        st = rnn.get_state_from_params(params, B, scale=0.02)
        m.train()
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
        logits_s, loss_s, _, _, _ = m(x.clone(), y, enc_mask, ca_mask, logits_mask=ym, init_state=st)
        m.eval()
        ost = [tuple(t.clone() for t in s) for s in rnn.get_state_from_params(params, B, scale=0.02).states]
        l3, loss3, _ = LO.lina_forward(sd, cfg, x, y, enc_mask, ca_mask, ym, state=ost, training=True)
    close(logits_s, l3, 1e-4, "model.logits(init_state)")
    # greedy generation, prompt continuation (modeling_lina.py:112-192, k=1)
    xt = torch.randint(3, 32, (12,))
    prompt = torch.randint(0, 64, (1, 1, 5))
    qs, atts, stop_tokens, cuts = m.generate_batch(xt, batch_size=3, prompt=prompt, max_seqlen=24, k=1,
                                                   force_max_seqlen=True)
    qs2, atts2, step_logits = LO.lina_generate_greedy(sd, cfg, xt, 3, prompt, 24)
    assert torch.equal(qs, qs2), "greedy tokens differ between oracle and reference"
    close(atts, atts2, 1e-5, "generate.atts")
    print(f"  model: loss {loss.item():.5f}, greedy tokens identical ({qs.shape})")
    out = {"w." + k: v for k, v in sd.items()}
    out.update(x=x, y=y, enc_mask=enc_mask, ca_mask=ca_mask, y_mask=ym, logits=logits, loss=loss, att=att,
               logits_init_state=logits_s, loss_init_state=loss_s, xt=xt, prompt=prompt, qs=qs, atts=atts,
               step_logits=step_logits)
    out.update({f"tune_k{i}": p[0] for i, p in enumerate(params)})
    out.update({f"tune_v{i}": p[1] for i, p in enumerate(params)})
    save("lina_tiny.npz", **out)


# ----------------------------------------------------------------------------
# 4. codec: small WavTokenizer (real n_fft=1280 / hop=320 head) decode
# ----------------------------------------------------------------------------
def gen_codec():
    from decoder.pretrained import WavTokenizer
    from decoder.feature_extractors import EncodecFeatures
    from decoder.models import VocosBackbone
    from decoder.heads import ISTFTHead
    torch.manual_seed(2)
    bins = 64
    fe = EncodecFeatures(num_quantizers=1, dowmsamples=[8, 5, 4, 2], vq_bins=bins, vq_kmeans=10)
    bb = VocosBackbone(input_channels=512, dim=64, intermediate_dim=128, num_layers=2, adanorm_num_embeddings=4)
    head = ISTFTHead(dim=64, n_fft=1280, hop_length=320, padding="same")
    wt = WavTokenizer(fe, bb, head).eval()
    with torch.no_grad():
        fe.encodec.quantizer.vq.layers[0]._codebook.embed.normal_()
        for p in list(bb.parameters()) + list(head.parameters()):
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
            else:
                p.mul_(8.0)       # the 0.02-std init gives near-silent output; make every stage matter
        bb.norm.scale.weight.add_(0.2 * torch.randn(4, 64)); bb.norm.shift.weight.add_(0.2 * torch.randn(4, 64))
        for blk in bb.convnext:
            blk.norm.scale.weight.add_(0.2 * torch.randn(4, 64)); blk.norm.shift.weight.add_(0.2 * torch.randn(4, 64))
    keep = ("backbone.", "head.", "feature_extractor.encodec.quantizer.vq.layers.0._codebook.embed")
    sd = {k: v.detach().clone() for k, v in wt.state_dict().items() if k.startswith(keep)}
    out = {"w." + k: v for k, v in sd.items()}
    for L, B in ((1, 1), (7, 2), (40, 2)):
        codes = torch.randint(0, bins, (1, B, L))
        bw = torch.tensor([1])
        with torch.no_grad():
            feats = wt.codes_to_features(codes)
            wav = wt.decode(feats, bandwidth_id=bw)
        f2 = CO.codes_to_features(sd, codes)
        w2 = CO.decode(sd, f2, bw)
        assert wav.shape == (B, 320 * L)
        close(feats, f2, 0, "codec.features"); e = close(wav, w2, 1e-5 * max(1.0, wav.abs().max().item()), "codec.wav")
        print(f"  codec L={L} B={B}: wav absmax {wav.abs().max():.3f}, oracle err {e:.2e}")
        out.update({f"L{L}_codes": codes, f"L{L}_wav": wav, f"L{L}_bw": bw})
    save("codec_small.npz", **out)


# ----------------------------------------------------------------------------
# 5. RWKV6 recurrence (secondary row a13): FLA/fla/ops/rwkv6/recurrent_naive.py
# ----------------------------------------------------------------------------
def gen_rwkv6():
    from fla.ops.rwkv6.recurrent_naive import naive_recurrent_rwkv6
    out = {}
    for ci, (B, H, T, K, V, use_h0) in enumerate([(2, 2, 1, 32, 64, True), (1, 3, 37, 64, 64, False), (2, 4, 70, 64, 128, True)]):
        torch.manual_seed(77 + ci)
        r, k, v = torch.randn(B, H, T, K), torch.randn(B, H, T, K), torch.randn(B, H, T, V)
        w = -torch.exp(torch.randn(B, H, T, K) - 1.0)                    # RWKV6 decays: w = -exp(.)  (FLA/fla/layers/rwkv6.py)
        u = torch.randn(H, K)
        h0 = torch.randn(B, H, K, V) if use_h0 else None
        o, ht = naive_recurrent_rwkv6(r, k, v, w, u, initial_state=h0.clone() if use_h0 else None, output_final_state=True)
        o2, ht2 = GO.recurrent_rwkv6(r, k, v, w, u, initial_state=h0)
        close(o, o2, 1e-5, f"rwkv6[{ci}].o"); close(ht, ht2, 1e-4, f"rwkv6[{ci}].ht")
        p = f"c{ci}_"
        out.update({p + "r": r, p + "k": k, p + "v": v, p + "w": w, p + "u": u, p + "o": o, p + "ht": ht})
        if use_h0:
            out[p + "h0"] = h0
    out["n_cases"] = 3
    save("rwkv6_ops.npz", **out)


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    which = sys.argv[1:] or ["ops", "layer", "model", "codec", "rwkv6"]
    for w in which:
        print(f"[{w}]")
        {"ops": gen_ops, "layer": gen_layer, "model": gen_model, "codec": gen_codec, "rwkv6": gen_rwkv6}[w]()
