"""CPU: host-side logic that needs no kernel -- token tools, the state cache, module / state-dict layout."""
import pytest
import torch

from lina_speech_b200.fla_api import Cache
from lina_speech_b200.model import tools
import lina_speech_b200.model as m
from lina_speech_b200.codec import WavTokenizer


def test_delay_undelay_roundtrip():
    code = torch.randint(3, 50, (3, 11))
    d = tools.delay_rvq(code, head_token=1, tail_token=2)
    assert d.shape == (3, 11 + 4)
    assert (d[0, 0] == 1) and (d[2, :3] == 1).all() and (d[0, -3:] == 2).all()
    u = tools.undelay_rvq(d.unsqueeze(1))
    assert torch.equal(u[:, 0], code)


def test_topk_sampling_greedy_is_argmax():
    torch.manual_seed(0)
    x = torch.randn(5, 40)
    assert torch.equal(tools.topk_sampling(x.clone(), k=1).squeeze(-1), x.argmax(-1))
    s = tools.topk_sampling(x.clone(), k=5, temp=0.7).squeeze(-1)
    top5 = x.topk(5, dim=-1).indices
    assert all(s[i] in top5[i] for i in range(5))


def test_cache_update_is_in_place_and_skips_self_copies():
    c = Cache()
    a, b = torch.zeros(2, 3), torch.zeros(2, 4)
    c.update((a, b), 0, offset=0)
    assert c[0][0] is a and len(c) == 1
    c.update((torch.ones(2, 3), b), 0)            # b is the same storage: no copy, a receives the new values
    assert torch.equal(a, torch.ones(2, 3)) and c.get_seq_length() == 1
    try:
        c[3]
        assert False
    except KeyError:
        pass


def test_state_layout_matches_reference_init_state():
    rnn = m.AttentiveGLA(64, 2, 2, blind=True, use_short_conv=True, pos_type="convolutional")
    cache = rnn.init_state(batch_size=3)
    assert len(cache) == 5                                          # 2 enc + 2 dec + pos_net (model/gla.py:302-313)
    shapes = [tuple(t.shape) for t in cache[0]]
    assert shapes == [(3, 64, 4), (3, 64, 4), (3, 128, 4), (3, 2, 32, 64)]
    params = rnn.get_init_state_tuning_params(lora=1)
    assert len(params) == 4 and params[0][0].shape == (1, 1, 2, 32, 1) and params[0][1].shape == (1, 1, 2, 1, 64)
    st = rnn.get_state_from_params(params, 3, scale=0.02)
    assert st[0][-1].shape == (3, 2, 32, 64) and st[0][-1].requires_grad


def test_flagship_parameter_count_is_the_readme_169M():
    """AttentiveGLA(d1024, n_layer=6, blind) = 169.35 M parameters (README.md:36)."""
    with torch.device("meta"):
        rnn = m.AttentiveGLA(1024, 6, 4, blind=True, use_short_conv=True, pos_type="convolutional")
    n = sum(p.numel() for p in rnn.parameters())
    assert abs(n / 1e6 - 169.35) < 0.3, n


def test_codec_module_tree_has_reference_key_names():
    wt = WavTokenizer.from_hparams(vq_bins=64, dim=64, intermediate_dim=128, num_layers=2)
    keys = set(wt.state_dict().keys())
    for k in ("backbone.embed.weight", "backbone.pos_net.0.norm1.weight", "backbone.pos_net.2.q.weight",
              "backbone.pos_net.5.bias", "backbone.norm.scale.weight", "backbone.convnext.1.dwconv.weight",
              "backbone.convnext.0.gamma", "backbone.final_layer_norm.weight", "head.out.weight",
              "head.istft.window", "feature_extractor.encodec.quantizer.vq.layers.0._codebook.embed"):
        assert k in keys, k


@pytest.mark.parametrize("B,H,T,K,V,with_state", [(1, 2, 128, 16, 32, True), (2, 1, 150, 16, 600, False), (1, 1, 200, 8, 512, True)])
def test_backward_through_the_forward_kernel_host_logic(B, H, T, K, V, with_state):
    """fla_api.ops._bwd_tc_reference (the scheme of the product's _bwd_tc) regroups the chunked GLA backward into five runs of the pre-gated forward kernel (role swaps,
    time reversal, V split, row decay).  With the oracle's restatement of that kernel's contract plugged in as ``run`` the
    result must equal the explicit backward of the recurrence (FLA/fla/ops/common/fused_recurrent.py:172-257,335-342),
    incl. dh0, the dht terms and a ragged last chunk.  (CPU: host logic only; the kernel itself is checked on the GPU.)"""
    import torch.nn.functional as F
    from oracle import gla_oracle as GO
    from lina_speech_b200.fla_api import ops

    def run(qg, kg, v, decay, h0, o, ht, row):
        oo, S = GO.pregated_chunk_fwd(qg, kg, v, decay, h0, row, acc_dtype=torch.float64)
        o.copy_(oo)
        if ht is not None:
            ht.copy_(S)

    torch.manual_seed(T)
    f64 = torch.float64
    q, k = (torch.randn(B, H, T, K, dtype=f64) for _ in range(2))
    v, do = (torch.randn(B, H, T, V, dtype=f64) for _ in range(2))
    gk = F.logsigmoid(torch.randn(B, H, T, K, dtype=f64)) / 4
    h0 = torch.randn(B, H, K, V, dtype=f64) if with_state else None
    dht = torch.randn(B, H, K, V, dtype=f64) if with_state else None
    ref = GO.recurrent_gla_bwd(q, k, v, gk, h0, do, dht)
    got = ops._bwd_tc_reference(q, k, v, gk, h0, do, dht, K ** -0.5, True, run=run)
    for name, a, b in zip(("dq", "dk", "dv", "dgk", "dh0"), got, ref):
        err = (a.double() - b).abs().max().item() / b.abs().max().item()
        assert err < 5e-6, f"{name}: relative error {err:.2e}"       # fp32 intermediates inside _bwd_tc


def test_istft_fft640_device_phases_on_the_host(tmp_path):
    """csrc/fft640.cuh (the warp-per-frame ISTFT kernel's phases) compiled for the host and run lane by lane:
    polar -> 640-point complex FFT (10*4*4*4) -> windowed frame, against numpy's irfft (DEC/spectral_ops.py:57-58)."""
    import ctypes, os, shutil, subprocess
    import numpy as np
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    here = os.path.dirname(os.path.abspath(__file__))
    so = str(tmp_path / "libfft640.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(here, "host_fft640.cpp")], check=True)
    lib = ctypes.CDLL(so)
    rng = np.random.default_rng(1)
    F = 9
    h = np.concatenate([rng.standard_normal((F, 641)) * 1.5, rng.uniform(-6, 6, (F, 641))], 1).astype(np.float32)
    h[0, :641] = 8.0                                    # exercises the clip at 100
    win = torch.hann_window(1280).numpy().astype(np.float32)
    out = np.zeros((F, 1280), np.float32)
    P = ctypes.POINTER(ctypes.c_float)
    assert lib.fft640_frames(h.ctypes.data_as(P), win.ctypes.data_as(P), out.ctypes.data_as(P), F) == 0
    mag = np.minimum(np.exp(h[:, :641].astype(np.float64)), 100.0)
    X = mag * np.exp(1j * h[:, 641:].astype(np.float64))
    ref = np.fft.irfft(X, n=1280, axis=1) * win
    assert np.abs(out - ref).max() <= 2e-6 * np.abs(ref).max() + 2e-6


def test_reverse_and_rwkv6_host_logic_with_the_kernel_stubbed_by_the_oracle(monkeypatch):
    """fused_recurrent_gla(reverse=True / gk=None) and the RWKV6-through-GLA composition are pure host logic around the GLA
    operator: run the GPU tests' own bodies on the CPU with _GLAFunction replaced by the oracle recurrence (autograd through
    torch), so flips, the one-step query shift, the bonus term and every gradient path are checked without a device."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle import gla_oracle as GO
    import lina_speech_b200.fla_api.ops as O
    import test_gla_ops_gpu as T

    class OracleGLA:
        @staticmethod
        def apply(q, k, v, gk, scale, h0, want_ht, kind):
            o, ht = GO.recurrent_gla(q, k, v, gk, scale=scale, initial_state=h0)
            return o, (ht if want_ht else None)

    monkeypatch.setattr(O, "_GLAFunction", OracleGLA)
    monkeypatch.setattr(T, "DEV", "cpu")
    T.test_fused_recurrent_reverse_and_ungated_forms()
    for op in ("fused_recurrent_rwkv6", "chunk_rwkv6"):
        T.test_rwkv6_gradients(op)


def test_short_convolution_with_bias_host_composition(monkeypatch):
    """ShortConvolution(bias=True) (convolution.py:84-139; unused by Lina) = kernel without activation + bias + SiLU in torch:
    checked against F.conv1d with the kernel call replaced by the oracle's conv."""
    import torch.nn.functional as F
    from oracle import gla_oracle as GO
    import lina_speech_b200.fla_api.modules as M

    class OracleConv:
        @staticmethod
        def apply(x, weight, cache, silu):
            return GO.short_conv_prefill(x, weight.squeeze(1), cache, "silu" if silu else None)

    monkeypatch.setattr(M, "_ShortConvFn", OracleConv)
    torch.manual_seed(0)
    conv = M.ShortConvolution(24, 4, bias=True, activation="silu")
    x = torch.randn(2, 11, 24)
    cache = torch.zeros(2, 24, 4)
    y = conv(x, cache=cache)
    ref = F.silu(F.conv1d(F.pad(x.transpose(1, 2), (3, 0)), conv.weight, conv.bias, groups=24)).transpose(1, 2)
    assert torch.allclose(y, ref, atol=1e-6)
    assert torch.equal(cache, x.transpose(1, 2)[..., -4:])


def test_norm_gate_residual_and_prenorm_host_composition(monkeypatch):
    """FusedRMSNormSwishGate(x, o, residual, prenorm, residual_in_fp32) (fused_norm_gate.py:100-111,460-480; unused by Lina):
    fp32 sum, normalised un-rounded, residual_out in the residual's dtype -- with the kernel replaced by the oracle."""
    from oracle import gla_oracle as GO
    import lina_speech_b200.fla_api.modules as M

    class OracleNormGate:
        @staticmethod
        def apply(x, g, weight, eps):
            return GO.rmsnorm_swish_gate(x, g.to(x.dtype), weight, eps)

    monkeypatch.setattr(M, "_NormGateFn", OracleNormGate)
    torch.manual_seed(0)
    m = M.FusedRMSNormSwishGate(32)
    with torch.no_grad():
        m.weight.uniform_(0.5, 1.5)
    x, o, r = torch.randn(3, 5, 32).bfloat16(), torch.randn(3, 5, 32).bfloat16(), torch.randn(3, 5, 32)
    y, res = m(x, o, residual=r, prenorm=True)
    xs = x.float() + r
    ref = xs * torch.rsqrt(xs.pow(2).mean(-1, keepdim=True) + 1e-5) * m.weight * o.float() * torch.sigmoid(o.float())
    assert y.dtype == torch.bfloat16 and res.dtype == torch.float32 and torch.equal(res, xs)
    assert torch.allclose(y.float(), ref, atol=2e-2, rtol=2e-2)
    y2, res2 = m(x, o, prenorm=True, residual_in_fp32=True)
    assert res2.dtype == torch.float32 and torch.equal(res2, x.float()) and y2.dtype == torch.bfloat16
    y3, res3 = m(x, o, prenorm=True)
    assert res3 is x


def test_cache_reorder_selects_batch_rows_of_every_state_tensor():
    c = Cache()
    c.update((torch.arange(6.).view(3, 2), torch.arange(12.).view(3, 4)), 0)
    c.update((torch.arange(3.).view(3, 1),), 1)
    c.reorder_cache(torch.tensor([2, 0, 0]))
    assert torch.equal(c[0][0], torch.tensor([[4., 5.], [0., 1.], [0., 1.]]))
    assert torch.equal(c[0][1][:, 0], torch.tensor([8., 0., 0.])) and torch.equal(c[1][0].flatten(), torch.tensor([2., 0., 0.]))


def test_small_reference_helpers():
    """model/tools.py:8-15 pad_2d_sequence, initial_state.py:13-18 filter_unk (imported by the notebook), base_blocks.unpack_ignore."""
    from lina_speech_b200.initial_state import filter_unk, filter_except
    from lina_speech_b200.model.base_blocks import unpack_ignore
    p = tools.pad_2d_sequence([torch.ones(2, 3), torch.ones(4, 1)], padding_value=7)
    assert p.shape == (2, 4, 3) and p[0, 2:].eq(7).all() and p[1, :, 1:].eq(7).all() and p[1, :, 0].eq(1).all()

    class Tok:
        def encode(self, x):
            if "?" in x:
                raise KeyError(x)
            return [1]

    assert filter_unk("abc", Tok()) and not filter_unk("a?c", Tok()) and not filter_except("abc")
    assert unpack_ignore((1, 2)) == 1 and unpack_ignore(3) == 3


def test_wavtokenizer_from_pretrained0802_reads_yaml_and_checkpoint(tmp_path):
    """DEC/pretrained.py:81-115: the notebook's constructor.  A yaml with the shipped structure + a checkpoint holding a
    reference-keyed state dict (with encoder-side keys that must be ignored) round-trips into the decode-side module."""
    import yaml
    cfg = {"model": {"init_args": {
        "feature_extractor": {"class_path": "decoder.feature_extractors.EncodecFeatures",
                              "init_args": {"encodec_model": "encodec_24khz", "bandwidths": [6.6], "train_codebooks": True,
                                            "num_quantizers": 1, "dowmsamples": [8, 5, 4, 2], "vq_bins": 64, "vq_kmeans": 200}},
        "backbone": {"class_path": "decoder.models.VocosBackbone",
                     "init_args": {"input_channels": 32, "dim": 64, "intermediate_dim": 96, "num_layers": 2,
                                   "adanorm_num_embeddings": 4}},
        "head": {"class_path": "decoder.heads.ISTFTHead", "init_args": {"dim": 64, "n_fft": 64, "hop_length": 16, "padding": "same"}}}}}
    (tmp_path / "c.yaml").write_text(yaml.safe_dump(cfg))
    src = WavTokenizer.from_hparams0802(str(tmp_path / "c.yaml"))
    with torch.no_grad():
        for p in src.parameters():
            p.normal_()
        src.feature_extractor.encodec.quantizer.vq.layers[0]._codebook.embed.normal_()
    sd = dict(src.state_dict())
    sd["feature_extractor.encodec.encoder.model.0.conv.conv.weight"] = torch.zeros(3)      # encoder side: ignored
    sd["discriminator.x"] = torch.zeros(1)
    torch.save({"state_dict": sd}, tmp_path / "m.ckpt")
    wt = WavTokenizer.from_pretrained0802(str(tmp_path / "c.yaml"), str(tmp_path / "m.ckpt"))
    assert not wt.training
    for k, v in src.state_dict().items():
        assert torch.equal(wt.state_dict()[k], v), k


def test_train_lina_mirror_checkpoint_round_trip_and_optimizer(tmp_path):
    """train_lina.py:12-120 without Lightning: constructor, a Lightning-style checkpoint (hyper_parameters + state_dict with
    ``model.*`` keys) loads back bit-exactly, configure_optimizers gives AdamW + the cosine warm-up schedule."""
    from lina_speech_b200.train_lina import TrainLina

    def parts():
        return dict(attentive_rnn=m.AttentiveGLA(32, 1, 2, blind=True, use_short_conv=True, pos_type="convolutional"),
                    txt_encoder=m.TextEncoder(32, 2, n_layers=1, dropout=0.0, rotary=False))

    hp = dict(d_model=32, quant_layer=[0], n_codebook=64, n_special_token_in=3, n_special_token_out=3, n_txt_vocab=40,
              learning_rate=2e-4, betas=(0.9, 0.95), n_warmup_steps=5, n_training_steps=50)
    torch.manual_seed(0)
    a = TrainLina(**parts(), **hp)
    assert all(k.startswith("model.") for k in a.state_dict())
    torch.save({"state_dict": a.state_dict(), "hyper_parameters": dict(hp, **parts())}, tmp_path / "last.ckpt")
    torch.manual_seed(1)
    b = TrainLina.load_from_checkpoint(str(tmp_path / "last.ckpt"))
    for k, v in a.state_dict().items():
        assert torch.equal(b.state_dict()[k], v), k
    (opt,), (sch,) = b.configure_optimizers()
    assert isinstance(opt, torch.optim.AdamW) and opt.defaults["betas"] == (0.9, 0.95) and sch["interval"] == "step"
    assert opt.param_groups[0]["weight_decay"] == 0.1 and abs(opt.param_groups[0]["lr"]) < 1e-12      # warm-up starts at 0


def test_reference_import_names_resolve_after_install_aliases(tmp_path):
    """InferenceLina.ipynb cell 1 imports + un-pickling of reference-named classes, in a subprocess (sys.modules is global)."""
    import subprocess, sys, os, textwrap
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = textwrap.dedent(f"""
        import sys, pickle
        sys.path.insert(0, {root!r})
        from lina_speech_b200.compat import install_reference_aliases
        done = install_reference_aliases()
        from train_lina import TrainLina
        from decoder.pretrained import WavTokenizer
        from initial_state import train_initial_state, filter_unk
        from model.gla import AttentiveGLA
        from model.encoder import TextEncoder
        from model.multiembed import MultiEmbedding
        from model.attentive_rnn import AttentiveRNN
        from model.accuracy import MulticlassAccuracy
        import torch
        # a pickle that names the reference's module path un-pickles into the mirror class
        real = TextEncoder.__module__
        TextEncoder.__module__ = "model.encoder"
        blob = pickle.dumps(TextEncoder(32, 2, n_layers=1, dropout=0.0, rotary=False))
        TextEncoder.__module__ = real
        assert b"model.encoder" in blob and type(pickle.loads(blob)) is TextEncoder
        acc = MulticlassAccuracy(5, top_k=2, ignore_index=[0])
        p = torch.tensor([[0.1, 0.9, 0.5, 0., 0.], [0.9, 0.1, 0., 0., 0.], [0., 0., 0., 1., 0.5]])
        assert float(acc(p, torch.tensor([2, 0, 1]))) == 0.5
        assert issubclass(AttentiveGLA, AttentiveRNN) and hasattr(WavTokenizer, "from_pretrained0802") and len(done) >= 12
        print("ok")
    """)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]


def test_gate_envelope_routing_of_the_chunk_operators():
    """fla_api/ops.py:_route -- the tensor-core chunk kernels hold one pivot per 64-token chunk, so inputs whose summed
    log-gate inside a chunk passes -80 on any channel are served by the exact recurrence (the reference is exact for any
    gate, chunk_fuse.py:196-238).  Pure host logic: checked here on CPU tensors, both memory layouts, ragged T."""
    import torch.nn.functional as F
    import lina_speech_b200.fla_api.ops as O
    torch.manual_seed(0)
    B, H, T, K = 2, 3, 150, 16
    mild = (F.logsigmoid(torch.randn(B, H, T, K)) / 16).bfloat16()                  # model/gla.py:174-181
    fla = F.logsigmoid(torch.randn(B, H, T, K)).clamp_min(-3)                       # FLA/tests/ops/test_gla.py:27 (sum ~ -60)
    ref = torch.stack([c.float().sum(2) for c in mild.split(64, dim=2)], 2).amin()
    assert torch.allclose(O._min_chunk_gate_sum(mild), ref)
    bthd = mild.transpose(1, 2).contiguous().transpose(1, 2)                         # [B,H,T,K] view of [B,T,H,K] memory
    assert not bthd.is_contiguous() and torch.allclose(O._min_chunk_gate_sum(bthd), ref)
    assert O._gates_in_envelope(mild) and O._gates_in_envelope(fla)
    strong = mild.clone()
    strong[1, 2, 128:150, 5] = -4.0                                                  # 22 steps of -4 in the ragged tail chunk
    assert not O._gates_in_envelope(strong)
    nan = mild.clone(); nan[0, 0, 3, 0] = float("nan")
    assert not O._gates_in_envelope(nan)
    q, v = torch.empty(B, H, T, K, dtype=torch.bfloat16), torch.empty(B, H, T, 32, dtype=torch.bfloat16)
    yes, no = (lambda *a: 1), (lambda *a: 0)
    assert O._route("chunk", q, v, mild, yes) == ("chunk", True)
    assert O._route("fused_chunk", q, v, strong, yes) == ("recurrent", False)
    assert O._route("chunk", q, v, strong, no) == ("chunk", None)                    # CUDA-core chunk path: exact already
    assert O._route("recurrent", q, v, strong, yes) == ("recurrent", None)           # the backward looks at the gates


def test_cross_attention_memo_never_serves_another_utterance():
    """the text-side memo was keyed on ctx.data_ptr(); under inference_mode the allocator hands a freed
    text tensor's address to the next utterance and the old k / v came back (197 of 200 calls).  The memo now holds the
    tensor itself, so its storage cannot be recycled while the entry lives."""
    import gc
    from lina_speech_b200.model.crossatt import BlindCrossAttention
    torch.manual_seed(0)
    ca = BlindCrossAttention(32, 32, 32, 1, torch.nn.Identity(), pos_dim=32, pos_type="sinusoidal").eval()
    wrong = 0
    with torch.inference_mode():
        for i in range(200):
            ctx = torch.randn(2, 11, 32)
            k, v, pe = ca._text_side(ctx, None)
            k2, v2, _ = ca._text_side(ctx, None)                   # same tensor again: served from the memo
            assert k2 is k and v2 is v
            want = ca.ln_k(ca.k(ctx)).unsqueeze(1)
            wrong += int(not torch.allclose(k, want))
            del ctx, k, v, pe, k2, v2, want
            gc.collect()
    assert wrong == 0
    # parameters that feed k / v / pos_emb are part of the key: an in-place update invalidates the entry
    ctx = torch.randn(2, 11, 32)
    with torch.no_grad():
        k, _, _ = ca._text_side(ctx, None)
        ca.ln_k.bias.add_(1.0)
        k3, _, _ = ca._text_side(ctx, None)
    assert not torch.allclose(k, k3)
    ca.clear_memo()
    assert ca._memo is None


def test_modules_unpickled_without_init_have_their_lazy_caches():
    """TrainLina.load_from_checkpoint un-pickles the reference's module instances into the mirror
    classes: __dict__ is restored without running __init__, so attributes that only the mirrors create must have class-level
    defaults; fla's module paths inside the pickle must resolve as well."""
    import pickle
    import lina_speech_b200.model as M
    from lina_speech_b200 import compat
    from lina_speech_b200.model.base_blocks import SwiGLU
    rnn = M.AttentiveGLA(64, 1, 2, blind=True, use_short_conv=True, pos_type="convolutional")
    for mod in rnn.modules():                                       # what an instance pickled by the reference looks like
        for name in ("_wcat", "_wcat4", "_memo", "_padded"):
            mod.__dict__.pop(name, None)
    rnn2 = pickle.loads(pickle.dumps(rnn))
    t = rnn2.encoder[0].tmix
    assert t._wcat is None and t._wcat4 is None and rnn2.cross_att._memo is None
    assert [m for m in rnn2.modules() if isinstance(m, SwiGLU)][0]._padded is None
    assert t._cat_weight().shape[0] == 2 * t.key_dim + 2 * t.value_dim + 16
    done = compat.install_reference_aliases()
    try:
        import importlib
        assert importlib.import_module("fla.modules.convolution").ShortConvolution is M.gla.ShortConvolution
        assert importlib.import_module("fla.modules.fused_norm_gate").FusedRMSNormSwishGate is M.gla.FusedRMSNormSwishGate
        assert importlib.import_module("fla.models.utils").Cache is M.gla.Cache
        from fla.ops.gla import fused_chunk_gla  # noqa: F401
    finally:
        import sys
        for name in done:
            sys.modules.pop(name, None)


def test_gate_certificate_from_weights_alone():
    """GatedLinearAttention.gates_certified: ||(W2 W1)_c|| * ||LN output|| + |b_c| bounds the gate pre-activation; when the
    bound keeps 64 * logsigmoid(-bound) / normalizer above -80 the inference path needs no device flag and no host read."""
    import lina_speech_b200.model as M
    torch.manual_seed(0)
    blk = M.AttentiveGLA(256, 1, 4, blind=True, use_short_conv=True, pos_type="convolutional").encoder[0]
    bound = blk._ln_output_norm_bound()
    assert abs(bound - 16.0) < 1e-5                                   # default LayerNorm: gamma 1, beta 0 -> sqrt(256)
    assert blk.tmix.gates_certified(bound)                           # xavier(gain 2^-2.5) gate projections: tiny pre-activations
    assert 2 * blk.tmix.gate_preactivation_bound(bound) < 20         # ... with the 2x staleness margin
    assert not blk.tmix.gates_certified(None)
    with torch.no_grad():
        blk.tmix.gk_proj[1].bias[3] = -48.0                          # logsigmoid(-48) / 16 * 64 = -192 < -80
    assert not blk.tmix.gates_certified(bound)
    with torch.no_grad():
        blk.tmix.gk_proj[1].bias[3] = 0.0
        blk.norm1.weight.mul_(400.0)                                 # a huge LayerNorm gain voids the certificate too
    assert not blk.tmix.gates_certified(blk._ln_output_norm_bound())
