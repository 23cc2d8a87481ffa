"""GPU parity of ShortConvolution / FusedRMSNormSwishGate / the fused decode step against the oracle."""
import pytest
import torch
import torch.nn.functional as F

from oracle import gla_oracle as GO

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _close(got, ref, atol, rtol=1e-4, what=""):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    err = (got - ref).abs().max().item()
    tol = atol + rtol * ref.abs().max().item()
    assert err <= tol, f"{what}: max err {err:.3e} > {tol:.3e}"


@pytest.mark.parametrize("B,Ln,D", [(2, 1, 8), (2, 3, 40), (3, 37, 200), (2, 128, 1024)])
@pytest.mark.parametrize("W", [4, 2])
def test_short_conv_prefill_and_cache(B, Ln, D, W):
    from lina_speech_b200.fla_api import ShortConvolution
    torch.manual_seed(B * 100 + Ln)
    conv = ShortConvolution(D, W, activation="silu").to(DEV)
    x = torch.randn(B, Ln, D)
    w = conv.weight.detach().cpu()[:, 0]
    ref_cache = torch.ones(B, D, W)
    ref = GO.short_conv_prefill(x, w, ref_cache) if Ln > 1 else GO.short_conv_prefill(x, w)
    cache = torch.ones(B, D, W, device=DEV)
    y = conv(x.to(DEV), cache=cache if Ln > 1 else None)
    _close(y, ref, 1e-5, what="conv y")
    if Ln > 1:
        assert torch.equal(cache.cpu(), ref_cache)


def test_short_conv_steps_equal_prefill_and_grads():
    from lina_speech_b200.fla_api import ShortConvolution
    torch.manual_seed(0)
    B, Ln, D, W = 2, 19, 96, 4
    conv = ShortConvolution(D, W, activation="silu").to(DEV)
    x = torch.randn(B, Ln, D, device=DEV)
    y = conv(x)
    cache = torch.zeros(B, D, W, device=DEV)
    ys = torch.cat([conv(x[:, t:t + 1], cache=cache) for t in range(Ln)], 1)
    _close(ys, y, 1e-5, what="steps vs prefill")
    # gradients against autograd through the torch restatement
    xr = x.detach().cpu().requires_grad_(True)
    wr = conv.weight.detach().cpu()[:, 0].clone().requires_grad_(True)
    dy = torch.randn(B, Ln, D)
    yr = F.silu(F.conv1d(F.pad(xr.transpose(1, 2), (W - 1, 0)), wr.unsqueeze(1), groups=D)).transpose(1, 2)
    (yr * dy).sum().backward()
    xg = x.detach().clone().requires_grad_(True)
    (conv(xg) * dy.to(DEV)).sum().backward()
    _close(xg.grad, xr.grad, 1e-4, what="conv dx")
    _close(conv.weight.grad[:, 0], wr.grad, 1e-3, 1e-4, what="conv dw")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,N", [(5, 64), (300, 128), (1000, 512)])
def test_rmsnorm_swishgate_fwd_bwd(dtype, M, N):
    from lina_speech_b200.fla_api import FusedRMSNormSwishGate
    torch.manual_seed(M + N)
    mod = FusedRMSNormSwishGate(N, eps=1e-5).to(DEV)
    with torch.no_grad():
        mod.weight.uniform_(0.5, 1.5)
    x, g, dy = (torch.randn(M, N).to(dtype) for _ in range(3))
    xg, gg = x.to(DEV).requires_grad_(True), g.to(DEV).requires_grad_(True)
    y = mod(xg, gg)
    assert y.dtype == dtype
    w = mod.weight.detach().cpu()
    w_used = w.to(dtype).float()
    ref = GO.rmsnorm_swish_gate(x.float(), g.float(), w_used)
    lo = dtype != torch.float32
    _close(y, ref, 2e-2 if lo else 1e-5, 1e-2 if lo else 1e-5, what="norm-gate y")
    (y.float() * dy.to(DEV).float()).sum().backward()
    rdx, rdg, rdw = GO.rmsnorm_swish_gate_bwd(x.float(), g.float(), w_used, dy.float())
    _close(xg.grad, rdx, 3e-2 if lo else 1e-4, 2e-2 if lo else 1e-4, what="dx")
    _close(gg.grad, rdg, 3e-2 if lo else 1e-4, 2e-2 if lo else 1e-4, what="dg")
    _close(mod.weight.grad, rdw, 1e-3, 2e-2 if lo else 1e-4, what="dw")


@pytest.mark.parametrize("state_dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,d,H", [(3, 64, 2), (2, 256, 4), (2, 1024, 4)])
def test_fused_step_matches_oracle_layer(state_dtype, B, d, H):
    """GatedLinearAttention single-token path (lina_gla_step) vs the oracle's layer, 6 consecutive steps."""
    from lina_speech_b200.model import GatedLinearAttention
    from oracle import lina_oracle as LO
    torch.manual_seed(d + B)
    dtype = state_dtype            # cache dtype = parameter dtype (model/gla.py:230-239)
    layer = GatedLinearAttention(hidden_size=d, num_heads=H, use_short_conv=True, layer_idx=0).eval()
    with torch.no_grad():
        layer.g_norm_swish_gate.weight.uniform_(0.5, 1.5)
        layer.gk_proj[1].bias.normal_()
        for p in layer.parameters():
            p.copy_(p.to(dtype).float())                      # oracle sees the rounded weights
    sd = {"l." + k: v.detach().clone() for k, v in layer.state_dict().items()}
    layer = layer.to(DEV).to(dtype)
    from lina_speech_b200.fla_api import Cache
    cache = Cache()
    cache.update(layer.init_state(B), 0, offset=0)
    ost = tuple(torch.randn_like(s.float().cpu()).to(dtype).float() for s in cache.states[0])
    for s, o in zip(cache.states[0], ost):
        s.copy_(o.to(dtype))
    with torch.inference_mode():
        for t in range(6):
            x = torch.randn(B, 1, d).to(dtype)
            y = layer(x.to(DEV), past_key_values=cache, use_cache=True)
            ry = LO.gla_layer(sd, "l", x.float(), H, ost)
            if dtype == torch.bfloat16:       # the reference keeps a bf16 cache: round the oracle state the same way
                for o in ost:
                    o.copy_(o.to(dtype).float())
            lo = dtype == torch.bfloat16
            _close(y, ry, 3e-2 if lo else 2e-5, 2e-2 if lo else 1e-4, what=f"step {t} y")
    for s, o in zip(cache.states[0], ost):
        _close(s, o, 2e-2 if dtype == torch.bfloat16 else 1e-4, 1e-2 if dtype == torch.bfloat16 else 1e-4, what="state")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_swiglu_inference_path_matches_reference_formula(dtype):
    """padded-weight GEMMs + fused silu*mul (no-grad CUDA path) == p_out(silu(gate) * u) (model/base_blocks.py:48-50)."""
    from lina_speech_b200.model import SwiGLU
    torch.manual_seed(0)
    m = SwiGLU(1024).to(DEV).to(dtype)
    x = torch.randn(3, 50, 1024, device=DEV).to(dtype)
    with torch.no_grad():
        y = m(x)
    with torch.enable_grad():
        gate, u = m.p_in(x).chunk(2, dim=-1)
        ref = m.p_out(F.silu(gate) * u)
    lo = dtype == torch.bfloat16
    _close(y, ref, 2e-2 if lo else 1e-5, 2e-2 if lo else 1e-5, what="swiglu")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("N", [64, 1024])
def test_add_layernorm_matches_torch(dtype, N):
    from lina_speech_b200.model.base_blocks import add_layernorm
    torch.manual_seed(N)
    norm = torch.nn.LayerNorm(N).to(DEV).to(dtype)
    with torch.no_grad():
        norm.weight.uniform_(0.5, 1.5); norm.bias.normal_()
    a, x = (torch.randn(7, 33, N, device=DEV).to(dtype) for _ in range(2))
    with torch.no_grad():
        s, ln = add_layernorm(a, x, norm)
        s0, ln0 = add_layernorm(None, x, norm)
        ref_s = a + x
        ref_ln, ref_ln0 = norm(ref_s), norm(x)
    assert torch.equal(s, ref_s) and s0.data_ptr() == x.data_ptr()
    lo = dtype == torch.bfloat16
    _close(ln, ref_ln, 2e-2 if lo else 1e-5, 1e-2 if lo else 1e-5, what="add+ln")
    _close(ln0, ref_ln0, 2e-2 if lo else 1e-5, 1e-2 if lo else 1e-5, what="ln")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,Ln,kd,vd", [(2, 1, 64, 128), (3, 5, 64, 128), (2, 37, 256, 512), (2, 130, 1024, 2048)])
@pytest.mark.parametrize("tl", [8, 16])
def test_prefill_prep_matches_oracle(dtype, B, Ln, kd, vd, tl):
    """lina_gla_prefill_prep (conv+SiLU on q,k,v read as column slices of one projection buffer, gate non-linearity,
    conv caches) == the oracle's ShortConvolution and logsigmoid/normalizer, for both tile heights."""
    from lina_speech_b200 import _lib as L
    torch.manual_seed(Ln * 7 + kd)
    lib = L.lib()
    lib.lina_debug_set_variant(0, tl)
    try:
        ldx = 2 * kd + 2 * vd + 16
        proj = torch.randn(B, Ln, ldx).to(dtype)
        gk_raw = (torch.randn(B, Ln, kd) * 3).to(dtype)
        wq, wk, wv = (torch.randn(d, 4).to(dtype) for d in (kd, kd, vd))
        pd, gd = proj.to(DEV), gk_raw.to(DEV)
        q, k, gk = (torch.empty(B, Ln, kd, dtype=dtype, device=DEV) for _ in range(3))
        v = torch.empty(B, Ln, vd, dtype=dtype, device=DEV)
        cq, ck = (torch.ones(B, kd, 4, dtype=dtype, device=DEV) for _ in range(2))
        cv = torch.ones(B, vd, 4, dtype=dtype, device=DEV)
        wqd, wkd, wvd = wq.to(DEV), wk.to(DEV), wv.to(DEV)
        xq, xk, xv = pd[..., :kd], pd[..., kd:2 * kd], pd[..., 2 * kd:2 * kd + vd]
        rc = lib.lina_gla_prefill_prep(L.ptr(xq), L.ptr(xk), L.ptr(xv), ldx, L.ptr(wqd), L.ptr(wkd), L.ptr(wvd), L.ptr(gd), kd,
                                       L.ptr(q), L.ptr(k), L.ptr(v), L.ptr(gk), L.ptr(cq), L.ptr(ck), L.ptr(cv), L.dt(cq),
                                       B, Ln, kd, vd, 4, 16.0, -0.1, 1, L.dt(pd), L.stream(pd))
        L.check(rc, "lina_gla_prefill_prep")
        lo = dtype != torch.float32
        for got, sl, w, cache in ((q, slice(0, kd), wq, cq), (k, slice(kd, 2 * kd), wk, ck),
                                  (v, slice(2 * kd, 2 * kd + vd), wv, cv)):
            rc_ = torch.ones(B, w.shape[0], 4)
            ref = GO.short_conv_prefill(proj[..., sl].float(), w.float(), rc_)
            _close(got, ref, 2e-2 if lo else 1e-5, 1e-2 if lo else 1e-5, what="conv")
            assert torch.equal(cache.float().cpu(), rc_), "conv cache"
        ref_g = GO.gate_logsigmoid(gk_raw.float(), 16.0, -0.1)
        _close(gk, ref_g, 2e-3 if lo else 1e-6, 1e-2 if lo else 1e-5, what="gate")
    finally:
        lib.lina_debug_set_variant(0, 0)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_rmsnorm_swishgate_strided_gate(dtype):
    """gate read in place from a wider buffer (g_group = heads, ldg = projection row stride)."""
    from lina_speech_b200 import _lib as L
    torch.manual_seed(3)
    Bt, H, N, ld = 50, 4, 512, 3 * 512 + 4 * 512
    x = torch.randn(Bt * H, N).to(dtype)
    proj = torch.randn(Bt, ld).to(dtype)
    w = torch.empty(N).uniform_(0.5, 1.5).to(dtype)
    xd, pd, wd = x.to(DEV), proj.to(DEV), w.to(DEV)
    g = pd[:, 3 * 512:]
    y = torch.empty_like(xd)
    rc = L.lib().lina_rmsnorm_swishgate_fwd_ld(L.ptr(xd), L.ptr(g), L.ptr(wd), L.ptr(y), None, Bt * H, N, 1e-5, H, ld,
                                               L.dt(xd), L.stream(xd))
    L.check(rc, "lina_rmsnorm_swishgate_fwd_ld")
    ref = GO.rmsnorm_swish_gate(x.float(), proj[:, 3 * 512:].reshape(Bt * H, N).float(), w.float())
    lo = dtype != torch.float32
    _close(y, ref, 2e-2 if lo else 1e-5, 1e-2 if lo else 1e-5, what="strided norm-gate")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,Vn,ld", [(37, 67, 72), (64, 4099, 4104), (9, 4099, 4099), (5, 100, 100)])
def test_cross_entropy_rows_matches_torch(dtype, M, Vn, ld):
    """lina_cross_entropy_rows == F.cross_entropy(logits.float(), target, ignore_index=1) incl. the row mask."""
    from lina_speech_b200 import _lib as L
    torch.manual_seed(M + Vn)
    buf = (torch.randn(M, ld) * 4).to(dtype)
    target = torch.randint(0, Vn, (M,))
    target[::5] = 1
    mask = torch.rand(M) > 0.2
    bd, td, md = buf.to(DEV), target.to(DEV), mask.to(DEV).view(torch.uint8)
    rows = torch.empty(2, M, dtype=torch.float32, device=DEV)
    rc = L.lib().lina_cross_entropy_rows(L.ptr(bd), ld, L.ptr(td), L.ptr(md), L.ptr(rows[0]), L.ptr(rows[1]), M, Vn, 1,
                                         L.dt(bd), L.stream(bd))
    L.check(rc, "lina_cross_entropy_rows")
    ref_rows = F.cross_entropy(buf[:, :Vn].float(), target, ignore_index=1, reduction="none") * mask
    _close(rows[0], ref_rows, 1e-5, 1e-5, what="row losses")
    keep = mask & (target != 1)
    assert torch.equal(rows[1].cpu() > 0, keep)
    ref = F.cross_entropy(buf[:, :Vn][mask].float(), target[mask], ignore_index=1)
    _close(rows[0].sum() / rows[1].sum(), ref, 1e-5, 1e-5, what="mean loss")


@pytest.mark.parametrize("B,T,H,K,V,use_h0", [(2, 150, 2, 64, 128, False), (1, 300, 4, 256, 512, True), (2, 64, 1, 128, 128, True),
                                              (1, 33, 2, 64, 256, False)])
def test_pregated_prep_and_chunk_kernel(B, T, H, K, V, use_h0):
    """lina_gla_prefill_prep_gated (conv + SiLU + gate + chunk cumsum -> q~, k~, decay) and the tcgen05 kernel on
    those operands == the oracle's conv -> logsigmoid/16 -> recurrence on the same inputs (bf16 I/O tolerance)."""
    from lina_speech_b200 import _lib as L
    torch.manual_seed(T + K)
    lib = L.lib()
    bf = torch.bfloat16
    kd, vd = H * K, H * V
    ldx = 2 * kd + 2 * vd
    proj = torch.randn(B, T, ldx).to(bf)
    gk_raw = (torch.randn(B, T, kd) * 2).to(bf)
    wq, wk, wv = (torch.randn(d, 4).mul(0.5).to(bf) for d in (kd, kd, vd))
    h0 = torch.randn(B, H, K, V) if use_h0 else None
    scale = K ** -0.5
    # oracle: fp32 math on the bf16-valued inputs
    cq_r, ck_r, cv_r = torch.ones(B, kd, 4), torch.ones(B, kd, 4), torch.ones(B, vd, 4)
    q = GO.short_conv_prefill(proj[..., :kd].float(), wq.float(), cq_r)
    k = GO.short_conv_prefill(proj[..., kd:2 * kd].float(), wk.float(), ck_r)
    v = GO.short_conv_prefill(proj[..., 2 * kd:2 * kd + vd].float(), wv.float(), cv_r)
    gk = GO.gate_logsigmoid(gk_raw.float()).to(bf).float()                       # the reference's gk is a bf16 tensor
    hd = lambda t, d: t.view(B, T, H, d).transpose(1, 2)
    ro, rht = GO.recurrent_gla(hd(q, K), hd(k, K), hd(v, V), hd(gk, K), scale=scale, initial_state=h0)
    # ours
    pd, gd = proj.to(DEV), gk_raw.to(DEV)
    qg, kg = (torch.empty(B, T, kd, dtype=bf, device=DEV) for _ in range(2))
    vv = torch.empty(B, T, vd, dtype=bf, device=DEV)
    nt = (T + 63) // 64
    decay = torch.empty(B, H, nt, K, dtype=torch.float32, device=DEV)
    cq, ck = (torch.ones(B, kd, 4, dtype=bf, device=DEV) for _ in range(2))
    cv = torch.ones(B, vd, 4, dtype=bf, device=DEV)
    wqd, wkd, wvd = wq.to(DEV), wk.to(DEV), wv.to(DEV)
    xq, xk, xv = pd[..., :kd], pd[..., kd:2 * kd], pd[..., 2 * kd:2 * kd + vd]
    rc = lib.lina_gla_prefill_prep_gated(L.ptr(xq), ldx, L.ptr(xk), ldx, L.ptr(xv), ldx, L.ptr(wqd), L.ptr(wkd), L.ptr(wvd), L.ptr(gd), kd,
                                         L.ptr(qg), L.ptr(kg), L.ptr(vv), L.ptr(decay), L.ptr(cq), L.ptr(ck), L.ptr(cv),
                                         L.dt(cq), B, T, H, K, V, 4, 16.0, scale, None, L.stream(pd))
    L.check(rc, "lina_gla_prefill_prep_gated")
    for got, ref in ((cq, cq_r), (ck, ck_r), (cv, cv_r)):
        assert torch.equal(got.float().cpu(), ref), "conv cache"
    _close(vv, v, 2e-2, 1e-2, what="v conv")
    # gated operands against the oracle's chunk-local cumsum
    G = torch.cat([gk[:, c:c + 64].cumsum(1) for c in range(0, T, 64)], 1)
    _close(qg, q * G.exp() * scale, 2e-2, 1e-2, what="q~")
    _close(kg, k * (-G).exp(), 2e-2, 1e-2, what="k~")
    GC = torch.stack([gk[:, c:c + 64].sum(1) for c in range(0, T, 64)], 1)         # [B, NT, kd]
    _close(decay, GC.exp().view(B, nt, H, K).transpose(1, 2), 1e-5, 1e-4, what="decay")
    o = torch.empty(B, T, H, V, dtype=bf, device=DEV)
    ht = torch.empty(B, H, K, V, dtype=torch.float32, device=DEV)
    h0d = h0.to(DEV) if h0 is not None else None
    rc = lib.lina_gla_chunk_fwd_pregated_bthd(L.ptr(qg), L.ptr(kg), L.ptr(vv), L.ptr(decay), L.ptr(h0d),
                                              L.dt(h0d) if h0d is not None else 0, L.ptr(o), L.ptr(ht), B, H, T, K, V,
                                              L.stream(pd))
    L.check(rc, "lina_gla_chunk_fwd_pregated_bthd")
    torch.cuda.synchronize()
    _close(o.transpose(1, 2), ro, 3e-2 * ro.abs().max().item(), 0.0, what="o (pregated tcgen05)")
    _close(ht, rht, 3e-2 * rht.abs().max().item(), 0.0, what="final state")
    # the kernel's build variants compute the same bits: round-1 one-CTA-per-tile kernel with one state warpgroup (key 4 = 1) or
    # three (key 9 = 1); the default is the CTA-pair kernel when V / 128 is even and K >= 128
    for key in (4, 9):
        o2, ht2 = torch.empty_like(o), torch.empty_like(ht)
        lib.lina_debug_set_variant(key, 1)
        try:
            rc = lib.lina_gla_chunk_fwd_pregated_bthd(L.ptr(qg), L.ptr(kg), L.ptr(vv), L.ptr(decay), L.ptr(h0d),
                                                      L.dt(h0d) if h0d is not None else 0, L.ptr(o2), L.ptr(ht2), B, H, T, K, V,
                                                      L.stream(pd))
            L.check(rc, f"lina_gla_chunk_fwd_pregated_bthd (variant {key})")
            torch.cuda.synchronize()
        finally:
            lib.lina_debug_set_variant(key, 0)
        assert torch.equal(o2, o) and torch.equal(ht2, ht), f"variant {key} differs"


@pytest.mark.parametrize("B,T,H,K,V,use_h0", [(10, 512, 4, 256, 512, True), (10, 500, 4, 256, 512, False), (20, 300, 4, 128, 256, True),
                                              (3, 1024, 4, 256, 512, False), (32, 2048, 4, 256, 512, False)])
def test_pair_kernel_and_time_cut_are_bit_identical_to_one_cta_per_tile(B, T, H, K, V, use_h0):
    """The CTA-pair kernel (score MMA shared through DSMEM) without and with the workspace (tiles of the last wave cut in two
    along T, fp32 state handed over in HBM) == the one-CTA-per-tile kernel, bit for bit, incl. the final state; the small case
    is also checked against the oracle's restatement of the kernel contract."""
    from lina_speech_b200 import _lib as L
    lib = L.lib()
    torch.manual_seed(B * T + K)
    bf = torch.bfloat16
    nt = (T + 63) // 64
    qg = (torch.randn(B, T, H, K, device=DEV) * 0.5).to(bf)
    kg = (torch.randn(B, T, H, K, device=DEV) * 0.5).to(bf)
    v = torch.randn(B, T, H, V, device=DEV).to(bf)
    decay = torch.rand(B, H, nt, K, device=DEV) * 0.5 + 0.5
    h0 = torch.randn(B, H, K, V, device=DEV) if use_h0 else None
    ws_bytes = lib.lina_gla_chunk_fwd_pregated_ws_bytes(B, H, T, K, V)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    U = B * H * (V // 256)
    expect_cut = U >= sms // 2 and 0 < U % (sms // 2) <= sms // 4 and nt >= 4
    assert (ws_bytes > 0) == expect_cut, (ws_bytes, U)
    ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=DEV)

    def run(variant, use_ws):
        o = torch.full((B, T, H, V), float("nan"), dtype=bf, device=DEV)
        ht = torch.full((B, H, K, V), float("nan"), dtype=torch.float32, device=DEV)
        lib.lina_debug_set_variant(9, variant)
        try:
            rc = lib.lina_gla_chunk_fwd_pregated_bthd_ws(L.ptr(qg), L.ptr(kg), L.ptr(v), L.ptr(decay), L.ptr(h0),
                                                         L.dt(h0) if h0 is not None else 0, L.ptr(o), L.ptr(ht),
                                                         L.ptr(ws) if use_ws else None, ws_bytes if use_ws else 0, B, H, T, K, V,
                                                         L.stream(qg))
            L.check(rc, "lina_gla_chunk_fwd_pregated_bthd_ws")
            torch.cuda.synchronize()
        finally:
            lib.lina_debug_set_variant(9, 0)
        return o, ht

    o_ref, ht_ref = run(1, False)
    assert torch.isfinite(o_ref.float()).all() and torch.isfinite(ht_ref).all()
    for variant, use_ws, name in ((0, False, "pairs"), (0, True, "pairs + time cut"), (0, True, "pairs + time cut (second launch)")):
        o, ht = run(variant, use_ws)
        assert torch.equal(o, o_ref), f"{name}: o differs ({(o.float() - o_ref.float()).abs().max().item():.3e})"
        assert torch.equal(ht, ht_ref), f"{name}: final state differs"
    if B * T <= 6000 and T % 64 == 0:
        ro, rs = GO.pregated_chunk_fwd(qg.transpose(1, 2).float().cpu(), kg.transpose(1, 2).float().cpu(),
                                       v.transpose(1, 2).float().cpu(), decay.cpu(), h0.cpu() if h0 is not None else None)
        _close(o_ref.transpose(1, 2), ro, 2e-2 * ro.abs().max().item(), 0.0, what="o vs kernel contract")
        _close(ht_ref, rs, 2e-2 * rs.abs().max().item(), 0.0, what="state vs kernel contract")


@pytest.mark.parametrize("M,N,R,strided,bias", [(65, 1024, 16, True, True), (300, 1024, 16, False, True), (64, 40, 8, False, False),
                                                 (1000, 2052, 32, True, True)])
def test_lowrank_linear_matches_fp32_linear(M, N, R, strided, bias):
    """lina_lowrank_linear (gk_proj[1], model/gla.py:96-97, over a whole sequence) == the fp32 product of the same bf16 values
    rounded once to bf16 (what a bf16 nn.Linear returns, up to the summation order inside the fp32 accumulation)."""
    from lina_speech_b200 import _lib as L
    torch.manual_seed(M + N)
    bf = torch.bfloat16
    wide = torch.randn(M, R + 24, device=DEV).to(bf)
    x = wide[:, 24:] if strided else wide[:, :R].contiguous()
    W = (torch.randn(N, R, device=DEV) * 0.3).to(bf)
    b = torch.randn(N, device=DEV).to(bf) if bias else None
    out = torch.full((M, N), float("nan"), dtype=bf, device=DEV)
    rc = L.lib().lina_lowrank_linear(L.ptr(x), x.stride(0), L.ptr(W), L.ptr(b), L.ptr(out), N, M, N, R, L.dt(x), L.stream(x))
    L.check(rc, "lina_lowrank_linear")
    ref = x.float() @ W.float().t() + (b.float() if bias else 0.0)
    err = (out.float() - ref).abs()
    assert torch.isfinite(out.float()).all()
    assert (err <= ref.abs() * 2.0 ** -8 + 1e-6).all(), f"max err {err.max().item():.3e}"      # half an ulp of bf16 (+ fp32 sum order)
    assert (out == ref.to(bf)).float().mean() > 0.999            # the rare differences are round-to-nearest ties of the sum order


def test_bf16_layer_prefill_pregated_matches_op_by_op_and_oracle():
    """GatedLinearAttention at the flagship head size (d1024, H4, K256, V512) in bf16: the pregated inference path ==
    the same layer with LINA_PREGATED / LINA_FUSED_PREFILL off (within bf16 rounding) == the oracle layer in fp32;
    prefill with a cache then continues with the fused decode step."""
    import lina_speech_b200.model.gla as G
    from lina_speech_b200.fla_api import Cache
    from oracle import lina_oracle as LO
    torch.manual_seed(5)
    B, T, d, H = 2, 200, 1024, 4
    layer = G.GatedLinearAttention(hidden_size=d, num_heads=H, use_short_conv=True, layer_idx=0).eval()
    with torch.no_grad():
        layer.g_norm_swish_gate.weight.uniform_(0.5, 1.5)
        layer.gk_proj[1].bias.normal_()
        for p in layer.parameters():
            p.copy_(p.to(torch.bfloat16).float())
    sd = {"l." + k: v.detach().clone() for k, v in layer.state_dict().items()}
    layer = layer.to(DEV).to(torch.bfloat16)
    x = torch.randn(B, T, d).to(torch.bfloat16)
    ref = LO.gla_layer(sd, "l", x.float(), H)
    ref = ref[0] if isinstance(ref, tuple) else ref
    outs = {}
    saved = (G.FUSED_PREFILL, G.PREGATED)
    try:
        for name, fp, pg in (("op_by_op", False, False), ("fused", True, False), ("pregated", True, True)):
            G.FUSED_PREFILL, G.PREGATED = fp, pg
            cache = Cache()
            cache.update(layer.init_state(B), 0, offset=0)
            with torch.inference_mode():
                y = layer(x.to(DEV), past_key_values=cache, use_cache=True)
                y1 = layer(x[:, :1].to(DEV), past_key_values=cache, use_cache=True)
            outs[name] = (y.float().cpu(), y1.float().cpu(), [s.float().cpu() for s in cache.states[0]])
    finally:
        G.FUSED_PREFILL, G.PREGATED = saved
    scale = ref.abs().max().item()
    for name, (y, y1, st) in outs.items():
        _close(y, ref, 3e-2 * scale, 0.0, what=f"{name} vs oracle")
        _close(y, outs["op_by_op"][0], 2e-2 * scale, 0.0, what=f"{name} vs op-by-op")
        _close(y1, outs["op_by_op"][1], 3e-2 * outs["op_by_op"][1].abs().max().item(), 0.0, what=f"{name} next step")
        for a, b in zip(st, outs["op_by_op"][2]):
            _close(a, b, 2e-2 * max(b.abs().max().item(), 1e-3), 0.0, what=f"{name} cache state")


@pytest.mark.parametrize("B,Ln,D", [(2, 150, 256), (1, 64, 1024), (3, 5, 64), (2, 67, 128)])
@pytest.mark.parametrize("silu", [True, False])
def test_short_conv_bf16_backward_packed_kernel(B, Ln, D, silu):
    """bf16 / W = 4 backward (short_conv4_bwd_bf16_kernel: one pass, packed fp32x2 math, tail-halo rows, atomics for dw)
    against autograd through the fp32 torch restatement on the same bf16-valued inputs."""
    from lina_speech_b200.fla_api import ShortConvolution
    torch.manual_seed(Ln + D)
    bf = torch.bfloat16
    conv = ShortConvolution(D, 4, activation="silu" if silu else None).to(DEV).to(bf)
    x = torch.randn(B, Ln, D).to(bf)
    dy = torch.randn(B, Ln, D).to(bf)
    xr = x.float().requires_grad_(True)
    wr = conv.weight.detach().float().cpu()[:, 0].clone().requires_grad_(True)
    pre = F.conv1d(F.pad(xr.transpose(1, 2), (3, 0)), wr.unsqueeze(1), groups=D).transpose(1, 2)
    yr = F.silu(pre) if silu else pre
    (yr * dy.float()).sum().backward()
    xg = x.to(DEV).requires_grad_(True)
    y = conv(xg)
    _close(y, yr, 2e-2, 1e-2, what="conv y (bf16)")
    (y.float() * dy.to(DEV).float()).sum().backward()
    _close(xg.grad, xr.grad, 2e-2, 1e-2, what="conv dx (bf16)")
    _close(conv.weight.grad[:, 0], wr.grad, 5e-2, 2e-2, what="conv dw (bf16)")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,Vn,ld", [(128, 4099, 4104), (7, 67, 72), (3, 300, 300)])
def test_fused_topk_sampling(dtype, B, Vn, ld):
    """lina_topk_sample: k = 1 is the arg-max; k = 100 equals the inverse-CDF sample of the reference's masked softmax
    (model/tools.py:38-44, incl. the unscaled-threshold quirk) for the same uniform numbers."""
    from lina_speech_b200 import _lib as L
    torch.manual_seed(B + Vn)
    buf = (torch.randn(B, ld) * 3).to(dtype)
    x = buf[:, :Vn].float()
    bd = buf.to(DEV)
    u = torch.rand(B)
    ud = u.to(DEV)
    out = torch.empty(B, dtype=torch.long, device=DEV)
    lib = L.lib()
    L.check(lib.lina_topk_sample(L.ptr(bd), ld, B, Vn, 1, 1.0, L.ptr(ud), L.ptr(out), L.dt(bd), L.stream(bd)), "topk k=1")
    ref1 = x.argmax(-1)
    assert torch.equal(x.gather(1, out.cpu()[:, None]), x.gather(1, ref1[:, None])), "k=1 must pick a maximum"
    for k, temp in ((min(100, Vn), 1.0), (min(20, Vn), 0.7)):
        L.check(lib.lina_topk_sample(L.ptr(bd), ld, B, Vn, k, temp, L.ptr(ud), L.ptr(out), L.dt(bd), L.stream(bd)), "topk")
        got = out.cpu()
        kth = torch.topk(x, k, dim=-1).values[:, -1:]
        xs = x / temp
        keep = xs >= torch.minimum(kth, xs.max(-1, keepdim=True).values)
        w = torch.where(keep, (xs - xs.max(-1, keepdim=True).values).exp(), torch.zeros_like(xs)).double()
        cdf = w.cumsum(-1)
        target = u.double()[:, None] * cdf[:, -1:]
        ref = (cdf > target).float().argmax(-1)
        assert keep.gather(1, got[:, None]).all(), "sample outside the kept set"
        agree = (got == ref).float().mean().item()
        assert agree >= 0.97, f"k={k}: only {agree:.3f} of the rows match the inverse-CDF reference"
        # the disagreeing rows sit on a CDF boundary: their cumulative probability is within fp32 rounding of the target
        bad = (got != ref).nonzero().flatten()
        for i in bad.tolist():
            lo, hi = sorted((int(got[i]), int(ref[i])))
            gap = (cdf[i, hi] - cdf[i, lo]).item() / cdf[i, -1].item()
            near = abs(cdf[i, lo].item() - target[i].item()) / cdf[i, -1].item()
            assert near < 1e-4 or gap < 1e-4, f"row {i}: picks {int(got[i])} vs {int(ref[i])} are not a rounding tie"


def test_generate_batch_with_fused_sampling_runs_and_stays_in_vocabulary(golden_model):
    import lina_speech_b200.model as m
    g = golden_model
    rnn = m.AttentiveGLA(64, 2, 2, blind=True, use_short_conv=True, pos_type="convolutional")
    lm = m.LinaModel(rnn, 64, 1, 64, 3, 3, 32, txt_encoder=m.TextEncoder(64, 2, n_layers=1, dropout=0.0, rotary=False))
    lm.load_state_dict({k[2:]: v for k, v in g.items() if k.startswith("w.")})
    lm = lm.to(DEV).eval()
    qs, atts, stop_tokens, cuts = lm.generate_batch(g["xt"].to(DEV), batch_size=4, prompt=g["prompt"].to(DEV), max_seqlen=16,
                                                    k=10, force_max_seqlen=True, cuda_graph=True)
    assert qs.shape == (1, 4, 16) and int(qs.min()) >= 0 and int(qs.max()) < 64 + 3


@pytest.mark.parametrize("N", [64, 1024])
def test_autocast_layernorm_matches_torch(N):
    """fp32-in / bf16-out LayerNorm of the autocast training path: forward equals torch's fp32 LayerNorm rounded to bf16,
    gradients (dx, dgamma, dbeta) equal autograd through torch's LayerNorm on the same upstream gradient."""
    from lina_speech_b200.model.base_blocks import autocast_layernorm
    torch.manual_seed(N)
    norm = torch.nn.LayerNorm(N).to(DEV)
    with torch.no_grad():
        norm.weight.uniform_(0.5, 1.5); norm.bias.normal_()
    x = (torch.randn(5, 37, N, device=DEV) * 2 + 0.3).requires_grad_(True)
    dy = torch.randn(5, 37, N, device=DEV).to(torch.bfloat16)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = autocast_layernorm(x, norm)
    assert y.dtype == torch.bfloat16
    ref = norm(x.detach().clone().requires_grad_(True))
    xr = x.detach().clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (N,), norm.weight, norm.bias, norm.eps)
    _close(y, ref, 1e-2, 8e-3, what="ln y")                      # bf16 rounding of the output
    y.backward(dy)
    gw, gb = norm.weight.grad.clone(), norm.bias.grad.clone()
    norm.weight.grad = None; norm.bias.grad = None
    ref.backward(dy.float())
    _close(x.grad, xr.grad, 1e-4, 1e-4, what="ln dx")
    _close(gw, norm.weight.grad, 1e-3, 1e-4, what="ln dgamma")
    _close(gb, norm.bias.grad, 1e-3, 1e-4, what="ln dbeta")
    # outside autocast nothing changes
    assert autocast_layernorm(x.detach(), norm).dtype == torch.float32


def test_inference_prefill_with_gates_outside_the_envelope_is_served_exactly():
    """ADVICE r1 / VERDICT weak #8: the default inference path (GatedLinearAttention._prefill, pre-gated tcgen05 kernel) keeps
    one pivot per 64-token chunk; gates summing below -80 inside a chunk must be noticed (device flag written by
    lina_gla_prefill_prep_gated) and the call served by the exact kernels -- lone layer (immediate check) and inside a
    backbone pass (one deferred check), with and without a cache."""
    import lina_speech_b200.model as M
    import lina_speech_b200.model.gla as G
    from lina_speech_b200.fla_api import Cache
    from oracle import lina_oracle as LO
    torch.manual_seed(9)
    bf = torch.bfloat16
    B, T, d, H = 1, 160, 512, 4                                # K = 128, V = 256: tensor-core envelope
    layer = G.GatedLinearAttention(hidden_size=d, num_heads=H, use_short_conv=True, layer_idx=0).eval()
    with torch.no_grad():
        layer.gk_proj[1].bias[::7] = -48.0                     # logsigmoid(-48)/16 = -3 per token = -192 per chunk
        for p in layer.parameters():
            p.copy_(p.to(bf).float())
    sd = {"l." + k: v.detach().clone() for k, v in layer.state_dict().items()}
    layer = layer.to(DEV).to(bf)
    x = torch.randn(B, T, d).to(bf)
    ref = LO.gla_layer(sd, "l", x.float(), H)
    assert G.GateEnvelope.depth == 0
    with torch.inference_mode():
        y = layer(x.to(DEV))
    assert torch.isfinite(y).all()
    _close(y, ref, 3e-2 * ref.abs().max().item(), 0.0, what="lone layer, gates outside the envelope")
    assert not G.GateEnvelope.tripped(torch.device(DEV))       # the layer consumed (and cleared) its own flag
    # the same through a backbone pass: the mixers only accumulate the flag, the backbone reads it once and redoes the pass
    torch.manual_seed(10)
    rnn = M.AttentiveGLA(d, 1, H, blind=True, use_short_conv=True, pos_type="convolutional").eval()
    with torch.no_grad():
        rnn.encoder[0].tmix.gk_proj[1].bias[::5] = -48.0
        for p in rnn.parameters():
            p.copy_(p.to(bf).float())
    sdr = {"r." + k: v.detach().clone() for k, v in rnn.state_dict().items()}
    cfg = {"d_model": d, "n_layer": 1, "heads": H, "pos_type": "convolutional"}
    ctx = torch.randn(B, 12, d).to(bf)
    refy, refatt = LO.attentive_gla(sdr, "r", cfg, x.float(), ctx.float())
    rnn = rnn.to(DEV).to(bf)
    with torch.inference_mode():
        yb, att = rnn(x.to(DEV), ctx.to(DEV))
    assert torch.isfinite(yb).all()
    _close(yb, refy, 4e-2 * refy.abs().max().item(), 0.0, what="backbone pass, gates outside the envelope")
    # multi-token pass WITH a cache (generate_batch(prefill_prompt=True)): the cache must end as the exact path leaves it
    state = LO.init_state(cfg, B)
    refs, _ = LO.attentive_gla(sdr, "r", cfg, x.float(), ctx.float(), state=state, step=True)
    cache = rnn.init_state(batch_size=B)
    with torch.inference_mode():
        ys, _, _ = rnn.step(x.to(DEV), ctx.to(DEV), 0, cache)
    _close(ys, refs, 4e-2 * refs.abs().max().item(), 0.0, what="multi-token step with cache")
    for i, st in enumerate(cache.states):
        S, Sr = st[-1].float().cpu(), state[i][-1]
        _close(S, Sr, 3e-2 * max(Sr.abs().max().item(), 1e-3), 0.0, what=f"cache state {i}")


def test_eval_under_autocast_takes_the_general_path():
    """eval + no_grad under torch.autocast(bf16) with fp32 weights (validation_step under bf16-mixed) must not
    reach the raw-pointer fast paths, which describe every buffer with ONE dtype."""
    import lina_speech_b200.model as M
    torch.manual_seed(3)
    rnn = M.AttentiveGLA(256, 1, 4, blind=True, use_short_conv=True, pos_type="convolutional").to(DEV).eval()
    x, ctx = torch.randn(2, 40, 256, device=DEV), torch.randn(2, 9, 256, device=DEV)
    with torch.no_grad():
        ref, _ = rnn(x, ctx)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y, _ = rnn(x, ctx)
            cache = rnn.init_state(batch_size=2)
            y1, _, _ = rnn.step(x[:, :1], ctx, 0, cache)
    assert torch.isfinite(y).all() and torch.isfinite(y1).all()
    _close(y, ref, 5e-2 * ref.abs().max().item(), 0.0, what="autocast eval vs fp32 eval")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_cross_attention_fused_step_equals_the_op_by_op_path(dtype):
    """BlindCrossAttention for one token: lina_cross_att_step (LayerNorm + softmax(q k^T) + weighted read, twice) against
    the torch sequence of model/crossatt.py:105-155 (eval branch) -- outputs, both attention rows, and the pos_net state."""
    import lina_speech_b200.model as M
    import lina_speech_b200.model.crossatt as CA
    torch.manual_seed(2)
    d, B, n = 256, 3, 37
    rnn = M.AttentiveGLA(d, 1, 4, blind=True, use_short_conv=True, pos_type="convolutional").to(DEV).to(dtype).eval()
    ca = rnn.cross_att
    ctx = torch.randn(B, n, d, device=DEV, dtype=dtype)
    ys = torch.randn(B, 4, d, device=DEV, dtype=dtype)
    res = {}
    saved = CA.FUSED_STEP
    try:
        for name, flag in (("fused", True), ("torch", False)):
            CA.FUSED_STEP = flag
            ca.clear_memo()
            cache = rnn.init_state(batch_size=B)
            outs, atts = [], []
            with torch.inference_mode():
                for t in range(ys.shape[1]):
                    o, a = ca(ys[:, t:t + 1], ctx, time_step=t, past_key_values=cache, use_cache=True)
                    outs.append(o); atts.append(a)
            res[name] = (torch.cat(outs, 1).float().cpu(), torch.cat(atts, 2).float().cpu(), cache.states[2][-1].float().cpu())
    finally:
        CA.FUSED_STEP = saved
    tol = 2e-5 if dtype == torch.float32 else 2e-2
    for i, what in enumerate(("output", "attention rows", "pos_net state")):
        a, b = res["fused"][i], res["torch"][i]
        assert a.shape == b.shape
        err = (a - b).abs().max().item()
        assert err <= tol * max(1.0, b.abs().max().item()), f"{what} ({dtype}): max diff {err:.3e}"
    assert abs(res["fused"][1].sum(-1) - 1).max() < 2e-2
