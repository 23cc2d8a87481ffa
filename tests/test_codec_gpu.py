"""GPU parity of the WavTokenizer decode path against the reference's golden waveforms and the oracle."""
import os
import pytest
import torch

from oracle import codec_oracle as CO

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _wt(golden_codec):
    from lina_speech_b200.codec import WavTokenizer
    sd = {k[2:]: v for k, v in golden_codec.items() if k.startswith("w.")}
    wt = WavTokenizer.from_hparams(vq_bins=64, dim=64, intermediate_dim=128, num_layers=2)
    wt.load_reference_state_dict(sd)
    return wt.to(DEV).eval(), sd


@pytest.mark.parametrize("Ln", [1, 7, 40])
def test_decode_matches_reference_golden(golden_codec, Ln):
    wt, _ = _wt(golden_codec)
    codes, bw = golden_codec[f"L{Ln}_codes"].to(DEV), golden_codec[f"L{Ln}_bw"].to(DEV)
    feats = wt.codes_to_features(codes)
    wav = wt.decode(feats, bandwidth_id=bw)
    ref = golden_codec[f"L{Ln}_wav"]
    assert wav.shape == ref.shape == (codes.shape[1], 320 * Ln)
    err = (wav.cpu() - ref).abs().max().item()
    assert err <= 1e-4 * max(1.0, ref.abs().max().item()), f"wav max err {err:.3e}"


@pytest.mark.parametrize("B,Ln", [(1, 75), (4, 225), (2, 750)])
def test_decode_matches_oracle_other_lengths(golden_codec, B, Ln):
    wt, sd = _wt(golden_codec)
    torch.manual_seed(Ln)
    codes = torch.randint(0, 64, (1, B, Ln))
    bw = torch.tensor([2])
    ref = CO.decode(sd, CO.codes_to_features(sd, codes), bw)
    feats = wt.codes_to_features(codes.to(DEV))
    assert torch.equal(feats.cpu(), CO.codes_to_features(sd, codes))          # gather + sum of one codebook: exact
    wav = wt.decode(feats, bandwidth_id=bw.to(DEV))
    assert wav.shape == (B, 320 * Ln)
    err = (wav.cpu() - ref).abs().max().item()
    assert err <= 2e-4 * max(1.0, ref.abs().max().item()), f"wav max err {err:.3e}"


def test_istft_head_alone_against_torch_irfft():
    """the FFT/OLA kernel in isolation at the real head width (n_fft 1280, hop 320), large phases included."""
    from lina_speech_b200.codec import ISTFTHead
    torch.manual_seed(0)
    head = ISTFTHead(32, 1280, 320).to(DEV)
    with torch.no_grad():
        head.out.weight.mul_(30.0)
        head.out.bias.normal_()
    x = torch.randn(3, 50, 32, device=DEV)
    wav = head(x)
    sd = {"head.out.weight": head.out.weight.detach().cpu(), "head.out.bias": head.out.bias.detach().cpu(),
          "head.istft.window": head.istft.window.cpu()}
    ref = CO.istft_head(sd, x.cpu(), 320)
    err = (wav.cpu() - ref).abs().max().item()
    assert err <= 2e-4 * max(1.0, ref.abs().max().item()), f"istft max err {err:.3e} (ref absmax {ref.abs().max():.2f})"


@pytest.mark.parametrize("B,Ln", [(5, 77), (4, 700)])      # 2800 frames: more than one pass of the persistent grid
def test_istft_warp_per_frame_kernel_matches_the_generic_one(B, Ln):
    """the fixed-radix warp-per-frame FFT (csrc/fft640.cuh; host-checked in tests/test_host.py), the default for n_fft = 1280,
    against the generic shared-memory FFT kernel."""
    from lina_speech_b200.codec import ISTFTHead
    from lina_speech_b200 import _lib as L
    torch.manual_seed(1)
    head = ISTFTHead(32, 1280, 320).to(DEV)
    with torch.no_grad():
        head.out.weight.mul_(30.0)
        head.out.bias.normal_()
    x = torch.randn(B, Ln, 32, device=DEV)
    wav = head(x)                                   # default: the warp-per-frame kernel
    L.lib().lina_debug_set_variant(8, 2)            # key 8 = 2: the generic shared-memory FFT
    try:
        base = head(x)
    finally:
        L.lib().lina_debug_set_variant(8, 0)
    err = (wav - base).abs().max().item()
    assert err <= 2e-5 * max(1.0, base.abs().max().item()), f"max diff {err:.3e}"


def test_two_part_products_stay_close(golden_codec):
    """gemm_precision='bf16x2' (three part products, 16 significand bits per operand) against the reference golden: the
    head's exp() amplifies operand error, so this mode is reported, and bounded at 100 x the fp32-equivalent mode's tolerance."""
    wt, sd = _wt(golden_codec)
    wt.gemm_precision = "bf16x2"
    codes, bw = golden_codec["L40_codes"].to(DEV), golden_codec["L40_bw"].to(DEV)
    wav = wt.decode(wt.codes_to_features(codes), bandwidth_id=bw)
    ref = golden_codec["L40_wav"]
    err = (wav.cpu() - ref).abs().max().item()
    print(f"bf16x2 waveform max abs error {err:.3e} (ref absmax {ref.abs().max().item():.3f})")
    assert torch.isfinite(wav).all() and err <= 1e-2 * max(1.0, ref.abs().max().item())
    wt.gemm_precision = "nope"
    with pytest.raises(ValueError):
        wt.decode(wt.codes_to_features(codes), bandwidth_id=bw)


def test_features_keep_the_reference_shape_and_decode_accepts_plain_tensors(golden_codec):
    """codes_to_features returns [B, C, L] like the reference (a channels-last buffer underneath); decode gives the same
    waveform for that tensor and for an ordinary contiguous [B, C, L] copy of it (any caller-made features)."""
    wt, sd = _wt(golden_codec)
    codes, bw = golden_codec["L40_codes"].to(DEV), golden_codec["L40_bw"].to(DEV)
    feats = wt.codes_to_features(codes)
    ref_feats = CO.codes_to_features(sd, golden_codec["L40_codes"])
    assert feats.shape == ref_feats.shape and torch.equal(feats.cpu(), ref_feats)
    a = wt.decode(feats, bandwidth_id=bw)
    b = wt.decode(feats.contiguous().clone(), bandwidth_id=bw)
    assert torch.equal(a, b)


@pytest.fixture(scope="module")
def full_size_codec():
    """The shipped configuration (dim 768, intermediate 2304, 12 ConvNeXt blocks, 4096 x 512 codebook, n_fft 1280 / hop 320;
    SURVEY 3d) with seeded random weights: torch's default initialisers, a N(0,1) codebook (the reference's buffer is
    zeros until k-means fills it) and non-trivial AdaLayerNorm tables."""
    from lina_speech_b200.codec import WavTokenizer
    torch.manual_seed(7)
    wt = WavTokenizer.from_hparams().eval()
    with torch.no_grad():
        wt.feature_extractor.encodec.quantizer.vq.layers[0]._codebook.embed.normal_()
        for m in wt.modules():
            if m.__class__.__name__ == "AdaLayerNorm":
                m.scale.weight.add_(0.1 * torch.randn_like(m.scale.weight))
                m.shift.weight.add_(0.1 * torch.randn_like(m.shift.weight))
    sd = {k: v.detach().clone() for k, v in wt.state_dict().items()}
    return wt.to(DEV), sd


@pytest.mark.parametrize("B,Ln", [(4, 750), (4, 2000)])
def test_full_size_decode_matches_oracle(full_size_codec, B, Ln):
    """SURVEY 4's shape matrix for the codec: the real widths, 10 s and 26.7 s of audio, against the CPU fp32 oracle
    (DEC/pretrained.py:192-239 restated); waveform atol 1e-4 relative to max(1, |ref|max) as BASELINE.md 3 states."""
    wt, sd = full_size_codec
    g = torch.Generator().manual_seed(Ln)
    codes = torch.randint(0, 4096, (1, B, Ln), generator=g)
    bw = torch.tensor([0])
    feats_ref = CO.codes_to_features(sd, codes)
    ref = CO.decode(sd, feats_ref, bw)
    feats = wt.codes_to_features(codes.to(DEV))
    assert torch.equal(feats.cpu(), feats_ref)
    wav = wt.decode(feats, bandwidth_id=bw.to(DEV))
    assert wav.shape == ref.shape == (B, 320 * Ln)
    err = (wav.cpu() - ref).abs().max().item()
    assert err <= 1e-4 * max(1.0, ref.abs().max().item()), f"wav max err {err:.3e} (ref absmax {ref.abs().max().item():.3f})"
