#!/usr/bin/env python
"""Headline benchmark: codec-tokens/sec/GPU of LinaModel d1024 l12 (13 GLA blocks) at bs32 x seq2048.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one teacher-forced pass of LinaModel.forward (the chunkwise-parallel GLA path) over one
batch of 32 synthetic (text, codec-token) sequences of 2048 tokens -> 65 536 codec tokens per step per
GPU.  ``value`` is timed with inputs resident in HBM; ``e2e`` is the same pass through the public API
with HOST (pinned) token ids copied in and the loss read back inside the timed region.  The same run
also times the autoregressive decode loop (generate_batch, CUDA-graphed step) for the RTF figure.

``--impl reference`` times the reference's CPU algorithm (the oracle port in oracle/, torch fp32 on all
host cores) on a bounded sample of the same workload.  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

# NCCL writes its "NCCL version ..." banner (NCCL_DEBUG=VERSION/WARN on these boxes) to stdout by default; stdout
# must carry exactly one JSON line, so NCCL's own output goes to stderr.
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = ("LinaModel d1024 l12 (AttentiveGLA n_layer=6: 13 GLA blocks, 207M params) teacher-forced forward, "
            "bs32 x seq2048 per GPU (BASELINE configs[1])")
METRIC = "codec_tokens_per_sec"
UNIT = "tokens/s"
CFG = dict(d_model=1024, n_layer=6, heads=4, n_codebook=4096, n_txt_vocab=256, txt_layers=4, txt_heads=4,
           batch=32, seq=2048, txt_len=128)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled through NVML during the timed region (B200_PROFILING.md).  NVML is
    initialised in the constructor (main thread, outside the timed region) so the first sample is immediate."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.max_mhz, self.reasons, self.nv, self.h = None, set(), None, None
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv, self.h = nv, nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception as e:      # noqa: BLE001  (clock sampling must never kill the bench)
            self.reasons.add(f"sampler_error:{type(e).__name__}")

    def _sample(self):
        nv = self.nv
        self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        for n, bit in (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown),
                       ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                       ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown),
                       ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap)):
            if r & bit:
                self.reasons.add(n)

    def run(self):
        if self.nv is None:
            return
        try:
            while not self.stop_flag:
                self._sample()
                time.sleep(0.01)
        except Exception as e:      # noqa: BLE001
            self.reasons.add(f"sampler_error:{type(e).__name__}")

    def result(self):
        self.stop_flag = True
        self.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def build_model(device, dtype):
    import lina_speech_b200.model as m
    torch.manual_seed(0)
    c = CFG
    rnn = m.AttentiveGLA(c["d_model"], c["n_layer"], c["heads"], blind=True, use_short_conv=True,
                         pos_type="convolutional")
    lm = m.LinaModel(rnn, c["d_model"], 1, c["n_codebook"], 3, 3, c["n_txt_vocab"],
                     txt_encoder=m.TextEncoder(c["d_model"], c["txt_heads"], n_layers=c["txt_layers"], dropout=0.0,
                                               rotary=False))
    return lm.to(device=device, dtype=dtype).eval()


def synth_inputs(B, T, Tx, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randint(3, CFG["n_txt_vocab"], (B, Tx), generator=g)
    y = torch.randint(3, CFG["n_codebook"] + 3, (B, T + 1, 1), generator=g)
    y[:, 0] = 1
    enc_mask = torch.ones(B, Tx, Tx, dtype=torch.bool)
    ca_mask = torch.ones(B, T + 1, Tx, dtype=torch.bool)
    return x, y, enc_mask, ca_mask


# --------------------------------------------------------------------------------------------------
_CPU_REF = {}


def _cpu_reference_setup():
    """Oracle-side model (reference-keyed fp32 state dict) + the host thread count that runs it fastest; built once."""
    if _CPU_REF:
        return _CPU_REF
    from oracle import lina_oracle as LO
    c = CFG
    cfg = {"d_model": c["d_model"], "n_layer": c["n_layer"], "heads": c["heads"], "txt_heads": c["txt_heads"],
           "txt_layers": c["txt_layers"], "pos_type": "convolutional"}
    # same architecture, default-initialiser scales, built by the oracle itself: nothing of the product package on this path
    sd = LO.random_state_dict(cfg, c["n_codebook"], 3, c["n_txt_vocab"], seed=0)
    B = 2

    def run(T):
        x, y, em, cm = synth_inputs(B, T, c["txt_len"], 1)
        t0 = time.perf_counter()
        with torch.no_grad():
            LO.lina_forward(sd, cfg, x, y, em, cm)
        return time.perf_counter() - t0

    cores_all = os.cpu_count() or 1
    probe_T, best = 8, None
    for nt in sorted({cores_all, min(cores_all, 32), min(cores_all, 16)}, reverse=True):   # many-core hosts: all threads is not always fastest
        torch.set_num_threads(nt)
        run(probe_T)                               # warm-up (thread pools, allocator)
        dt = min(run(probe_T), run(probe_T))
        if best is None or dt < best[0]:
            best = (dt, nt)
    torch.set_num_threads(best[1])
    _CPU_REF.update(run=run, B=B, cores=best[1], probe_rate=B * probe_T / best[0])
    return _CPU_REF


def cpu_reference_rate(budget_s: float, repeats: int = 1):
    """tokens/s of the reference's CPU algorithm (oracle port, torch fp32, host threads) on a bounded sample of the
    same workload: same architecture and weight init, B=2, T sized to the time budget."""
    ref = _cpu_reference_setup()
    B, c = ref["B"], CFG
    T = int(max(8, min(256, ref["probe_rate"] * budget_s / B)))
    times = [ref["run"](T) for _ in range(repeats)]
    _CPU_REF["last_ms"] = statistics.median(times) * 1e3
    return (B * T / statistics.median(times), ref["cores"],
            f"oracle port, fp32, B={B} T={T} of {c['batch']}x{c['seq']}, {repeats} pass(es)")


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = max(1, args.steps + args.warmup)
    val, cores, sample = cpu_reference_rate(budget_s=max(1.0, args.cpu_budget / n), repeats=1)
    vals = [val]
    for _ in range(args.steps - 1):
        vals.append(cpu_reference_rate(budget_s=max(1.0, args.cpu_budget / n))[0])
    v = statistics.median(vals)
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": _CPU_REF.get("last_ms"), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOAD, "global_batch": CFG["batch"], "seq_len": CFG["seq"], "text_len": CFG["txt_len"],
                      "parallelism": "cpu", "cpu_sample": sample},
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def train_step_metrics(dev, B=8, T=4096, Tx=256, steps=3):
    torch.cuda.empty_cache()
    lm = build_model(dev, torch.float32).train()
    opt = torch.optim.AdamW(lm.parameters(), lr=2e-4, betas=(0.9, 0.95), weight_decay=0.1)
    x, y, em, cm = synth_inputs(B, T, Tx, seed=7)
    xd, yd, emd, cmd = x.to(dev), y.to(dev), em.to(dev), cm.to(dev)

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = lm(xd, yd, emd, cmd)[1]
        loss.backward()
        opt.step()
        return loss

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out = {"workload": "LinaModel d1024 l12 train step (fwd + bwd + AdamW), bf16 autocast, bs8 x seq4096 (BASELINE configs[3])",
           "batch": B, "seq_len": T, "text_len": Tx, "steps": steps, "ms_per_step": ms, "tokens_per_s": B * T / (ms * 1e-3),
           "loss": float(loss.detach()), "gla_backward": "tensor cores (5 runs of the pre-gated tcgen05 kernel)",
           "peak_mem_GB": torch.cuda.max_memory_allocated() / 2 ** 30}
    del lm, opt
    torch.cuda.empty_cache()
    return out


def codec_metrics(dev, B=32, Ln=750, steps=5):
    """SURVEY 8d cfg 5's tail: WavTokenizer ``codes_to_features`` + ``decode`` of [1, B, 750] codes -> [B, 240000] fp32
    waveform (10 s of 24 kHz audio per sequence) at the reference's fp32 precision, shipped widths, random weights."""
    from lina_speech_b200.codec import WavTokenizer
    from lina_speech_b200.codec import wavtokenizer as WT
    torch.manual_seed(0)
    wt = WavTokenizer.from_hparams().eval()
    with torch.no_grad():
        wt.feature_extractor.encodec.quantizer.vq.layers[0]._codebook.embed.normal_()
    wt = wt.to(dev)
    g = torch.Generator().manual_seed(2)
    codes_h = torch.randint(0, 4096, (1, B, Ln), generator=g).pin_memory()
    codes = codes_h.to(dev)
    bw = torch.tensor([0], device=dev)

    def run(c):
        return wt.decode(wt.codes_to_features(c), bandwidth_id=bw)

    for _ in range(3):
        run(codes)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        run(codes)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    wav_h = torch.empty(B, 320 * Ln, dtype=torch.float32).pin_memory()
    e0.record()
    for _ in range(steps):                                    # end to end: host codes in, host waveform out
        wav_h.copy_(run(codes_h.to(dev, non_blocking=True)), non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1) / steps
    WT.PROFILE = []
    run(codes)
    torch.cuda.synchronize()
    prof, WT.PROFILE = WT.PROFILE, None
    pk = peaks()
    stages = {}
    for name, nbytes, flops, a, b in prof:
        d = stages.setdefault(name, {"launch_groups": 0, "ms": 0.0, "bytes": 0, "flops": 0})
        d["launch_groups"] += 1
        d["ms"] += a.elapsed_time(b)
        d["bytes"] += nbytes
        d["flops"] += flops
    for d in stages.values():
        if d["ms"] > 0 and d["bytes"]:
            d["hbm_frac"] = d["bytes"] / (d["ms"] * 1e-3) / 1e9 / pk["hbm_gbs"]
        if d["ms"] > 0 and d["flops"]:          # bf16 tensor-core work actually issued (all part products)
            d["tflops_bf16"] = d["flops"] / (d["ms"] * 1e-3) / 1e12
            d["tensor_frac_of_burst"] = d["tflops_bf16"] / pk["bf16_tflops"]
    own_ms = sum(d["ms"] for d in stages.values())
    frames = B * Ln
    # the same decode with two-part operands (16 significand bits per operand): reported beside the default, not instead of it
    alt = {}
    try:
        wt.gemm_precision = "bf16x2"
        ref_wav = None
        for _ in range(2):
            run(codes)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            w2 = run(codes)
        e1.record()
        torch.cuda.synchronize()
        wt.gemm_precision = "bf16x3"
        ref_wav = run(codes)
        alt = {"precision": "bf16x2", "ms": e0.elapsed_time(e1) / steps,
               "max_abs_diff_vs_bf16x3": float((w2 - ref_wav).abs().max()), "wav_absmax": float(ref_wav.abs().max())}
    except Exception as e:      # noqa: BLE001
        alt = {"error": repr(e)[:200]}
        wt.gemm_precision = "bf16x3"
    return {"workload": f"WavTokenizer codes_to_features + decode, {B} x {Ln} frames (10 s each), dim 768 / 2304, 12 ConvNeXt, "
                        "n_fft 1280 hop 320, fp32-equivalent arithmetic (SURVEY 8d cfg 5)",
            "precision": getattr(wt, "gemm_precision", "fp32") + " (six bf16 part products per contraction: 24 significand bits "
                         "per operand, fp32 accumulate; no cuBLAS / cuDNN call)", "two_part_mode": alt,
            "batch": B, "frames": Ln, "ms": ms,
            "frames_per_s": frames / (ms * 1e-3), "x_realtime": frames / 75.0 / (ms * 1e-3),
            "e2e_ms": ms_e2e, "e2e_frames_per_s": frames / (ms_e2e * 1e-3), "h2d_bytes": codes_h.numel() * 8,
            "d2h_bytes": wav_h.numel() * 4, "algorithmic_flops": 125.6e6 * frames,
            "tflops": 125.6e6 * frames / (ms * 1e-3) / 1e12,
            "layer_boundary_bytes": 220 * 1024 * frames,
            "layer_boundary_hbm_frac": 220 * 1024 * frames / (ms * 1e-3) / 1e9 / pk["hbm_gbs"],
            "own_kernels_ms": own_ms, "stages": stages}


def decode_prompt_metrics(lm, dev, x_txt, B=128, p_len=225, gen=750, grp=None, world=1):
    """BASELINE configs[2] / SURVEY 8d cfg 3: bs 128, a 225-token (3 s) prompt continued for 750 tokens (10 s), k = 100,
    ``force_max_seqlen``; both ways of consuming the prompt (token by token as the reference, one chunkwise prefill pass)."""
    import torch.distributed as dist
    g = torch.Generator().manual_seed(9)
    prompt = torch.randint(3, 4099, (1, B, p_len), generator=g).to(dev)      # already ids (+3), one per sequence
    out = {"batch": B, "prompt_tokens": p_len, "generated_tokens": gen, "k": 100, "state_dtype": "bf16"}
    for name, pre in (("token_by_token", False), ("prefill", True)):
        tm = {}
        lm.generate_batch(x_txt, batch_size=B, prompt=prompt, max_seqlen=p_len + 8, k=100, force_max_seqlen=True,
                          cuda_graph=True, dist_group=grp, prefill_prompt=pre, stop_check_interval=1 << 30)
        torch.cuda.synchronize()
        lm.generate_batch(x_txt, batch_size=B, prompt=prompt, max_seqlen=p_len + gen, k=100, force_max_seqlen=True,
                          cuda_graph=True, dist_group=grp, prefill_prompt=pre, stop_check_interval=1 << 30, _timing=tm)
        torch.cuda.synchronize()
        ms = tm["start"].elapsed_time(tm["end"])
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        out[name] = {"total_ms": ms, "generated_tokens_per_s": world * B * gen / (ms * 1e-3),
                     "rtf_24khz": (ms * 1e-3) / (gen / 75.0), "ms_per_loop_step": ms / max(1, tm["steps"])}
    return out


class _ByteTokenizer:
    """Stand-in for the notebook's bpe256 tokenizer: ids in [3, 256)."""

    def encode(self, s):
        return [3 + (b % 253) for b in s.encode()]


def tuning_metrics(dev, n_samples=96):
    """InferenceLina.ipynb cell 15 / initial_state.py:85-160: ``train_initial_state`` through the fp32 d1024 l12 model, batch 2,
    grad_acc 4, rank 1, ``fused_recurrent`` forward + backward (dh0).  The reference's one published figure is 12.28 it/s on
    the author's GPU (120 iterations in 9 s, utterances of the Expresso set: here 3-10 s synthetic utterances)."""
    from lina_speech_b200.tuning import train_initial_state
    lm = build_model(dev, torch.float32)
    g = torch.Generator().manual_seed(4)
    ds = []
    for i in range(16):
        n = int(torch.randint(225, 751, (1,), generator=g))
        ds.append({"audio_token": torch.randint(0, 4096, (1, n), generator=g),
                   "text": "".join(chr(97 + int(c)) for c in torch.randint(0, 26, (60 + 5 * i,), generator=g))})
    tok = _ByteTokenizer()
    train_initial_state(lm, ds, tok, 16)                      # warm-up: 8 iterations (allocator, cuBLAS handles)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    params, losses = train_initial_state(lm, ds, tok, n_samples)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    its = n_samples // 2
    out = {"workload": "train_initial_state, fp32 d1024 l12, batch 2, grad_acc 4, rank 1, utterances of 225-750 codec tokens",
           "iterations": its, "seconds": dt, "it_per_s": its / dt, "utterances_per_s": 2 * its / dt,
           "loss_first": losses[0], "loss_last": losses[-1],
           "reference_published_it_per_s": 12.28, "reference_source": "InferenceLina.ipynb cell 15 (GPU not stated)",
           "timing": "wall clock around the call (host loop with a .item() per iteration, as the reference's)"}
    del lm
    torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------------------------------
def main_ours(args):
    import torch.distributed as dist
    from lina_speech_b200 import _lib
    from lina_speech_b200.fla_api import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout must carry exactly ONE JSON line.  Native libraries write to file descriptor 1 behind Python's back (NCCL prints
    # its version banner there at communicator creation, whatever NCCL_DEBUG_FILE says), so fd 1 is pointed at stderr for the
    # whole run and restored only to emit the line.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line: str):
        import ctypes
        sys.stdout.flush()
        try:
            ctypes.CDLL(None).fflush(None)          # C stdio buffers of native libraries, while fd 1 is still stderr
        except Exception:       # noqa: BLE001
            pass
        os.write(real_stdout, (line + "\n").encode())

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (ours) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()

    c = CFG
    B, T, Tx = c["batch"], c["seq"], c["txt_len"]
    dtype = torch.bfloat16
    lm = build_model(dev, dtype)
    x, y, em, cm = synth_inputs(B, T, Tx, seed=1000 + rank)
    xh, yh = x.pin_memory(), y.pin_memory()
    xd, yd, emd, cmd = x.to(dev), y.to(dev), em.to(dev), cm.to(dev)
    tokens_per_step = B * T

    def step_resident():                 # eager launches (~380 per pass): the region in which the GLA kernel is event-timed
        with torch.inference_mode():
            return lm(xd, yd, emd, cmd)[1]

    def step_graphed():                  # the same pass replayed from its CUDA graph, inputs resident in HBM
        with torch.inference_mode():
            return lm.forward_graphed(xd, yd, emd, cmd)[1]

    def step_e2e():
        # the public end-to-end call: LinaModel.forward_graphed copies the pinned HOST token ids into the static inputs of a CUDA
        # graph of the same teacher-forced pass (captured on the first call, before the timed region) and replays it
        with torch.inference_mode():
            loss = lm.forward_graphed(xh, yh, emd, cmd)[1]
            return float(loss.item())                      # D2H read of the step's result

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, mid=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn()
            if mid is not None:
                mid()               # NVML query from the launching thread: the GPU is busy with the queued steps
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    if args.profile:                       # for ncu: just W + K resident steps, nothing else
        for _ in range(args.warmup + args.steps):
            step_resident()
        torch.cuda.synchronize()
        return
    for _ in range(max(3, args.warmup)):
        step_resident()
    step_graphed()                       # capture (outside every timed region)
    for _ in range(max(3, args.warmup)):
        step_graphed()
    step_e2e()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = _lib.launches()
    def mid_sample():
        # one cheap NVML query from the launching thread (the GPU is busy with the queued step); the throttle reasons are
        # read by the background thread only -- a slow NVML call here would drain the launch queue inside the timed region
        if sampler.nv is not None:
            try:
                sampler.samples.append(sampler.nv.nvmlDeviceGetClockInfo(sampler.h, sampler.nv.NVML_CLOCK_SM))
            except Exception:       # noqa: BLE001
                pass

    # headline: K replays of the pass's CUDA graph (one launch per step; on these boxes the eager pass's ~380 launches cost about
    # as much host time as the pass takes on the GPU, so an eager timed region measures the host on the slower ones)
    ms = timed(step_graphed, args.steps, mid_sample)
    launches = _lib.launches() - l0
    clocks = sampler.result()
    graphed = bool(lm._fwd_graphs) and all(g is not False for g in lm._fwd_graphs.values())
    # second timed region, eager launches of the same K steps: CUDA events around each GLA-kernel launch (events cannot be
    # recorded inside a graph replay) -> roofline.launch_ms
    ops.PROFILE = []
    ms_eager = timed(step_resident, args.steps)
    prof, ops.PROFILE = ops.PROFILE, None
    ms_e2e = timed(step_e2e, args.steps)
    graphed_e2e = graphed

    # dominant kernel of ours: the GLA chunk-forward launch (13 per step)
    H, K, V = c["heads"], c["d_model"] // c["heads"], 2 * c["d_model"] // c["heads"]
    kern_ms = [a.elapsed_time(b) for (_, a, b) in prof]
    k_ms = statistics.mean(kern_ms) if kern_ms else float("nan")
    pregated = any(kind == "chunk_pregated" for (kind, _, _) in prof)
    if pregated:     # q~, k~, v in + o out (bf16) + the per-chunk decay vectors (fp32); gk is consumed by the prep pass
        alg_bytes = B * H * T * (2 * K + 2 * V) * 2 + B * H * ((T + 63) // 64) * K * 4
    else:
        alg_bytes = B * H * T * (3 * K + 2 * V) * 2                  # bf16 q,k,gk + v,o per launch
    alg_flops = B * H * T * (4 * K * V + 64 * (K + V))               # SURVEY 8(d), C = 64
    pk = peaks()
    uses_tc = bool(_lib.lib().lina_gla_chunk_fwd_uses_tensor_cores(B, H, T, K, V, _lib.BF16))
    traffic = None                                   # DRAM bytes per launch from the committed ncu --set full capture
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic_r02.json" if pregated else "ncu_traffic_r01.json")
    if uses_tc and os.path.exists(tpath) and (B, H, T, K, V) == (32, 4, 2048, 256, 512):
        with open(tpath) as f:
            t = json.load(f).get("gla_chunk_fwd_sm100_kernel<256,292>" if pregated else "gla_chunk_fwd_sm100_kernel<256>")
        if t:
            traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
    roofline = {"kernel": ("lina_gla_chunk_fwd_pregated_bthd_ws (tcgen05, CTA pairs + T cut, operands gated by lina_gla_prefill_prep_gated)" if pregated
                           else "lina_gla_chunk_fwd (tcgen05)" if uses_tc else "lina_gla_chunk_fwd (CUDA-core recurrence)"),
                "bound": "hbm", "achieved": alg_bytes / (k_ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": alg_bytes / (k_ms * 1e-3) / 1e9 / pk["hbm_gbs"], "traffic": traffic,
                "traffic_source": (os.path.relpath(tpath, ROOT) + " (ncu --set full of this kernel at this shape; not re-captured per run)"
                                   if traffic is not None else None),
                "peak_source": pk["src"],
                "launch_ms": k_ms, "launches_timed": len(kern_ms), "share_of_step": sum(kern_ms) / ms_eager,
                "timed_in": "the eager-launch region of the same K steps (eager_ms_per_step), right after the graph-replay region",
                "algorithmic_bytes": alg_bytes, "tensor_achieved_tflops": alg_flops / (k_ms * 1e-3) / 1e12,
                "tensor_frac_of_sustained": alg_flops / (k_ms * 1e-3) / 1e12 / pk["bf16_tflops_sustained"]}

    grp = dist.group.WORLD if world > 1 else None          # batch-sharded generation: token all-gather per step
    # autoregressive decode loop (generate_batch) for tokens/s per stream and RTF
    dec_steps = 750                                   # 10 s of audio per stream
    xt = x[0]
    tm = {}
    lm.generate_batch(xt.to(dev), batch_size=B, max_seqlen=8, k=100, force_max_seqlen=True, cuda_graph=True, dist_group=grp)
    l1 = _lib.launches()
    lm.generate_batch(xt.to(dev), batch_size=B, max_seqlen=dec_steps, k=100, force_max_seqlen=True, cuda_graph=True,
                      dist_group=grp, stop_check_interval=1 << 30, _timing=tm)
    torch.cuda.synchronize()
    dec_ms = tm["start"].elapsed_time(tm["end"]) / tm["steps"]
    if world > 1:
        t = torch.tensor([dec_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dec_ms = float(t.item())
    n_blocks = 2 * c["n_layer"] + 1
    state_bytes = B * n_blocks * 2 * H * K * V * 2                    # bf16 cache, read + write per step
    decode = {"batch": B, "steps": tm["steps"], "ms_per_step": dec_ms, "tokens_per_s": world * B / (dec_ms * 1e-3),
              "tokens_per_s_per_stream": 1.0 / (dec_ms * 1e-3), "rtf_24khz": 75.0 / (1.0 / (dec_ms * 1e-3)),
              "state_dtype": "bf16", "state_bytes_per_step": state_bytes,
              "state_hbm_frac": state_bytes / (dec_ms * 1e-3) / 1e9 / pk["hbm_gbs"], "cuda_graph": True,
              "token_all_gather": world > 1}

    # same loop at the batch where the state traffic dominates (BASELINE configs[2]: bs 128)
    B2 = 128
    tm2 = {}
    lm.generate_batch(xt.to(dev), batch_size=B2, max_seqlen=8, k=100, force_max_seqlen=True, cuda_graph=True, dist_group=grp)
    lm.generate_batch(xt.to(dev), batch_size=B2, max_seqlen=750, k=100, force_max_seqlen=True, cuda_graph=True,
                      dist_group=grp, stop_check_interval=1 << 30, _timing=tm2)
    torch.cuda.synchronize()
    dec2_ms = tm2["start"].elapsed_time(tm2["end"]) / tm2["steps"]
    if world > 1:
        t = torch.tensor([dec2_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dec2_ms = float(t.item())
    sb2 = B2 * n_blocks * 2 * H * K * V * 2
    decode_bs128 = {"batch": B2, "steps": tm2["steps"], "ms_per_step": dec2_ms, "tokens_per_s": world * B2 / (dec2_ms * 1e-3),
                    "rtf_24khz": 75.0 * dec2_ms * 1e-3, "state_bytes_per_step": sb2,
                    "state_hbm_frac": sb2 / (dec2_ms * 1e-3) / 1e9 / pk["hbm_gbs"]}

    extras = {}

    def leg(name, fn, *a, **kw):
        try:
            extras[name] = fn(*a, **kw)
        except Exception as e:      # noqa: BLE001  (an extra section must never take the headline line down)
            extras[name] = {"error": repr(e)[:300]}

    if not args.no_extras:
        leg("decode_prompt", decode_prompt_metrics, lm, dev, xt.to(dev), 128, 225, 750, grp, world)
        leg("codec", codec_metrics, dev)
        if world == 1:
            leg("init_state_tuning", tuning_metrics, dev)

    # BASELINE configs[3]: one training step (fwd + bwd + AdamW, bf16 autocast over fp32 parameters) at bs8 x seq4096,
    # GLA backward on the tensor-core path (five runs of the pre-gated tcgen05 kernel); rank 0 at N = 1 only
    train = None
    if world == 1 and not args.no_train:
        try:
            train = train_step_metrics(dev)
        except Exception as e:      # noqa: BLE001  (an extra section must never take the headline line down)
            train = {"error": repr(e)[:300]}

    if rank == 0:
        cpu_v, cores, sample = cpu_reference_rate(budget_s=20.0) if world == 1 else (None, None, None)
        per_step = ms / args.steps
        out = {"metric": METRIC, "value": world * tokens_per_step / (per_step * 1e-3), "unit": UNIT, "n_gpus": world,
               "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": per_step, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
               "config": {"workload": WORKLOAD,
                          "global_batch": world * B, "seq_len": T, "text_len": Tx, "parallelism": f"dp{world}",
                          "l2": "per-step working set (>= 64 MB per activation, 0.4 GB weights) exceeds the 126 MB L2; no flush"},
               "e2e": {"value": world * tokens_per_step / (ms_e2e / args.steps * 1e-3), "unit": UNIT,
                       "h2d_bytes_per_step": xh.numel() * 8 + yh.numel() * 8, "d2h_bytes_per_step": 4,
                       "path": ("LinaModel.forward_graphed: pinned host ids -> static graph inputs, one CUDA-graph replay of the same "
                                "teacher-forced pass, loss read back" if graphed_e2e else
                                "LinaModel.forward (eager; the graph path declined: gates not certified by the weights)"),
                       "value_path": ("LinaModel.forward_graphed, inputs resident (graph replay)" if graphed else
                                      "LinaModel.forward (eager; the graph path declined)")},
               "eager_ms_per_step": ms_eager / args.steps,
               "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "decode": decode,
               "decode_bs128": decode_bs128}
        out.update(extras)
        if train is not None:
            out["train_step"] = train
        if cpu_v is not None:
            out["cpu_baseline"] = {"value": cpu_v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-budget", type=float, default=150.0, help="seconds of CPU work for the whole --impl reference run")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step section")
    ap.add_argument("--no-extras", action="store_true", help="skip the codec / prompted-decode / state-tuning sections")
    ap.add_argument("--profile", action="store_true", help="run only W+K resident steps (for ncu); prints nothing")
    a = ap.parse_args()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_ours(a)
