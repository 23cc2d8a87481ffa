"""LinaModel: embeddings, teacher-forced forward, batched autoregressive sampling loop.

Mirrors ``model/modeling_lina.py`` of the reference (``LinaModel.forward`` :61-108,
``generate_batch`` :112-192) -- same constructor, parameter names, arguments and return values --
on top of the B200 GLA backbone.  Additions (all opt-in, defaults reproduce the reference):

  * ``stop_check_interval``: poll the "all sequences stopped" flag every n steps instead of forcing a
    device->host sync per token (modeling_lina.py:172);
  * ``cuda_graph``: capture one decode step (13 GLA blocks + cross attention + logits + sampling +
    embedding) in a CUDA graph and replay it;
  * ``dist_group``: batch-sharded generation over the GPUs of one node -- every rank decodes its own
    slice of the batch and the sampled token ids are all-gathered once per step (NCCL over NVLink).
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.nn.functional as F
from einops import rearrange, reduce, repeat
from torch import Tensor, nn

from .contracts import AttentiveRNN
from .embeddings import MultiEmbedding
from ..parallel import gather_tokens
from .tools import topk_sampling, undelay_rvq


FUSED_SAMPLING = os.environ.get("LINA_FUSED_SAMPLING", "1") != "0"


def exists(x):
    return x is not None


class LogitsHead(nn.Module):
    """EinMix("b n d -> b n q l", weight_shape="q l d") of the reference (modeling_lina.py:51-57):
    a single ``weight[q, l, d]`` parameter, no bias."""

    def __init__(self, q: int, l: int, d: int):
        super().__init__()
        bound = (3.0 / d) ** 0.5
        self.weight = nn.Parameter(torch.empty(q, l, d).uniform_(-bound, bound))

    def forward(self, x):
        q, l, d = self.weight.shape
        if q == 1 and l % 8 and x.is_cuda and not torch.is_grad_enabled():
            # l = n_codebook + 3 = 4099 is odd: the unpadded GEMM runs on cuBLAS' unaligned legacy kernels.
            # Inference pads the vocabulary dim of the weight once to a multiple of 8 and slices the result.
            key = (self.weight.data_ptr(), getattr(self.weight, "_version", 0) if not self.weight.is_inference() else -1,
                   self.weight.dtype)
            if getattr(self, "_wpad", None) is None or self._wpad[0] != key:
                lp = (l + 7) // 8 * 8
                wp = self.weight.new_zeros(lp, d)
                wp[:l] = self.weight.detach()[0]
                self._wpad = (key, wp)
            return F.linear(x, self._wpad[1])[..., :l].unsqueeze(-2)
        if q == 1 and l % 8 and x.is_cuda:      # autograd path: pad differentiably (aligned GEMM kernels), slice the logits
            return F.linear(x, F.pad(self.weight[0], (0, 0, 0, (l + 7) // 8 * 8 - l)))[..., :l].unsqueeze(-2)
        return F.linear(x, self.weight.view(q * l, d)).view(*x.shape[:-1], q, l)


class LinaModel(nn.Module):
    def __init__(self, attentive_rnn: AttentiveRNN, d_model: int, n_quant: int, n_codebook: int,
                 n_special_token_in: int, n_special_token_out: int, n_txt_vocab: int, tie_embed: bool = False,
                 txt_encoder: Optional[nn.Module] = None, spk_encoder: Optional[nn.Module] = None,
                 mask_text_p: float = 0.0):
        super().__init__()
        self.n_quant, self.n_codebook = n_quant, n_codebook
        self.n_special_token_in, self.n_special_token_out = n_special_token_in, n_special_token_out
        self.mask_text_p = mask_text_p
        self.n_txt_vocab = n_txt_vocab + int(mask_text_p > 0.0)
        self.n_target_vocab = n_codebook + n_special_token_out
        self.txt_encoder, self.spk_encoder, self.attentive_rnn = txt_encoder, spk_encoder, attentive_rnn
        self.txt_embed = nn.Embedding(n_txt_vocab, d_model, padding_idx=0)
        self.rvq_embed = MultiEmbedding(n_quant, n_codebook + n_special_token_in, d_model, padding_idx=0)
        self.logits_head = LogitsHead(n_quant, self.n_target_vocab, d_model)
        if tie_embed:
            self.logits_head.weight = self.rvq_embed.weight

    def forward(self, x, y, encoder_mask, crossatt_mask, logits_mask=None, attention_only=False,
                forced_attention=None, init_state=None, crossatt_pos=None):
        """Teacher-forced pass (modeling_lina.py:61-108). x [b,n_txt] ids; y [b,n,q] ids.
        Returns (logits [b,n-1,q,l], loss, att, masked_logits, masked_target)."""
        if self.mask_text_p > 0.0:
            m = torch.empty(x.shape[0]).bernoulli_(self.mask_text_p).bool()
            x[m] = self.n_txt_vocab - 1
        x_embd = self.txt_embed(x)
        y_embd = self.rvq_embed(rearrange(y, "b n q -> q b n")).sum(0)
        x_enc = self.txt_encoder(x_embd, mask=encoder_mask)
        self._new_text()
        if self.spk_encoder is not None:
            y_embd[:, 0] = self.spk_encoder(y_embd)
        y_hat, att = self.attentive_rnn(
            y_embd[:, :-1, :], x_enc, mask=crossatt_mask[:, :-1],
            forced_attention=forced_attention[:, :, :y_embd.shape[1] - 1] if forced_attention is not None else None,
            attention_only=attention_only, init_state=init_state, crossatt_pos=crossatt_pos)
        if attention_only:
            return att
        logits = self.logits_head(y_hat)
        if logits_mask is not None:
            masked_logits = logits[logits_mask[:, 1:], :, :]
            masked_target = y[:, 1:][logits_mask[:, 1:], :]
        else:
            masked_logits, masked_target = logits, y[:, 1:]
        loss = self._fused_cross_entropy(logits, y, logits_mask)
        if loss is None:
            loss = F.cross_entropy(masked_logits.reshape(-1, masked_logits.shape[-1]).float(),
                                   masked_target.reshape(-1), ignore_index=1)
        return logits, loss, att, masked_logits, masked_target

    # ---------------------------------------------------------------------------------------------
    _fwd_graphs = None       # class-level default (un-pickled reference instances never ran this __init__)

    def forward_graphed(self, x, y, encoder_mask, crossatt_mask):
        """``forward(x, y, encoder_mask, crossatt_mask)`` without autograd, replayed from a CUDA graph captured once per input
        signature (shapes / dtypes / device): the ~380 launches of a teacher-forced pass (scoring, evaluation, teacher-forced
        alignment at a fixed batch shape) become one graph launch, which removes the launch gaps between the kernels (1.4 of
        35 ms at bs32 x seq2048).  Inputs may live on the host (pinned): they are copied into the graph's static input buffers.
        Returns the same tuple as ``forward``; the tensors are the graph's static outputs, overwritten by the next call with the
        same signature.  Falls back to the eager pass when the graph cannot be used safely: autograd on, text masking (host
        randomness), or gates that are not certified by the weights (the eager pass reads the device-side envelope flag and
        re-serves the call exactly; a replay cannot)."""
        dev = next(self.parameters()).device
        rnn = self.attentive_rnn
        certified = getattr(rnn, "_all_certified", None)
        if (torch.is_grad_enabled() or dev.type != "cuda" or self.mask_text_p > 0.0 or self.spk_encoder is not None
                or certified is None or torch.cuda.is_current_stream_capturing()):
            return self.forward(x.to(dev), y.to(dev), encoder_mask.to(dev), crossatt_mask.to(dev))
        key = tuple((tuple(t.shape), t.dtype) for t in (x, y, encoder_mask, crossatt_mask)) + (dev.index,)
        if self._fwd_graphs is None:
            self._fwd_graphs = {}
        g = self._fwd_graphs.get(key)
        versions = tuple(p._version for p in self.parameters())
        if g is not None and g is not False and g["versions"] != versions:
            g = None                                          # a parameter was updated in place: tensors derived from the weights
            self._fwd_graphs.pop(key)                         # (padded / concatenated copies) are baked into the graph -> re-capture
        if g is None:
            g = self._fwd_graphs[key] = self._capture_forward(x, y, encoder_mask, crossatt_mask, dev)
            if g is not False:
                g["versions"] = versions
        if g is False or not certified():                     # gates not provably inside the tensor-core envelope: eager pass
            return self.forward(x.to(dev), y.to(dev), encoder_mask.to(dev), crossatt_mask.to(dev))
        for buf, src in zip(g["inputs"], (x, y, encoder_mask, crossatt_mask)):
            buf.copy_(src, non_blocking=True)
        g["graph"].replay()
        from .. import _lib as L
        L.count_launches(g["launches"])
        return g["out"]

    def _capture_forward(self, x, y, encoder_mask, crossatt_mask, dev):
        from .. import _lib as L
        from ..fla_api import ops as fla_ops
        inputs = [torch.empty(t.shape, dtype=t.dtype, device=dev) for t in (x, y, encoder_mask, crossatt_mask)]
        for buf, src in zip(inputs, (x, y, encoder_mask, crossatt_mask)):
            buf.copy_(src)
        prof, fla_ops.PROFILE = fla_ops.PROFILE, None          # timing events cannot be recorded inside a capture
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side), torch.inference_mode():
                for _ in range(2):                             # warm-up: allocations, cuBLAS handles, weight caches, certificates
                    self.forward(*inputs)
            torch.cuda.current_stream(dev).wait_stream(side)
            if not self.attentive_rnn._all_certified():
                return False                                   # gates not provably inside the tensor-core envelope: stay eager
            graph = torch.cuda.CUDAGraph()
            l0 = L.launches()
            with torch.inference_mode(), torch.cuda.graph(graph):
                out = self.forward(*inputs)
            return {"graph": graph, "inputs": inputs, "out": out, "launches": L.launches() - l0}
        finally:
            fla_ops.PROFILE = prof

    def _new_text(self):
        """A new text tensor enters the backbone: drop the cross attention's memo of the previous one."""
        ca = getattr(self.attentive_rnn, "cross_att", None)
        if ca is not None and hasattr(ca, "clear_memo"):
            ca.clear_memo()

    @staticmethod
    def _fused_cross_entropy(logits, y, logits_mask):
        """Inference-only: the mean cross entropy of modeling_lina.py:104-106 in one pass over the (possibly
        vocabulary-padded, bf16) logits -- no contiguous copy, no fp32 copy, no log-softmax tensor.  None = not
        applicable (autograd on, several quantizers, exotic strides): the caller takes the reference's route."""
        if torch.is_grad_enabled() or not logits.is_cuda or logits.dim() != 4 or logits.shape[2] != 1:
            return None
        if logits.dtype not in (torch.float32, torch.bfloat16, torch.float16) or logits.stride(-1) != 1:
            return None
        b, n1, _, l = logits.shape
        ld = logits.stride(1)
        if ld < l or (b > 1 and logits.stride(0) != n1 * ld):
            return None
        from .. import _lib as L
        target = y[:, 1:, 0].long().contiguous()
        mask = logits_mask[:, 1:].bool().contiguous().view(torch.uint8) if logits_mask is not None else None
        rows = torch.empty(2, b * n1, dtype=torch.float32, device=logits.device)
        rc = L.lib().lina_cross_entropy_rows(L.ptr(logits), ld, L.ptr(target), L.ptr(mask), L.ptr(rows[0]), L.ptr(rows[1]),
                                             b * n1, l, 1, L.dt(logits), L.stream(logits))
        L.count_launches(1)
        L.check(rc, "lina_cross_entropy_rows")
        tot = rows.sum(dim=1)
        return tot[0] / tot[1]

    # ---------------------------------------------------------------------------------------------
    def _sample(self, logits, k, first_greedy_quant, temp):
        """logits [b,1,q,l] -> ids [q,b,1] (modeling_lina.py:156-165; NB quantizers with index
        < first_greedy_quant are the *sampled* ones, the rest greedy -- the name is inverted upstream)."""
        if FUSED_SAMPLING and logits.is_cuda and not torch.is_grad_enabled() and logits.stride(-1) == 1 \
                and logits.dtype in (torch.float32, torch.bfloat16, torch.float16) and logits.shape[-1] <= 8192:
            # one launch per quantizer (lina_topk_sample) instead of topk (a sort) + div + compare + masked_fill + softmax +
            # multinomial; same distribution, greedy ids identical
            from .. import _lib as L
            b, _, q, l = logits.shape
            out = torch.empty(q, b, 1, dtype=torch.long, device=logits.device)
            u = torch.rand(q, b, device=logits.device, dtype=torch.float32)
            for i in range(q):
                row = logits[:, 0, i]
                ki, ti = (k, temp) if i < first_greedy_quant else (1, 1.0)
                rc = L.lib().lina_topk_sample(L.ptr(row), row.stride(0), b, l, min(int(ki), l), float(ti), L.ptr(u[i]),
                                              L.ptr(out[i]), L.dt(row), L.stream(row))
                L.count_launches(1)
                L.check(rc, "lina_topk_sample")
            return out
        lg = rearrange(logits, "b 1 q l -> q b l").float()
        out = [topk_sampling(qq, k=k, temp=temp) if i < first_greedy_quant else topk_sampling(qq, k=1)
               for i, qq in enumerate(lg)]
        return torch.stack(out)

    @torch.inference_mode()
    def generate_batch(self, x: Tensor, batch_size: int = 3, prompt: Optional[Tensor] = None, device: str = "cpu",
                       max_seqlen: int = 1000, k: int = 100, first_greedy_quant: int = 1, temp: float = 1.0,
                       init_state=None, force_max_seqlen: bool = False, stop_check_interval: int = 1,
                       cuda_graph: bool = False, dist_group=None, prefill_prompt: bool = False,
                       _timing: Optional[dict] = None):
        """modeling_lina.py:112-192.  Returns (qs [q,b,steps], atts [b,2,steps,n], stop_tokens, cuts).

        With ``dist_group`` every rank passes its LOCAL ``batch_size``; qs / stop_tokens / cuts come back
        for the GLOBAL batch (rank-major), atts stay local.

        ``prefill_prompt`` (new, opt-in; the reference teacher-forces the prompt one token per step): the start
        token and the prompt go through ONE multi-token pass of every block with the cache -- the chunkwise GLA kernels,
        conv tails and the cross attention's pos_net state end where the token-by-token loop would leave them -- and the
        loop starts at the first free position.  Outputs keep their shape (the positions inside the prompt are sampled
        from the same logits as in the loop)."""
        if device == "cpu":
            device = next(self.parameters()).device            # no CPU path: follow the weights
        x = repeat(x, "n -> b n", b=batch_size).to(device)
        stop_token = torch.full((self.n_quant, 1, 1), 2, device=device, dtype=torch.long)
        y_start = torch.ones(self.n_quant, batch_size, 1, device=device, dtype=torch.long)
        x_embd = self.txt_embed(x)
        y_embd = self.rvq_embed(y_start).sum(0)
        p_len = -1
        if exists(prompt):
            if prompt.shape[1] != batch_size:
                prompt = repeat(prompt, "q 1 n -> q b n", b=batch_size) + 3
            prompt = self.rvq_embed(prompt.to(device)).sum(0)
            p_len = prompt.shape[1]
            if self.spk_encoder is not None:
                prompt[:, 0] = self.spk_encoder(prompt)
        x_enc = self.txt_encoder(x_embd)
        self._new_text()
        state = init_state
        if state is None:
            state = self.attentive_rnn.init_state(max_seqlen=max_seqlen, batch_size=batch_size)

        world = 1
        if dist_group is not None:
            import torch.distributed as dist
            world = dist.get_world_size(dist_group)

        def one_step(y_in, t):
            y_out, att, _ = self.attentive_rnn.step(y_in, x_enc, t, state)
            q_s = self._sample(self.logits_head(y_out), k, first_greedy_quant, temp)
            return q_s, att, self.rvq_embed(q_s).sum(0)

        world_b = batch_size * world
        if cuda_graph:
            # ONE graph replay per token and nothing else on the host: besides the model step the graph writes the sampled ids
            # (all-gathered over ``dist_group`` INSIDE the graph) and the attention rows into their slot of the output
            # buffers (device-side step counter), keeps the "all stopped" flag, and picks the next input (prompt embedding
            # while teacher-forcing, the sampled token's embedding afterwards).
            n_txt = x_enc.shape[1]
            qs_out = torch.zeros(self.n_quant, world_b, max_seqlen, device=device, dtype=torch.long)
            t_dev = torch.zeros(1, device=device, dtype=torch.long)
            y_buf = y_embd.clone()
            all_stop = torch.zeros(world_b, 1, device=device, dtype=torch.bool)
            p_tab = prompt if (exists(prompt) and p_len > 0) else None
            p_len_dev = torch.full((1,), max(p_len, 0), device=device, dtype=torch.long)
            atts_out = None

            def graph_step():
                nonlocal atts_out
                q_s, att, emb = one_step(y_buf, 0)
                q_all = gather_tokens(q_s, dist_group) if dist_group is not None else q_s
                qs_out.index_copy_(2, t_dev, q_all)
                if att is not None:
                    if atts_out is None:
                        atts_out = torch.zeros(att.shape[0], att.shape[1], max_seqlen, att.shape[3], device=device, dtype=att.dtype)
                    atts_out.index_copy_(2, t_dev, att)
                all_stop.logical_or_((q_all == stop_token).prod(dim=0).bool())
                if p_tab is not None:                          # modeling_lina.py:175-178: the prompt is teacher-forced
                    forced = p_tab.index_select(1, torch.minimum(t_dev, p_len_dev - 1))
                    emb = torch.where(t_dev < p_len_dev, forced, emb)
                y_buf.copy_(emb)
                t_dev.add_(1)

            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                      # warm-up outside capture (allocations, cuBLAS / NCCL handles)
                snap = [tuple(s_.clone() for s_ in st) for st in state.states]
                y_snap = y_buf.clone()
                for _ in range(2):
                    graph_step()

                def restore():
                    for st, sn in zip(state.states, snap):
                        for a, b in zip(st, sn):
                            a.copy_(b)
                    y_buf.copy_(y_snap)
                    t_dev.zero_()
                    all_stop.zero_()
                    qs_out.zero_()
                restore()
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                graph_step()
            restore()                                          # capture does not execute; restore anyway

        if _timing is not None:                                # bench hook: device time of the steady-state loop
            _timing["start"], _timing["end"] = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            _timing["start"].record()
        qs, atts, stop_tokens = [], [], []
        t_start = 0
        if not cuda_graph:
            all_stop = torch.zeros(world_b, 1, device=device, dtype=torch.bool)
        if prefill_prompt and exists(prompt) and p_len > 0:
            n_pre = min(p_len + 1, max_seqlen)                      # inputs of steps 0 .. p_len: start token, then the prompt
            y_seq = torch.cat([y_embd, prompt[:, :n_pre - 1]], dim=1)
            y_out, att_pre, _ = self.attentive_rnn.step(y_seq, x_enc, 0, state)
            logits_pre = self.logits_head(y_out)
            for t in range(n_pre):
                q_sampled = self._sample(logits_pre[:, t:t + 1], k, first_greedy_quant, temp)
                atts.append(att_pre[:, :, t:t + 1] if att_pre is not None else None)
                q_all = gather_tokens(q_sampled, dist_group) if dist_group is not None else q_sampled
                qs.append(q_all)
                all_stop.logical_or_((q_all == stop_token).prod(dim=0).bool())
            y_embd = self.rvq_embed(q_sampled).sum(0)               # first free position: fed with the last sampled token
            t_start = n_pre
        if cuda_graph:
            if t_start > 0:                                         # hand the prefilled positions to the graph's buffers
                qs_out[:, :, :t_start] = torch.cat(qs, dim=2)
                if atts_out is not None and atts[0] is not None:
                    atts_out[:, :, :t_start] = torch.cat(atts, dim=2)
                y_buf.copy_(y_embd)
                t_dev.fill_(t_start)
            n_steps = t_start
            for t in range(t_start, max_seqlen):
                graph.replay()
                n_steps = t + 1
                if not force_max_seqlen and n_steps % stop_check_interval == 0 and bool(all_stop.all()):
                    break
            if _timing is not None:
                _timing["end"].record()
                _timing["steps"] = n_steps
            qs = qs_out[:, :, :n_steps]
            atts = atts_out[:, :, :n_steps] if atts_out is not None else None
            is_stop = (qs == stop_token).prod(dim=0)                # [b_global, steps]
            stop_tokens = torch.cat([is_stop, torch.ones(world_b, 1, device=device, dtype=is_stop.dtype)], dim=1).float()
        else:
            for t in range(t_start, max_seqlen):
                q_sampled, att, emb = one_step(y_embd, t)
                atts.append(att)
                if dist_group is not None:
                    # the one exchange of the data path: [q, b_local, 1] int64 ids -> [q, b_global, 1]
                    q_all = gather_tokens(q_sampled, dist_group)
                else:
                    q_all = q_sampled
                qs.append(q_all)
                all_stop.logical_or_((q_all == stop_token).prod(dim=0).bool())
                if not force_max_seqlen and (t + 1) % stop_check_interval == 0 and bool(all_stop.all()):
                    break
                y_embd = prompt[:, [t]] if (exists(prompt) and t < p_len) else emb
            if _timing is not None:
                _timing["end"].record()
                _timing["steps"] = len(qs)
            atts = torch.cat(atts, dim=2) if exists(atts[0]) else None
            qs = torch.stack(qs, dim=2).squeeze(-1)
            is_stop = (qs == stop_token).prod(dim=0)
            stop_tokens = torch.cat([is_stop, torch.ones(world_b, 1, device=device, dtype=is_stop.dtype)], dim=1).float()
        bg = world_b
        n = stop_tokens.shape[1]
        rvq = (undelay_rvq(qs) - self.n_special_token_in).clamp_min(0)
        stop_idx = (stop_tokens * torch.arange(n, device=device).unsqueeze(0)).long()
        cuts = []
        for i in range(bg):
            idx = torch.unique(stop_idx[i])[1]
            a = atts[i, :, :idx] if (atts is not None and i < atts.shape[0] and world == 1) else None
            cuts.append((rvq[:, [i], :idx - self.n_quant], a))
        self._new_text()                                        # release the memo's hold on this utterance's text
        return qs, atts, stop_tokens, cuts
