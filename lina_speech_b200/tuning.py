"""Initial-state tuning (voice prompting by learning the recurrent state) -- mirror of the reference's
``initial_state.py``: ``train_initial_state`` (:85-160), ``simple_collate`` (:51-82), ``speaker_state_dict`` /
``parse_speaker_state`` (:20-48; the reference's versions miss the ``safe_open`` / ``deepcopy`` imports).

Only the 2N rank-r state factors (k [1,r,H,K,1], v [1,r,H,1,V] per encoder / decoder block) are trained;
the model runs teacher-forced in ``fused_recurrent`` mode so that the GLA op returns d(loss)/d(initial state)
(``lina_gla_recurrent_bwd`` -> dh0) -- the one path of the shipped model that needs dh0.
"""
from __future__ import annotations

import random
from copy import deepcopy
from functools import reduce
from typing import Dict, List, Sequence, Tuple

import torch
from torch.nn.utils.rnn import pad_sequence

from .model.tools import delay_rvq, sequence_mask


def filter_unk(x, tokenizer) -> bool:
    """initial_state.py:13-18 (the notebook's dataset filter, InferenceLina.ipynb cell "expresso_ds.filter"): True when the
    tokenizer can encode the transcript."""
    try:
        tokenizer.encode(x)
        return True
    except Exception:
        return False


def filter_except(x, tokenizer=None) -> bool:
    """initial_state.py:32-37 reads a module-level ``tokenizer`` that the reference never defines, so it always answers
    False there; here the tokenizer can be passed, and without one the answer is the reference's."""
    return False if tokenizer is None else filter_unk(x, tokenizer)


def speaker_state_dict(params) -> Dict[str, torch.Tensor]:
    """initial_state.py:20-30 -- flat dict ready for safetensors.save_file."""
    out = {}
    for i, layer in enumerate(params):
        if len(layer) == 2:
            out[f"layer{i}_k"], out[f"layer{i}_v"] = layer[0], layer[1]
        else:
            out[f"layer{i}"] = layer
    return out


def parse_speaker_state(path, device="cpu") -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """initial_state.py:39-48."""
    from safetensors import safe_open
    with safe_open(path, framework="pt", device=device) as state:
        keys = sorted((k for k in state.keys() if k.endswith("_k")),
                      key=lambda x: int("".join(ch for ch in x if ch.isdigit())))
        return [(state.get_tensor(k), state.get_tensor(k[:-2] + "_v")) for k in keys]


def simple_collate(batch: Sequence[dict], tokenizer) -> dict:
    """initial_state.py:51-82: codes -> ids (+3), delay pattern with start=1 / stop=2, text ids, masks."""
    audio_token, text = zip(*[(x["audio_token"], x["text"]) for x in batch])
    delayed = []
    for x in audio_token:
        x = torch.as_tensor(x).squeeze()
        if x.dim() == 1:
            x = x.unsqueeze(0)
        delayed.append(delay_rvq(x + 3, head_token=1, tail_token=2).transpose(-1, -2))
    text_token = [torch.LongTensor(tokenizer.encode("[BOS]" + t + "[EOS]")) for t in text]
    xlen, ylen = [t.shape[0] for t in text_token], [t.shape[0] for t in delayed]
    x_mask = sequence_mask(torch.tensor(xlen), device="cpu")
    y_mask = sequence_mask(torch.tensor(ylen), device="cpu")
    audio = pad_sequence(delayed, batch_first=True, padding_value=0)
    textp = pad_sequence(text_token, batch_first=True, padding_value=0)
    encoder_mask = x_mask.unsqueeze(1) * x_mask.unsqueeze(2)
    crossatt_mask = x_mask.unsqueeze(1) * y_mask.unsqueeze(2)
    crossatt_mask[:, :, 0] = True
    return {"text_token": textp, "audio_token": audio, "orig_token": audio_token, "crossatt_mask": crossatt_mask,
            "encoder_mask": encoder_mask, "text": text, "y_mask": y_mask, "x_len": xlen, "y_len": ylen}


class _EndlessShuffle:
    """Indices 0..n-1 in shuffled epochs forever (initial_state.py:106-112: ``inf_sampler_wo_replacement``)."""

    def __init__(self, n: int, seed: int):
        self.n, self.rng, self.pending = n, random.Random(seed), []

    def __next__(self) -> int:
        if not self.pending:
            self.pending = list(range(self.n))
            self.rng.shuffle(self.pending)
            self.pending.reverse()
        return self.pending.pop()


class StateTuner:
    """Adam on the rank-r initial-state factors of every encoder / decoder block; model weights stay frozen."""

    def __init__(self, model, rank: int = 1, scale: float = 0.02, lr: float = 0.1):
        self.model, self.scale = model, scale
        self.device = next(model.parameters()).device
        model.attentive_rnn.to_mode("fused_recurrent")          # the mode whose backward returns dh0
        model.train()
        # Only the state factors are optimised (initial_state.py:99-104 hands Adam nothing else); the reference leaves
        # requires_grad on the model weights, so its backward also computes -- and never reads -- every weight gradient.
        # Here the weights are frozen for the duration of the tuning (restored by release()): same losses, same factors,
        # no dW GEMMs.
        self._frozen = [p for p in model.parameters() if p.requires_grad]
        for p in self._frozen:
            p.requires_grad_(False)
        self.factors = model.attentive_rnn.get_init_state_tuning_params(lora=rank, device=self.device)
        self.opt = torch.optim.Adam(reduce(tuple.__add__, self.factors), lr=lr)

    def release(self) -> None:
        for p in self._frozen:
            p.requires_grad_(True)
        self._frozen = []

    def loss_on(self, batch: dict, batch_size: int) -> torch.Tensor:
        on_dev = {name: (val.to(self.device) if torch.is_tensor(val) else val) for name, val in batch.items()}
        state = self.model.attentive_rnn.get_state_from_params(self.factors, batch_size, scale=self.scale)
        out = self.model(on_dev["text_token"], on_dev["audio_token"], on_dev["encoder_mask"], on_dev["crossatt_mask"],
                         logits_mask=on_dev["y_mask"], init_state=state)
        return out[1]


def train_initial_state(model, dataset, tokenizer, n_samples: int, lr: float = 0.1, grad_acc: int = 4,
                        batch_size: int = 2, scale: float = 0.02, save_every_k_steps: int = 0, seed: int = 123,
                        rank: int = 1, progress: bool = False):
    """initial_state.py:85-160, same arguments and return value: (parameters, train_losses); the model is left in
    eval mode.  ``dataset`` is any indexable of {"audio_token": [Q,T] codes, "text": str}."""
    tuner = StateTuner(model, rank=rank, scale=scale, lr=lr)
    picks = _EndlessShuffle(len(dataset), seed)
    history, checkpoints, optimizer_steps = [], [], 0
    rounds = range(n_samples // batch_size)
    if progress:
        from tqdm import tqdm
        rounds = tqdm(rounds)
    try:
        for it in rounds:
            examples = [dataset[next(picks)] for _ in range(batch_size)]
            loss = tuner.loss_on(simple_collate(examples, tokenizer), batch_size)
            history.append(loss.item())
            loss.backward()
            if (it + 1) % grad_acc == 0:
                tuner.opt.step()
                tuner.opt.zero_grad()
                optimizer_steps += 1
                if save_every_k_steps > 0 and optimizer_steps % save_every_k_steps == 0:
                    checkpoints.append(deepcopy(tuner.factors))
    finally:
        tuner.release()
    model.eval()
    if save_every_k_steps > 0:
        return checkpoints + [tuner.factors], history
    return tuner.factors, history
