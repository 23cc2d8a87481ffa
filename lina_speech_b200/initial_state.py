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


def train_initial_state(model, dataset, tokenizer, n_samples: int, lr: float = 0.1, grad_acc: int = 4,
                        batch_size: int = 2, scale: float = 0.02, save_every_k_steps: int = 0, seed: int = 123,
                        rank: int = 1, progress: bool = False):
    """initial_state.py:85-160.  ``dataset`` is any indexable of {"audio_token": [Q,T] codes, "text": str}.
    Returns (parameters, train_losses); leaves the model in eval mode."""
    device = next(model.parameters()).device
    snapshots = []
    model.attentive_rnn.to_mode("fused_recurrent")
    model = model.train()
    parameters = model.attentive_rnn.get_init_state_tuning_params(lora=rank, device=device)
    optimizer = torch.optim.Adam(reduce(tuple.__add__, parameters), lr=lr)

    def sampler(length):
        rng = random.Random(seed)
        while True:
            idx = list(range(length))
            rng.shuffle(idx)
            yield from idx

    order = sampler(len(dataset))
    losses, k_steps = [], 0
    n_iter = n_samples // batch_size
    it = range(n_iter)
    if progress:
        from tqdm import tqdm
        it = tqdm(it)
    for i in it:
        batch = simple_collate([dataset[next(order)] for _ in range(batch_size)], tokenizer)
        batch = {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in batch.items()}
        init_state = model.attentive_rnn.get_state_from_params(parameters, batch_size, scale=scale)
        _, loss, _, _, _ = model(batch["text_token"], batch["audio_token"], batch["encoder_mask"], batch["crossatt_mask"],
                                 logits_mask=batch["y_mask"], init_state=init_state)
        losses.append(loss.item())
        loss.backward()
        if i % grad_acc == grad_acc - 1:
            optimizer.step()
            optimizer.zero_grad()
            k_steps += 1
            if save_every_k_steps > 0 and k_steps % save_every_k_steps == 0:
                snapshots.append(deepcopy(parameters))
    if save_every_k_steps > 0:
        snapshots.append(parameters)
        parameters = snapshots
    model.eval()
    return parameters, losses
