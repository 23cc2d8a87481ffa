"""Host-side glue kept as in the reference (model/tools.py): sampling, RVQ delay pattern, masks."""
from itertools import accumulate

import torch


def pad_2d_sequence(seq, padding_value=0):
    """model/tools.py:8-15: right/bottom-pad 2-D tensors to the largest height and width and stack them."""
    rows, cols = max(t.shape[0] for t in seq), max(t.shape[1] for t in seq)
    return torch.stack([torch.nn.functional.pad(t, (0, cols - t.shape[1], 0, rows - t.shape[0]), value=padding_value)
                        for t in seq])


def topk_sampling(seq, k=1, temp=1.0):
    """model/tools.py:38-44, including its quirk: the k-th largest *unscaled* logit is the
    threshold applied to the temperature-scaled logits."""
    kth = torch.topk(seq, k, dim=-1).values[:, -1:]
    logits = seq / temp
    logits = logits.masked_fill(logits < kth, -float("inf"))
    return torch.multinomial(torch.softmax(logits, dim=-1), num_samples=1)


def delay_rvq(code, head_token: int = -2, tail_token: int = -3):
    """model/tools.py:46-59: quantizer i is shifted right by i+1, padded with head / tail tokens."""
    q, _ = code.shape
    ext = torch.ones((q, q + 1)).tril() * head_token
    ext += torch.ones((q + 1, q)).tril(diagonal=-1).T * tail_token
    ext = torch.flip(ext, (1,))
    out = torch.cat((code, ext.to(code.device)), dim=1)
    for i in range(q):
        out[i, :] = torch.roll(out[i, :], i + 1)
    return out.long()


def undelay_rvq(extended_code):
    """model/tools.py:61-67."""
    q, _, n = extended_code.shape
    out = torch.stack([torch.roll(extended_code[i], -(i + 1), dims=1) for i in range(q)], dim=0)
    return out[:, :, :-(q + 1)]


def sequence_mask(lengths, max_len=None, device=None):
    """model/tools.py:69-77."""
    if max_len is None:
        max_len = int(torch.max(lengths).item())
    ids = torch.arange(0, max_len, device=device if device is not None else lengths.device)
    return ids.unsqueeze(0) < lengths.to(ids.device).unsqueeze(1)


def packmask_2d(xlen, ylen, offset: int = 0) -> torch.Tensor:
    """model/tools.py:17-35: block mask for packed sequences."""
    ybound = [0] + list(accumulate(ylen))
    lb, hb = [], []
    for n, lo, hi in zip(xlen, ybound[:-1], ybound[1:]):
        lb += [lo] * n
        hb += [hi] * n
    lb, hb = torch.tensor(lb), torch.tensor(hb)
    if offset:
        lb -= offset
        hb += offset
    rge = torch.arange(ybound[-1])
    return (rge.unsqueeze(0) >= lb.unsqueeze(1)) * (rge.unsqueeze(0) < hb.unsqueeze(1))
