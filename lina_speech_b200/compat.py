"""Module aliases that let the reference's notebook / scripts run unchanged on this package.

``install_reference_aliases()`` registers this package's mirrors under the reference's import names
(``model``, ``model.gla`` ..., ``train_lina``, ``initial_state``, ``decoder.pretrained``), so that

    from train_lina import TrainLina
    from decoder.pretrained import WavTokenizer
    from initial_state import train_initial_state, filter_unk
    TrainLina.load_from_checkpoint(".../last.ckpt")          # un-pickles model.gla.AttentiveGLA, model.encoder.TextEncoder ...

(InferenceLina.ipynb cells 1-3) resolve to the B200 implementations.  Nothing is installed unless this is called."""
from __future__ import annotations

import importlib
import sys
import types


class MulticlassAccuracy:
    """model/accuracy.py:11-33: top-k accuracy over rows whose target is not in ``ignore_index``."""

    def __init__(self, num_classes: int, top_k: int = 1, ignore_index=None):
        self.num_classes, self.top_k, self.ignore_index = num_classes, top_k, ignore_index

    def __call__(self, preds, targets):
        import torch
        if self.ignore_index is not None:
            keep = ~torch.isin(targets, torch.tensor(self.ignore_index).to(targets))
            preds, targets = preds[keep], targets[keep]
        hit = (preds.topk(self.top_k, dim=-1).indices == targets.unsqueeze(1)).any(dim=1)
        return hit.sum() / len(hit)


def _module(name: str, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    return mod


def install_reference_aliases(force: bool = False) -> dict:
    """Register the aliases in ``sys.modules``; existing entries are kept unless ``force``.  Returns what was installed."""
    from . import codec, initial_state, model, train_lina
    from .model import contracts, embeddings
    table = {"model": model, "train_lina": train_lina, "initial_state": initial_state}
    for sub in ("gla", "base_blocks", "crossatt", "modeling_lina", "tools", "encoder"):
        table[f"model.{sub}"] = importlib.import_module(f"{model.__name__}.{sub}")
    table["model.attentive_rnn"] = _module("model.attentive_rnn", AttentiveRNN=contracts.AttentiveRNN)
    table["model.multiembed"] = _module("model.multiembed", MultiEmbedding=embeddings.MultiEmbedding)
    table["model.accuracy"] = _module("model.accuracy", MulticlassAccuracy=MulticlassAccuracy,
                                      exists=lambda x: x is not None)
    pretrained = _module("decoder.pretrained", WavTokenizer=codec.WavTokenizer)
    table["decoder"] = _module("decoder", pretrained=pretrained, __path__=[])
    table["decoder.pretrained"] = pretrained
    # the sub-modules pickled inside a reference checkpoint name fla's classes by module path
    # (fla.modules.convolution.ShortConvolution, fla.modules.fused_norm_gate.FusedRMSNormSwishGate, fla.models.utils.Cache)
    from . import fla_api
    from .fla_api import modules as fm, ops as fo
    gla_ops = _module("fla.ops.gla", fused_recurrent_gla=fo.fused_recurrent_gla, fused_chunk_gla=fo.fused_chunk_gla,
                      chunk_gla=fo.chunk_gla, __path__=[])
    rwkv_ops = _module("fla.ops.rwkv6", fused_recurrent_rwkv6=fo.fused_recurrent_rwkv6, chunk_rwkv6=fo.chunk_rwkv6, __path__=[])
    fla_ops = _module("fla.ops", gla=gla_ops, rwkv6=rwkv_ops, __path__=[])
    conv = _module("fla.modules.convolution", ShortConvolution=fm.ShortConvolution)
    fng = _module("fla.modules.fused_norm_gate", FusedRMSNormSwishGate=fm.FusedRMSNormSwishGate)
    fmods = _module("fla.modules", ShortConvolution=fm.ShortConvolution, FusedRMSNormSwishGate=fm.FusedRMSNormSwishGate,
                    convolution=conv, fused_norm_gate=fng, __path__=[])
    futils = _module("fla.models.utils", Cache=fm.Cache)
    fmodels = _module("fla.models", utils=futils, __path__=[])
    table.update({"fla": _module("fla", ops=fla_ops, modules=fmods, models=fmodels, __path__=[], __version__="lina_speech_b200"),
                  "fla.ops": fla_ops, "fla.ops.gla": gla_ops, "fla.ops.rwkv6": rwkv_ops, "fla.modules": fmods,
                  "fla.modules.convolution": conv, "fla.modules.fused_norm_gate": fng, "fla.models": fmodels,
                  "fla.models.utils": futils})
    done = {}
    for name, mod in table.items():
        if force or name not in sys.modules:
            sys.modules[name] = mod
            done[name] = mod
    return done
