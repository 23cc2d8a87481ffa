"""Host mirror of the reference's ``train_lina.py`` (TrainLina, :12-120) without the Lightning dependency: same constructor
arguments, ``.model`` (the LinaModel), ``step`` / ``training_step`` / ``validation_step`` / ``configure_optimizers`` and
``load_from_checkpoint`` -- what InferenceLina.ipynb and ``train_initial_state`` touch.  A Lightning Trainer can still drive it
through a thin ``LightningModule`` subclass; a plain loop is ``loss = m.training_step(batch, i); loss.backward(); opt.step()``.

The arithmetic of a step is ``LinaModel.forward`` (model/modeling_lina.py): with bf16 autocast the GLA blocks run forward AND
backward on the tcgen05 kernel."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
from torch import nn

from .model.modeling_lina import LinaModel


class TrainLina(nn.Module):
    def __init__(self, attentive_rnn: nn.Module, d_model: int, quant_layer: List[int], n_codebook: int,
                 n_special_token_in: int, n_special_token_out: int, n_txt_vocab: int, tie_embed: bool = False,
                 txt_encoder: Optional[nn.Module] = None, spk_encoder: Optional[nn.Module] = None,
                 learning_rate: float = 5e-4, weight_decay: float = 0.1, betas: Tuple[float, float] = (0.9, 0.999),
                 n_warmup_steps: int = 500, n_training_steps: int = 300000, mask_text_p: float = 0.,
                 load_weights: Optional[str] = None):
        super().__init__()
        self.learning_rate, self.weight_decay, self.betas = learning_rate, weight_decay, tuple(betas)
        self.n_warmup_steps, self.n_training_steps = n_warmup_steps, n_training_steps
        self.model = LinaModel(attentive_rnn, d_model, len(quant_layer), n_codebook, n_special_token_in,
                               n_special_token_out, n_txt_vocab, tie_embed=tie_embed, txt_encoder=txt_encoder,
                               spk_encoder=spk_encoder, mask_text_p=mask_text_p)
        if load_weights is not None:                                       # train_lina.py:62-64
            self.load_state_dict(torch.load(load_weights, map_location="cpu", weights_only=False)["state_dict"])

    def step(self, batch):
        """train_lina.py:72-86."""
        logits, loss, att, masked_logits, masked_target = self.model(
            batch["text_token"], batch["audio_token"], batch["encoder_mask"], batch["crossatt_mask"],
            logits_mask=batch["y_mask"], crossatt_pos=batch["crossatt_pos"])
        return logits, loss, att, []

    def training_step(self, batch, idx=0):
        return self.step(batch)[1]

    def validation_step(self, batch, idx=0):
        return self.step(batch)[1]

    def configure_optimizers(self):
        """train_lina.py:105-120: AdamW over all model parameters + cosine schedule with warm-up, stepped every step."""
        from transformers import get_cosine_schedule_with_warmup
        opt = torch.optim.AdamW([{"params": self.model.parameters(), "weight_decay": self.weight_decay}],
                                lr=self.learning_rate, betas=self.betas)
        sched = get_cosine_schedule_with_warmup(opt, num_warmup_steps=self.n_warmup_steps,
                                                num_training_steps=self.n_training_steps)
        return [opt], [{"scheduler": sched, "interval": "step"}]

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path: str, map_location="cpu", strict: bool = True, **overrides) -> "TrainLina":
        """Reads a Lightning checkpoint written by the reference: ``hyper_parameters`` (the constructor arguments saved by
        ``save_hyperparameters()``, train_lina.py:55) + ``state_dict`` (keys ``model.*``).  The pickled sub-modules name the
        reference's classes (``model.gla.AttentiveGLA`` ...): alias the package first, e.g.
        ``sys.modules['model'] = lina_speech_b200.model``, or pass the modules as keyword overrides."""
        ckpt = torch.load(checkpoint_path, map_location=map_location, weights_only=False)
        hp = dict(ckpt.get("hyper_parameters", {}))
        hp.update(overrides)
        hp.pop("load_weights", None)
        self = cls(**hp)
        self.load_state_dict(ckpt["state_dict"], strict=strict)
        return self
