"""TextEncoder (model/encoder.py:14-43): bidirectional self-attention blocks, once per utterance."""
import torch

from .base_blocks import MixingBlock, SelfAttention, SwiGLU


class TextEncoder(torch.nn.Module):
    def __init__(self, dim: int, heads: int, n_layers=4, dropout=0.1, rotary=True):
        super().__init__()
        self.sa = torch.nn.ModuleList([
            MixingBlock(lambda: SelfAttention(dim, heads, rotary=rotary), lambda: SwiGLU(dim),
                        lambda: torch.nn.LayerNorm(dim), dropout) for _ in range(n_layers)])

    def forward(self, x, mask=None, pos=None):
        if mask is not None:
            eye = torch.eye(mask.shape[-1], device=x.device, dtype=torch.bool)
            mask = mask.unsqueeze(1).bool() | eye.view(1, 1, *eye.shape)
        for block in self.sa:
            x = block(x, mask=mask, pos=pos)
        return x


from .speaker import SimpleSpeakerEncoder  # noqa: E402,F401  (model/encoder.py:45-84 lives in speaker.py here)
