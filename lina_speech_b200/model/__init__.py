from .gla import GatedLinearAttention, AttentiveGLA
from .modeling_lina import LinaModel
from .encoder import TextEncoder
from .base_blocks import MixingBlock, SwiGLU, SelfAttention
from .crossatt import BlindCrossAttention, CrossAttention, ConvPos, SinPos
from .multiembed import MultiEmbedding
from . import tools

__all__ = ["GatedLinearAttention", "AttentiveGLA", "LinaModel", "TextEncoder", "MixingBlock", "SwiGLU",
           "SelfAttention", "BlindCrossAttention", "CrossAttention", "ConvPos", "SinPos", "MultiEmbedding", "tools"]
