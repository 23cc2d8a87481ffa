from .contracts import AttentiveRNN
from .gla import GatedLinearAttention, AttentiveGLA
from .modeling_lina import LinaModel
from .encoder import TextEncoder, SimpleSpeakerEncoder
from .base_blocks import MixingBlock, SwiGLU, SelfAttention
from .crossatt import BlindCrossAttention, CrossAttention
from .positional import ConvPos, SinPos
from .embeddings import MultiEmbedding
from . import tools

__all__ = ["AttentiveRNN", "GatedLinearAttention", "AttentiveGLA", "LinaModel", "TextEncoder", "SimpleSpeakerEncoder", "MixingBlock", "SwiGLU",
           "SelfAttention", "BlindCrossAttention", "CrossAttention", "ConvPos", "SinPos", "MultiEmbedding", "tools"]
