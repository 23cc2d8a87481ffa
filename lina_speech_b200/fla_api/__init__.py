from .ops import fused_recurrent_gla, fused_chunk_gla, chunk_gla, fused_recurrent_rwkv6, chunk_rwkv6
from .modules import ShortConvolution, FusedRMSNormSwishGate, Cache

__all__ = ["fused_recurrent_gla", "fused_chunk_gla", "chunk_gla", "fused_recurrent_rwkv6", "chunk_rwkv6", "ShortConvolution",
           "FusedRMSNormSwishGate", "Cache"]
