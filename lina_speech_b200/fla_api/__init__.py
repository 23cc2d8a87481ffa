from .ops import fused_recurrent_gla, fused_chunk_gla, chunk_gla
from .modules import ShortConvolution, FusedRMSNormSwishGate, Cache

__all__ = ["fused_recurrent_gla", "fused_chunk_gla", "chunk_gla", "ShortConvolution",
           "FusedRMSNormSwishGate", "Cache"]
