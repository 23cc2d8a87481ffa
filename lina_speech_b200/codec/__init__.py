from .wavtokenizer import WavTokenizer, VocosBackbone, ISTFTHead, CodebookFeatures

__all__ = ["WavTokenizer", "VocosBackbone", "ISTFTHead", "CodebookFeatures"]
