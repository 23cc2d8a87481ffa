"""lina_speech_b200 -- B200 (sm_100a) implementation of Lina-Speech's GLA hot path and
WavTokenizer decode, behind the reference's own operator API.  See DESIGN.md."""
__version__ = "0.1.0"

from . import _lib  # noqa: F401
