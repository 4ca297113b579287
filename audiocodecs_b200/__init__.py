"""B200-native (sm_100a) tokenize/detokenize path behind the `audiocodecs` Codec API.

Drop-in for `audiocodecs.{Encodec,DAC,Mimi}` `sig_to_toks` / `toks_to_sig` (R/audiocodecs/codec.py:57-66,90-100).
"""
from .codec import Codec
from .dac import DAC
from .encodec import Encodec
from .mimi import Mimi
from . import shard  # noqa: F401  (clip sharding across GPUs)
from .graphs import GraphedCodec
from .consumers import CodebookUtil, MultiHeadEmbedding

__version__ = "0.1.0"
__all__ = ["Codec", "Encodec", "DAC", "Mimi", "GraphedCodec", "CodebookUtil", "MultiHeadEmbedding"]
