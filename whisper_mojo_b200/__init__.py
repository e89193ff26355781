"""whisper_mojo_b200 -- B200-native Whisper transcription path behind the API of antonvice/whisper.Mojo.

(The task names the package `whisper.mojo_b200`; a dot is not importable, hence the underscore.)

    from whisper_mojo_b200 import Whisper, WeightLoader, Tensor
    w = Whisper(); w.load(WeightLoader("whisper_tiny_weights.bin")); tokens = w.transcribe(mel)

Importing the package does not need a GPU; the first compute call does (no CPU fallback).
"""
from .config import WhisperConfig  # noqa: F401
from .loader import WeightLoader  # noqa: F401
from .tokenizer import Tokenizer  # noqa: F401
from .whisper_tensor import Tensor  # noqa: F401
from .layers import KVCache, LayerCache, MultiHeadAttention, ResidualAttentionBlock  # noqa: F401
from .whisper import DeviceKVCache, Whisper, WhisperDecoder, WhisperEncoder  # noqa: F401
from . import audio, export  # noqa: F401,E402  (host-side ingest and checkpoint export)
