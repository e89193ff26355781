"""Runtime model configuration and the flat weight-file layout.

The reference fixes every dimension at compile time (config.mojo:4-17, WhisperConfig.tiny()
whisper.mojo:29-31) and hard-codes the rest as literals (whisper.mojo:61-69,123-128).  Here the
same numbers are one runtime struct so the Small-shaped config and a micro test config run through
the same kernels.  The field order of `WhisperConfig.as_c_array()` is the `wm_config` struct of
include/whisper_b200.h.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
from typing import List, Tuple

# Prompt / stop ids: whisper.mojo:187-191,206
SOT, LANG_EN, TRANSCRIBE, NO_TIMESTAMPS, EOT = 50258, 50259, 50359, 50363, 50257
MAX_DECODE_ITERS = 195  # whisper.mojo:205
SAMPLE_RATE = 16000
N_FFT = 400
HOP = 160


@dataclass(frozen=True)
class WhisperConfig:
    d_model: int = 384
    n_heads: int = 6
    n_layers: int = 4
    vocab_size: int = 51865
    n_audio_ctx: int = 1500  # MAX_SEQ_LEN, config.mojo:10
    n_text_ctx: int = 448  # MAX_TOKENS, config.mojo:11; KVCache max_len whisper.mojo:193
    n_mels: int = 80
    prompt: Tuple[int, int, int, int] = (SOT, LANG_EN, TRANSCRIBE, NO_TIMESTAMPS)
    eot: int = EOT
    max_iters: int = MAX_DECODE_ITERS
    pos_quirk: int = 1  # 1 = reference's start_pos = current_len - 1 (whisper.mojo:217); 0 = HF positions

    @staticmethod
    def tiny() -> "WhisperConfig":
        """whisper.mojo:29-31 + config.mojo."""
        return WhisperConfig()

    @staticmethod
    def small_shaped() -> "WhisperConfig":
        """BASELINE.json configs[4]: 12 layers, d=768, 12 heads (head_dim stays 64)."""
        return WhisperConfig(d_model=768, n_heads=12, n_layers=12)

    @staticmethod
    def micro() -> "WhisperConfig":
        """Small shapes for fast CPU/GPU unit tests; not a reference config."""
        return WhisperConfig(d_model=128, n_heads=2, n_layers=2, vocab_size=1000, n_audio_ctx=96,
                             n_text_ctx=64, prompt=(990, 991, 992, 993), eot=989, max_iters=20)

    @property
    def head_dim(self) -> int:
        return self.d_model // self.n_heads

    @property
    def n_frames(self) -> int:
        return 2 * self.n_audio_ctx

    @property
    def n_samples(self) -> int:
        return self.n_frames * HOP

    @property
    def max_tokens(self) -> int:
        """4 prompt ids + 1 + max_iters generated (whisper.mojo:200-221)."""
        return 5 + self.max_iters

    def weight_layout(self) -> List[Tuple[str, Tuple[int, ...]]]:
        """Tensors of the flat fp32 weight file, in file order (export_weights.py:19-90; read order
        whisper.mojo:60-69,122-128 and layers.mojo:96-103,418-433).  Linear weights are [out, in]."""
        D, F = self.d_model, 4 * self.d_model
        out: List[Tuple[str, Tuple[int, ...]]] = []

        def attn(p):
            out.extend([(p + "q.w", (D, D)), (p + "q.b", (D,)), (p + "k.w", (D, D)), (p + "v.w", (D, D)),
                        (p + "v.b", (D,)), (p + "o.w", (D, D)), (p + "o.b", (D,))])

        def mlp(p):
            out.extend([(p + "fc1.w", (F, D)), (p + "fc1.b", (F,)), (p + "fc2.w", (D, F)), (p + "fc2.b", (D,)),
                        (p + "mlp_ln.w", (D,)), (p + "mlp_ln.b", (D,))])

        out.extend([("enc.conv1.w", (D, self.n_mels, 3)), ("enc.conv1.b", (D,)),
                    ("enc.conv2.w", (D, D, 3)), ("enc.conv2.b", (D,)), ("enc.pos", (self.n_audio_ctx, D))])
        for i in range(self.n_layers):
            p = f"enc.{i}."
            attn(p + "attn.")
            out.extend([(p + "attn_ln.w", (D,)), (p + "attn_ln.b", (D,))])
            mlp(p)
        out.extend([("enc.ln_post.w", (D,)), ("enc.ln_post.b", (D,))])
        out.extend([("dec.token_emb", (self.vocab_size, D)), ("dec.pos", (self.n_text_ctx, D))])
        for i in range(self.n_layers):
            p = f"dec.{i}."
            attn(p + "attn.")
            out.extend([(p + "attn_ln.w", (D,)), (p + "attn_ln.b", (D,))])
            attn(p + "cross.")
            out.extend([(p + "cross_ln.w", (D,)), (p + "cross_ln.b", (D,))])
            mlp(p)
        out.extend([("dec.ln_post.w", (D,)), ("dec.ln_post.b", (D,))])
        return out

    def weight_offsets(self):
        """name -> (offset_in_floats, shape)."""
        off, table = 0, {}
        for name, shape in self.weight_layout():
            n = 1
            for s in shape:
                n *= s
            table[name] = (off, shape)
            off += n
        return table

    def weight_count(self) -> int:
        n = 0
        for _, shape in self.weight_layout():
            k = 1
            for s in shape:
                k *= s
            n += k
        return n

    def as_c_array(self):
        """int32[16] image of `wm_config` (include/whisper_b200.h)."""
        vals = [self.d_model, self.n_heads, self.n_layers, self.vocab_size, self.n_audio_ctx, self.n_text_ctx,
                self.n_mels, self.prompt[0], self.prompt[1], self.prompt[2], self.prompt[3], self.eot,
                self.max_iters, self.pos_quirk, 0, 0]
        return (ctypes.c_int32 * 16)(*vals)
