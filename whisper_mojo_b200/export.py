"""Exporter for other checkpoints: HF `WhisperForConditionalGeneration` state_dict ->
the reference's headerless little-endian fp32 weight stream (export_weights.py:11-92) for any layer count /
width (tiny ... small ...), plus vocab.txt (export_weights.py:135-142).  Needs `transformers`/`torch` only when
called with a live model; `export_state_dict` works on any mapping of numpy-convertible tensors."""
from __future__ import annotations

from typing import Dict, List, Mapping

import numpy as np

from .config import WhisperConfig

__all__ = ["tensor_order", "export_state_dict", "write_weights", "config_from_hf", "write_vocab"]

_ATTN = ["q_proj.weight", "q_proj.bias", "k_proj.weight", "v_proj.weight", "v_proj.bias", "out_proj.weight",
         "out_proj.bias"]  # no k_proj.bias in the file (export_weights.py:28-34)


def tensor_order(n_layers: int) -> List[str]:
    """HF state_dict keys in file order (export_weights.py:19-90; read back by whisper.mojo:60-69,122-128
    and layers.mojo:96-103,418-433).  proj_out is tied to embed_tokens and not written."""
    keys = ["model.encoder.conv1.weight", "model.encoder.conv1.bias", "model.encoder.conv2.weight",
            "model.encoder.conv2.bias", "model.encoder.embed_positions.weight"]
    for i in range(n_layers):
        p = f"model.encoder.layers.{i}."
        keys += [p + "self_attn." + a for a in _ATTN]
        keys += [p + "self_attn_layer_norm.weight", p + "self_attn_layer_norm.bias"]
        keys += [p + "fc1.weight", p + "fc1.bias", p + "fc2.weight", p + "fc2.bias"]
        keys += [p + "final_layer_norm.weight", p + "final_layer_norm.bias"]
    keys += ["model.encoder.layer_norm.weight", "model.encoder.layer_norm.bias"]
    keys += ["model.decoder.embed_tokens.weight", "model.decoder.embed_positions.weight"]
    for i in range(n_layers):
        p = f"model.decoder.layers.{i}."
        keys += [p + "self_attn." + a for a in _ATTN]
        keys += [p + "self_attn_layer_norm.weight", p + "self_attn_layer_norm.bias"]
        keys += [p + "encoder_attn." + a for a in _ATTN]
        keys += [p + "encoder_attn_layer_norm.weight", p + "encoder_attn_layer_norm.bias"]
        keys += [p + "fc1.weight", p + "fc1.bias", p + "fc2.weight", p + "fc2.bias"]
        keys += [p + "final_layer_norm.weight", p + "final_layer_norm.bias"]
    keys += ["model.decoder.layer_norm.weight", "model.decoder.layer_norm.bias"]
    return keys


def _np(t) -> np.ndarray:
    if hasattr(t, "detach"):
        t = t.detach().cpu().float().numpy()
    return np.ascontiguousarray(t, np.float32)


def export_state_dict(state_dict: Mapping[str, object], n_layers: int) -> np.ndarray:
    """Flat fp32 array in the reference's file order."""
    parts = []
    for k in tensor_order(n_layers):
        if k not in state_dict:
            raise KeyError(f"state_dict has no {k}")
        parts.append(_np(state_dict[k]).reshape(-1))
    return np.concatenate(parts)


def write_weights(path: str, state_dict: Mapping[str, object], n_layers: int) -> int:
    flat = export_state_dict(state_dict, n_layers)
    flat.astype("<f4").tofile(path)
    return int(flat.size)


def config_from_hf(hf_config, **overrides) -> WhisperConfig:
    """WhisperConfig for an HF WhisperConfig (encoder and decoder must have the same depth / width, as in every
    released Whisper size; the reference hard-codes tiny)."""
    if hf_config.encoder_layers != hf_config.decoder_layers or \
            hf_config.encoder_attention_heads != hf_config.decoder_attention_heads:
        raise ValueError("encoder / decoder shapes differ")
    base = WhisperConfig.tiny().__dict__
    base.update(d_model=hf_config.d_model, n_heads=hf_config.encoder_attention_heads, n_layers=hf_config.encoder_layers,
                vocab_size=hf_config.vocab_size, n_mels=hf_config.num_mel_bins,
                n_audio_ctx=hf_config.max_source_positions, n_text_ctx=hf_config.max_target_positions)
    base.update(overrides)
    return WhisperConfig(**base)


def write_vocab(path: str, vocab: Dict[str, int]) -> None:
    """export_weights.py:135-142: tokens sorted by id, one per line, newlines escaped."""
    with open(path, "w", encoding="utf-8") as f:
        for token, _ in sorted(vocab.items(), key=lambda kv: kv[1]):
            f.write(token.replace("\n", "\\n") + "\n")
