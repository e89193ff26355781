"""Generate tests/golden/*.npz from independent Python references available in THIS container.

TEST INFRASTRUCTURE ONLY.  Run from the repo root:  python -m oracle.make_golden

What is generated, and from what:
  logmel_hf.npz   HF transformers WhisperFeatureExtractor (the package the reference itself calls,
                  export_weights.py:116) on seeded synthetic audio -> pins oracle/logmel_oracle.py
                  and the CUDA frontend.
  hf_micro.npz    HF WhisperForConditionalGeneration (tanh GELU, reference prompt/positions, no
  hf_tiny.npz     logits processors; oracle/hf_crosscheck.py) loaded with seeded weights in the
                  reference's file order -> pins oracle/whisper_oracle.c (encoder output,
                  teacher-forced logits, greedy tokens).
  reference_expected_tokens.json   the 89 ids of /root/reference/expected_tokens.txt parsed to
                  JSON (the reference's only golden vector; usable only when the real
                  whisper_tiny_weights.bin + sample_input.bin are supplied).

Inputs are NOT stored: they are regenerated from seeds by whisper_mojo_b200/synth.py (numpy PCG64,
bit-stable across machines for a given numpy), which keeps the fixtures small.
"""
from __future__ import annotations

import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from whisper_mojo_b200 import synth  # noqa: E402
from whisper_mojo_b200.config import WhisperConfig  # noqa: E402
from oracle import hf_crosscheck as H  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# Parameters shared with the tests (tests/golden_params.py re-reads them from the npz files).
LOGMEL_SEED, LOGMEL_CHUNKS, LOGMEL_FRAME_STRIDE = 0, 2, 8
WEIGHT_SEED, MEL_SEED = 0, 0
TINY_ROW_STRIDE, TINY_FORCED, TINY_GREEDY_ITERS, TINY_LOGIT_STRIDE = 50, 12, 24, 97


def gen_logmel():
    from transformers import WhisperFeatureExtractor

    fe = WhisperFeatureExtractor()
    audio = synth.make_audio(LOGMEL_CHUNKS, seed=LOGMEL_SEED)
    mel = fe(list(audio), sampling_rate=16000, return_tensors="np").input_features.astype(np.float32)
    short = np.random.default_rng(3).standard_normal(16000 * 5).astype(np.float32)  # export_weights.py:147
    mel_short = fe(short, sampling_rate=16000, return_tensors="np").input_features[0].astype(np.float32)
    np.savez_compressed(
        os.path.join(GOLD, "logmel_hf.npz"),
        seed=LOGMEL_SEED, n_chunks=LOGMEL_CHUNKS, frame_stride=LOGMEL_FRAME_STRIDE,
        mel_sub=mel[:, :, ::LOGMEL_FRAME_STRIDE].astype(np.float16),  # f16 storage: abs err <= 1e-3; exact stats below
        mel_first_frames=mel[:, :, :64], mel_last_frames=mel[:, :, -64:],
        mel_row_sums=mel.astype(np.float64).sum(axis=2), mel_max=mel.max(axis=(1, 2)), mel_min=mel.min(axis=(1, 2)),
        short_first_frames=mel_short[:, :64], short_row_sums=mel_short.astype(np.float64).sum(axis=1),
        mel_filters=np.asarray(fe.mel_filters, np.float32),
        transformers_version=np.array(__import__("transformers").__version__),
    )


def gen_model(cfg: WhisperConfig, name: str, row_stride: int, n_forced: int, greedy_iters: int, logit_stride: int):
    w = synth.make_weights(cfg, seed=WEIGHT_SEED)
    mel = synth.make_mel(1, cfg, MEL_SEED)[0]
    hf = H.build_hf(cfg, w)
    enc = H.hf_encode(hf, mel)
    forced = np.concatenate([np.array(cfg.prompt),
                             np.random.default_rng(1).integers(0, cfg.vocab_size, n_forced)]).astype(np.int32)
    out = {"weight_seed": WEIGHT_SEED, "mel_seed": MEL_SEED, "row_stride": row_stride,
           "logit_stride": logit_stride, "enc_rows": enc[::row_stride], "forced": forced}
    for q in (1, 0):
        lg = H.hf_teacher_forced(hf, cfg, enc, forced, q)
        out[f"tf_logits_q{q}"] = lg[:, ::logit_stride]
        out[f"tf_argmax_q{q}"] = lg.argmax(axis=1).astype(np.int32)
        out[f"tf_max_q{q}"] = lg.max(axis=1)
        out[f"greedy_q{q}"] = H.hf_greedy(hf, cfg, enc, q, greedy_iters)
    np.savez_compressed(os.path.join(GOLD, name), **out)


def gen_expected_tokens():
    src = "/root/reference/expected_tokens.txt"
    ids = [int(x) for x in re.findall(r"np\.int64\((\d+)\)", open(src).read())]
    assert len(ids) == 89
    with open(os.path.join(GOLD, "reference_expected_tokens.json"), "w") as f:
        json.dump({"source": "expected_tokens.txt:1 (HF model.generate on the reference's sample clip; "
                             "no 4-id prompt prefix, no EOT suffix)", "ids": ids}, f)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    gen_logmel()
    gen_model(WhisperConfig.micro(), "hf_micro.npz", 1, 12, 20, 1)
    gen_model(WhisperConfig.tiny(), "hf_tiny.npz", TINY_ROW_STRIDE, TINY_FORCED, TINY_GREEDY_ITERS, TINY_LOGIT_STRIDE)
    gen_expected_tokens()
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))
