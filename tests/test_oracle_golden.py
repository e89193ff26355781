"""CPU: pin the oracle against the committed golden fixtures (tests/golden, made by
oracle/make_golden.py from HF transformers -- the package the reference exports from / claims to match)."""
import json
import os

import numpy as np
import pytest

from oracle import logmel_oracle as LM
from oracle import oracle as O
from whisper_mojo_b200 import synth
from whisper_mojo_b200.config import WhisperConfig

from conftest import GOLDEN


def test_mel_filters_match_hf():
    g = np.load(os.path.join(GOLDEN, "logmel_hf.npz"))
    assert np.array_equal(LM.mel_filters(), g["mel_filters"].T)


def test_logmel_oracle_matches_hf_golden():
    g = np.load(os.path.join(GOLDEN, "logmel_hf.npz"))
    audio = synth.make_audio(int(g["n_chunks"]), seed=int(g["seed"]))
    mel = LM.log_mel(audio)
    s = int(g["frame_stride"])
    # tolerance: north_star "log-mel within 1e-4 relative"; formula: max|a-b| / (max(b)-min(b))
    rng = float(g["mel_max"].max() - g["mel_min"].min())
    assert np.abs(mel[:, :, :64] - g["mel_first_frames"]).max() / rng <= 1e-4
    assert np.abs(mel[:, :, -64:] - g["mel_last_frames"]).max() / rng <= 1e-4
    assert np.abs(mel[:, :, ::s] - g["mel_sub"].astype(np.float32)).max() <= 2e-3  # f16-stored subsample
    assert np.abs(mel.astype(np.float64).sum(axis=2) - g["mel_row_sums"]).max() / 3000 <= 1e-5
    assert np.allclose(mel.max(axis=(1, 2)), g["mel_max"], atol=1e-5)
    assert np.allclose(mel.min(axis=(1, 2)), g["mel_min"], atol=1e-5)


def test_logmel_short_audio_is_zero_padded():
    g = np.load(os.path.join(GOLDEN, "logmel_hf.npz"))
    short = np.random.default_rng(3).standard_normal(16000 * 5).astype(np.float32)
    mel = LM.log_mel(short)
    assert mel.shape == (80, 3000)
    assert np.abs(mel[:, :64] - g["short_first_frames"]).max() <= 1e-4
    assert np.abs(mel.astype(np.float64).sum(axis=1) - g["short_row_sums"]).max() / 3000 <= 1e-5


@pytest.mark.parametrize("name,cfg", [("hf_micro.npz", WhisperConfig.micro()), ("hf_tiny.npz", WhisperConfig.tiny())])
def test_oracle_matches_hf_golden(name, cfg):
    g = np.load(os.path.join(GOLDEN, name))
    w = synth.make_weights(cfg, seed=int(g["weight_seed"]))
    mel = synth.make_mel(1, cfg, int(g["mel_seed"]))[0]
    om = O.OracleWhisper(cfg, w)
    enc = om.encode(mel)
    assert np.abs(enc[:: int(g["row_stride"])] - g["enc_rows"]).max() <= 1e-4
    forced = g["forced"]
    ls = int(g["logit_stride"])
    for q in (1, 0):
        lg = om.teacher_forced(enc, forced, pos_quirk=q)
        assert np.abs(lg[:, ::ls] - g[f"tf_logits_q{q}"]).max() <= 1e-4
        assert np.array_equal(lg.argmax(axis=1), g[f"tf_argmax_q{q}"])
        assert np.abs(lg.max(axis=1) - g[f"tf_max_q{q}"]).max() <= 1e-4
        gold = g[f"greedy_q{q}"]
        toks = om.greedy(enc, pos_quirk=q, max_iters=len(gold) - 5)
        assert np.array_equal(toks, gold)
    # the reference's position quirk must actually change the logits (otherwise the switch is dead)
    assert np.abs(om.teacher_forced(enc, forced, 1) - om.teacher_forced(enc, forced, 0)).max() > 1e-2


def test_weight_layout_matches_reference_format():
    cfg = WhisperConfig.tiny()
    assert len(cfg.weight_layout()) == 167
    assert cfg.weight_count() == 37_760_640  # 151 042 560 bytes
    off = cfg.weight_offsets()
    assert off["enc.pos"][0] == 535_296
    assert off["enc.0.attn.q.w"][0] == 1_111_296
    assert off["enc.1.attn.q.w"][0] - off["enc.0.attn.q.w"][0] == 1_774_080
    assert off["enc.ln_post.w"][0] == 8_207_616
    assert off["dec.token_emb"][0] == 8_208_384
    assert off["dec.pos"][0] == 28_124_544
    assert off["dec.0.attn.q.w"][0] == 28_296_576
    assert off["dec.1.attn.q.w"][0] - off["dec.0.attn.q.w"][0] == 2_365_824
    assert off["dec.ln_post.w"][0] == 37_759_872
    assert len(WhisperConfig.small_shaped().weight_layout()) == 479


def test_reference_expected_tokens_fixture():
    d = json.load(open(os.path.join(GOLDEN, "reference_expected_tokens.json")))
    assert len(d["ids"]) == 89 and d["ids"][:4] == [639, 307, 452, 3177]
