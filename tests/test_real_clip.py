"""BASELINE.json configs[0]: the reference's real clip (`main.mojo:16-37`, `expected_tokens.txt:1`).

The assets are NOT shipped with the reference (`.gitignore:1` = `*.bin`; `export_weights.py` downloads the
openai/whisper-tiny checkpoint and a wav) and cannot be fetched here, so these tests are gated on their presence:
drop `whisper_tiny_weights.bin` and `sample_input.bin` (f32 [80, 3000]) into `$WHISPER_ASSETS` (default
`<repo>/assets/`) and they run; otherwise they SKIP with the reason (never faked).

`expected_tokens.txt` is HF `model.generate` output (89 ids, no prompt, no EOT), not the Mojo binary's: the reference
differs from HF by tanh GELU, no logits processors and the position off-by-one, so neither `pos_quirk`
setting is REQUIRED to reproduce it; the tests report which (if either) does and assert what must hold either way:
the GPU path's ids equal the CPU restatement's under the strict rule of gpu_util.tokens_agree_up_to_margin.
"""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from oracle import oracle as O
from whisper_mojo_b200 import WhisperConfig

ASSETS = os.environ.get("WHISPER_ASSETS", os.path.join(ROOT, "assets"))
W_PATH, MEL_PATH = os.path.join(ASSETS, "whisper_tiny_weights.bin"), os.path.join(ASSETS, "sample_input.bin")
needs_assets = pytest.mark.skipif(
    not (os.path.exists(W_PATH) and os.path.exists(MEL_PATH)),
    reason=f"reference assets not shipped (export_weights.py needs the HF hub): put whisper_tiny_weights.bin and "
           f"sample_input.bin into {ASSETS} (or set WHISPER_ASSETS) to run the expected_tokens.txt replay")


def golden_ids():
    return np.array(json.load(open(os.path.join(GOLDEN, "reference_expected_tokens.json")))["ids"], np.int32)


def load_assets():
    cfg = WhisperConfig.tiny()
    w = np.fromfile(W_PATH, "<f4")
    assert w.size == cfg.weight_count(), f"{W_PATH}: {w.size} floats, Tiny needs {cfg.weight_count()} (loader.mojo never checks)"
    mel = np.fromfile(MEL_PATH, "<f4")
    assert mel.size == cfg.n_mels * cfg.n_frames, f"{MEL_PATH}: {mel.size} floats, expected 80 x 3000 (main.mojo:23-27)"
    return cfg, w, mel.reshape(cfg.n_mels, cfg.n_frames)


def body(ids, cfg):
    """Strip the 4-id prompt and the trailing EOT the reference appends (whisper.mojo:187-223) -> comparable to
    expected_tokens.txt."""
    ids = [int(t) for t in ids][4:]
    return np.array(ids[:-1] if ids and ids[-1] == cfg.eot else ids, np.int32)


def test_golden_fixture_is_the_references_file():
    ids = golden_ids()
    assert len(ids) == 89 and ids.min() >= 0 and ids.max() < 50257  # text tokens only: no prompt, no EOT (F6)


@needs_assets
def test_oracle_replays_expected_tokens():
    cfg, w, mel = load_assets()
    om = O.OracleWhisper(cfg, w)
    gold = golden_ids()
    res = {}
    for quirk in (1, 0):
        got = body(om.transcribe(mel, pos_quirk=quirk), cfg)
        n = min(len(got), len(gold))
        first = np.nonzero(got[:n] != gold[:n])[0]
        res[quirk] = (len(got) == len(gold) and len(first) == 0, int(first[0]) if len(first) else n)
        print(f"oracle pos_quirk={quirk}: {len(got)} ids, matches expected_tokens.txt up to id {res[quirk][1]} of {len(gold)}"
              + (" -- IDENTICAL" if res[quirk][0] else ""))
    # readme.md:19 claims identical results to HF on this clip; report, and require at least the reference's own
    # setting (pos_quirk = 1) to reproduce a substantial prefix -- a restatement bug would diverge at once
    assert res[1][1] >= 8 or res[0][1] >= 8, res


@needs_assets
@pytest.mark.gpu
def test_gpu_path_replays_the_real_clip():
    from gpu_util import tokens_agree_up_to_margin, tolerances
    from whisper_mojo_b200 import WeightLoader, Whisper

    cfg, w, mel = load_assets()
    gold = golden_ids()
    for quirk in (1, 0):
        c = WhisperConfig(**{**cfg.__dict__, "pos_quirk": quirk})
        m = Whisper(c)
        m.load(WeightLoader(data=w))
        got = np.array(m.transcribe(mel), np.int32)
        om = O.OracleWhisper(c, w)
        ref, mg = om.greedy(om.encode(mel), margins=True)
        ok, msg = tokens_agree_up_to_margin(got, ref, mg, tolerances()[4])
        b = body(got, c)
        n = min(len(b), len(gold))
        first = np.nonzero(b[:n] != gold[:n])[0]
        print(f"gpu pos_quirk={quirk}: vs oracle: {msg}; vs expected_tokens.txt: "
              + ("IDENTICAL" if len(b) == len(gold) and len(first) == 0 else f"first difference at id {int(first[0]) if len(first) else n}"))
        assert ok, msg
