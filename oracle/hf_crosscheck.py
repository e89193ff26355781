"""Independent cross-check of the oracle: HF transformers Whisper driven the reference's way.

TEST INFRASTRUCTURE ONLY.  The reference claims to match HF `model.generate` on its clip
(readme.md:19, export_weights.py:125-131) and defines its weight file from an HF state_dict
(export_weights.py:19-90).  This module loads a flat weight array in that file order INTO an HF
`WhisperForConditionalGeneration` built from a config (no hub access), selects tanh GELU
('gelu_new' -- the reference's whisper_tensor.mojo:288-308 formula), and runs a manual greedy loop
with the reference's prompt, no logits processors and (optionally) the reference's position shift
(whisper.mojo:217).  Agreement of the C restatement with this to ~1e-4 is what pins the oracle,
since the reference's own golden (expected_tokens.txt) needs assets that are not shipped.
"""
from __future__ import annotations

import numpy as np
import torch


def hf_name_map(cfg):
    """our layout name -> HF state_dict key (export_weights.py:19-90)."""
    m = {
        "enc.conv1.w": "model.encoder.conv1.weight", "enc.conv1.b": "model.encoder.conv1.bias",
        "enc.conv2.w": "model.encoder.conv2.weight", "enc.conv2.b": "model.encoder.conv2.bias",
        "enc.pos": "model.encoder.embed_positions.weight",
        "enc.ln_post.w": "model.encoder.layer_norm.weight", "enc.ln_post.b": "model.encoder.layer_norm.bias",
        "dec.token_emb": "model.decoder.embed_tokens.weight", "dec.pos": "model.decoder.embed_positions.weight",
        "dec.ln_post.w": "model.decoder.layer_norm.weight", "dec.ln_post.b": "model.decoder.layer_norm.bias",
    }
    proj = {"q.w": "q_proj.weight", "q.b": "q_proj.bias", "k.w": "k_proj.weight", "v.w": "v_proj.weight",
            "v.b": "v_proj.bias", "o.w": "out_proj.weight", "o.b": "out_proj.bias"}
    for side, hs in (("enc", "encoder"), ("dec", "decoder")):
        for i in range(cfg.n_layers):
            p, h = f"{side}.{i}.", f"model.{hs}.layers.{i}."
            for a, b in proj.items():
                m[p + "attn." + a] = h + "self_attn." + b
                if side == "dec":
                    m[p + "cross." + a] = h + "encoder_attn." + b
            m[p + "attn_ln.w"] = h + "self_attn_layer_norm.weight"
            m[p + "attn_ln.b"] = h + "self_attn_layer_norm.bias"
            if side == "dec":
                m[p + "cross_ln.w"] = h + "encoder_attn_layer_norm.weight"
                m[p + "cross_ln.b"] = h + "encoder_attn_layer_norm.bias"
            for a, b in (("fc1.w", "fc1.weight"), ("fc1.b", "fc1.bias"), ("fc2.w", "fc2.weight"),
                         ("fc2.b", "fc2.bias"), ("mlp_ln.w", "final_layer_norm.weight"),
                         ("mlp_ln.b", "final_layer_norm.bias")):
                m[p + a] = h + b
    return m


def build_hf(cfg, flat: np.ndarray, activation: str = "gelu_new"):
    from transformers import WhisperConfig as HFConfig, WhisperForConditionalGeneration

    hc = HFConfig(vocab_size=cfg.vocab_size, num_mel_bins=cfg.n_mels, d_model=cfg.d_model,
                  encoder_layers=cfg.n_layers, decoder_layers=cfg.n_layers,
                  encoder_attention_heads=cfg.n_heads, decoder_attention_heads=cfg.n_heads,
                  encoder_ffn_dim=4 * cfg.d_model, decoder_ffn_dim=4 * cfg.d_model,
                  max_source_positions=cfg.n_audio_ctx, max_target_positions=cfg.n_text_ctx,
                  activation_function=activation, pad_token_id=0, bos_token_id=1, eos_token_id=2,
                  decoder_start_token_id=1, attn_implementation="eager")
    model = WhisperForConditionalGeneration(hc).eval()
    sd = model.state_dict()
    table = cfg.weight_offsets()
    names = hf_name_map(cfg)
    new = {}
    for ours, (off, shape) in table.items():
        n = int(np.prod(shape))
        new[names[ours]] = torch.from_numpy(flat[off:off + n].reshape(shape).copy())
    missing = [k for k in sd if k not in new and k != "proj_out.weight" and not k.endswith("k_proj.bias")]
    assert not missing, missing
    new["proj_out.weight"] = new["model.decoder.embed_tokens.weight"]
    res = model.load_state_dict(new, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    return model


@torch.no_grad()
def hf_encode(model, mel: np.ndarray) -> np.ndarray:
    """HF's encoder modules driven by hand: HF hard-codes erf-GELU after the two convs whatever
    `activation_function` says, while the reference uses its tanh GELU there too
    (whisper.mojo:75,80), so the stem is applied explicitly with approximate='tanh'."""
    enc = model.model.encoder
    x = torch.from_numpy(np.ascontiguousarray(mel, np.float32))[None]
    x = torch.nn.functional.gelu(enc.conv1(x), approximate="tanh")
    x = torch.nn.functional.gelu(enc.conv2(x), approximate="tanh")
    x = x.permute(0, 2, 1) + enc.embed_positions.weight
    for layer in enc.layers:
        x = layer(x, None)
        x = x[0] if isinstance(x, tuple) else x
    return enc.layer_norm(x)[0].numpy()


def _decoder_last_hidden(model, enc, toks, pos_quirk):
    """Full-prefix decoder pass with the reference's positions: index n uses position n for n < 4
    and n - pos_quirk afterwards (whisper.mojo:212-218).  HF derives its causal mask from
    position_ids, so the shift is applied through inputs_embeds instead: HF adds pos[n] itself and
    the embeds carry tok + (pos[n - quirk] - pos[n])."""
    dec = model.model.decoder
    n = len(toks)
    ids = torch.as_tensor(np.asarray(toks, np.int64))[None]
    emb = dec.embed_tokens(ids)
    if pos_quirk and n > 4:
        P = dec.embed_positions.weight
        idx = torch.arange(4, n)
        emb = emb.clone()
        emb[0, 4:] += P[idx - pos_quirk] - P[idx]
    out = dec(inputs_embeds=emb, encoder_hidden_states=enc, use_cache=False)
    return out.last_hidden_state[0, -1]


@torch.no_grad()
def hf_teacher_forced(model, cfg, enc_out: np.ndarray, forced: np.ndarray, pos_quirk: int) -> np.ndarray:
    """Logits [len(forced)-3, vocab]; row 0 after the 4-token prefill, then one per forced token.
    Runs the full prefix each time (no cache); positions per _decoder_last_hidden."""
    enc = torch.from_numpy(np.ascontiguousarray(enc_out, np.float32))[None]
    rows = []
    for n in range(4, len(forced) + 1):
        h = _decoder_last_hidden(model, enc, forced[:n], pos_quirk)
        rows.append((model.proj_out.weight @ h).numpy())
    return np.stack(rows)


@torch.no_grad()
def hf_greedy(model, cfg, enc_out: np.ndarray, pos_quirk: int, max_iters: int) -> np.ndarray:
    toks = list(cfg.prompt)
    enc = torch.from_numpy(np.ascontiguousarray(enc_out, np.float32))[None]
    for it in range(max_iters + 1):
        logits = model.proj_out.weight @ _decoder_last_hidden(model, enc, toks, pos_quirk)
        nxt = int(torch.argmax(logits))
        toks.append(nxt)
        if nxt == cfg.eot:
            break
    return np.array(toks, np.int32)


def hf_generate_rate(cfg, flat: np.ndarray, mel: np.ndarray, n_new: int = 196):
    """The reference's second CPU program, benchmark_python.py:8-30, restated for assets that exist here:
    HF `WhisperForConditionalGeneration` (built from a config, same flat weights, HF's own erf GELU and
    KV-cached `model.generate`, greedy, exactly `n_new` new tokens: EOS suppressed so the work is the same
    196 decoder forwards the GPU path runs) on one [n_mels, n_frames] chunk, one warm-up + one timed call.
    Returns (seconds, torch threads).  CPU baseline only: never on the product path."""
    import time
    from transformers.utils import logging as hf_logging

    hf_logging.set_verbosity_error()
    model = build_hf(cfg, flat, activation="gelu")
    x = torch.from_numpy(np.ascontiguousarray(mel, np.float32))[None]
    prompt = torch.as_tensor([list(cfg.prompt)], dtype=torch.long)
    kw = dict(decoder_input_ids=prompt, max_new_tokens=n_new, min_new_tokens=n_new, do_sample=False, num_beams=1)
    with torch.no_grad():
        model.generate(input_features=x, **kw)  # warm-up, benchmark_python.py:25-26
        t0 = time.perf_counter()
        out = model.generate(input_features=x, **kw)
        dt = time.perf_counter() - t0
    assert out.shape[1] >= n_new, out.shape
    return dt, torch.get_num_threads()
