#!/usr/bin/env python
"""Where does the 16-bit error of the fast path come from?  (VERDICT r01 "next" 1b)

An emulation of the fast path's arithmetic on the CPU (torch fp32 matmuls standing in for the fp32
tensor-core accumulators) in which every class of 16-bit rounding can be switched on by itself:

  W     GEMM weights (conv, q/k/v/o, fc1/fc2, token_emb)                 model.cu:model_load
  A     GEMM A operands and 16-bit activations between kernels           ln_* outputs, qkv, attention
        (LayerNorm outputs, q/k/v, attention outputs, GELU(fc1), P)      outputs, h, P tiles
  KV    what the decode step streams per chunk: self K/V cache,          Cache::self_kv, cross_enc, qp, ctx
        the 16-bit enc_out cache, folded query q' and context ctx
  FOLD  the folded cross projections Wqk = Wk_h^T Wq_h, Wov = Wo Wv_h    kernels.cu:fold_cross_weights
        (rounded AFTER folding; without W the unfolded fp32 product)

Each run is compared with the fp32 oracle (oracle/whisper_oracle.c): enc_out max-abs and
teacher-forced logits max-abs on the synthetic weights the parity tests use and on HF-init
weights, for bf16 (8 significand bits) and fp16 (11 bits).  Test infrastructure: imports oracle/.

    python tools/error_attribution.py [--chunks 2] [--steps 12] [--json out.json]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from whisper_mojo_b200 import WhisperConfig, synth  # noqa: E402


def rounder(dtype):
    if dtype is None:
        return lambda t: t
    return lambda t: t.to(dtype).to(torch.float32)


def gelu(x):  # whisper_tensor.mojo:288-308
    return 0.5 * x * (1.0 + torch.tanh(0.79788456 * (x + 0.044715 * x * x * x)))


def ln(x, g, b):  # one-pass variance, whisper_tensor.mojo:249-285
    mean = x.mean(-1, keepdim=True)
    var = (x * x).mean(-1, keepdim=True) - mean * mean
    return (x - mean) / torch.sqrt(var + 1e-5) * g + b


class Emu:
    def __init__(self, cfg, flat, dtype, on):
        self.cfg, self.on = cfg, on
        r = rounder(dtype)
        self.rW = r if "W" in on else (lambda t: t)
        self.rA = r if "A" in on else (lambda t: t)
        self.rKV = r if "KV" in on else (lambda t: t)
        self.rF = r if "FOLD" in on else (lambda t: t)
        self.fold = "FOLD" in on or "FOLDFORM" in on
        tab = cfg.weight_offsets()
        self.w = {k: torch.from_numpy(flat[o:o + int(np.prod(s))].reshape(s).copy()) for k, (o, s) in tab.items()}

    def W(self, name):
        return self.rW(self.w[name])

    def encode(self, mel):
        c, w, rA = self.cfg, self.w, self.rA
        x = rA(torch.from_numpy(mel))[None]  # [1, 80, 3000]
        x = gelu(torch.nn.functional.conv1d(x, self.W("enc.conv1.w"), w["enc.conv1.b"], padding=1))
        x = rA(x)
        x = gelu(torch.nn.functional.conv1d(x, self.W("enc.conv2.w"), w["enc.conv2.b"], stride=2, padding=1))
        x = x[0].T + w["enc.pos"]
        H = c.n_heads
        for i in range(c.n_layers):
            p = f"enc.{i}."
            xn = rA(ln(x, w[p + "attn_ln.w"], w[p + "attn_ln.b"]))
            q = rA(xn @ self.W(p + "attn.q.w").T + w[p + "attn.q.b"])
            k = rA(xn @ self.W(p + "attn.k.w").T)
            v = rA(xn @ self.W(p + "attn.v.w").T + w[p + "attn.v.b"])
            outs = []
            for h in range(H):
                sl = slice(h * 64, h * 64 + 64)
                s = (q[:, sl] @ k[:, sl].T) * 0.125
                e = torch.exp(s - s.max(-1, keepdim=True).values)
                outs.append((rA(e) @ v[:, sl]) / e.sum(-1, keepdim=True))  # P rounded, row sum in fp32 (attn_tc.cu)
            a = rA(torch.cat(outs, 1))
            x = x + a @ self.W(p + "attn.o.w").T + w[p + "attn.o.b"]
            xn = rA(ln(x, w[p + "mlp_ln.w"], w[p + "mlp_ln.b"]))
            h1 = rA(gelu(xn @ self.W(p + "fc1.w").T + w[p + "fc1.b"]))
            x = x + h1 @ self.W(p + "fc2.w").T + w[p + "fc2.b"]
        return ln(x, w["enc.ln_post.w"], w["enc.ln_post.b"])

    def teacher_forced(self, enc_out, forced):
        c, w, rA, rKV = self.cfg, self.w, self.rA, self.rKV
        D, H, L = c.d_model, c.n_heads, c.n_layers
        enc = rKV(torch.from_numpy(enc_out))
        cross = []
        for i in range(L):
            p = f"dec.{i}.cross."
            if self.fold:
                Wq, Wk, Wv, Wo = w[p + "q.w"], w[p + "k.w"], w[p + "v.w"], w[p + "o.w"]
                wqk = torch.cat([Wk[h * 64:h * 64 + 64].T @ Wq[h * 64:h * 64 + 64] for h in range(H)], 0) * 0.125  # [H*D, D]
                bqk = torch.cat([Wk[h * 64:h * 64 + 64].T @ w[p + "q.b"][h * 64:h * 64 + 64] for h in range(H)], 0) * 0.125
                wov = torch.cat([Wo[:, h * 64:h * 64 + 64] @ Wv[h * 64:h * 64 + 64] for h in range(H)], 1)  # [D, H*D]
                bov = w[p + "o.b"] + Wo @ w[p + "v.b"]
                r2 = self.rF if "FOLD" in self.on else self.rW
                cross.append((r2(wqk), bqk, r2(wov), bov))
            else:
                K = rKV(enc @ self.W(p + "k.w").T)
                V = rKV(enc @ self.W(p + "v.w").T + w[p + "v.b"])
                cross.append((K, V))
        ks = [[] for _ in range(L)]
        vs = [[] for _ in range(L)]
        out = []
        for t, tok in enumerate(forced):
            pos = t if t < 4 else t - c.pos_quirk
            x = (w["dec.token_emb"][tok] + w["dec.pos"][pos])[None]
            for i in range(L):
                p = f"dec.{i}."
                xn = rA(ln(x, w[p + "attn_ln.w"], w[p + "attn_ln.b"]))
                q = rKV(rA(xn @ self.W(p + "attn.q.w").T + w[p + "attn.q.b"]))
                ks[i].append(rKV(rA(xn @ self.W(p + "attn.k.w").T)))
                vs[i].append(rKV(rA(xn @ self.W(p + "attn.v.w").T + w[p + "attn.v.b"])))
                K, V = torch.cat(ks[i], 0), torch.cat(vs[i], 0)
                a = []
                for h in range(H):
                    sl = slice(h * 64, h * 64 + 64)
                    pr = torch.softmax((q[:, sl] @ K[:, sl].T) * 0.125, -1)
                    a.append(pr @ V[:, sl])
                x = x + rA(torch.cat(a, 1)) @ self.W(p + "attn.o.w").T + w[p + "attn.o.b"]
                xn = rA(ln(x, w[p + "cross_ln.w"], w[p + "cross_ln.b"]))
                if self.fold:
                    wqk, bqk, wov, bov = cross[i]
                    qp = rKV(rA(xn @ wqk.T + bqk)).reshape(H, D)
                    s = qp @ enc.T  # [H, S]
                    e = torch.exp(s - s.max(-1, keepdim=True).values)
                    ctx = rKV(rA((rA(e) @ enc) / e.sum(-1, keepdim=True))).reshape(1, H * D)
                    x = x + ctx @ wov.T + bov
                else:
                    K, V = cross[i]
                    q = rKV(rA(xn @ self.W(p + "cross.q.w").T + w[p + "cross.q.b"]))
                    a = []
                    for h in range(H):
                        sl = slice(h * 64, h * 64 + 64)
                        pr = torch.softmax((q[:, sl] @ K[:, sl].T) * 0.125, -1)
                        a.append(pr @ V[:, sl])
                    x = x + rA(torch.cat(a, 1)) @ self.W(p + "cross.o.w").T + w[p + "cross.o.b"]
                xn = rA(ln(x, w[p + "mlp_ln.w"], w[p + "mlp_ln.b"]))
                h1 = rA(gelu(xn @ self.W(p + "fc1.w").T + w[p + "fc1.b"]))
                x = x + h1 @ self.W(p + "fc2.w").T + w[p + "fc2.b"]
            if t >= 3:
                xn = rA(ln(x, w["dec.ln_post.w"], w["dec.ln_post.b"]))
                out.append(xn @ self.W("dec.token_emb").T)
        return torch.cat(out, 0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", type=int, default=2)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = WhisperConfig.tiny()
    cases = [("none (fp32 emulation vs oracle)", None, set()), ("absorbed form in fp32", None, {"FOLDFORM"})]
    for name, dt in (("bf16", torch.bfloat16), ("fp16", torch.float16)):
        for on in ({"W"}, {"A"}, {"KV"}, {"FOLD"}, {"W", "A", "KV", "FOLD"}):
            cases.append((f"{name}: " + "+".join(sorted(on)), dt, on))
    results = {}
    for wname, maker in (("synthetic (tests, bench)", lambda: synth.make_weights(cfg, seed=0)),
                         ("hf-init N(0,0.02)", lambda: synth.make_weights_hf_init(cfg, seed=0))):
        flat = maker()
        om = O.OracleWhisper(cfg, flat)
        mel = synth.make_mel(a.chunks, cfg, 0)
        enc_ref = [om.encode(mel[i]) for i in range(a.chunks)]
        forced = [np.concatenate([np.array(cfg.prompt), np.random.default_rng(10 + i).integers(0, cfg.vocab_size, a.steps)]).astype(np.int32)
                  for i in range(a.chunks)]
        lg_ref = [om.teacher_forced(enc_ref[i], forced[i]) for i in range(a.chunks)]
        print(f"\n== weights: {wname}; |enc_out| max {max(np.abs(e).max() for e in enc_ref):.2f}, |logit| max "
              f"{max(np.abs(l).max() for l in lg_ref):.2f}")
        print(f"{'rounding':38s} {'enc_out max':>12s} {'enc mean':>10s} {'logits max':>12s} {'logits med':>11s}")
        results[wname] = {}
        for cname, dt, on in cases:
            e = Emu(cfg, flat, dt, on)
            em, ea, lm, lmed = 0.0, 0.0, 0.0, 0.0
            with torch.no_grad():
                for i in range(a.chunks):
                    if on & {"W", "A"} or not on:
                        d = np.abs(e.encode(mel[i]).numpy() - enc_ref[i])
                        em, ea = max(em, float(d.max())), max(ea, float(d.mean()))
                    # decoder error is measured from the ORACLE's enc_out, as the teacher-forced parity test does
                    d = np.abs(e.teacher_forced(enc_ref[i], forced[i]).numpy() - lg_ref[i])
                    lm, lmed = max(lm, float(d.max())), max(lmed, float(np.median(d)))
            results[wname][cname] = {"enc_max": em, "enc_mean": ea, "logit_max": lm, "logit_median": lmed}
            print(f"{cname:38s} {em:12.3e} {ea:10.2e} {lm:12.3e} {lmed:11.2e}", flush=True)
    if a.json:
        json.dump(results, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
