"""GPU bring-up diagnostics: prints parity numbers section by section (each guarded so one failure
does not hide the rest).  Run on the GPU box:  python tools/gpu_diag.py [sections...]"""
import ctypes
import os
import sys
import time
import traceback
from ctypes import c_void_p

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from whisper_mojo_b200 import _lib, synth  # noqa: E402
from whisper_mojo_b200 import Whisper, WeightLoader, WhisperConfig  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle import logmel_oracle as LM  # noqa: E402


def bf16_round(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).bfloat16().float().numpy()


def debug_gemm(impl, A, W, bias, epi, batches=1, taps=1, conv_stride=1, pad=0, rows_per_batch=None, out0=None):
    A = np.ascontiguousarray(A, np.float32)
    W = np.ascontiguousarray(W, np.float32)
    src_rows, lda = A.shape[-2], A.shape[-1]
    Cin = lda
    N = W.shape[0]
    rows = rows_per_batch or src_rows
    out = np.zeros((batches * rows, N), np.float32) if out0 is None else np.ascontiguousarray(out0, np.float32).copy()
    b = None if bias is None else np.ascontiguousarray(bias, np.float32)
    rc = _lib.load().wb_debug_gemm(impl, A.ctypes.data_as(c_void_p), batches, src_rows, lda, Cin, taps, conv_stride, pad,
                                   rows, W.ctypes.data_as(c_void_p), N, b.ctypes.data_as(c_void_p) if b is not None else None,
                                   epi, out.ctypes.data_as(c_void_p))
    _lib.check(rc)
    return out


def sec_gemm():
    rng = np.random.default_rng(0)
    for (M, N, K) in [(128, 128, 64), (256, 384, 384), (200, 1152, 384), (1500, 384, 1536), (4, 51865 // 16, 384), (1, 1000, 128)]:
        A = rng.standard_normal((M, K), dtype=np.float32)
        W = rng.standard_normal((N, K), dtype=np.float32) / np.sqrt(K)
        b = rng.standard_normal(N, dtype=np.float32)
        ref = bf16_round(A).astype(np.float64) @ bf16_round(W).astype(np.float64).T + b
        for impl in (0, 1):
            t = time.time()
            out = debug_gemm(impl, A, W, b, 3)
            print(f"gemm impl={impl} M={M} N={N} K={K}: maxabs err {np.abs(out - ref).max():.3e}  ({time.time()-t:.2f}s)", flush=True)
        outb = debug_gemm(1, A, W, b, 0)
        print(f"   bf16 store err {np.abs(outb - ref).max():.3e}; gelu:", end=" ")
        import torch
        g = torch.nn.functional.gelu(torch.from_numpy(ref), approximate="tanh").numpy()
        print(f"{np.abs(debug_gemm(1, A, W, b, 1) - g).max():.3e}; resid:", end=" ")
        x0 = rng.standard_normal((M, N), dtype=np.float32)
        print(f"{np.abs(debug_gemm(1, A, W, b, 2, out0=x0) - (x0 + ref)).max():.3e}; argmax:", end=" ")
        am = debug_gemm(1, A, W, None, 4)
        lg = bf16_round(A).astype(np.float64) @ bf16_round(W).astype(np.float64).T
        print("match" if np.array_equal(am[:, -1].astype(np.int64), lg.argmax(1)) else f"MISMATCH {am[:4,-1]} vs {lg.argmax(1)[:4]}",
              f"logits err {np.abs(am[:, :-1] - lg[:, :-1]).max():.3e}", flush=True)
    # conv-style: 3 taps, stride 1 and 2, batches
    for (cs, Cin, L, N) in [(1, 128, 300, 128), (2, 128, 300, 256), (2, 384, 3000, 384)]:
        B = 2
        A = rng.standard_normal((B, L, Cin), dtype=np.float32)
        W = rng.standard_normal((N, 3 * Cin), dtype=np.float32) / np.sqrt(3 * Cin)
        Lo = (L + 2 - 3) // cs + 1
        Ab, Wb = bf16_round(A).astype(np.float64), bf16_round(W).astype(np.float64)
        ref = np.zeros((B, Lo, N))
        for t in range(3):
            for m in range(Lo):
                r = m * cs + t - 1
                if 0 <= r < L:
                    ref[:, m] += Ab[:, r] @ Wb[:, t * Cin:(t + 1) * Cin].T
        for impl in (0, 1):
            out = debug_gemm(impl, A.reshape(B * L, Cin).reshape(B, L, Cin), W, None, 3, batches=B, taps=3, conv_stride=cs, pad=1, rows_per_batch=Lo)
            print(f"conv-gemm impl={impl} stride={cs} Cin={Cin} L={L} N={N}: maxabs err {np.abs(out.reshape(B, Lo, N) - ref).max():.3e}", flush=True)


def make_model(cfg, seed=0, **opts):
    w = synth.make_weights(cfg, seed=seed)
    m = Whisper(cfg)
    for k, v in opts.items():
        m.set_option(k, v)
    m.load(WeightLoader(data=w))
    return m, w


def sec_logmel():
    m = Whisper(WhisperConfig.tiny())
    a = synth.make_audio(3, seed=0)
    t = time.time(); mel = m.log_mel(a); dt = time.time() - t
    ref = LM.log_mel(a)
    rng = ref.max() - ref.min()
    print(f"logmel: maxabs {np.abs(mel - ref).max():.3e} rel(range) {np.abs(mel - ref).max() / rng:.3e}  ({dt:.2f}s)")
    g = np.load(os.path.join(ROOT, "tests", "golden", "logmel_hf.npz"))
    print(f"logmel vs HF golden first frames: {np.abs(mel[:2, :, :64] - g['mel_first_frames']).max():.3e} last: {np.abs(mel[:2, :, -64:] - g['mel_last_frames']).max():.3e}")


def sec_model(cfg_name):
    import torch
    cfg = WhisperConfig.micro() if cfg_name == "micro" else WhisperConfig.tiny()
    n = 3
    mel = synth.make_mel(n, cfg, 0)
    w = synth.make_weights(cfg, seed=0)
    om = O.OracleWhisper(cfg, w)
    enc_ref = np.stack([om.encode(mel[i]) for i in range(n)])
    forced = np.stack([np.concatenate([np.array(cfg.prompt), np.random.default_rng(10 + i).integers(0, cfg.vocab_size, 12)]) for i in range(n)]).astype(np.int32)
    lg_ref = np.stack([om.teacher_forced(enc_ref[i], forced[i]) for i in range(n)])
    tok_ref = [om.greedy(enc_ref[i], margins=True) for i in range(n)]
    for impl in (0, 1):
        try:
            m, _ = make_model(cfg, gemm_impl=impl)
            t = time.time(); enc = m.encode(mel); dt = time.time() - t
            print(f"[{cfg_name} gemm_impl={impl}] enc_out maxabs err {np.abs(enc - enc_ref).max():.3e} (|ref|max {np.abs(enc_ref).max():.2f}) {dt:.2f}s", flush=True)
            lg = m.teacher_forced(torch.from_numpy(enc_ref).cuda(), forced)
            print(f"   teacher-forced logits (oracle enc) maxabs err {np.abs(lg - lg_ref).max():.3e} (|ref|max {np.abs(lg_ref).max():.2f}) argmax agree {np.mean(lg.argmax(-1) == lg_ref.argmax(-1)):.3f}", flush=True)
            t = time.time(); toks, lens = m.transcribe_batch(mel); dt = time.time() - t
            for i in range(n):
                rt, mg = tok_ref[i]
                got = toks[i, :lens[i]]
                k = min(len(rt), len(got))
                neq = np.nonzero(rt[:k] != got[:k])[0]
                first = int(neq[0]) if len(neq) else -1
                print(f"   chunk {i}: len {lens[i]} vs {len(rt)}; first mismatch idx {first}" + (f" (oracle margin there {mg[first - 4]:.4f})" if first >= 4 else "") + f"; min margin {mg.min():.4f}", flush=True)
            print(f"   transcribe {dt:.2f}s timing {m.last_timing()}", flush=True)
            t1, l1 = m.transcribe_batch(mel[1:2])
            print("   batch invariance (chunk 1 alone == in batch):", np.array_equal(t1[0], toks[1]), flush=True)
            del m
        except Exception:
            traceback.print_exc()
    try:
        if cfg_name == "micro":
            mo = Whisper(cfg, engine="ops"); mo.load(WeightLoader(data=w))
            t = time.time(); tk = mo.transcribe(mel[0]); dt = time.time() - t
            print(f"[micro ops engine] tokens equal oracle: {np.array_equal(np.array(tk), tok_ref[0][0])} ({dt:.1f}s)")
            from whisper_mojo_b200 import Tensor
            e = mo.encoder.forward(Tensor.from_numpy(mel[0])).numpy()
            print(f"   ops enc_out maxabs err {np.abs(e - enc_ref[0]).max():.3e}")
    except Exception:
        traceback.print_exc()


def sec_attn():
    rng = np.random.default_rng(0)
    for (B, H, ln, splits) in [(3, 6, 1500, 1), (3, 6, 1500, 11), (2, 2, 96, 1), (5, 6, 7, 1), (2, 12, 200, 1), (1, 6, 1, 1)]:
        D = H * 64
        q = bf16_round(rng.standard_normal((B, D), dtype=np.float32) * 1.5)
        K = bf16_round(rng.standard_normal((B, ln, D), dtype=np.float32) * 1.5)
        V = bf16_round(rng.standard_normal((B, ln, D), dtype=np.float32))
        out = np.zeros((B, D), np.float32)
        _lib.check(_lib.load().wb_debug_decode_attention(q.ctypes.data_as(c_void_p), K.ctypes.data_as(c_void_p), V.ctypes.data_as(c_void_p), B, H, ln, splits, out.ctypes.data_as(c_void_p)))
        qh = q.reshape(B, H, 64).astype(np.float64); Kh = K.reshape(B, ln, H, 64).astype(np.float64); Vh = V.reshape(B, ln, H, 64).astype(np.float64)
        s_ = np.einsum("bhd,bjhd->bhj", qh, Kh) * 0.125
        p_ = np.exp(s_ - s_.max(-1, keepdims=True)); p_ /= p_.sum(-1, keepdims=True)
        ref = np.einsum("bhj,bjhd->bhd", p_, Vh).reshape(B, D)
        print(f"decode_attn B={B} H={H} len={ln} splits={splits}: maxabs err {np.abs(out - ref).max():.3e} (|ref| max {np.abs(ref).max():.2f})", flush=True)


def sec_tfdetail(cfg_name="tiny"):
    import torch
    cfg = WhisperConfig.micro() if cfg_name.endswith("micro") else WhisperConfig.tiny()
    w = synth.make_weights(cfg, seed=0)
    mel = synth.make_mel(1, cfg, 0)
    om = O.OracleWhisper(cfg, w)
    enc_ref = om.encode(mel[0])
    nat, _ = om.greedy(enc_ref, margins=True)
    rnd = np.concatenate([np.array(cfg.prompt), np.random.default_rng(10).integers(0, cfg.vocab_size, 12)]).astype(np.int32)
    m, _ = make_model(cfg)
    for name, forced in (("random", rnd), ("natural", nat[:16].astype(np.int32))):
        ref = om.teacher_forced(enc_ref, forced)
        lg = m.teacher_forced(torch.from_numpy(enc_ref[None]).cuda(), forced[None])[0]
        err = np.abs(lg - ref)
        print(f"[{cfg_name} {name}] per-step max err:", np.array2string(err.max(1), precision=3), flush=True)
        print(f"    err at fed token id:", np.array2string(np.array([err[i, forced[i + 3]] for i in range(len(ref))]), precision=3))
        print(f"    ref logit at fed token:", np.array2string(np.array([ref[i, forced[i + 3]] for i in range(len(ref))]), precision=2))
        print(f"    err excluding fed token: {np.max([np.delete(err[i], forced[i + 3]).max() for i in range(len(ref))]):.3e}; median err {np.median(err):.3e}", flush=True)
    # step API vs teacher forced
    from whisper_mojo_b200 import DeviceKVCache
    c = DeviceKVCache(m, 1, 32)
    e = torch.from_numpy(enc_ref[None]).cuda()
    c.set_encoder(e.data_ptr())
    ref = om.teacher_forced(enc_ref, rnd)
    for i in range(len(rnd)):
        pos = i if i < 4 else i - 1
        nx, lg = m.decode_step(c, [int(rnd[i])], pos, want_logits=True)
        if i >= 3:
            print(f"    step api i={i}: err {np.abs(lg[0] - ref[i - 3]).max():.3e} next {nx[0]} ref {ref[i - 3].argmax()}", flush=True)


def sec_pcm():
    cfg = WhisperConfig.tiny()
    m, w = make_model(cfg)
    a = synth.make_audio(2, seed=0)
    toks, lens = m.transcribe_pcm_batch(a)
    print("pcm transcribe lens", lens, "timing", m.last_timing())
    mel = m.log_mel(a)
    t2, l2 = m.transcribe_batch(mel)
    print("pcm path == mel path:", np.array_equal(toks, t2))


if __name__ == "__main__":
    secs = sys.argv[1:] or ["gemm", "logmel", "micro", "tiny", "pcm"]
    for s in secs:
        print(f"===== {s} =====", flush=True)
        try:
            if s == "gemm": sec_gemm()
            elif s == "logmel": sec_logmel()
            elif s in ("micro", "tiny"): sec_model(s)
            elif s == "pcm": sec_pcm()
            elif s == "attn": sec_attn()
            elif s.startswith("tfdetail"): sec_tfdetail(s)
        except Exception:
            traceback.print_exc()
    print("launches", _lib.launch_count())
