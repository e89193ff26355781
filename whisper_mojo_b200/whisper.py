"""Mirror of whisper.mojo: WhisperConfig, WhisperEncoder, WhisperDecoder, Whisper.

Two engines behind the same API:
  engine="fast" (default)  the batched B200 path (wm_* C ABI): bf16 tcgen05 GEMMs, KV-cached batched
                           greedy decode with fused logits+argmax.  `transcribe(mel)` with one
                           [80, 3000] mel is exactly main.mojo's call (main.mojo:30).
  engine="ops"             op-by-op orchestration over the wt_* kernels in fp32, structured like the
                           reference (layers.py) -- the path a Mojo host would drive through FFI.
"""
from __future__ import annotations

import ctypes
from ctypes import c_float, c_int, c_int64, c_uint64, c_void_p
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from . import whisper_tensor as wt
from .config import WhisperConfig
from .layers import KVCache, LayerCache, ResidualAttentionBlock
from .loader import WeightLoader
from .whisper_tensor import Tensor

__all__ = ["WhisperConfig", "WhisperEncoder", "WhisperDecoder", "Whisper", "DeviceKVCache"]


def _is_cuda_tensor(x) -> bool:
    return hasattr(x, "data_ptr") and getattr(x, "is_cuda", False)


class WhisperEncoder:
    """whisper.mojo:34-99 (op-level engine)."""

    def __init__(self, config: WhisperConfig):
        self.config = config
        self.blocks = [ResidualAttentionBlock(config.d_model, config.n_heads, is_decoder=False)
                       for _ in range(config.n_layers)]

    def load(self, loader: WeightLoader):
        c = self.config
        d = c.d_model
        self.conv1_w = wt.transpose_conv_weights(loader.next_tensor(d, c.n_mels * 3), d, c.n_mels, 3)
        self.conv1_b = loader.next_tensor(1, d)
        self.conv2_w = wt.transpose_conv_weights(loader.next_tensor(d, d * 3), d, d, 3)
        self.conv2_b = loader.next_tensor(1, d)
        self.pos_emb = loader.next_tensor(c.n_audio_ctx, d)
        for b in self.blocks:
            b.load(loader, is_decoder_block=False)
        self.ln_post_w = loader.next_tensor(1, d)
        self.ln_post_b = loader.next_tensor(1, d)

    def forward(self, mel: Tensor) -> Tensor:
        c = self.config
        d, S = c.d_model, c.n_audio_ctx
        x1 = Tensor(d, 2 * S)
        wt.conv1d(x1, mel, self.conv1_w, self.conv1_b, stride=1, padding=1)
        wt.gelu(x1)
        x2 = Tensor(S, d)
        wt.conv1d(x2, x1, self.conv2_w, self.conv2_b, stride=2, padding=1, out_T=True)
        wt.gelu(x2)
        x = Tensor(S, d)
        wt.add(x, x2, self.pos_emb)
        for b in self.blocks:
            x = b.forward(x, Tensor(0, 0), LayerCache(), use_cache=False)
        out = Tensor(x.rows, x.cols)
        wt.layer_norm(out, x, self.ln_post_w, self.ln_post_b)
        return out


class WhisperDecoder:
    """whisper.mojo:102-167 (op-level engine)."""

    def __init__(self, config: WhisperConfig):
        self.config = config
        self.blocks = [ResidualAttentionBlock(config.d_model, config.n_heads, is_decoder=True)
                       for _ in range(config.n_layers)]

    def load(self, loader: WeightLoader):
        c = self.config
        self.token_emb = loader.next_tensor(c.vocab_size, c.d_model)
        self.pos_emb = loader.next_tensor(c.n_text_ctx, c.d_model)
        for b in self.blocks:
            b.load(loader, is_decoder_block=True)
        self.ln_post_w = loader.next_tensor(1, c.d_model)
        self.ln_post_b = loader.next_tensor(1, c.d_model)

    def forward(self, tokens: Sequence[int], enc_out: Tensor, cache: KVCache, use_cache: bool = False,
                start_pos: int = 0) -> Tensor:
        c = self.config
        L_tgt = len(tokens)
        x = Tensor(L_tgt, c.d_model)
        wt.embed(x, self.token_emb, self.pos_emb, tokens, start_pos)
        for i, b in enumerate(self.blocks):
            x = b.forward(x, enc_out, cache.layers[i], use_cache)
        out = Tensor(x.rows, x.cols)
        wt.layer_norm(out, x, self.ln_post_w, self.ln_post_b)
        last_hidden = Tensor.view(out, 1, c.d_model, offset=(L_tgt - 1) * c.d_model)
        logits = Tensor(1, c.vocab_size)
        wt.matmul(logits, last_hidden, self.token_emb, Tensor(0, 0))
        return logits


class DeviceKVCache:
    """KVCache for `n_chunks` sequences on the fast engine (wm_kvcache_*)."""

    def __init__(self, model: "Whisper", n_chunks: int, max_len: Optional[int] = None):
        self._model = model
        self.n_chunks = n_chunks
        h = c_uint64(0)
        _lib.check(_lib.load().wm_kvcache_create(model._h, n_chunks, max_len or model.config.n_text_ctx,
                                                 ctypes.byref(h)))
        self._h = h.value

    @property
    def current_len(self) -> int:
        n = c_int(0)
        _lib.check(_lib.load().wm_kvcache_len(self._h, ctypes.byref(n)))
        return n.value

    def set_encoder(self, enc_out_dev_ptr: int) -> None:
        """enc_out f32 [n_chunks, ctx, d_model] on the device (raw pointer).  The model's stream is ordered after
        torch's current stream first, so a tensor still being produced there is not read early."""
        self._model._order_after_torch()
        _lib.check(_lib.load().wm_kvcache_set_encoder_dev(self._model._h, self._h, c_void_p(enc_out_dev_ptr)))

    def reset(self) -> None:
        _lib.check(_lib.load().wm_kvcache_reset(self._h))

    def __del__(self):
        h, self._h = getattr(self, "_h", 0), 0
        if h:
            try:
                _lib.load().wm_kvcache_destroy(h)
            except Exception:
                pass


class Whisper:
    """whisper.mojo:170-223."""

    def __init__(self, config: Optional[WhisperConfig] = None, engine: str = "fast", stream: Optional[int] = None):
        self.config = config or WhisperConfig.tiny()
        assert engine in ("fast", "ops")
        self.engine = engine
        self._h = 0
        self.encoder = WhisperEncoder(self.config)
        self.decoder = WhisperDecoder(self.config)
        if engine == "fast":
            h = c_uint64(0)
            self._cfg_c = self.config.as_c_array()
            _lib.check(_lib.load().wm_create(ctypes.cast(self._cfg_c, c_void_p), c_void_p(stream or 0),
                                             ctypes.byref(h)))
            self._h = h.value

    def __del__(self):
        h, self._h = getattr(self, "_h", 0), 0
        if h:
            try:
                _lib.load().wm_destroy(h)
            except Exception:
                pass

    # ---- loading ---------------------------------------------------------------------------
    def load(self, loader: WeightLoader) -> None:
        """Whisper.load(loader) (whisper.mojo:180-182)."""
        if self.engine == "ops":
            self.encoder.load(loader)
            self.decoder.load(loader)
            return
        w = loader.raw_data
        _lib.check(_lib.load().wm_load_weights(self._h, w.ctypes.data_as(c_void_p), w.size))
        loader.offset = loader.size

    def load_file(self, path: str) -> None:
        assert self.engine == "fast"
        _lib.check(_lib.load().wm_load_weights_file(self._h, path.encode()))

    def set_option(self, key: str, value: int) -> None:
        _lib.check(_lib.load().wm_set_option(self._h, key.encode(), int(value)))

    def set_stop_lengths(self, lens=None) -> None:
        """Declare per-chunk lengths for later transcribe calls (wm_set_stop_lengths; None clears)."""
        if lens is None or len(lens) == 0:
            _lib.check(_lib.load().wm_set_stop_lengths(self._h, None, 0))
            return
        a = np.ascontiguousarray(lens, np.int32)
        _lib.check(_lib.load().wm_set_stop_lengths(self._h, a.ctypes.data_as(c_void_p), a.size))

    def stream_ptr(self) -> int:
        """cudaStream_t the model enqueues on (wm_stream)."""
        p = c_void_p(0)
        _lib.check(_lib.load().wm_stream(self._h, ctypes.byref(p)))
        return p.value or 0

    def synchronize(self) -> None:
        _lib.check(_lib.load().wm_synchronize(self._h))

    def _order_after_torch(self, device=None) -> None:
        """Make the model's stream wait for everything queued so far on torch's current stream (CUDA-tensor inputs may
        still be in flight there; the library runs on its own non-blocking stream unless one was passed to Whisper())."""
        import torch

        cur = torch.cuda.current_stream(device)
        ptr = self.stream_ptr()
        if ptr != cur.cuda_stream:
            torch.cuda.ExternalStream(ptr, device=cur.device).wait_stream(cur)

    # ---- main.mojo's call ------------------------------------------------------------------
    def transcribe(self, mel) -> List[int]:
        """transcribe(mel: Tensor[80, 3000]) -> List[Int] (whisper.mojo:184-223)."""
        if self.engine == "ops":
            return self._transcribe_ops(mel)
        a = mel.numpy() if isinstance(mel, Tensor) else np.asarray(mel, np.float32)
        toks, lens = self.transcribe_batch(a[None])
        return [int(t) for t in toks[0, :lens[0]]]

    def _transcribe_ops(self, mel) -> List[int]:
        c = self.config
        if not isinstance(mel, Tensor):
            mel = Tensor.from_numpy(np.asarray(mel, np.float32))
        enc_out = self.encoder.forward(mel)
        tokens = list(c.prompt)
        cache = KVCache(c.n_layers, c.d_model, c.n_text_ctx, c.n_audio_ctx)
        logits = self.decoder.forward(tokens, enc_out, cache, use_cache=True, start_pos=0)
        nxt = wt.argmax(logits)
        all_tokens = tokens + [nxt]
        for _ in range(c.max_iters):
            if nxt == c.eot:
                break
            start = cache.layers[0].current_len - (1 if c.pos_quirk else 0)  # whisper.mojo:217
            logits = self.decoder.forward([nxt], enc_out, cache, use_cache=True, start_pos=start)
            nxt = wt.argmax(logits)
            all_tokens.append(nxt)
        return all_tokens

    # ---- batched fast path -----------------------------------------------------------------
    def _out_buffers(self, n):
        T_out = self.config.max_tokens
        return np.full((n, T_out), -1, np.int32), np.zeros(n, np.int32)

    def transcribe_batch(self, mel):
        """mel f32 [n, n_mels, n_frames] (numpy, or a CUDA torch tensor) -> (tokens int32 [n, 5+max_iters]
        padded with -1, lengths int32 [n]).  With a CUDA tensor the outputs are CUDA tensors too."""
        c = self.config
        lib = _lib.load()
        if _is_cuda_tensor(mel):
            import torch

            assert mel.dtype == torch.float32 and mel.is_contiguous() and tuple(mel.shape[1:]) == (c.n_mels, c.n_frames)
            n = mel.shape[0]
            toks = torch.empty((n, c.max_tokens), dtype=torch.int32, device=mel.device)
            lens = torch.empty((n,), dtype=torch.int32, device=mel.device)
            self._order_after_torch(mel.device)  # wm_transcribe*_dev synchronise before returning: outputs are ready
            _lib.check(lib.wm_transcribe_dev(self._h, c_void_p(mel.data_ptr()), n, c_void_p(toks.data_ptr()),
                                             c_void_p(lens.data_ptr())))
            return toks, lens
        a = np.ascontiguousarray(mel, np.float32)
        assert a.ndim == 3 and a.shape[1:] == (c.n_mels, c.n_frames), a.shape
        toks, lens = self._out_buffers(a.shape[0])
        _lib.check(lib.wm_transcribe(self._h, a.ctypes.data_as(c_void_p), a.shape[0], toks.ctypes.data_as(c_void_p),
                                     lens.ctypes.data_as(c_void_p)))
        return toks, lens

    def transcribe_pcm_batch(self, pcm):
        """pcm f32 [n, n_samples] 16 kHz (numpy or CUDA torch tensor) -> (tokens, lengths); the
        log-mel frontend runs on the GPU in front of the encoder."""
        c = self.config
        lib = _lib.load()
        if _is_cuda_tensor(pcm):
            import torch

            assert pcm.dtype == torch.float32 and pcm.is_contiguous() and pcm.shape[1] == c.n_samples
            n = pcm.shape[0]
            toks = torch.empty((n, c.max_tokens), dtype=torch.int32, device=pcm.device)
            lens = torch.empty((n,), dtype=torch.int32, device=pcm.device)
            self._order_after_torch(pcm.device)
            _lib.check(lib.wm_transcribe_pcm_dev(self._h, c_void_p(pcm.data_ptr()), n, c_void_p(toks.data_ptr()),
                                                 c_void_p(lens.data_ptr())))
            return toks, lens
        a = np.ascontiguousarray(pcm, np.float32)
        assert a.ndim == 2 and a.shape[1] == c.n_samples, a.shape
        toks, lens = self._out_buffers(a.shape[0])
        _lib.check(lib.wm_transcribe_pcm(self._h, a.ctypes.data_as(c_void_p), a.shape[0],
                                         toks.ctypes.data_as(c_void_p), lens.ctypes.data_as(c_void_p)))
        return toks, lens

    def log_mel(self, pcm: np.ndarray) -> np.ndarray:
        """pcm f32 [n, n_samples] -> log-mel f32 [n, n_mels, n_frames] (export_weights.py:116 on the GPU)."""
        c = self.config
        a = np.ascontiguousarray(pcm, np.float32)
        assert a.ndim == 2 and a.shape[1] == c.n_samples, a.shape
        out = np.empty((a.shape[0], c.n_mels, c.n_frames), np.float32)
        _lib.check(_lib.load().wm_logmel(self._h, a.ctypes.data_as(c_void_p), a.shape[0], out.ctypes.data_as(c_void_p)))
        return out

    def encode(self, mel: np.ndarray) -> np.ndarray:
        """WhisperEncoder.forward batched: mel [n, n_mels, n_frames] -> enc_out [n, n_audio_ctx, d_model]."""
        c = self.config
        a = np.ascontiguousarray(mel, np.float32)
        assert a.ndim == 3 and a.shape[1:] == (c.n_mels, c.n_frames), a.shape
        out = np.empty((a.shape[0], c.n_audio_ctx, c.d_model), np.float32)
        _lib.check(_lib.load().wm_encode(self._h, a.ctypes.data_as(c_void_p), a.shape[0], out.ctypes.data_as(c_void_p)))
        return out

    def teacher_forced(self, enc_out, forced: np.ndarray) -> np.ndarray:
        """enc_out [n, ctx, d] (CUDA torch tensor) and forced int32 [n, n_forced] -> logits
        [n, n_forced - 3, vocab] with the greedy loop's position rule (parity tests)."""
        assert _is_cuda_tensor(enc_out)
        self._order_after_torch(enc_out.device)
        f = np.ascontiguousarray(forced, np.int32)
        n, nf = f.shape
        out = np.empty((n, nf - 3, self.config.vocab_size), np.float32)
        _lib.check(_lib.load().wm_teacher_forced(self._h, c_void_p(enc_out.data_ptr()), n, f.ctypes.data_as(c_void_p),
                                                 nf, out.ctypes.data_as(c_void_p)))
        return out

    def decode_step(self, cache: DeviceKVCache, tokens: Sequence[int], start_pos: int, want_logits: bool = False):
        """WhisperDecoder.forward([tok], enc_out, cache, use_cache=True, start_pos) for every chunk of
        the cache; returns (next_tokens int32 [n], logits or None)."""
        t = np.ascontiguousarray(tokens, np.int32)
        assert t.size == cache.n_chunks
        nxt = np.empty(cache.n_chunks, np.int32)
        logits = np.empty((cache.n_chunks, self.config.vocab_size), np.float32) if want_logits else None
        _lib.check(_lib.load().wm_decode_step(self._h, cache._h, t.ctypes.data_as(c_void_p), int(start_pos),
                                              logits.ctypes.data_as(c_void_p) if want_logits else None,
                                              nxt.ctypes.data_as(c_void_p)))
        return nxt, logits

    def last_timing(self) -> dict:
        ms = (c_float * 5)()
        _lib.check(_lib.load().wm_last_timing(self._h, ms))
        return dict(zip(("frontend_ms", "encoder_ms", "cross_kv_ms", "decode_ms", "total_ms"), [float(x) for x in ms]))

    def last_kernel_timing(self, kernel: str = "cross_attention"):
        """(total ms, launches) of one decode-kernel category of the last transcribe call (profile_attn option)."""
        ms, n = c_float(0), c_int64(0)
        _lib.check(_lib.load().wm_last_kernel_timing(self._h, kernel.encode(), ctypes.byref(ms), ctypes.byref(n)))
        return float(ms.value), int(n.value)

    def last_cross_attention_timing(self):
        return self.last_kernel_timing("cross_attention")
