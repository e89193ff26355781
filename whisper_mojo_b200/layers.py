"""Mirror of layers.mojo on the op-level C ABI: LayerCache, KVCache, MultiHeadAttention,
ResidualAttentionBlock.  Same structure and call order as the reference, every numeric step a
wt_* kernel in fp32 -- this is the "Mojo host code keeps orchestrating" path of the drop-in
boundary, and a near-exact fp32 GPU twin of the CPU oracle.  The batched bf16 fast path lives in
whisper.py (`Whisper.transcribe`).
"""
from __future__ import annotations

import math
from typing import List

from . import whisper_tensor as wt
from .loader import WeightLoader
from .whisper_tensor import Tensor


class LayerCache:
    """layers.mojo:14-52"""

    def __init__(self):
        self.self_k = Tensor(0, 0)
        self.self_v = Tensor(0, 0)
        self.cross_k = Tensor(0, 0)
        self.cross_v = Tensor(0, 0)
        self.current_len = 0
        self.has_cross = False

    def reset(self, d_model: int, max_len: int, n_audio_ctx: int = 1500):
        self.self_k = Tensor(max_len, d_model)
        self.self_v = Tensor(max_len, d_model)
        self.cross_k = Tensor(n_audio_ctx, d_model)  # 1500 hard-coded in layers.mojo:33-34
        self.cross_v = Tensor(n_audio_ctx, d_model)
        self.current_len = 0
        self.has_cross = False


class KVCache:
    """layers.mojo:55-69"""

    def __init__(self, n_layers: int, d_model: int, max_len: int, n_audio_ctx: int = 1500):
        self.layers: List[LayerCache] = []
        for _ in range(n_layers):
            layer = LayerCache()
            layer.reset(d_model, max_len, n_audio_ctx)
            self.layers.append(layer)


class MultiHeadAttention:
    """layers.mojo:72-359"""

    def __init__(self, d_model: int, n_heads: int):
        self.d_model, self.n_heads, self.head_dim = d_model, n_heads, d_model // n_heads
        self.q_proj_w = self.q_proj_b = self.k_proj_w = self.v_proj_w = None
        self.v_proj_b = self.out_proj_w = self.out_proj_b = None

    def load(self, loader: WeightLoader, is_self_attn: bool):
        d = self.d_model
        self.q_proj_w = loader.next_tensor(d, d)
        self.q_proj_b = loader.next_tensor(1, d)
        self.k_proj_w = loader.next_tensor(d, d)
        self.v_proj_w = loader.next_tensor(d, d)
        self.v_proj_b = loader.next_tensor(1, d)
        self.out_proj_w = loader.next_tensor(d, d)
        self.out_proj_b = loader.next_tensor(1, d)

    def forward(self, query: Tensor, key: Tensor, value: Tensor, mask: bool, cache: LayerCache,
                is_self_attn: bool, use_cache: bool) -> Tensor:
        d, hd = self.d_model, self.head_dim
        q_len, k_len = query.rows, key.rows
        empty = Tensor(0, 0)
        q = Tensor(q_len, d)
        wt.matmul(q, query, self.q_proj_w, self.q_proj_b)
        if use_cache:
            if is_self_attn:  # :131-147
                new_k = Tensor(q_len, d)
                wt.matmul(new_k, key, self.k_proj_w, empty)
                new_v = Tensor(q_len, d)
                wt.matmul(new_v, value, self.v_proj_w, self.v_proj_b)
                dest = cache.current_len * d
                wt.memcpy(cache.self_k, dest, new_k, 0, q_len * d)
                wt.memcpy(cache.self_v, dest, new_v, 0, q_len * d)
                cache.current_len += q_len
                k = Tensor.view(cache.self_k, cache.current_len, d)
                v = Tensor.view(cache.self_v, cache.current_len, d)
            else:  # :148-157
                if not cache.has_cross:
                    wt.matmul(cache.cross_k, key, self.k_proj_w, empty)
                    wt.matmul(cache.cross_v, value, self.v_proj_w, self.v_proj_b)
                    cache.has_cross = True
                k = Tensor.view(cache.cross_k, k_len, d)
                v = Tensor.view(cache.cross_v, k_len, d)
        else:  # :158-176
            k = Tensor(k_len, d)
            wt.matmul(k, key, self.k_proj_w, empty)
            v = Tensor(k_len, d)
            wt.matmul(v, value, self.v_proj_w, self.v_proj_b)
        final_k_len = k.rows
        out = Tensor(q_len, d)
        scale = 1.0 / math.sqrt(float(hd))
        # The reference has a register-resident path for q_len == 1 (:186-272) and a matmul-based
        # path otherwise (:273-342); they compute the same softmax(q k^T * scale) v, and on the GPU
        # one op sequence serves both.  Mask threshold: j > current_len - q_len + i (:311) which for
        # q_len == 1 is j > current_len - 1, i.e. never (:213).
        base = (cache.current_len - q_len) if (use_cache and is_self_attn) else 0
        kT_free = Tensor(final_k_len, hd)
        for h in range(self.n_heads):
            q_h = Tensor(q_len, hd)
            k_h = kT_free
            v_h = Tensor(final_k_len, hd)
            for i in range(q_len):  # :280-291 head gather
                wt.memcpy(q_h, i * hd, q, i * d + h * hd, hd)
            _gather_head(k_h, k, h, hd, d)
            _gather_head(v_h, v, h, hd, d)
            scores = Tensor(q_len, final_k_len)
            wt.matmul(scores, q_h, k_h, empty)
            wt.scale_mask(scores, scale, mask, base)  # :304-320
            wt.softmax(scores)
            v_h_T = Tensor(hd, final_k_len)
            wt.transpose(v_h_T, v_h)  # :324-327
            out_h = Tensor(q_len, hd)
            wt.matmul(out_h, scores, v_h_T, empty)
            for i in range(q_len):  # :338-342 scatter
                wt.memcpy(out, i * d + h * hd, out_h, i * hd, hd)
        final_out = Tensor(q_len, d)
        wt.matmul(final_out, out, self.out_proj_w, self.out_proj_b)
        return final_out


def _gather_head(dst: Tensor, src: Tensor, h: int, hd: int, d: int) -> None:
    """rows of `src` [n, d] restricted to head h -> dst [n, hd]: one strided copy expressed as
    transpose(view) so it stays a single kernel for n = 1500."""
    n = src.rows
    # view src as [n, d] -> transpose to [d, n] -> rows h*hd .. (h+1)*hd are contiguous -> transpose back
    srcT = Tensor(d, n)
    wt.transpose(srcT, Tensor.view(src, n, d))
    wt.transpose(dst, Tensor.view(srcT, hd, n, offset=h * hd * n))


class ResidualAttentionBlock:
    """layers.mojo:386-519"""

    def __init__(self, d_model: int, n_heads: int, is_decoder: bool):
        self.d_model, self.is_decoder = d_model, is_decoder
        self.attn = MultiHeadAttention(d_model, n_heads)
        self.cross_attn = MultiHeadAttention(d_model, n_heads)
        self.attn_ln_w = self.attn_ln_b = self.cross_attn_ln_w = self.cross_attn_ln_b = None
        self.mlp_fc1_w = self.mlp_fc1_b = self.mlp_fc2_w = self.mlp_fc2_b = self.mlp_ln_w = self.mlp_ln_b = None

    def load(self, loader: WeightLoader, is_decoder_block: bool):
        d = self.d_model
        self.attn.load(loader, is_self_attn=True)
        self.attn_ln_w = loader.next_tensor(1, d)
        self.attn_ln_b = loader.next_tensor(1, d)
        if is_decoder_block:
            self.cross_attn.load(loader, is_self_attn=False)
            self.cross_attn_ln_w = loader.next_tensor(1, d)
            self.cross_attn_ln_b = loader.next_tensor(1, d)
        self.mlp_fc1_w = loader.next_tensor(d * 4, d)
        self.mlp_fc1_b = loader.next_tensor(1, d * 4)
        self.mlp_fc2_w = loader.next_tensor(d, d * 4)
        self.mlp_fc2_b = loader.next_tensor(1, d)
        self.mlp_ln_w = loader.next_tensor(1, d)
        self.mlp_ln_b = loader.next_tensor(1, d)

    def forward(self, x: Tensor, enc_out: Tensor, cache: LayerCache, use_cache: bool) -> Tensor:
        x_norm = Tensor(x.rows, x.cols)
        wt.layer_norm(x_norm, x, self.attn_ln_w, self.attn_ln_b)
        self_attn_out = self.attn.forward(x_norm, x_norm, x_norm, mask=self.is_decoder, cache=cache,
                                          is_self_attn=True, use_cache=use_cache)
        current_x = Tensor(x.rows, x.cols)
        wt.add(current_x, x, self_attn_out)
        if self.is_decoder and enc_out.size > 0:
            x_norm_cross = Tensor(current_x.rows, current_x.cols)
            wt.layer_norm(x_norm_cross, current_x, self.cross_attn_ln_w, self.cross_attn_ln_b)
            cross = self.cross_attn.forward(x_norm_cross, enc_out, enc_out, mask=False, cache=cache,
                                            is_self_attn=False, use_cache=use_cache)
            x_res2 = Tensor(current_x.rows, current_x.cols)
            wt.add(x_res2, current_x, cross)
            current_x = x_res2
        x_norm_mlp = Tensor(current_x.rows, current_x.cols)
        wt.layer_norm(x_norm_mlp, current_x, self.mlp_ln_w, self.mlp_ln_b)
        hidden = Tensor(current_x.rows, self.d_model * 4)
        wt.matmul(hidden, x_norm_mlp, self.mlp_fc1_w, self.mlp_fc1_b)
        wt.gelu(hidden)
        mlp_out = Tensor(current_x.rows, current_x.cols)
        wt.matmul(mlp_out, hidden, self.mlp_fc2_w, self.mlp_fc2_b)
        final_out = Tensor(current_x.rows, current_x.cols)
        wt.add(final_out, current_x, mlp_out)
        return final_out
