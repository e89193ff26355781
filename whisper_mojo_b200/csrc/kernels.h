// kernels.h -- launchers for the non-GEMM kernels of the batched fast path (kernels.cu, frontend.cu).
#pragma once
#include "dtype.h"
#include <cuda_runtime.h>
#include <stdint.h>

namespace wb {

// LayerNorm (whisper_tensor.mojo:249-285, one-pass variance) fp32 in -> bf16 out (+ optional fp32 out).
int ln_h16(cudaStream_t st, const float *x, const float *gamma, const float *beta, int rows, int D,
            h16 *out_h16, float *out_f32);

// Second half of a split-K residual GEMM fused with the LayerNorm that follows it (layers.mojo:456-461,486-491,
// 515-517 + the next block's LayerNorm): x[row] += bias + part[0][row] + ... + part[n_split-1][row] in that
// fixed order, then out_h16[row] = LN(x[row]) unless gamma is null.  part is [n_split][rows][D] fp32.
int resid_ln(cudaStream_t st, float *x, const float *part, int n_split, const float *bias, const float *gamma,
             const float *beta, int rows, int D, h16 *out_h16);

// Decoder input (whisper.mojo:138-149) fused with the first LayerNorm of layer 0:
//   x[b] = token_emb[cur_tok[b]] + pos_emb[*pos];  xn[b] = LN(x[b]) in bf16.
int embed_ln(cudaStream_t st, const float *tok_emb, const float *pos_emb, const int *cur_tok, const int *pos_dev,
             int B, int D, int vocab, int n_pos, const float *gamma, const float *beta, float *x,
             h16 *xn);

// Prefill (whisper.mojo:195-197): rows b * q_len + p = token_emb[prompt[p]] + pos_emb[p] for every chunk b, then the
// same LayerNorm; *set_len = q_len - 1 (cur_len once the bookkeeping kernel after the logits has added the last 1).
int embed_ln_prefill(cudaStream_t st, const float *tok_emb, const float *pos_emb, const int *prompt4_host, int q_len, int B,
                     int D, int vocab, int n_pos, const float *gamma, const float *beta, float *x, h16 *xn, int *set_len);
// k / v [B * q_len][D] of a q_len-token forward -> rows 0 .. q_len-1 of the chunks' self K/V cache (chunk stride
// kv_batch_stride elements).
int kv_scatter(cudaStream_t st, const h16 *k, const h16 *v, h16 *Kc, h16 *Vc, int B, int q_len, int D, int64_t kv_batch_stride);

// Single-query attention for one decode step (layers.mojo:186-272), batched over chunks and heads.
//   q   bf16 [B][D]                         (head h = columns h*64 .. h*64+63)
//   K,V bf16 [B][rows][D], chunk stride `kv_batch_stride` elements
//   len = len_const, or *len_dev + len_add when len_dev != nullptr
//   out bf16 [B][D]
// `ws` is scratch for the split-K partials: floats [B][splits][H][66].
struct DecodeAttnArgs {
    const h16 *q, *K, *V;
    h16 *out;
    int64_t kv_batch_stride;
    int B, H, D;
    int len_const;
    const int *len_dev;
    int len_add;
    int max_len;  // upper bound of len (sizes shared memory)
    int splits;
    float *ws;
    const int *done = nullptr;  // optional [B]: chunks whose flag is set are skipped (their output row keeps its old value)
    // q_len > 1 (prefill, layers.mojo:273-342 with the causal fill :304-320): B counts query ROWS, row vb = query
    // vb % q_len of chunk vb / q_len; with len_const == 0 and no len_dev it attends over keys 0 .. vb % q_len.
    int q_len = 1;
};
int decode_attention(cudaStream_t st, const DecodeAttnArgs &a);
int decode_attention_splits(int B, int len, int H);

// Encoder self-attention, bring-up implementation on CUDA cores (layers.mojo:273-342, no mask):
// qkv bf16 [B*S][3D] -> out bf16 [B*S][D].
int encoder_attention_ref(cudaStream_t st, const h16 *qkv, h16 *out, int B, int S, int H, int D);
// Same contract on tcgen05 tensor cores, flash-attention style (attn_tc.cu).
int encoder_attention_tc(cudaStream_t st, const h16 *qkv, h16 *out, int B, int S, int H, int D);

// Decode cross-attention over enc_out itself (cross_attn_tc.cu): q' bf16 [B][H*D], enc bf16 [B][S][D]
// -> ctx bf16 [B][H*D].  Needs the folded weights below.
// live / n_live (device, optional): attend only for the chunks live[0 .. *n_live) (finished chunks are skipped).
// q_rows > 1 (prefill): q' and ctx hold q_rows rows per chunk, [B][q_rows][H*D]; every row attends over the chunk's enc.
int cross_attention_absorbed(cudaStream_t st, const h16 *qp, const h16 *enc, h16 *ctx,
                             int B, int S, int D, int H, const int *live = nullptr, const int *n_live = nullptr, int q_rows = 1);
bool cross_attn_absorbed_supported(int D, int H);
extern unsigned long long *g_xa_dbg;  // development aid: timestamp buffer for CTA 0 (normally null)
// Load-time folding (fp32 math, bf16 result):
//   Wqk[h*D + c][i] = (log2(e)/8) * sum_d Wk[h*64+d][c] * Wq[h*64+d][i],  bqk[h*D + c] = (log2(e)/8) * sum_d Wk[h*64+d][c] * bq[h*64+d]
//   Wov[n][h*D + c] = sum_d Wo[n][h*64+d] * Wv[h*64+d][c],                  bov[n] = bo[n] + sum_j Wo[n][j] * bv[j]
int fold_cross_weights(cudaStream_t st, const float *Wq, const float *bq, const float *Wk, const float *Wv,
                       const float *bv, const float *Wo, const float *bo, int D, int H, h16 *Wqk,
                       float *bqk, h16 *Wov, float *bov);

// Greedy bookkeeping after a logits step (whisper.mojo:198-221): append next token unless the chunk
// has finished, mark EOT, set the next input token, advance cur_len / pos.
struct GreedyState {
    int *tokens_out;  // [B][T_out], -1 filled
    int *out_len;     // [B]
    int *cur_tok;     // [B]
    int *done;        // [B]
    int *scalars;     // [0]=cur_len, [1]=pos, [2]=n_done, [3]=n_live
    int T_out, eot, pos_quirk;
    int *live = nullptr;     // [B] indices of the chunks still decoding, in order (rebuilt by greedy_rebuild_live)
    int *stop_at = nullptr;  // [B] forced length: the id that brings a chunk's count to stop_at[b] is replaced by EOT
                             // (INT_MAX = never; a test / bench hook, since EOT never wins with random weights)
};
int greedy_init(cudaStream_t st, const GreedyState &g, int B, const int *prompt4_host);
// live[] = indices b with done[b] == 0 in increasing order, scalars[3] = their count (one CTA: B is at most a few thousand).
int greedy_rebuild_live(cudaStream_t st, const GreedyState &g, int B);
// mode 0: prefill advance (feed prompt[next_prompt_idx] next); mode 1: greedy append from `next`.
int greedy_advance(cudaStream_t st, const GreedyState &g, int B, int mode, int next_prompt_token, const int *next);
// argmax over the EPI_ARGMAX partials of the logits GEMM + the mode-1 bookkeeping above, in one kernel.
int greedy_argmax_advance(cudaStream_t st, const GreedyState &g, int B, const float *part_val, const int *part_idx,
                          int tiles_n, int *next);

// ---- frontend (frontend.cu) ------------------------------------------------------------------
struct FrontendTables {
    float *tw_cos = nullptr, *tw_sin = nullptr;  // [200][104] twiddles cos/sin(2*pi*k*n/400), k < 101
    float *window = nullptr;                     // [400] periodic Hann
    float *mel_w = nullptr;                      // sparse filterbank weights
    int *mel_start = nullptr, *mel_len = nullptr, *mel_off = nullptr;  // per mel bin
    int n_mels = 0;
    // tensor-core frontend (frontend_tc.cu): TF32 hi / lo twiddles [416][416], per-DFT-bin filter view, workspace
    bool tc_ok = false;
    float *tc_b_hi = nullptr, *tc_b_lo = nullptr, *tc_bin_wa = nullptr, *tc_bin_wb = nullptr;
    int *tc_bin_j = nullptr;
    float *tc_ws = nullptr;  // reflect-padded hi / lo copies of the pcm, grow-only
    size_t tc_ws_cap = 0;
};
int frontend_tables_create(FrontendTables *t, int n_mels);
void frontend_tables_destroy(FrontendTables *t);
// pcm f32 [B][n_frames*160] -> raw log10 mel f32 [B][n_mels][n_frames] and per-chunk max (ordered-int encoded).
// impl: 0 = fp32 FMA DFT (frontend.cu), 1 = TF32x3 tensor-core DFT (frontend_tc.cu) when the tables allow it.
int logmel_raw(cudaStream_t st, FrontendTables &t, const float *pcm, int B, int n_frames, float *mel_raw,
               int *chunk_max_enc, int impl = 1);
int logmel_raw_tc(cudaStream_t st, FrontendTables &t, const float *pcm, int B, int n_frames, float *mel_raw,
                  int *chunk_max_enc);
// Finalise in place: v = (max(v, chunk_max - 8) + 4) / 4.
int logmel_finalize(cudaStream_t st, float *mel, const int *chunk_max_enc, int B, int n_mels, int n_frames);
// mel f32 [B][n_mels][n_frames] -> bf16 [B][n_frames][128] (channels >= n_mels zero) for the conv1 GEMM.
int mel_to_h16_T(cudaStream_t st, const float *mel, h16 *out, int B, int n_mels, int n_frames);

// fp32 -> bf16 conversion helpers used at weight load.
int convert_f32_h16(cudaStream_t st, const float *src, h16 *dst, size_t n);
// conv weight [C_out][C_in][3] fp32 -> bf16 [C_out][3*C_in_pad] in the (tap, ci) order of
// transpose_conv_weights (whisper_tensor.mojo:358-364), channels padded with zeros to C_in_pad.
int convert_conv_weight(cudaStream_t st, const float *w, h16 *dst, int C_out, int C_in, int C_in_pad);

}  // namespace wb
