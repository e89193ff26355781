/*
 * whisper_b200.h -- C ABI of the B200-native Whisper transcription path.
 *
 * This is the drop-in boundary for antonvice/whisper.Mojo: the reference has no FFI seam today
 * (everything is Mojo calling Mojo); the seam this library provides is the one its L2 op set
 * (whisper_tensor.mojo) and its model objects (whisper.mojo, layers.mojo, loader.mojo) would bind
 * through Mojo's `external_call` / `DLHandle` (see INTEGRATION.md for the Mojo-side stubs).
 * Plain C: opaque 64-bit handles, raw pointers and sizes, int status returns (0 = WB_OK).
 * No C++ types, no exceptions and no torch types cross this boundary.  There is NO CPU fallback:
 * every compute entry point fails with WB_ERR_CUDA when no sm_100 device / kernel image is usable.
 *
 * Two levels:
 *   wt_*  op level, 1:1 with the reference's L2 routines, fp32 device tensors (Tensor
 *         whisper_tensor.mojo:10-69).  Lets layers.mojo / whisper.mojo keep orchestrating.
 *   wm_*  model level, the batched fast path: weights uploaded once, 16-bit (fp16 by default) tensor-core encoder,
 *         KV-cached batched greedy decode, fused logits+argmax.  `Whisper.transcribe(mel)`
 *         (whisper.mojo:184) == wm_transcribe with n_chunks = 1.
 *
 * Threading: one host thread per model; calls enqueue on the model's stream; functions that return
 * host data synchronise that stream.  Independent models (one per device / process) are safe.
 */
#ifndef WHISPER_B200_H
#define WHISPER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WB_OK 0
#define WB_ERR_ARG 1    /* bad handle / shape / size (the reference does no such checks) */
#define WB_ERR_CUDA 2   /* CUDA runtime / driver failure, or no usable device */
#define WB_ERR_IO 3     /* weight file cannot be opened or has the wrong byte count */
#define WB_ERR_STATE 4  /* call order (e.g. decode before weights are loaded) */

/* Last error text of the calling thread (NUL-terminated, truncated to n). */
int wb_last_error(char *buf, size_t n);
/* ABI version of this header (bumped on any signature change). */
int wb_abi_version(void);
/* Number of this library's kernels launched by the calling process so far (bench evidence). */
int64_t wb_kernel_launch_count(void);
/* 16-bit storage / tensor-core operand type this build of the library computes in: "fp16" (default build,
 * libwhisper_b200.so: IEEE half, 11 significand bits, fp32 accumulation) or "bf16" (libwhisper_b200_bf16.so). */
const char *wb_precision(void);

/* ------------------------------------------------------------------------------------------ */
/* Op level: `Tensor` and the L2 routines                                                     */
/* ------------------------------------------------------------------------------------------ */

typedef uint64_t wt_tensor; /* 0 is the "absent" tensor, i.e. Tensor(0,0) (empty bias / no enc_out) */

/* Tensor(rows, cols): owning, zero-filled, 2-D row-major fp32 on the device.  whisper_tensor.mojo:17-23 */
int wt_tensor_alloc(int64_t rows, int64_t cols, wt_tensor *out);
/* Tensor.view(ptr, rows, cols): non-owning window over `base` starting at element `offset`.  :25-33 */
int wt_tensor_view(wt_tensor base, int64_t offset, int64_t rows, int64_t cols, wt_tensor *out);
/* deinit: frees owned storage (views free only the handle).  :55-57 */
int wt_tensor_free(wt_tensor t);
int wt_tensor_shape(wt_tensor t, int64_t *rows, int64_t *cols);
/* Raw device pointer of the storage (for callers that share CUDA memory, e.g. torch). */
int wt_tensor_data(wt_tensor t, void **dev_ptr);
/* host -> device / device -> host, n fp32 values starting at element `offset` (store/load :59-69). */
int wt_tensor_upload(wt_tensor t, int64_t offset, const float *host, int64_t n);
int wt_tensor_download(wt_tensor t, int64_t offset, float *host, int64_t n);
/* memcpy between tensors (layers.mojo:141-142 KV append, :284-289 head gather). */
int wt_tensor_copy(wt_tensor dst, int64_t dst_off, wt_tensor src, int64_t src_off, int64_t n);

/* matmul(C, A, B, bias): C[M,N] = A[M,K] * B[N,K]^T (+ bias[N]); bias = 0 for none.
 * whisper_tensor.mojo:151-246; also the contract of the MAX wrappers :74-146. */
int wt_matmul(wt_tensor C, wt_tensor A, wt_tensor B, wt_tensor bias);
/* layer_norm(out, inp, gamma, beta, eps): one-pass variance.  :249-285 */
int wt_layer_norm(wt_tensor out, wt_tensor inp, wt_tensor gamma, wt_tensor beta, float eps);
/* gelu(t): in-place tanh GELU with the reference's constants.  :288-308 */
int wt_gelu(wt_tensor t);
/* softmax(t): in-place row softmax.  :311-355 */
int wt_softmax(wt_tensor t);
/* transpose_conv_weights(w, C_out, C_in, K): [C_out, C_in*K] -> new [C_out*K, C_in].  :358-364 */
int wt_transpose_conv_weights(wt_tensor w, int C_out, int C_in, int K, wt_tensor *out);
/* conv1d(out, inp, weight, bias, stride, padding, out_T): K = 3, weight in the transposed layout.  :367-428 */
int wt_conv1d(wt_tensor out, wt_tensor inp, wt_tensor weight, wt_tensor bias, int stride, int padding, int out_T);
/* argmax(t): index of the first maximum.  :431-439 */
int wt_argmax(wt_tensor t, int64_t *idx);
/* out = a + b (residual adds layers.mojo:455-461,482-488,512-517; pos-emb add whisper.mojo:83-89). */
int wt_add(wt_tensor out, wt_tensor a, wt_tensor b);
/* scale + causal fill of attention scores in place: s = s*scale; s[i][j] = -1e10 where j > base+i
 * (layers.mojo:304-320); mask = 0 scales only. */
int wt_scale_mask(wt_tensor scores, float scale, int mask, int64_t base);
/* x[i] = token_emb[tokens[i]] + pos_emb[start_pos + i]   (whisper.mojo:138-149) */
int wt_embed(wt_tensor out, wt_tensor token_emb, wt_tensor pos_emb, const int32_t *tokens, int n, int start_pos);
/* out[c, r] = in[r, c] (V^T build layers.mojo:324-327) */
int wt_transpose(wt_tensor out, wt_tensor in);

/* ------------------------------------------------------------------------------------------ */
/* Model level                                                                                */
/* ------------------------------------------------------------------------------------------ */

/* WhisperConfig (whisper.mojo:15-31) + config.mojo constants + the loop constants of
 * whisper.mojo:187-217, as 16 int32 in this order. */
typedef struct wm_config {
    int32_t d_model, n_heads, n_layers, vocab_size; /* tiny(): 384, 6, 4, 51865 */
    int32_t n_audio_ctx;                            /* 1500 */
    int32_t n_text_ctx;                             /* 448 */
    int32_t n_mels;                                 /* 80 */
    int32_t prompt[4];                              /* 50258, 50259, 50359, 50363 */
    int32_t eot;                                    /* 50257 */
    int32_t max_iters;                              /* 195 */
    int32_t pos_quirk;                              /* 1 = start_pos = current_len - 1 (reference), 0 = HF */
    int32_t reserved[2];
} wm_config;

typedef uint64_t wm_model;
typedef uint64_t wm_cache;

/* Whisper() (whisper.mojo:175-178).  `stream` = a cudaStream_t to enqueue on, or NULL for a
 * library-owned stream on the current device. */
int wm_create(const wm_config *cfg, void *stream, wm_model *out);
/* Fails with WB_ERR_ARG while wm_kvcache handles of the model are alive (destroy those first). */
int wm_destroy(wm_model m);
/* Stream ordering.  Every wm_* call enqueues on the model's stream (the one passed to wm_create, else a
 * library-owned non-blocking stream).  Entry points that return HOST data synchronise it; the *_dev entry points
 * (wm_logmel_dev, wm_encode_dev, wm_kvcache_set_encoder_dev) return with work still queued and READ their inputs on
 * that stream: a caller that produces inputs or consumes outputs on another stream must order the two -- get the
 * stream with wm_stream and use events, or call wm_synchronize (wm_transcribe*_dev synchronise before returning).
 * Host buffers may be pageable or pinned; pinned (cudaMallocHost / cudaHostRegister) buffers make the uploads of
 * wm_transcribe / wm_transcribe_pcm asynchronous DMA that fully overlaps the compute, pageable ones are staged by
 * the driver one encoder sub-batch at a time (still overlapped with the previous sub-batch's compute, but the
 * PCIe copy then bounds the call; see DESIGN.md section 6 for both numbers). */
int wm_stream(wm_model m, void **stream);
int wm_synchronize(wm_model m);
/* Number of fp32 values the flat weight file must hold for this config (export_weights.py:19-90). */
int64_t wm_weight_count(const wm_config *cfg);
/* Whisper.load(WeightLoader(path)) (loader.mojo:10-27, whisper.mojo:180-182): validates the byte
 * count, uploads once, builds the device-side 16-bit / fused copies. */
int wm_load_weights_file(wm_model m, const char *path);
/* Same from host memory (n_floats fp32 values in file order). */
int wm_load_weights(wm_model m, const float *host, int64_t n_floats);
/* Device pointer + element count of one tensor of the uploaded fp32 weight image by file-order
 * index (0 .. n_tensors-1); what WeightLoader.next_tensor hands out. */
int wm_weight_tensor(wm_model m, int index, void **dev_ptr, int64_t *n_floats);

/* Options: "gemm_impl" 0 = CUDA-core bring-up kernel, 1 = tcgen05/TMA kernels (default: CTA-pair kernel for the
 * large GEMMs, single-CTA kernel otherwise), 2 / 3 = force the single-CTA / CTA-pair kernel; "attn_impl" 0 = CUDA-core
 * encoder attention, 1 = tcgen05 flash attention (default); "frontend_impl" 0 = fp32 FMA DFT, 1 = TF32x3 tensor-core
 * DFT (default); "cross_impl" 0 = per-layer cross K/V cache (reference form), 1 = absorbed form over enc_out
 * (default when d_model <= 768 and <= 16 heads); "use_graph", "decode_lanes", "enc_batch", "wave_max", "profile_attn" (1 = time the
 * cross-attention launches, 2 = every decode kernel by category), "pdl" (programmatic dependent launch, default 0), "small_batch" (waves of at most this many chunks use the
 * latency-oriented decode: K/V-form cross-attention split over the SMs + programmatic dependent launch; default 0 = off,
 * 8 is a good value for batch-1 use), "decode_split_k" (split-K residual GEMMs + fused residual/LayerNorm in
 * the decode step: 0 = off, 1 = on (default); never a function of the batch size, so a chunk's ids do not depend
 * on how many chunks share its wave), "decode_fused" (1 = persistent chain kernels, 4 L + 3 launches per step;
 * 0 = one kernel per op, 12 L + 4; 2 (default) = chain kernels for waves of <= 1280 chunks, where the step is latency
 * bound, kernel per op above -- the two forms produce the same bits), "prefill_impl" (1 (default) = the 4 prompt ids run
 * as one q_len = 4 forward with the causal block path, whisper.mojo:195-197; 0 = fed one by one through the cached step:
 * same ids, 3 more forwards), "skip_done" (see wm_set_stop_lengths). */
int wm_set_option(wm_model m, const char *key, int64_t value);

/* Log-mel frontend (HF WhisperFeatureExtractor via export_weights.py:116): pcm f32 [n_chunks,
 * n_samples] (16 kHz; shorter audio must be zero-padded by the caller) -> mel f32
 * [n_chunks, n_mels, 2*n_audio_ctx].  *_dev variants take device pointers. */
int wm_logmel(wm_model m, const float *pcm_host, int n_chunks, float *mel_host);
int wm_logmel_dev(wm_model m, const float *pcm_dev, int n_chunks, float *mel_dev);

/* WhisperEncoder.forward (whisper.mojo:71-99), batched: mel f32 [n_chunks, n_mels, 2*ctx] ->
 * enc_out f32 [n_chunks, ctx, d_model]. */
int wm_encode(wm_model m, const float *mel_host, int n_chunks, float *enc_out_host);
int wm_encode_dev(wm_model m, const float *mel_dev, int n_chunks, float *enc_out_dev);

/* KVCache(n_layers, d_model, max_len) for `n_chunks` sequences (layers.mojo:55-63). */
int wm_kvcache_create(wm_model m, int n_chunks, int max_len, wm_cache *out);
int wm_kvcache_destroy(wm_cache c);
int wm_kvcache_reset(wm_cache c);
int wm_kvcache_len(wm_cache c, int *current_len);
/* Fill the cross-attention K/V of every layer from enc_out (layers.mojo:148-157; the reference
 * does this lazily inside the first decoder forward).  enc_out_dev f32 [n_chunks, ctx, d_model]. */
int wm_kvcache_set_encoder_dev(wm_model m, wm_cache c, const float *enc_out_dev);

/* WhisperDecoder.forward(tokens, enc_out, cache, use_cache=True, start_pos) for q_len = 1
 * (whisper.mojo:130-167), batched over the cache's chunks: tokens int32 [n_chunks] (host);
 * appends K/V at current_len, returns next = argmax(logits) per chunk (host int32 [n_chunks],
 * may be NULL) and optionally the logits (host f32 [n_chunks, vocab], may be NULL). */
int wm_decode_step(wm_model m, wm_cache c, const int32_t *tokens_host, int start_pos, float *logits_host,
                   int32_t *next_host);

/* Whisper.transcribe (whisper.mojo:184-223), batched: mel f32 [n_chunks, n_mels, 2*ctx] ->
 * out_tokens int32 [n_chunks, 5 + max_iters] (prompt + generated, EOT included when produced,
 * unused tail = -1) and out_len int32 [n_chunks]. */
int wm_transcribe(wm_model m, const float *mel_host, int n_chunks, int32_t *out_tokens_host, int32_t *out_len_host);
int wm_transcribe_dev(wm_model m, const float *mel_dev, int n_chunks, int32_t *out_tokens_dev, int32_t *out_len_dev);
/* Same with the frontend in front: pcm f32 [n_chunks, n_samples] -> tokens. */
int wm_transcribe_pcm(wm_model m, const float *pcm_host, int n_chunks, int32_t *out_tokens_host, int32_t *out_len_host);
int wm_transcribe_pcm_dev(wm_model m, const float *pcm_dev, int n_chunks, int32_t *out_tokens_dev,
                          int32_t *out_len_dev);

/* Test / bench hook for `if next_token == 50257: break` (whisper.mojo:206-207).  With random weights EOT never wins,
 * so chunk lengths can be DECLARED: for every later wm_transcribe* call the id that would bring chunk i's count to
 * lens_host[i] (>= 5: 4 prompt ids + EOT) is replaced by EOT and the chunk finishes there.  n = 0 clears the schedule;
 * chunks beyond n are unaffected.  Finished chunks drop out of the attention kernels (option "skip_done", default 1;
 * 0 keeps streaming them until the whole wave is done, round 1's behaviour) -- ids are identical either way. */
int wm_set_stop_lengths(wm_model m, const int32_t *lens_host, int n_chunks);

/* Teacher-forced decode for parity tests: forced int32 [n_chunks, n_forced] (host); prefill with
 * forced[:, 0:4], then feed forced[:, 4:] with the greedy loop's start_pos rule; logits_host f32
 * [n_chunks, n_forced - 3, vocab].  Encoder output of earlier calls is NOT reused:
 * enc_out_dev f32 [n_chunks, ctx, d_model] must be passed. */
int wm_teacher_forced(wm_model m, const float *enc_out_dev, int n_chunks, const int32_t *forced_host, int n_forced,
                      float *logits_host);

/* Phase timers of the last wm_transcribe* call, CUDA-event milliseconds on the model's stream:
 * [0] frontend, [1] encoder, [2] cross-KV projection, [3] prefill + decode loop, [4] total. */
int wm_last_timing(wm_model m, float ms[5]);
/* Device-time of decode-step kernels accumulated over the last wm_transcribe* call (CUDA events around each
 * launch; roofline evidence for bench.py): total milliseconds and launch count of category `kernel` --
 * "cross_attention" (option profile_attn >= 1), and with profile_attn = 2 also "self_attention", "gemm_qkv",
 * "gemm_o", "gemm_cross_q", "gemm_cross_o", "gemm_fc1", "gemm_fc2", "layer_norm", "gemm_logits", "misc". */
int wm_last_kernel_timing(wm_model m, const char *kernel, float *total_ms, int64_t *launches);

/* ------------------------------------------------------------------------------------------ */
/* Test hook (used by tests/ only): run one GEMM of the fast path on host fp32 data.             */
/*   A(b, m, tap*Cin + ci) = A_host[b][(m*conv_stride + tap - pad)][ci]  (zero outside [0, src_rows)) */
/*   out f32 [batches*rows_per_batch][N]; epi: 0 store(16-bit-rounded) 1 gelu(16-bit-rounded)        */
/*   2 out += (out pre-filled by the caller) 3 store f32 4 argmax (logits in out, the row's argmax    */
/*   index replaces the LAST column).  impl: 0 CUDA-core reference kernel, 1 tcgen05/TMA kernel.                */
/* ------------------------------------------------------------------------------------------ */
int wb_debug_gemm(int impl, const float *A_host, int batches, int src_rows, int lda, int Cin, int taps,
                  int conv_stride, int pad, int rows_per_batch, const float *W_host, int N, const float *bias_host,
                  int epi, float *out_host);

/* Test hook: single-query attention of one decode step (layers.mojo:186-272) on host fp32 data
 * (rounded to the 16-bit type on the device): q [B][D], K/V [B][len][D] -> out [B][D]; D = H*64. */
int wb_debug_decode_attention(const float *q_host, const float *K_host, const float *V_host, int B, int H, int len,
                              int splits, float *out_host);

/* Test hook: encoder self-attention (layers.mojo:273-342, no mask) on host fp32 data rounded to the 16-bit type:
 * qkv [B*S][3*D] (q | k | v) -> out [B*S][D].  impl: 0 CUDA-core kernel, 1 tcgen05 flash-attention kernel. */
int wb_debug_encoder_attention(int impl, const float *qkv_host, int B, int S, int H, float *out_host);

/* Test hook: absorbed cross-attention (cross_attn_tc.cu) on host fp32 data rounded to the 16-bit type:
 * qp [B][H*D] (scores are used in base 2: p = 2^(s - max)), enc [B][S][D] -> ctx [B][H*D]. */
int wb_debug_cross_attention_absorbed(const float *qp_host, const float *enc_host, int B, int S, int D, int H,
                                      float *ctx_host);

#ifdef __cplusplus
}
#endif
#endif /* WHISPER_B200_H */
