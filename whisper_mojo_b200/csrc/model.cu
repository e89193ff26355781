// model.cu -- batched Whisper fast path: weight upload, encoder, cross-K/V, KV-cached greedy decode.
//
// Mirrors whisper.mojo (Whisper / WhisperEncoder / WhisperDecoder), layers.mojo
// (ResidualAttentionBlock, MultiHeadAttention, KVCache) and loader.mojo, restructured for a batch of
// independent 30 s chunks resident in HBM:
//   * weights: the flat fp32 file image is uploaded once; bf16 copies of every matrix are made on
//     the device (q/k/v fused into one [3D, D] matrix, the L cross-attention K/V projections fused
//     into one [L*2*D, D] matrix, conv weights in the (tap, channel) order of transpose_conv_weights).
//   * activations: fp32 residual stream, bf16 GEMM operands, fp32 accumulation.
//   * KV cache: bf16, [layer][k|v][chunk][position][D] -- a chunk's K (or V) rows are contiguous so
//     the decode attention kernels stream whole rows.
//   * decode: one step = ~48 small kernels; cur_len / position / tokens / EOT flags live in device
//     memory so a single CUDA graph of the step is replayed for the whole greedy loop.
#include "model.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "common.cuh"
#include "gemm.h"

namespace wb {

// ---------------------------------------------------------------------------------------------
// layout of the flat weight file (export_weights.py:19-90; read order whisper.mojo:60-69,122-128)
// ---------------------------------------------------------------------------------------------
Layout make_layout(const wm_config &c) {
    Layout l;
    int64_t off = 0;
    const int64_t D = c.d_model, F = 4 * D;
    auto take = [&](int64_t n) {
        int64_t o = off;
        l.tensors.push_back({o, n});
        off += n;
        return o;
    };
    auto attn = [&](AttnW &a) {
        a.q_w = take(D * D), a.q_b = take(D), a.k_w = take(D * D), a.v_w = take(D * D), a.v_b = take(D);
        a.o_w = take(D * D), a.o_b = take(D);
    };
    auto block = [&](BlockW &b, bool dec) {
        attn(b.attn);
        b.attn_ln_w = take(D), b.attn_ln_b = take(D);
        if (dec) {
            attn(b.cross);
            b.cross_ln_w = take(D), b.cross_ln_b = take(D);
        }
        b.fc1_w = take(F * D), b.fc1_b = take(F), b.fc2_w = take(D * F), b.fc2_b = take(D);
        b.mlp_ln_w = take(D), b.mlp_ln_b = take(D);
    };
    l.conv1_w = take(D * c.n_mels * 3), l.conv1_b = take(D);
    l.conv2_w = take(D * D * 3), l.conv2_b = take(D);
    l.enc_pos = take((int64_t)c.n_audio_ctx * D);
    l.enc.resize(c.n_layers);
    for (auto &b : l.enc) block(b, false);
    l.enc_ln_w = take(D), l.enc_ln_b = take(D);
    l.tok_emb = take((int64_t)c.vocab_size * D);
    l.dec_pos = take((int64_t)c.n_text_ctx * D);
    l.dec.resize(c.n_layers);
    for (auto &b : l.dec) block(b, true);
    l.dec_ln_w = take(D), l.dec_ln_b = take(D);
    l.total = off;
    return l;
}

template <typename T>
static int dalloc(Model *m, T **p, size_t n) {
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T));
    if (e != cudaSuccess) {
        set_error("cudaMalloc(%zu bytes) -> %s", n * sizeof(T), cudaGetErrorString(e));
        return WB_ERR_CUDA;
    }
    *p = reinterpret_cast<T *>(q);
    if (m) m->owned.push_back(q);
    return WB_OK;
}

int model_create(const wm_config *cfg, void *stream, Model **out) {
    WB_ARG(cfg && out, "wm_create: null argument");
    WB_ARG(cfg->d_model > 0 && cfg->n_heads > 0 && cfg->d_model == cfg->n_heads * 64,
           "wm_create: head_dim must be 64 (d_model=%d n_heads=%d)", cfg->d_model, cfg->n_heads);
    WB_ARG(cfg->d_model % 128 == 0 && cfg->d_model <= 1024, "wm_create: d_model must be a multiple of 128, <= 1024");
    WB_ARG(cfg->n_mels > 0 && cfg->n_mels <= 128 && cfg->n_layers > 0 && cfg->vocab_size > 0 &&
               cfg->n_audio_ctx > 0 && cfg->n_text_ctx >= 8 && cfg->max_iters >= 0,
           "wm_create: bad config");
    // the greedy loop writes K/V rows 0 .. 3 + max_iters and embeds positions up to 4 + max_iters (whisper.mojo:193
    // sizes the reference's cache at 448 for 4 + 195 rows): a longer loop would run past the cache
    WB_ARG(cfg->max_iters + 5 <= cfg->n_text_ctx, "wm_create: max_iters %d + 5 exceeds n_text_ctx %d", cfg->max_iters,
           cfg->n_text_ctx);
    // decode self-attention runs one warp per head in one CTA (kernels.cu: decode_attn_kernel, 12 warps)
    WB_ARG(cfg->n_heads <= 12, "wm_create: n_heads %d > 12 is not supported by the decode attention kernel", cfg->n_heads);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("wm_create: no CUDA device (this library has no CPU fallback)");
        return WB_ERR_CUDA;
    }
    Model *m = new Model();
    if (const char *e = getenv("WB_PDL")) m->pdl = atoi(e) != 0;  // A/B switch for programmatic dependent launch
    cudaGetDevice(&m->device);
    m->cfg = *cfg;
    m->D = cfg->d_model, m->H = cfg->n_heads, m->L = cfg->n_layers, m->V = cfg->vocab_size;
    m->S = cfg->n_audio_ctx, m->T = cfg->n_text_ctx, m->NM = cfg->n_mels, m->F = 4 * cfg->d_model;
    m->n_frames = 2 * m->S, m->n_samples = m->n_frames * 160;
    m->lay = make_layout(*cfg);
    if (!cross_attn_absorbed_supported(m->D, m->H)) m->cross_impl = 0;
    if (stream) {
        m->stream = reinterpret_cast<cudaStream_t>(stream);
    } else {
        cudaError_t e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            set_error("cudaStreamCreate -> %s", cudaGetErrorString(e));
            delete m;
            return WB_ERR_CUDA;
        }
        m->own_stream = true;
    }
    if (cudaStreamCreateWithFlags(&m->stream2, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&m->ev_join, cudaEventDisableTiming) != cudaSuccess) {
        set_error("wm_create: stream / event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
        model_destroy(m);
        return WB_ERR_CUDA;
    }
    chain_debug_init();
    int rc = frontend_tables_create(&m->ft, m->NM);
    if (rc != WB_OK) {
        model_destroy(m);
        return rc;
    }
    *out = m;
    return WB_OK;
}

void model_destroy(Model *m) {
    if (!m) return;
    cudaStreamSynchronize(m->stream);
    if (getenv("WB_CHAIN_DBG")) chain_debug_dump();
    if (m->stream2) cudaStreamSynchronize(m->stream2);
    if (m->tr_cache) cache_destroy(m->tr_cache);
    cudaFree(m->tr_mel);
    cudaFree(m->stop_sched);
    cudaFree(m->stage_in);
    cudaFree(m->stage_out);
    for (void *p : m->owned) cudaFree(p);
    for (cudaEvent_t e : m->cross_timer.ev) cudaEventDestroy(e);
    for (cudaEvent_t e : m->ev_plain) cudaEventDestroy(e);
    for (cudaEvent_t e : m->ev_timed) cudaEventDestroy(e);
    frontend_tables_destroy(&m->ft);
    if (m->stream2) cudaStreamDestroy(m->stream2);
    if (m->ev_fork) cudaEventDestroy(m->ev_fork);
    if (m->ev_join) cudaEventDestroy(m->ev_join);
    if (m->own_stream) cudaStreamDestroy(m->stream);
    delete m;
}

// Whisper.load (whisper.mojo:180-182) + WeightLoader (loader.mojo): one upload, device-side repack.
int model_load(Model *m, const float *host, int64_t n_floats) {
    WB_ARG(host, "wm_load_weights: null pointer");
    if (n_floats != m->lay.total) {
        set_error("weight image has %lld fp32 values, config needs %lld", (long long)n_floats,
                  (long long)m->lay.total);
        return WB_ERR_IO;
    }
    WB_ARG(!m->loaded, "weights already loaded");
    cudaStream_t st = m->stream;
    const int D = m->D, L = m->L, F = m->F;
    WB_CHECK(dalloc(m, &m->w32, (size_t)n_floats));
    WB_CUDA(cudaMemcpyAsync(m->w32, host, (size_t)n_floats * 4, cudaMemcpyHostToDevice, st));
    const float *W = m->w32;
    const Layout &ly = m->lay;
    WB_CHECK(dalloc(m, &m->conv1_w, (size_t)D * 3 * 128));
    WB_CHECK(dalloc(m, &m->conv2_w, (size_t)D * 3 * D));
    WB_CHECK(convert_conv_weight(st, W + ly.conv1_w, m->conv1_w, D, m->NM, 128));
    WB_CHECK(convert_conv_weight(st, W + ly.conv2_w, m->conv2_w, D, D, D));
    WB_CHECK(dalloc(m, &m->tok_emb_h16, (size_t)m->V * D));
    WB_CHECK(convert_f32_h16(st, W + ly.tok_emb, m->tok_emb_h16, (size_t)m->V * D));
    WB_CHECK(dalloc(m, &m->cross_wkv, (size_t)L * 2 * D * D));
    WB_CHECK(dalloc(m, &m->cross_bkv, (size_t)L * 2 * D));
    WB_CUDA(cudaMemsetAsync(m->cross_bkv, 0, (size_t)L * 2 * D * 4, st));
    m->enc.resize(L);
    m->dec.resize(L);
    const size_t DD = (size_t)D * D;
    for (int side = 0; side < 2; side++) {
        for (int i = 0; i < L; i++) {
            const BlockW &b = side ? ly.dec[i] : ly.enc[i];
            LayerDev &d = side ? m->dec[i] : m->enc[i];
            WB_CHECK(dalloc(m, &d.wqkv, 3 * DD));
            WB_CHECK(convert_f32_h16(st, W + b.attn.q_w, d.wqkv, DD));
            WB_CHECK(convert_f32_h16(st, W + b.attn.k_w, d.wqkv + DD, DD));
            WB_CHECK(convert_f32_h16(st, W + b.attn.v_w, d.wqkv + 2 * DD, DD));
            WB_CHECK(dalloc(m, &d.bqkv, (size_t)3 * D));
            WB_CUDA(cudaMemsetAsync(d.bqkv, 0, (size_t)3 * D * 4, st));
            WB_CUDA(cudaMemcpyAsync(d.bqkv, W + b.attn.q_b, (size_t)D * 4, cudaMemcpyDeviceToDevice, st));
            WB_CUDA(cudaMemcpyAsync(d.bqkv + 2 * D, W + b.attn.v_b, (size_t)D * 4, cudaMemcpyDeviceToDevice, st));
            WB_CHECK(dalloc(m, &d.wo, DD));
            WB_CHECK(convert_f32_h16(st, W + b.attn.o_w, d.wo, DD));
            d.bo = W + b.attn.o_b;
            WB_CHECK(dalloc(m, &d.w1, (size_t)F * D));
            WB_CHECK(convert_f32_h16(st, W + b.fc1_w, d.w1, (size_t)F * D));
            WB_CHECK(dalloc(m, &d.w2, (size_t)F * D));
            WB_CHECK(convert_f32_h16(st, W + b.fc2_w, d.w2, (size_t)F * D));
            d.b1 = W + b.fc1_b, d.b2 = W + b.fc2_b;
            d.ln1_g = W + b.attn_ln_w, d.ln1_b = W + b.attn_ln_b;
            d.ln3_g = W + b.mlp_ln_w, d.ln3_b = W + b.mlp_ln_b;
            d.ln2_g = d.ln2_b = d.cbq = d.cbo = nullptr;
            if (side) {
                WB_CHECK(dalloc(m, &d.cwq, DD));
                WB_CHECK(convert_f32_h16(st, W + b.cross.q_w, d.cwq, DD));
                WB_CHECK(dalloc(m, &d.cwo, DD));
                WB_CHECK(convert_f32_h16(st, W + b.cross.o_w, d.cwo, DD));
                d.cbq = W + b.cross.q_b, d.cbo = W + b.cross.o_b;
                d.ln2_g = W + b.cross_ln_w, d.ln2_b = W + b.cross_ln_b;
                // fused cross K/V projection: rows [(i*2+0)*D, +D) = Wk (no bias), [(i*2+1)*D, +D) = Wv (+bias)
                WB_CHECK(convert_f32_h16(st, W + b.cross.k_w, m->cross_wkv + (size_t)(i * 2) * DD, DD));
                WB_CHECK(convert_f32_h16(st, W + b.cross.v_w, m->cross_wkv + (size_t)(i * 2 + 1) * DD, DD));
                WB_CUDA(cudaMemcpyAsync(m->cross_bkv + (size_t)(i * 2 + 1) * D, W + b.cross.v_b, (size_t)D * 4,
                                        cudaMemcpyDeviceToDevice, st));
                if (cross_attn_absorbed_supported(D, m->H)) {
                    const size_t HD = (size_t)m->H * D;
                    WB_CHECK(dalloc(m, &d.wqk, HD * D));
                    WB_CHECK(dalloc(m, &d.bqk, HD));
                    WB_CHECK(dalloc(m, &d.wov, HD * D));
                    WB_CHECK(dalloc(m, &d.bov, (size_t)D));
                    WB_CHECK(fold_cross_weights(st, W + b.cross.q_w, W + b.cross.q_b, W + b.cross.k_w, W + b.cross.v_w,
                                                W + b.cross.v_b, W + b.cross.o_w, W + b.cross.o_b, D, m->H, d.wqk, d.bqk,
                                                d.wov, d.bov));
                }
            }
        }
    }
    WB_CUDA(cudaStreamSynchronize(st));
    m->loaded = true;
    return WB_OK;
}

// ---------------------------------------------------------------------------------------------
// frontend
// ---------------------------------------------------------------------------------------------
int model_logmel(Model *m, const float *pcm_dev, int n, float *mel_dev) {
    if (n <= 0) return WB_OK;
    if (n > m->cmax_cap) {  // per-chunk maxima (ordered-int encoding); grow-only, stays with the model
        int *p = nullptr;
        WB_CHECK(dalloc(m, &p, (size_t)n));
        m->cmax = p, m->cmax_cap = n;
    }
    WB_CHECK(logmel_raw(m->stream, m->ft, pcm_dev, n, m->n_frames, mel_dev, m->cmax, m->frontend_impl));
    return logmel_finalize(m->stream, mel_dev, m->cmax, n, m->NM, m->n_frames);
}

// ---------------------------------------------------------------------------------------------
// encoder (whisper.mojo:71-99) + cross-K/V projection (layers.mojo:148-157)
// ---------------------------------------------------------------------------------------------
static int ensure_encoder_ws(Model *m, int n) {
    if (n <= m->enc_cap) return WB_OK;
    // (re)allocate; old buffers stay owned by the model and are freed at destroy -- the capacity only
    // grows to enc_batch, so this happens at most a couple of times.
    const size_t M = (size_t)n * m->S, D = m->D;
    WB_CHECK(dalloc(m, &m->e_melT, (size_t)n * m->n_frames * 128));
    WB_CHECK(dalloc(m, &m->e_x1T, (size_t)n * m->n_frames * D));
    WB_CHECK(dalloc(m, &m->e_x, M * D));
    WB_CHECK(dalloc(m, &m->e_xn, M * D));
    WB_CHECK(dalloc(m, &m->e_qkv, M * 3 * D));
    WB_CHECK(dalloc(m, &m->e_attn, M * D));
    WB_CHECK(dalloc(m, &m->e_h, M * m->F));
    WB_CHECK(dalloc(m, &m->e_enc, M * D));
    m->enc_cap = n;
    return WB_OK;
}

static GemmDesc plain_gemm(const h16 *A, int M, int K, const h16 *W, int N, const float *bias, int epi, void *out,
                           int64_t out_ld) {
    GemmDesc g;
    g.A = A, g.lda = K, g.src_rows = M, g.Cin = K, g.rows_per_batch = M, g.batches = 1;
    g.W = W, g.N = N, g.bias = bias, g.epi = epi;
    g.out[0] = out, g.out_ld[0] = out_ld, g.n_seg_ptrs = 1;
    return g;
}

static int cross_kv_project(Model *m, const h16 *enc_h16, int n, Cache *c, int cache_off) {
    // out layout [L][2][B][S][D]; segment s = l*2+kv has stride B*S*D, rows are (chunk, position).
    GemmDesc g = plain_gemm(enc_h16, n * m->S, m->D, m->cross_wkv, m->L * 2 * m->D, m->cross_bkv, EPI_STORE_H16,
                            c->cross_kv + (size_t)cache_off * m->S * m->D, m->D);
    g.n_seg_ptrs = 0;
    g.seg_cols = m->D;
    g.seg_stride = (int64_t)c->B * m->S * m->D;
    return gemm_run(m->stream, g, m->gemm_impl);
}

// Encode n <= enc_batch chunks.  Outputs (any may be null): fp32 enc_out, cross K/V into `into`.
static int encode_batch(Model *m, const float *mel_dev, int n, float *enc_out_dev, Cache *into, int cache_off) {
    WB_CHECK(ensure_encoder_ws(m, n));
    cudaStream_t st = m->stream;
    const int D = m->D, S = m->S, NF = m->n_frames, M = n * S, impl = m->gemm_impl;
    const float *W = m->w32;
    WB_CHECK(mel_to_h16_T(st, mel_dev, m->e_melT, n, m->NM, NF));
    {  // conv1 + GELU (whisper.mojo:73-75) as a 3-tap implicit GEMM over [frames][128 padded channels]
        GemmDesc g;
        g.A = m->e_melT, g.a_batch_stride = (int64_t)NF * 128, g.lda = 128, g.src_rows = NF;
        g.conv_stride = 1, g.pad = 1, g.taps = 3, g.Cin = 128, g.batches = n, g.rows_per_batch = NF;
        g.W = m->conv1_w, g.N = D, g.bias = W + m->lay.conv1_b, g.epi = EPI_GELU_H16;
        g.out[0] = m->e_x1T, g.out_ld[0] = D;
        WB_CHECK(gemm_run(st, g, impl));
    }
    {  // conv2 (stride 2) + GELU + positional embedding (whisper.mojo:78-89)
        GemmDesc g;
        g.A = m->e_x1T, g.a_batch_stride = (int64_t)NF * D, g.lda = D, g.src_rows = NF;
        g.conv_stride = 2, g.pad = 1, g.taps = 3, g.Cin = D, g.batches = n, g.rows_per_batch = S;
        g.W = m->conv2_w, g.N = D, g.bias = W + m->lay.conv2_b, g.epi = EPI_GELU_POS_F32;
        g.out[0] = m->e_x, g.out_ld[0] = D, g.pos = W + m->lay.enc_pos;
        WB_CHECK(gemm_run(st, g, impl));
    }
    for (int l = 0; l < m->L; l++) {  // layers.mojo:435-519 with is_decoder = False
        const LayerDev &d = m->enc[l];
        WB_CHECK(ln_h16(st, m->e_x, d.ln1_g, d.ln1_b, M, D, m->e_xn, nullptr));
        WB_CHECK(gemm_run(st, plain_gemm(m->e_xn, M, D, d.wqkv, 3 * D, d.bqkv, EPI_STORE_H16, m->e_qkv, 3 * D), impl));
        if (m->attn_impl) WB_CHECK(encoder_attention_tc(st, m->e_qkv, m->e_attn, n, S, m->H, D));
        else WB_CHECK(encoder_attention_ref(st, m->e_qkv, m->e_attn, n, S, m->H, D));
        WB_CHECK(gemm_run(st, plain_gemm(m->e_attn, M, D, d.wo, D, d.bo, EPI_RESID_F32, m->e_x, D), impl));
        WB_CHECK(ln_h16(st, m->e_x, d.ln3_g, d.ln3_b, M, D, m->e_xn, nullptr));
        WB_CHECK(gemm_run(st, plain_gemm(m->e_xn, M, D, d.w1, m->F, d.b1, EPI_GELU_H16, m->e_h, m->F), impl));
        WB_CHECK(gemm_run(st, plain_gemm(m->e_h, M, m->F, d.w2, D, d.b2, EPI_RESID_F32, m->e_x, D), impl));
    }
    if (into && into->cross_impl == 1) {
        // absorbed form: the bf16 encoder output IS the cross-attention cache (one tensor for all layers)
        WB_CHECK(ln_h16(st, m->e_x, W + m->lay.enc_ln_w, W + m->lay.enc_ln_b, M, D,
                         into->cross_enc + (size_t)cache_off * S * D, enc_out_dev));
        return WB_OK;
    }
    WB_CHECK(ln_h16(st, m->e_x, W + m->lay.enc_ln_w, W + m->lay.enc_ln_b, M, D, m->e_enc, enc_out_dev));
    if (into) WB_CHECK(cross_kv_project(m, m->e_enc, n, into, cache_off));
    return WB_OK;
}

int model_encode(Model *m, const float *mel_dev, int n, float *enc_out_dev, Cache *into, int cache_off) {
    WB_ARG(m->loaded, "encode before weights are loaded");
    const size_t mel_per = (size_t)m->NM * m->n_frames, enc_per = (size_t)m->S * m->D;
    for (int i = 0; i < n; i += m->enc_batch) {
        int nb = std::min(m->enc_batch, n - i);
        WB_CHECK(encode_batch(m, mel_dev + i * mel_per, nb, enc_out_dev ? enc_out_dev + i * enc_per : nullptr, into,
                              cache_off + i));
    }
    if (into) into->has_cross = true;
    return WB_OK;
}

// ---------------------------------------------------------------------------------------------
// KV cache (layers.mojo:14-69) and decode workspace
// ---------------------------------------------------------------------------------------------
int cache_create(Model *m, int B, int max_len, bool want_logits, int n_lanes, Cache **out) {
    WB_ARG(B > 0 && max_len > 0 && max_len <= m->T, "kvcache: bad size (B=%d max_len=%d n_text_ctx=%d)", B, max_len,
           m->T);
    Cache *c = new Cache();
    c->m = m, c->B = B, c->T = max_len;
    const size_t D = m->D;
    bool ok = true;
    auto A = [&](auto **p, size_t n) {
        void *q = nullptr;
        if (!ok) return;
        if (cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(**p)) != cudaSuccess) {
            set_error("kvcache: cudaMalloc of %zu bytes failed: %s", n * sizeof(**p),
                      cudaGetErrorString(cudaGetLastError()));
            ok = false;
            return;
        }
        c->owned.push_back(q);
        *p = static_cast<std::remove_reference_t<decltype(*p)>>(q);
    };
    if (n_lanes < 1 || B < 256) n_lanes = 1;  // small batches are launch bound either way
    if (n_lanes > 2) n_lanes = 2;
    const int slots = gemm_tiles_n(m->V);
    const int T_out = 5 + m->cfg.max_iters;
    A(&c->self_kv, (size_t)m->L * 2 * B * max_len * D);
    c->cross_impl = m->cross_impl;
    if (c->cross_impl == 1) A(&c->cross_enc, (size_t)B * m->S * D);
    else A(&c->cross_kv, (size_t)m->L * 2 * B * m->S * D);
    A(&c->tokens_out, (size_t)B * T_out);
    A(&c->out_len, B);
    A(&c->cur_tok, B);
    A(&c->done, B);
    A(&c->live, B);
    A(&c->stop_at, B);
    A(&c->scalars, 4 * n_lanes);
    c->lanes.resize(n_lanes);
    for (int i = 0; i < n_lanes && ok; i++) {
        Lane &ln = c->lanes[i];
        ln.b_off = (int)((int64_t)B * i / n_lanes);
        ln.B = (int)((int64_t)B * (i + 1) / n_lanes) - ln.b_off;
        ln.cross_splits = decode_attention_splits(ln.B, m->S, m->H);
        // the row buffers hold PREFILL_LEN rows per chunk: the prompt runs as ONE q_len = 4 forward (whisper.mojo:195-197),
        // a decode step uses the first B rows
        const size_t R = (size_t)ln.B * PREFILL_LEN;
        A(&ln.x, R * D);
        A(&ln.xn, R * D);
        A(&ln.q, R * D);
        A(&ln.pf_k, R * D);
        A(&ln.pf_v, R * D);
        A(&ln.attn, R * D);
        A(&ln.h, R * m->F);
        A(&ln.part_val, (size_t)ln.B * slots);
        A(&ln.part_idx, (size_t)ln.B * slots);
        A(&ln.next, ln.B);
        A(&ln.attn_ws, R * ln.cross_splits * m->H * 66);
        // split-K partial products of the residual GEMMs: round 1's form uses up to 4 slices, the chain kernel
        // chain_split_k(K) of them for K = D (o), H*D or D (cross-o) and F (fc2)
        ln.part_splits = std::max({4, chain_split_k((int)D), chain_split_k(c->cross_impl == 1 ? m->H * (int)D : (int)D),
                                   chain_split_k(m->F)});
        A(&ln.part, (size_t)ln.part_splits * R * D);
        if (c->cross_impl == 1) {
            A(&ln.qp, R * m->H * D);
            A(&ln.ctx, R * m->H * D);
        }
        if (want_logits || m->gemm_impl == GEMM_IMPL_REF) A(&ln.logits, (size_t)ln.B * m->V);
        ln.g.tokens_out = c->tokens_out + (size_t)ln.b_off * T_out;
        ln.g.out_len = c->out_len + ln.b_off;
        ln.g.cur_tok = c->cur_tok + ln.b_off;
        ln.g.done = c->done + ln.b_off;
        ln.g.live = c->live + ln.b_off, ln.g.stop_at = c->stop_at + ln.b_off;
        ln.g.scalars = c->scalars + 4 * i;
        ln.g.T_out = T_out, ln.g.eot = m->cfg.eot, ln.g.pos_quirk = m->cfg.pos_quirk;
        ln.counter_bytes = (size_t)(2 * m->L + 2) * chain_counter_ints(ln.B) * sizeof(int);
        A(&ln.counters, ln.counter_bytes / sizeof(int));
    }
    if (ok && (cudaMallocHost((void **)&c->pinned_scalars, 16 * sizeof(int)) != cudaSuccess ||
               cudaEventCreateWithFlags(&c->poll_ev[0], cudaEventDisableTiming) != cudaSuccess ||
               cudaEventCreateWithFlags(&c->poll_ev[1], cudaEventDisableTiming) != cudaSuccess)) {
        set_error("kvcache: pinned allocation / event creation failed");
        ok = false;
    }
    if (!ok) {
        cache_destroy(c);
        return WB_ERR_CUDA;
    }
    *out = c;
    return cache_reset(c);
}

void cache_destroy(Cache *c) {
    if (!c) return;
    if (c->m) {
        cudaStreamSynchronize(c->m->stream);
        cudaStreamSynchronize(c->m->stream2);
    }
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    for (Lane &ln : c->lanes) {
        for (ChainPlan *p : ln.plans) chain_plan_destroy(p);
        chain_plan_destroy(ln.plan_last_nolog);
    }
    for (void *p : c->owned) cudaFree(p);
    if (c->pinned_scalars) cudaFreeHost(c->pinned_scalars);
    for (cudaEvent_t e : c->poll_ev)
        if (e) cudaEventDestroy(e);
    delete c;
}

int cache_reset(Cache *c) {
    Model *m = c->m;
    // the reference zero-fills its cache tensors (layers.mojo:30-36)
    WB_CUDA(cudaMemsetAsync(c->self_kv, 0, (size_t)m->L * 2 * c->B * c->T * m->D * sizeof(h16), m->stream));
    for (Lane &ln : c->lanes) WB_CHECK(greedy_init(m->stream, ln.g, ln.B, m->cfg.prompt));
    WB_CUDA(cudaMemsetAsync(c->stop_at, 0x7f, (size_t)c->B * sizeof(int), m->stream));  // no forced lengths
    c->host_len = 0;
    c->has_cross = false;
    return WB_OK;
}

int cache_set_encoder(Cache *c, const float *enc_out_dev) {
    Model *m = c->m;
    WB_ARG(m->loaded, "kvcache_set_encoder before weights are loaded");
    const size_t per = (size_t)m->S * m->D;
    if (c->cross_impl == 1) {
        WB_CHECK(convert_f32_h16(m->stream, enc_out_dev, c->cross_enc, (size_t)c->B * per));
        c->has_cross = true;
        return WB_OK;
    }
    for (int i = 0; i < c->B; i += m->enc_batch) {
        int nb = std::min(m->enc_batch, c->B - i);
        WB_CHECK(ensure_encoder_ws(m, nb));
        WB_CHECK(convert_f32_h16(m->stream, enc_out_dev + i * per, m->e_enc, nb * per));
        WB_CHECK(cross_kv_project(m, m->e_enc, nb, c, i));
    }
    c->has_cross = true;
    return WB_OK;
}

// The fused decode step: one ChainPlan per kernel (decode_chain.h).  Buffers and weights are fixed for the life of
// the cache, so the tensor maps and phase lists are built once here and the step only launches them.
static int build_chain_plans(Cache *c, Lane &ln) {
    Model *m = c->m;
    WB_ARG(m->loaded, "decode before weights are loaded");
    const int B = ln.B, D = m->D, L = m->L, F = m->F, HD = m->H * D;
    const float *W = m->w32;
    int *cur_len = ln.g.scalars, *pos = ln.g.scalars + 1;
    const size_t self_seg = (size_t)c->B * c->T * D, self_off = (size_t)ln.b_off * c->T * D;
    const size_t cstride = chain_counter_ints(B);
    auto qkv_phase = [&](ChainPlan *p, int l) {
        const LayerDev &d = m->dec[l];
        h16 *sk = c->self_kv + (size_t)(l * 2) * self_seg + self_off, *sv = sk + self_seg;
        ChainGemm g;  // q, k, v projections; k / v rows land in the cache at position cur_len (layers.mojo:131-143)
        g.A = ln.xn, g.K = D, g.W = d.wqkv, g.N = 3 * D, g.bias = d.bqkv, g.epi = EPI_STORE_H16;
        g.n_seg_ptrs = 3, g.seg_cols = D;
        g.out[0] = ln.q, g.out_ld[0] = D;
        g.out[1] = sk, g.out_ld[1] = (int64_t)c->T * D, g.dyn_mult[1] = D;
        g.out[2] = sv, g.out_ld[2] = (int64_t)c->T * D, g.dyn_mult[2] = D;
        g.dyn_off = cur_len;
        return chain_plan_add_gemm(p, g);
    };
    auto partial_phase = [&](ChainPlan *p, const h16 *A, int K, const h16 *Wt) {
        ChainGemm g;  // residual GEMM as split-K partial products; the row phase that follows adds them to x
        g.A = A, g.K = K, g.W = Wt, g.N = D, g.epi = EPI_STORE_F32, g.split_k = chain_split_k(K);
        g.out[0] = ln.part, g.out_ld[0] = D;
        return chain_plan_add_gemm(p, g);
    };
    auto ln_phase = [&](ChainPlan *p, int K, const float *bias, const float *g, const float *b) {
        ChainRows r;
        r.x = ln.x, r.part = ln.part, r.n_split = chain_split_k(K), r.bias = bias, r.gamma = g, r.beta = b, r.xn = ln.xn;
        return chain_plan_add_rows(p, r);
    };
    int idx = 0;
    auto new_plan = [&]() { return chain_plan_create(B, D, ln.counters + cstride * (idx++)); };
    {  // embed + attn_ln of layer 0 + qkv_0 (whisper.mojo:138-149)
        ChainPlan *p = new_plan();
        ln.plans.push_back(p);
        ChainRows r;
        r.embed = true, r.x = ln.x, r.xn = ln.xn, r.gamma = m->dec[0].ln1_g, r.beta = m->dec[0].ln1_b;
        r.tok_emb = W + m->lay.tok_emb, r.pos_emb = W + m->lay.dec_pos, r.cur_tok = ln.g.cur_tok, r.pos_dev = pos;
        r.vocab = m->V, r.n_pos = m->T;
        WB_CHECK(chain_plan_add_rows(p, r));
        WB_CHECK(qkv_phase(p, 0));
    }
    for (int l = 0; l < L; l++) {
        const LayerDev &d = m->dec[l];
        {  // self-attention output projection + residual, cross_attn_ln, cross query projection (layers.mojo:455-470)
            ChainPlan *p = new_plan();
            ln.plans.push_back(p);
            WB_CHECK(partial_phase(p, ln.attn, D, d.wo));
            WB_CHECK(ln_phase(p, D, d.bo, d.ln2_g, d.ln2_b));
            ChainGemm g;
            g.A = ln.xn, g.K = D, g.epi = EPI_STORE_H16;
            if (c->cross_impl == 1) g.W = d.wqk, g.N = HD, g.bias = d.bqk, g.out[0] = ln.qp, g.out_ld[0] = HD;
            else g.W = d.cwq, g.N = D, g.bias = d.cbq, g.out[0] = ln.q, g.out_ld[0] = D;
            WB_CHECK(chain_plan_add_gemm(p, g));
        }
        const bool last = l + 1 == L;
        for (int variant = 0; variant < (last ? 2 : 1); variant++) {
            // cross-attention output projection + residual, mlp_ln, MLP (layers.mojo:482-517), then the LayerNorm in
            // front of what follows -- the next layer's attn_ln + qkv, or the decoder's ln_post before the logits
            // (whisper.mojo:156-158), or nothing on a prompt step (variant 1)
            ChainPlan *p = new_plan();
            if (variant == 0) ln.plans.push_back(p);
            else ln.plan_last_nolog = p;
            if (c->cross_impl == 1) {
                WB_CHECK(partial_phase(p, ln.ctx, HD, d.wov));
                WB_CHECK(ln_phase(p, HD, d.bov, d.ln3_g, d.ln3_b));
            } else {
                WB_CHECK(partial_phase(p, ln.attn, D, d.cwo));
                WB_CHECK(ln_phase(p, D, d.cbo, d.ln3_g, d.ln3_b));
            }
            ChainGemm g;
            g.A = ln.xn, g.K = D, g.W = d.w1, g.N = F, g.bias = d.b1, g.epi = EPI_GELU_H16, g.out[0] = ln.h, g.out_ld[0] = F;
            WB_CHECK(chain_plan_add_gemm(p, g));
            WB_CHECK(partial_phase(p, ln.h, F, d.w2));
            if (!last) {
                WB_CHECK(ln_phase(p, F, d.b2, m->dec[l + 1].ln1_g, m->dec[l + 1].ln1_b));
                WB_CHECK(qkv_phase(p, l + 1));
            } else if (variant == 0) {
                WB_CHECK(ln_phase(p, F, d.b2, W + m->lay.dec_ln_w, W + m->lay.dec_ln_b));
            } else {
                WB_CHECK(ln_phase(p, F, d.b2, nullptr, nullptr));
            }
        }
    }
    // launch geometry (grid, per-CTA L2 prefetch slices) once every phase is in place
    for (ChainPlan *p : ln.plans) WB_CHECK(chain_plan_finalize(p));
    WB_CHECK(chain_plan_finalize(ln.plan_last_nolog));
    return WB_OK;
}

// Does a q_len = 1 step of this lane run the chain kernels?  (decode_fused = 2: by wave size, see decode_step)
static bool lane_is_fused(const Cache *c, const Lane &ln) {
    const Model *m = c->m;
    static const bool fused_lanes = getenv("WB_FUSED_LANES") != nullptr;  // experiment: chain kernels on two lanes
    const bool want_fused = m->decode_fused == 1 || (m->decode_fused == 2 && ln.B <= FUSED_MAX_WAVE);
    return want_fused && m->gemm_impl == GEMM_IMPL_TC && (c->lanes.size() == 1 || fused_lanes) && m->D <= 768;
}
// Build the lanes' chain plans if the step will use them.  Every entry point calls it BEFORE it queues / captures
// steps (greedy_loop, the step API, teacher forcing), so nothing but launches happens inside a stream capture.
static int prepare_chain_plans(Cache *c) {
    for (Lane &ln : c->lanes)
        if (lane_is_fused(c, ln) && ln.plans.empty()) WB_CHECK(build_chain_plans(c, ln));
    return WB_OK;
}

template <typename Fn>
static int timed_kernel(Model *m, cudaStream_t st, int category, Fn launch) {
    if (!(m->profile_attn == 2 || (m->profile_attn == 1 && category == TK_CROSS))) return launch();
    KernelTimer &t = m->cross_timer;
    while ((int)t.ev.size() < t.used + 2) {
        cudaEvent_t e;
        WB_CUDA(cudaEventCreate(&e));
        t.ev.push_back(e);
    }
    if ((int)t.cat.size() < t.used / 2 + 1) t.cat.resize(t.used / 2 + 1);
    t.cat[t.used / 2] = category;
    WB_CUDA(cudaEventRecord(t.ev[t.used], st));
    int rc = launch();
    WB_CUDA(cudaEventRecord(t.ev[t.used + 1], st));
    t.used += 2;
    return rc;
}

// One decoder forward with q_len = 1 for every chunk of the cache (whisper.mojo:130-167,
// layers.mojo:435-519 with is_decoder = True).  Token / position / cur_len come from device memory.
// q_len = 1: one cached decode step for the lane's chunks (layers.mojo:186-272 path).  q_len = PREFILL_LEN: the
// reference's prefill, `decoder.forward(prompt ids, enc, cache, 0)` (whisper.mojo:195-197), as ONE forward over
// R = q_len * B rows (row b * q_len + p = prompt position p of chunk b): the dense ops run on R rows, the self
// attention is the causal block path (layers.mojo:273-342, mask :304-320) and only the last position's logits are
// computed (whisper.mojo:159-166).  Every row's arithmetic is the one the cached single-token step at that position
// runs (same GEMM K order, same attention kernels), so the ids do not depend on which form ran -- tested.
int decode_step(Cache *c, Lane &ln, cudaStream_t st, bool with_logits, bool store_logits, bool advance, int q_len) {
    Model *m = c->m;
    const int B = ln.B, D = m->D, impl = m->gemm_impl;
    const int R = B * q_len;  // rows of the dense ops
    WB_ARG(q_len == 1 || (q_len == PREFILL_LEN && with_logits && c->host_len == 0), "decode_step: bad prefill call");
    const float *W = m->w32;
    int *cur_len = ln.g.scalars, *pos = ln.g.scalars + 1;
    // cache-wide segment sizes; this lane's chunks start b_off rows into every segment
    const size_t self_seg = (size_t)c->B * c->T * D, cross_seg = (size_t)c->B * m->S * D;
    const size_t self_off = (size_t)ln.b_off * c->T * D, cross_off = (size_t)ln.b_off * m->S * D;
    // finished chunks (EOT) drop out of the attention kernels: the self-attention CTA of a done chunk exits at once, the
    // cross-attention walks the live list (rebuilt every 16 steps next to the host's EOT poll, greedy_loop)
    const int *skip_done = m->skip_done ? ln.g.done : nullptr, *skip_live = m->skip_done ? ln.g.live : nullptr;
    // decode_fused = 2: by wave size.  The chain kernels win while the step is latency bound; from ~1.3 k chunks on the
    // phases are throughput bound and the CTA-pair GEMM kernels of the kernel-per-op form are ahead (B200, Tiny:
    // 1024 chunks 249.8 vs 252.8 ms, 1536: 349.8 vs 342.8, 2048: 445 vs 434; tools/fused_ab.py).  Same bits either way.
    const bool fused = q_len == 1 && lane_is_fused(c, ln);
    WB_ARG(!fused || !ln.plans.empty(), "decode_step: chain plans were not prepared");
    if (fused) {
        // 4 L + 3 kernels: chain | self-attn | chain | cross-attn | chain | ... | logits+argmax | argmax reduce
        WB_CUDA(cudaMemsetAsync(ln.counters, 0, ln.counter_bytes, st));  // arrival counters of this step's chains
        WB_CHECK(timed_kernel(m, st, TK_CHAIN_FIRST, [&] { return chain_launch(st, ln.plans[0]); }));
        for (int l = 0; l < m->L; l++) {
            h16 *sk = c->self_kv + (size_t)(l * 2) * self_seg + self_off, *sv = sk + self_seg;
            DecodeAttnArgs a;
            a.q = ln.q, a.K = sk, a.V = sv, a.out = ln.attn, a.kv_batch_stride = (int64_t)c->T * D;
            a.B = B, a.H = m->H, a.D = D, a.len_const = 0, a.len_dev = cur_len, a.len_add = 1, a.max_len = c->T;
            a.splits = 1, a.ws = nullptr, a.done = skip_done;
            WB_CHECK(timed_kernel(m, st, TK_SELF, [&] { return decode_attention(st, a); }));
            WB_CHECK(timed_kernel(m, st, TK_CHAIN_B, [&] { return chain_launch(st, ln.plans[1 + 2 * l]); }));
            if (c->cross_impl == 1) {
                const h16 *enc = c->cross_enc + (size_t)ln.b_off * m->S * D;
                WB_CHECK(timed_kernel(m, st, TK_CROSS, [&] {
                    return cross_attention_absorbed(st, ln.qp, enc, ln.ctx, B, m->S, D, m->H, skip_live, skip_live ? ln.g.scalars + 3 : nullptr);
                }));
            } else {
                a.q = ln.q, a.K = c->cross_kv + (size_t)(l * 2) * cross_seg + cross_off, a.V = a.K + cross_seg;
                a.kv_batch_stride = (int64_t)m->S * D;
                a.len_const = m->S, a.len_dev = nullptr, a.len_add = 0, a.max_len = m->S;
                a.splits = ln.cross_splits, a.ws = ln.attn_ws;
                WB_CHECK(timed_kernel(m, st, TK_CROSS, [&] { return decode_attention(st, a); }));
            }
            ChainPlan *ca = (l + 1 == m->L && !with_logits) ? ln.plan_last_nolog : ln.plans[2 + 2 * l];
            WB_CHECK(timed_kernel(m, st, TK_CHAIN_CA, [&] { return chain_launch(st, ca); }));
        }
    } else {
    WB_CHECK(timed_kernel(m, st, TK_LN, [&] {
        if (q_len > 1)  // rows b * q_len + p = prompt id p at position p; cur_len becomes q_len - 1 (+ 1 after the logits)
            return embed_ln_prefill(st, W + m->lay.tok_emb, W + m->lay.dec_pos, m->cfg.prompt, q_len, B, D, m->V, m->T,
                                    m->dec[0].ln1_g, m->dec[0].ln1_b, ln.x, ln.xn, cur_len);
        return embed_ln(st, W + m->lay.tok_emb, W + m->lay.dec_pos, ln.g.cur_tok, pos, B, D, m->V, m->T, m->dec[0].ln1_g,
                        m->dec[0].ln1_b, ln.x, ln.xn);
    }));
    // x += A W^T + bias followed by xn = LN(x; g, b)  (g == nullptr: no LayerNorm follows).
    //  * plain: residual-add GEMM epilogue, then the LayerNorm kernel;
    //  * split-K (tcgen05 path, batch large enough): the GEMM with M = batch and N = D only fills 32 CTAs and its K
    //    loop is what takes the time, so K is cut into slices that run as extra tiles (fp32 partials), and the
    //    LayerNorm kernel sums them into x in a fixed order first -- same number of launches, whole chip busy.
    auto resid_gemm_ln = [&](int cat, const h16 *A, int K, const h16 *Wt, const float *bias, const float *g,
                             const float *bb) -> int {
        // The same K slices, summed in the same order, as the chain kernels use (chain_split_k: a property of the
        // model, never of the batch size), including the one-slice case: x + bias + part[0] + part[1] + ..., so this
        // kernel-per-op form and the fused form produce the same bits, and a chunk's fp32 sums (hence its ids at
        // low-margin steps) do not depend on how many other chunks share the wave.
        if (impl == GEMM_IMPL_TC && ln.part && m->decode_split_k >= 1) {
            const int split = chain_split_k(K);
            GemmDesc gd = plain_gemm(A, R, K, Wt, D, nullptr, EPI_STORE_F32, ln.part, D);
            gd.split_k = split;
            WB_CHECK(timed_kernel(m, st, cat, [&] { return gemm_run(st, gd, impl); }));
            return timed_kernel(m, st, TK_LN, [&] { return resid_ln(st, ln.x, ln.part, split, bias, g, bb, R, D, ln.xn); });
        }
        WB_CHECK(timed_kernel(m, st, cat, [&] {
            return gemm_run(st, plain_gemm(A, R, K, Wt, D, bias, EPI_RESID_F32, ln.x, D), impl);
        }));
        if (!g) return WB_OK;
        return timed_kernel(m, st, TK_LN, [&] { return ln_h16(st, ln.x, g, bb, R, D, ln.xn, nullptr); });
    };
    for (int l = 0; l < m->L; l++) {
        const LayerDev &d = m->dec[l];
        h16 *sk = c->self_kv + (size_t)(l * 2) * self_seg + self_off, *sv = sk + self_seg;
        h16 *ck = c->cross_kv ? c->cross_kv + (size_t)(l * 2) * cross_seg + cross_off : nullptr;
        h16 *cv = ck ? ck + cross_seg : nullptr;
        // (xn = LN1(x) was produced by embed_ln / by the previous layer's fc2 step)
        {  // q, k, v projections; k / v rows land in the cache at position cur_len (layers.mojo:131-143)
            GemmDesc g = plain_gemm(ln.xn, R, D, d.wqkv, 3 * D, d.bqkv, EPI_STORE_H16, ln.q, D);
            g.n_seg_ptrs = 3, g.seg_cols = D;
            if (q_len == 1) {
                g.out[1] = sk, g.out_ld[1] = (int64_t)c->T * D, g.dyn_mult[1] = D;
                g.out[2] = sv, g.out_ld[2] = (int64_t)c->T * D, g.dyn_mult[2] = D;
                g.dyn_off = cur_len;
            } else {  // prefill: k / v rows [R][D], then into cache rows 0 .. q_len-1 of their chunks
                g.out[1] = ln.pf_k, g.out_ld[1] = D;
                g.out[2] = ln.pf_v, g.out_ld[2] = D;
            }
            WB_CHECK(timed_kernel(m, st, TK_QKV, [&] { return gemm_run(st, g, impl); }));
            if (q_len > 1)
                WB_CHECK(timed_kernel(m, st, TK_MISC, [&] { return kv_scatter(st, ln.pf_k, ln.pf_v, sk, sv, B, q_len, D, (int64_t)c->T * D); }));
        }
        DecodeAttnArgs a;
        a.q = ln.q, a.K = sk, a.V = sv, a.out = ln.attn, a.kv_batch_stride = (int64_t)c->T * D;
        a.B = R, a.H = m->H, a.D = D, a.len_const = 0, a.len_dev = cur_len, a.len_add = 1, a.max_len = c->T;
        a.splits = 1, a.ws = nullptr, a.done = skip_done;
        if (q_len > 1) a.len_dev = nullptr, a.len_add = 0, a.q_len = q_len, a.done = nullptr;  // causal: row p sees keys 0 .. p
        WB_CHECK(timed_kernel(m, st, TK_SELF, [&] { return decode_attention(st, a); }));
        WB_CHECK(resid_gemm_ln(TK_O, ln.attn, D, d.wo, d.bo, d.ln2_g, d.ln2_b));
        // cross attention over the encoder positions (layers.mojo:463-488)
        if (c->cross_impl == 1) {
            // absorbed form: q' = (Wk_h^T Wq_h) x + ..., attend over enc_out, out = (Wo Wv_h) ctx_h + ...
            const int HD = m->H * D;
            WB_CHECK(timed_kernel(m, st, TK_CQ, [&] {
                return gemm_run(st, plain_gemm(ln.xn, R, D, d.wqk, HD, d.bqk, EPI_STORE_H16, ln.qp, HD), impl);
            }));
            const h16 *enc = c->cross_enc + (size_t)ln.b_off * m->S * D;
            WB_CHECK(timed_kernel(m, st, TK_CROSS, [&] {
                if (q_len > 1) return cross_attention_absorbed(st, ln.qp, enc, ln.ctx, B, m->S, D, m->H, nullptr, nullptr, q_len);
                return cross_attention_absorbed(st, ln.qp, enc, ln.ctx, B, m->S, D, m->H, skip_live, skip_live ? ln.g.scalars + 3 : nullptr);
            }));
            WB_CHECK(resid_gemm_ln(TK_CO, ln.ctx, HD, d.wov, d.bov, d.ln3_g, d.ln3_b));
        } else {
            WB_CHECK(timed_kernel(m, st, TK_CQ, [&] {
                return gemm_run(st, plain_gemm(ln.xn, R, D, d.cwq, D, d.cbq, EPI_STORE_H16, ln.q, D), impl);
            }));
            a.K = ck, a.V = cv, a.kv_batch_stride = (int64_t)m->S * D;  // (q_len > 1: every row of a chunk sees all S keys)
            a.len_const = m->S, a.len_dev = nullptr, a.len_add = 0, a.max_len = m->S;
            a.splits = ln.cross_splits, a.ws = ln.attn_ws;
            WB_CHECK(timed_kernel(m, st, TK_CROSS, [&] { return decode_attention(st, a); }));
            WB_CHECK(resid_gemm_ln(TK_CO, ln.attn, D, d.cwo, d.cbo, d.ln3_g, d.ln3_b));
        }
        // MLP (layers.mojo:490-517); the LayerNorm after fc2 is the next layer's attn_ln, or the decoder's ln_post
        // in front of the logits (whisper.mojo:156-158), or none on a prefill step
        WB_CHECK(timed_kernel(m, st, TK_FC1, [&] {
            return gemm_run(st, plain_gemm(ln.xn, R, D, d.w1, m->F, d.b1, EPI_GELU_H16, ln.h, m->F), impl);
        }));
        const bool last = l + 1 == m->L;
        const float *ng = last ? (with_logits ? W + m->lay.dec_ln_w : nullptr) : m->dec[l + 1].ln1_g;
        const float *nb = last ? (with_logits ? W + m->lay.dec_ln_b : nullptr) : m->dec[l + 1].ln1_b;
        WB_CHECK(resid_gemm_ln(TK_FC2, ln.h, m->F, d.w2, d.b2, ng, nb));
    }
    }  // !fused
    if (with_logits) {  // whisper.mojo:156-166 + argmax :198,219
        // prefill: only the last position's logits (whisper.mojo:159-166 slices row q_len - 1): rows q_len - 1, 2 q_len - 1, ...
        GemmDesc g = plain_gemm(ln.xn + (size_t)(q_len - 1) * D, B, D, m->tok_emb_h16, m->V, nullptr, EPI_ARGMAX, nullptr, 0);
        g.lda = q_len * D;
        g.part_val = ln.part_val, g.part_idx = ln.part_idx;
        g.logits = (store_logits || impl == GEMM_IMPL_REF) ? ln.logits : nullptr;
        WB_ARG(!(store_logits || impl == GEMM_IMPL_REF) || ln.logits, "decode_step: cache has no logits buffer");
        WB_CHECK(timed_kernel(m, st, TK_LOGITS, [&] { return gemm_run(st, g, impl); }));
        WB_CHECK(timed_kernel(m, st, TK_MISC, [&] {
            // greedy loop: argmax + append / EOT / position bookkeeping in one kernel (whisper.mojo:198-221)
            if (advance) return greedy_argmax_advance(st, ln.g, B, ln.part_val, ln.part_idx, gemm_tiles_n(m->V), ln.next);
            return argmax_partials(st, ln.part_val, ln.part_idx, B, gemm_tiles_n(m->V), ln.next);
        }));
    }
    return WB_OK;
}

// ---------------------------------------------------------------------------------------------
// Whisper.transcribe (whisper.mojo:184-223), batched
// ---------------------------------------------------------------------------------------------
// One greedy step for every lane.  With two lanes the second one runs on stream2 between a fork and
// a join event, so the same code serves eager execution and stream capture into one graph.
// mode 0: prompt step (feed next_prompt_token next), 1: greedy step, 2: the whole prompt as one q_len = 4 forward
static int step_all_lanes(Cache *c, bool with_logits, int mode, int next_prompt_token) {
    Model *m = c->m;
    const bool two = c->lanes.size() > 1;
    if (two) {
        WB_CUDA(cudaEventRecord(m->ev_fork, m->stream));
        WB_CUDA(cudaStreamWaitEvent(m->stream2, m->ev_fork, 0));
    }
    for (size_t i = 0; i < c->lanes.size(); i++) {
        Lane &ln = c->lanes[i];
        cudaStream_t st = i == 0 ? m->stream : m->stream2;
        const bool fused_adv = with_logits && mode >= 1;
        WB_CHECK(decode_step(c, ln, st, with_logits, false, fused_adv, mode == 2 ? PREFILL_LEN : 1));
        if (!fused_adv) WB_CHECK(greedy_advance(st, ln.g, ln.B, mode, next_prompt_token, ln.next));
    }
    if (two) {
        WB_CUDA(cudaEventRecord(m->ev_join, m->stream2));
        WB_CUDA(cudaStreamWaitEvent(m->stream, m->ev_join, 0));
    }
    return WB_OK;
}

static int greedy_loop(Cache *c) {
    Model *m = c->m;
    cudaStream_t st = m->stream;
    WB_CHECK(prepare_chain_plans(c));
    const int n_lanes = (int)c->lanes.size();
    // prefill: the reference runs the 4 prompt ids as one q_len = 4 forward with a causal mask
    // (whisper.mojo:195-197) -- decode_step with q_len = 4.  prefill_impl = 0 feeds them one by one through the
    // cached step instead, which computes the same thing (masked scores are exp(-1e10 - max) = 0 there; only the
    // last position's logits are used) with 3 more forwards; the two forms produce the same ids (tested).
    if (m->prefill_impl) {
        WB_CHECK(step_all_lanes(c, true, 2, 0));
    } else {
        for (int i = 0; i < 4; i++) {
            if (i < 3) WB_CHECK(step_all_lanes(c, false, 0, m->cfg.prompt[i + 1]));
            else WB_CHECK(step_all_lanes(c, true, 1, 0));
        }
    }
    const bool graph = m->use_graph && !m->profile_attn;
    if (c->graph_exec && c->graph_key != (m->decode_fused | (m->skip_done << 2) | ((int)g_pdl << 3))) {
        cudaGraphExecDestroy(c->graph_exec);  // captured under other launch options: capture again
        c->graph_exec = nullptr;
    }
    if (graph && !c->graph_exec) {
        c->graph_key = m->decode_fused | (m->skip_done << 2) | ((int)g_pdl << 3);
        cudaGraph_t gr = nullptr;
        WB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        const int64_t l0 = g_launches.load();
        int rc = step_all_lanes(c, true, 1, 0);
        c->graph_kernels = (int)(g_launches.load() - l0);  // kernel nodes of one step (another thread's launches aside)
        cudaError_t e = cudaStreamEndCapture(st, &gr);
        if (rc != WB_OK) {
            if (gr) cudaGraphDestroy(gr);
            return rc;
        }
        WB_CUDA(e);
        WB_CUDA(cudaGraphInstantiate(&c->graph_exec, gr, 0));
        cudaGraphDestroy(gr);
    }
    // `if next_token == 50257: break` (whisper.mojo:207), batched: stop once every chunk has produced EOT.  The
    // done counters are snapshotted every 16 steps and the host looks at the PREVIOUS snapshot, so a block of
    // steps is always queued behind the one being checked and host jitter never idles the GPU (finished
    // chunks ignore the extra steps: greedy_advance is a no-op for them).
    int blk = 0;
    for (int it = 0; it < m->cfg.max_iters; it++) {  // whisper.mojo:205
        if ((it & 15) == 0) {
            if (m->skip_done && it > 0)  // chunks that finished during the last 16 steps leave the attention kernels
                for (Lane &ln : c->lanes) WB_CHECK(greedy_rebuild_live(st, ln.g, ln.B));
            WB_CUDA(cudaMemcpyAsync(c->pinned_scalars + 8 * (blk & 1), c->scalars, 4 * n_lanes * sizeof(int),
                                    cudaMemcpyDeviceToHost, st));
            WB_CUDA(cudaEventRecord(c->poll_ev[blk & 1], st));
            if (blk >= 1) {
                WB_CUDA(cudaEventSynchronize(c->poll_ev[(blk - 1) & 1]));
                int n_done = 0;
                for (int i = 0; i < n_lanes; i++) n_done += c->pinned_scalars[8 * ((blk - 1) & 1) + 4 * i + 2];
                if (n_done >= c->B) break;
            }
            blk++;
        }
        if (graph) {
            WB_CUDA(cudaGraphLaunch(c->graph_exec, st));
            g_launches.fetch_add(c->graph_kernels, std::memory_order_relaxed);  // kernels replayed by the graph
        } else {
            WB_CHECK(step_all_lanes(c, true, 1, 0));
        }
    }
    return WB_OK;
}

// Events used to pipeline / time the per-sub-batch work of one transcribe call (grow-only pool).
static int event_pool(Model *m, size_t n_plain, size_t n_timed) {
    while (m->ev_plain.size() < n_plain) {
        cudaEvent_t e;
        WB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        m->ev_plain.push_back(e);
    }
    while (m->ev_timed.size() < n_timed) {
        cudaEvent_t e;
        WB_CUDA(cudaEventCreate(&e));
        m->ev_timed.push_back(e);
    }
    return WB_OK;
}

// `in_host` (optional): the input still lives in host memory and `mel_dev` / `pcm_dev` is an empty staging
// buffer of the same size.  The upload is then cut into encoder sub-batches and issued on the copy stream,
// and the frontend + encoder of sub-batch i wait only for their own slice, so the PCIe transfer of the
// following slices runs under the compute of the earlier ones (pinned host memory makes the copies async).
static int model_transcribe_impl(Model *m, const float *mel_dev, const float *pcm_dev, int n, int32_t *out_tokens_dev,
                                 int32_t *out_len_dev, const float *in_host);

// Small batches (option "small_batch" = largest wave treated as small, default 0 = off): kernel-count latency, not
// bandwidth, decides there.  One CTA per chunk cannot fill the chip in the absorbed cross-attention (12 key blocks in
// sequence, 4 x 24 us of a 370 us step at batch 1), so such waves use the K/V form with the keys split over the SMs,
// and programmatic dependent launch for the decode-step kernels: 74 -> 57 ms per 30 s clip at batch 1.  Off by
// default because a chunk then no longer gets bit-identical ids alone and inside a large batch (two formulations).
int model_transcribe(Model *m, const float *mel_dev, const float *pcm_dev, int n, int32_t *out_tokens_dev,
                     int32_t *out_len_dev, const float *in_host) {
    const bool small = m->small_batch > 0 && n > 0 && std::min(n, m->wave_max) <= m->small_batch && m->cross_impl == 1;
    const int saved_impl = m->cross_impl, saved_fused = m->decode_fused;
    PdlScope pdl(small || m->pdl);  // this thread's launches only
    // (kernel-per-op decode for such waves: its launches can overlap through programmatic dependent launch, which the
    // cooperative chain kernels cannot -- 57 vs 68 ms per clip at batch 1; the two forms produce the same bits)
    if (small) m->cross_impl = 0, m->decode_fused = 0;
    const int rc = model_transcribe_impl(m, mel_dev, pcm_dev, n, out_tokens_dev, out_len_dev, in_host);
    m->cross_impl = saved_impl, m->decode_fused = saved_fused;
    return rc;
}

static int model_transcribe_impl(Model *m, const float *mel_dev, const float *pcm_dev, int n, int32_t *out_tokens_dev,
                                 int32_t *out_len_dev, const float *in_host) {
    WB_ARG(m->loaded, "transcribe before weights are loaded");
    WB_ARG(n > 0 && (mel_dev || pcm_dev) && out_tokens_dev && out_len_dev, "transcribe: bad arguments");
    cudaStream_t st = m->stream;
    const int T_out = 5 + m->cfg.max_iters;
    const int T_cache = std::min(m->T, (T_out + 7) & ~7);
    m->cross_timer.used = 0;
    const size_t mel_per = (size_t)m->NM * m->n_frames;
    const size_t in_per = mel_dev ? mel_per : (size_t)m->n_samples;
    float *in_dev = const_cast<float *>(mel_dev ? mel_dev : pcm_dev);
    int wave = std::min(n, m->wave_max);
    const bool reuse = m->tr_cache && m->tr_cache->B == wave && m->tr_cache->T == T_cache &&
                       m->tr_cache->cross_impl == m->cross_impl;  // same shape as last time: it fits
    if (!reuse) {  // bound the cross K/V cache to about half of the free HBM (cudaMemGetInfo costs ~3 ms)
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        size_t per_chunk = ((size_t)m->L * 2 * T_cache + (m->cross_impl == 1 ? (size_t)m->S : (size_t)m->L * 2 * m->S)) *
                           m->D * sizeof(h16);
        // + the decode workspace: PREFILL_LEN rows per chunk of x (fp32), xn / q / k / v / attn, h, the split-K partials
        // (at most K / 384 slices of fp32 [D]) and, in the absorbed form, q' and ctx [H * D]
        const size_t HD = (size_t)m->H * m->D;
        per_chunk += (size_t)PREFILL_LEN * ((size_t)m->D * (4 + 5 * sizeof(h16)) + (size_t)m->F * sizeof(h16) +
                                            (std::max<size_t>(4, HD / 384 + 1)) * m->D * 4 + 2 * HD * sizeof(h16));
        size_t cap = std::max<size_t>(1, (free_b / 2) / per_chunk);
        wave = (int)std::min<size_t>(wave, cap);
    }
    const int eb = m->enc_batch;
    // sub-batches never straddle a wave: enumerate them once so the uploads can all be queued up front
    std::vector<std::pair<int, int>> subs;  // (first chunk, count)
    for (int w0 = 0; w0 < n; w0 += wave)
        for (int i = w0; i < std::min(n, w0 + wave); i += eb) subs.push_back({i, std::min(eb, std::min(n, w0 + wave) - i)});
    const size_t n_waves = (size_t)cdiv(n, wave);
    WB_CHECK(event_pool(m, subs.size() + 1, 3 * subs.size() + 2 * n_waves));
    if (in_host) {
        // the staging buffer may still be read by work queued earlier on the compute stream
        WB_CUDA(cudaEventRecord(m->ev_plain[subs.size()], st));
        WB_CUDA(cudaStreamWaitEvent(m->stream2, m->ev_plain[subs.size()], 0));
    }
    // Upload of sub-batch k on the copy stream.  Sub-batch 0 goes first, and sub-batch k + 1 is issued right after
    // the compute of sub-batch k has been queued: with pinned host memory every copy is an asynchronous DMA and the
    // order of issue does not matter; with PAGEABLE memory cudaMemcpyAsync stages through the driver and blocks the
    // host until the slice is staged, so issuing all slices up front would keep the GPU idle for the whole transfer,
    // while this order hides each slice's staging under the previous slice's frontend + encoder.
    auto upload = [&](size_t k) -> int {
        if (!in_host || k >= subs.size()) return WB_OK;
        const size_t off = (size_t)subs[k].first * in_per, cnt = (size_t)subs[k].second * in_per;
        WB_CUDA(cudaMemcpyAsync(in_dev + off, in_host + off, cnt * 4, cudaMemcpyHostToDevice, m->stream2));
        WB_CUDA(cudaEventRecord(m->ev_plain[k], m->stream2));
        return WB_OK;
    };
    WB_CHECK(upload(0));
    if (!mel_dev && m->tr_mel_cap < (size_t)std::min(eb, n) * mel_per) {  // log-mel of one sub-batch
        cudaFree(m->tr_mel);
        m->tr_mel = nullptr, m->tr_mel_cap = 0;
        if (cudaMalloc((void **)&m->tr_mel, (size_t)eb * mel_per * 4) != cudaSuccess) {
            set_error("transcribe: cannot allocate the log-mel buffer");
            return WB_ERR_CUDA;
        }
        m->tr_mel_cap = (size_t)eb * mel_per;
    }
    Cache *c = m->tr_cache;  // reused across calls while the wave size stays the same
    int rc = WB_OK;
    size_t k = 0, wv = 0;
    for (int w0 = 0; w0 < n && rc == WB_OK; w0 += wave, wv++) {
        const int nb = std::min(wave, n - w0);
        if (!c || c->B != nb || c->T != T_cache || c->cross_impl != m->cross_impl) {
            if (c) cache_destroy(c);
            c = m->tr_cache = nullptr;
            rc = cache_create(m, nb, T_cache, false, m->decode_lanes, &c);
            if (rc != WB_OK) break;
            m->tr_cache = c;
        } else {
            rc = cache_reset(c);
            if (rc != WB_OK) break;
        }
        for (int i = w0; i < w0 + nb && rc == WB_OK; i += eb, k++) {
            const int ns = std::min(eb, w0 + nb - i);
            if (in_host) WB_CUDA(cudaStreamWaitEvent(st, m->ev_plain[k], 0));
            cudaEventRecord(m->ev_timed[3 * k], st);
            const float *mel_s = mel_dev ? mel_dev + (size_t)i * mel_per : m->tr_mel;
            if (!mel_dev) rc = model_logmel(m, pcm_dev + (size_t)i * m->n_samples, ns, m->tr_mel);
            cudaEventRecord(m->ev_timed[3 * k + 1], st);
            if (rc == WB_OK) rc = encode_batch(m, mel_s, ns, nullptr, c, i - w0);
            cudaEventRecord(m->ev_timed[3 * k + 2], st);
            if (rc == WB_OK) rc = upload(k + 1);
        }
        if (rc != WB_OK) break;
        c->has_cross = true;
        if (m->stop_sched) {  // forced lengths of this wave's chunks (chunks past the schedule: none)
            const int have = std::max(0, std::min(nb, m->stop_sched_n - w0));
            if (have) cudaMemcpyAsync(c->stop_at, m->stop_sched + w0, (size_t)have * sizeof(int), cudaMemcpyDeviceToDevice, st);
        }
        cudaEvent_t d0 = m->ev_timed[3 * subs.size() + 2 * wv], d1 = m->ev_timed[3 * subs.size() + 2 * wv + 1];
        cudaEventRecord(d0, st);
        rc = greedy_loop(c);
        if (rc != WB_OK) break;
        cudaEventRecord(d1, st);
        cudaMemcpyAsync(out_tokens_dev + (size_t)w0 * T_out, c->tokens_out, (size_t)nb * T_out * 4,
                        cudaMemcpyDeviceToDevice, st);
        cudaMemcpyAsync(out_len_dev + w0, c->out_len, (size_t)nb * 4, cudaMemcpyDeviceToDevice, st);
        if (cudaStreamSynchronize(st) != cudaSuccess) {  // the next wave reuses the cache
            set_error("transcribe: %s", cudaGetErrorString(cudaGetLastError()));
            rc = WB_ERR_CUDA;
            break;
        }
    }
    if (rc == WB_OK) {
        float acc[3] = {0, 0, 0}, ms = 0.f;
        for (size_t i = 0; i < subs.size(); i++) {
            if (cudaEventElapsedTime(&ms, m->ev_timed[3 * i], m->ev_timed[3 * i + 1]) == cudaSuccess) acc[0] += ms;
            if (cudaEventElapsedTime(&ms, m->ev_timed[3 * i + 1], m->ev_timed[3 * i + 2]) == cudaSuccess) acc[1] += ms;
        }
        for (size_t i = 0; i < n_waves; i++)
            if (cudaEventElapsedTime(&ms, m->ev_timed[3 * subs.size() + 2 * i], m->ev_timed[3 * subs.size() + 2 * i + 1]) ==
                cudaSuccess)
                acc[2] += ms;
        // cross-K/V projection time is part of encode_batch; report it inside [1] and leave [2] = 0
        m->timing[0] = acc[0], m->timing[1] = acc[1], m->timing[2] = 0.f, m->timing[3] = acc[2];
        m->timing[4] = acc[0] + acc[1] + acc[2];
        KernelTimer &t = m->cross_timer;
        for (int k = 0; k < TK_COUNT; k++) t.total_ms[k] = 0.f, t.launches[k] = 0;
        for (int i = 0; i + 1 < t.used; i += 2) {
            float ems = 0.f;
            if (cudaEventElapsedTime(&ems, t.ev[i], t.ev[i + 1]) == cudaSuccess)
                t.total_ms[t.cat[i / 2]] += ems, t.launches[t.cat[i / 2]]++;
        }
    }
    return rc;
}

// ---------------------------------------------------------------------------------------------
// step-wise API and teacher forcing (tests / WhisperDecoder.forward mirror)
// ---------------------------------------------------------------------------------------------
static int set_step_state(Cache *c, int cur_len, int pos, const int32_t *tokens_host) {
    Model *m = c->m;
    int sc[2] = {cur_len, pos};
    WB_CUDA(cudaMemcpyAsync(c->lanes[0].g.scalars, sc, sizeof sc, cudaMemcpyHostToDevice, m->stream));
    WB_CUDA(cudaMemcpyAsync(c->cur_tok, tokens_host, (size_t)c->B * 4, cudaMemcpyHostToDevice, m->stream));
    WB_CUDA(cudaStreamSynchronize(m->stream));  // sc is a stack variable
    return WB_OK;
}

// The step-wise entry points use caches created with a single lane.
int cache_step_api(Cache *c, const int32_t *tokens_host, int start_pos, float *logits_host, int32_t *next_host) {
    Model *m = c->m;
    WB_ARG(m->loaded && c->has_cross, "decode_step: weights / encoder output not set");
    WB_ARG(c->lanes.size() == 1, "decode_step: cache was not created through wm_kvcache_create");
    WB_ARG(tokens_host, "decode_step: null tokens");
    WB_ARG(c->host_len < c->T, "decode_step: cache full (%d)", c->T);
    WB_ARG(start_pos >= 0 && start_pos < m->T, "decode_step: start_pos %d out of range", start_pos);
    Lane &ln = c->lanes[0];
    WB_ARG(!logits_host || ln.logits, "decode_step: cache was created without a logits buffer");
    WB_CHECK(set_step_state(c, c->host_len, start_pos, tokens_host));
    WB_CHECK(prepare_chain_plans(c));
    WB_CHECK(decode_step(c, ln, m->stream, true, logits_host != nullptr, false));
    c->host_len++;
    if (logits_host)
        WB_CUDA(cudaMemcpyAsync(logits_host, ln.logits, (size_t)c->B * m->V * 4, cudaMemcpyDeviceToHost, m->stream));
    if (next_host) WB_CUDA(cudaMemcpyAsync(next_host, ln.next, (size_t)c->B * 4, cudaMemcpyDeviceToHost, m->stream));
    WB_CUDA(cudaStreamSynchronize(m->stream));
    return WB_OK;
}

int model_teacher_forced(Model *m, const float *enc_out_dev, int n, const int32_t *forced_host, int n_forced,
                         float *logits_host) {
    WB_ARG(m->loaded, "teacher_forced before weights are loaded");
    WB_ARG(n > 0 && n_forced >= 4 && n_forced <= m->T && enc_out_dev && forced_host && logits_host,
           "teacher_forced: bad arguments");
    Cache *c = nullptr;
    WB_CHECK(cache_create(m, n, std::min(m->T, (n_forced + 7) & ~7), true, 1, &c));
    Lane &ln = c->lanes[0];
    int rc = cache_set_encoder(c, enc_out_dev);
    if (rc == WB_OK) rc = prepare_chain_plans(c);
    std::vector<int32_t> col(n);
    for (int i = 0; i < n_forced && rc == WB_OK; i++) {
        for (int b = 0; b < n; b++) col[b] = forced_host[(size_t)b * n_forced + i];
        const int pos = i < 4 ? i : i - m->cfg.pos_quirk;  // whisper.mojo:196,217
        rc = set_step_state(c, i, pos, col.data());
        if (rc == WB_OK) rc = decode_step(c, ln, m->stream, i >= 3, i >= 3, false);
        if (rc == WB_OK && i >= 3) {
            // logits_host is [n][n_forced-3][V]; this step is row i-3 of every chunk
            cudaError_t e = cudaMemcpy2DAsync(logits_host + (size_t)(i - 3) * m->V,
                                              (size_t)(n_forced - 3) * m->V * 4, ln.logits, (size_t)m->V * 4,
                                              (size_t)m->V * 4, n, cudaMemcpyDeviceToHost, m->stream);
            if (e != cudaSuccess || cudaStreamSynchronize(m->stream) != cudaSuccess) {
                set_error("teacher_forced: copy failed: %s", cudaGetErrorString(cudaGetLastError()));
                rc = WB_ERR_CUDA;
            }
        }
    }
    cache_destroy(c);
    return rc;
}

}  // namespace wb
