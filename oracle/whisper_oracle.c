/*
 * whisper_oracle.c -- CPU fp32 restatement of antonvice/whisper.Mojo's transcription path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under whisper_mojo_b200/ (the product) may import, link or
 * execute this file.  Allowed users: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg
 * and `bench.py --impl reference`.
 *
 * Every routine follows the cited reference lines (paths relative to /root/reference):
 * the same loop nests, the same lane-strided partial sums (SIMD width = WO_W lanes, reduced as a
 * halving tree), one-pass LayerNorm variance, tanh-GELU with the reference's truncated constants,
 * -1e10 mask fill, lowest-index argmax, the positional off-by-one of the decode loop, and the
 * 4-id prompt / 195-iteration cap.  Dimensions the reference hard-codes (384, 1500, 51865, 448,
 * 80, K=3) are runtime config here so the Small-shaped and micro test configs run through the
 * same code; with the Tiny config it is the reference algorithm verbatim.
 *
 * What is NOT restated (third-party, source absent from /root/reference): Modular MAX
 * `linalg.matmul` 25.7.0 (whisper_tensor.mojo:74-146).  Its contract is C = A*B^T (+bias), the same
 * as the in-repo `matmul` the reference falls back to (layers.mojo:120-123); this file uses the
 * in-repo algorithm for those shapes too.  Mojo's SIMD `exp`/`tanh` are polynomial kernels; libm
 * expf/tanhf are used here (differences are ~1 ulp).
 *
 * Pinning status: the only golden vector the reference ships (expected_tokens.txt) needs the
 * un-shipped whisper_tiny_weights.bin + sample_input.bin, so it cannot be replayed here
 * ("parity unpinned" against the reference's own output).  The restatement is instead
 * cross-validated against an independent implementation (HF transformers Whisper with tanh GELU and
 * the same position shift; oracle/make_golden.py, tests/golden/).
 *
 * Build: see oracle/Makefile (gcc -O3 -mavx2 -ffp-contract=off -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define WO_W 8 /* simdwidthof[f32] stand-in; layers.mojo/whisper.mojo hard-code 8 as well */

typedef float v8f __attribute__((vector_size(32), aligned(4)));

static inline v8f ld8(const float *p) { v8f v; memcpy(&v, p, sizeof v); return v; }
static inline void st8(float *p, v8f v) { memcpy(p, &v, sizeof v); }
static inline v8f splat8(float x) { v8f v = {x, x, x, x, x, x, x, x}; return v; }
/* SIMD.reduce_add as a halving tree: (l[i]+l[i+4]) -> (+2) -> (+1) */
static inline float hsum8(v8f v) {
    float a0 = v[0] + v[4], a1 = v[1] + v[5], a2 = v[2] + v[6], a3 = v[3] + v[7];
    float b0 = a0 + a2, b1 = a1 + a3;
    return b0 + b1;
}
static inline float hmax8(v8f v) {
    float m = v[0];
    for (int i = 1; i < 8; i++) m = v[i] > m ? v[i] : m;
    return m;
}

/* ------------------------------------------------------------------------------------------- */
/* L2 ops (whisper_tensor.mojo)                                                                */
/* ------------------------------------------------------------------------------------------- */

/* whisper_tensor.mojo:151-246  C[M,N] = A[M,K] * B[N,K]^T (+ bias[N]); bias may be NULL. */
void wo_matmul(float *C, const float *A, const float *B, const float *bias, int M, int N, int K) {
    const int Kr = (K / WO_W) * WO_W;
    if (M <= 4) { /* :158-175 vector path, parallel over N */
#pragma omp parallel for schedule(static)
        for (int n = 0; n < N; n++) {
            for (int m = 0; m < M; m++) {
                v8f sum = splat8(0.f);
                const float *a = A + (size_t)m * K, *b = B + (size_t)n * K;
                for (int k = 0; k < Kr; k += WO_W) sum += ld8(a + k) * ld8(b + k);
                float f = hsum8(sum);
                for (int k = Kr; k < K; k++) f += a[k] * b[k];
                if (bias) f += bias[n];
                C[(size_t)m * N + n] = f;
            }
        }
        return;
    }
    /* :176-246 matrix path, parallel over M, 8-wide N tile */
#pragma omp parallel for schedule(static)
    for (int m = 0; m < M; m++) {
        const float *a = A + (size_t)m * K;
        int n0 = 0;
        for (; n0 + 8 <= N; n0 += 8) {
            v8f s[8];
            for (int t = 0; t < 8; t++) s[t] = splat8(0.f);
            for (int k = 0; k < Kr; k += WO_W) {
                v8f av = ld8(a + k);
                for (int t = 0; t < 8; t++) s[t] += av * ld8(B + (size_t)(n0 + t) * K + k);
            }
            float f[8];
            for (int t = 0; t < 8; t++) f[t] = hsum8(s[t]);
            for (int k = Kr; k < K; k++) {
                float av = a[k];
                for (int t = 0; t < 8; t++) f[t] += av * B[(size_t)(n0 + t) * K + k];
            }
            for (int t = 0; t < 8; t++) C[(size_t)m * N + n0 + t] = f[t] + (bias ? bias[n0 + t] : 0.f);
        }
        for (int n = n0; n < N; n++) { /* :235-244 remainder columns */
            v8f sum = splat8(0.f);
            const float *b = B + (size_t)n * K;
            for (int k = 0; k < Kr; k += WO_W) sum += ld8(a + k) * ld8(b + k);
            float f = hsum8(sum);
            for (int k = Kr; k < K; k++) f += a[k] * b[k];
            C[(size_t)m * N + n] = f + (bias ? bias[n] : 0.f);
        }
    }
}

/* whisper_tensor.mojo:249-285  one-pass mean / E[x^2]-mean^2; cols must be a multiple of 8
 * (the reference reads past the row otherwise). */
void wo_layer_norm(float *out, const float *inp, const float *gamma, const float *beta, int rows, int cols,
                   float eps) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < rows; i++) {
        const float *x = inp + (size_t)i * cols;
        v8f s = splat8(0.f), q = splat8(0.f);
        for (int j = 0; j < cols; j += WO_W) {
            v8f v = ld8(x + j);
            s += v;
            q += v * v;
        }
        float sum = hsum8(s), sq = hsum8(q);
        float mean = sum / (float)cols;
        float var = (sq / (float)cols) - (mean * mean);
        float inv_std = 1.0f / sqrtf(var + eps);
        v8f vm = splat8(mean), vi = splat8(inv_std);
        for (int j = 0; j < cols; j += WO_W) {
            v8f r = (ld8(x + j) - vm) * vi * ld8(gamma + j) + ld8(beta + j);
            st8(out + (size_t)i * cols + j, r);
        }
    }
}

/* whisper_tensor.mojo:288-308  tanh GELU, constants as written there; only size//W vectors. */
void wo_gelu(float *t, size_t size) {
    const float SQRT_2_PI = 0.79788456f, COEFF = 0.044715f;
    size_t nvec = size / WO_W;
#pragma omp parallel for schedule(static)
    for (size_t b = 0; b < nvec; b++) {
        float *p = t + b * WO_W;
        for (int l = 0; l < WO_W; l++) {
            float x = p[l];
            float x3 = x * x * x;
            float inner = SQRT_2_PI * (x + COEFF * x3);
            p[l] = 0.5f * x * (1.0f + tanhf(inner));
        }
    }
}

/* whisper_tensor.mojo:311-355  row softmax in place (SIMD body + scalar tail). */
void wo_softmax(float *t, int rows, int cols) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < rows; i++) {
        float *r = t + (size_t)i * cols;
        const int body = cols >= WO_W ? cols - cols % WO_W : 0;
        float max_val = r[0];
        if (cols >= WO_W) {
            v8f mv = splat8(max_val);
            for (int j = 0; j < body; j += WO_W) {
                v8f v = ld8(r + j);
                for (int l = 0; l < 8; l++) mv[l] = v[l] > mv[l] ? v[l] : mv[l];
            }
            max_val = hmax8(mv);
        }
        for (int j = body; j < cols; j++)
            if (r[j] > max_val) max_val = r[j];
        float sum_exp = 0.f;
        if (cols >= WO_W) {
            v8f sv = splat8(0.f);
            for (int j = 0; j < body; j += WO_W) {
                v8f e;
                for (int l = 0; l < 8; l++) e[l] = expf(r[j + l] - max_val);
                st8(r + j, e);
                sv += e;
            }
            sum_exp = hsum8(sv);
        }
        for (int j = body; j < cols; j++) {
            float e = expf(r[j] - max_val);
            r[j] = e;
            sum_exp += e;
        }
        for (int j = 0; j < cols; j++) r[j] = r[j] / sum_exp;
    }
}

/* whisper_tensor.mojo:358-364  w[co][ci][k] -> w'[(co*K+k)][ci] */
void wo_transpose_conv_weights(float *new_w, const float *w, int C_out, int C_in, int K) {
    for (int co = 0; co < C_out; co++)
        for (int ci = 0; ci < C_in; ci++)
            for (int k = 0; k < K; k++)
                new_w[((size_t)co * K + k) * C_in + ci] = w[(size_t)co * (C_in * K) + (size_t)ci * K + k];
}

/* whisper_tensor.mojo:367-428  K=3 conv, zero padding, weight in the transposed layout above.
 * out_T=0 -> out[C_out][L_out]; out_T=1 -> out[L_out][C_out].  C_in must be a multiple of 8. */
void wo_conv1d(float *out, const float *inp, const float *weight, const float *bias, int C_in, int L_in, int C_out,
               int stride, int padding, int out_T) {
    const int K = 3;
    const int L_out = (L_in + 2 * padding - K) / stride + 1;
    float *inp_T = (float *)malloc(sizeof(float) * (size_t)L_in * C_in); /* :383-388 */
#pragma omp parallel for schedule(static)
    for (int li = 0; li < L_in; li++)
        for (int ci = 0; ci < C_in; ci++) inp_T[(size_t)li * C_in + ci] = inp[(size_t)ci * L_in + li];
#pragma omp parallel for schedule(static)
    for (int co = 0; co < C_out; co++) {
        float b_val = bias[co];
        size_t w_base = (size_t)co * 3 * C_in;
        for (int lo = 0; lo < L_out; lo++) {
            v8f dot = splat8(0.f);
            int start_l = lo * stride - padding;
            for (int k = 0; k < K; k++) {
                int li = start_l + k;
                if (li >= 0 && li < L_in) {
                    const float *ip = inp_T + (size_t)li * C_in;
                    const float *wp = weight + w_base + (size_t)k * C_in;
                    for (int ci = 0; ci < C_in; ci += WO_W) dot += ld8(ip + ci) * ld8(wp + ci);
                }
            }
            float r = hsum8(dot) + b_val;
            if (out_T) out[(size_t)lo * C_out + co] = r;
            else out[(size_t)co * L_out + lo] = r;
        }
    }
    free(inp_T);
}

/* whisper_tensor.mojo:431-439  first maximum wins */
int wo_argmax(const float *t, int size) {
    float max_val = t[0];
    int max_idx = 0;
    for (int i = 1; i < size; i++) {
        if (t[i] > max_val) {
            max_val = t[i];
            max_idx = i;
        }
    }
    return max_idx;
}

/* ------------------------------------------------------------------------------------------- */
/* Model structures                                                                            */
/* ------------------------------------------------------------------------------------------- */

typedef struct {
    int d_model, n_heads, n_layers, vocab;
    int n_audio_ctx; /* 1500 */
    int n_text_ctx;  /* 448: decoder pos_emb rows and KVCache max_len (whisper.mojo:124,193) */
    int n_mels;      /* 80 */
} wo_config;

typedef struct { /* layers.mojo:72-103 */
    const float *q_w, *q_b, *k_w, *v_w, *v_b, *o_w, *o_b;
} wo_attn;

typedef struct { /* layers.mojo:386-433 */
    wo_attn attn;
    const float *attn_ln_w, *attn_ln_b;
    wo_attn cross;
    const float *cross_ln_w, *cross_ln_b;
    const float *fc1_w, *fc1_b, *fc2_w, *fc2_b, *mlp_ln_w, *mlp_ln_b;
    int is_decoder;
} wo_block;

typedef struct {
    wo_config cfg;
    float *conv1_wT, *conv2_wT; /* transposed at load, whisper.mojo:61,63 */
    const float *conv1_b, *conv2_b, *enc_pos;
    wo_block *enc_blocks;
    const float *enc_ln_w, *enc_ln_b;
    const float *token_emb, *dec_pos;
    wo_block *dec_blocks;
    const float *dec_ln_w, *dec_ln_b;
} wo_model;

typedef struct { /* layers.mojo:14-36 */
    float *self_k, *self_v, *cross_k, *cross_v;
    int current_len, has_cross;
} wo_layer_cache;

typedef struct { /* layers.mojo:55-63 */
    wo_layer_cache *layers;
    int n_layers, d_model, max_len, n_audio_ctx;
} wo_kvcache;

/* Number of fp32 values in the flat weight file for a config (export_weights.py:19-90). */
long long wo_weight_count(const wo_config *c) {
    long long D = c->d_model, F = 4 * D, n = 0;
    n += D * c->n_mels * 3 + D + D * D * 3 + D + (long long)c->n_audio_ctx * D;
    long long attn = D * D + D + D * D + D * D + D + D * D + D;
    long long mlp = F * D + F + D * F + D;
    n += c->n_layers * (attn + 2 * D + mlp + 2 * D) + 2 * D;
    n += (long long)c->vocab * D + (long long)c->n_text_ctx * D;
    n += c->n_layers * (attn + 2 * D + attn + 2 * D + mlp + 2 * D) + 2 * D;
    return n;
}

static const float *take(const float **cur, long long n) { /* loader.mojo:21-27 (without the copy) */
    const float *p = *cur;
    *cur += n;
    return p;
}
static void load_attn(wo_attn *a, const float **cur, int D) { /* layers.mojo:96-103 */
    a->q_w = take(cur, (long long)D * D);
    a->q_b = take(cur, D);
    a->k_w = take(cur, (long long)D * D);
    a->v_w = take(cur, (long long)D * D);
    a->v_b = take(cur, D);
    a->o_w = take(cur, (long long)D * D);
    a->o_b = take(cur, D);
}
static void load_block(wo_block *b, const float **cur, int D, int is_decoder) { /* layers.mojo:418-433 */
    memset(b, 0, sizeof *b);
    b->is_decoder = is_decoder;
    load_attn(&b->attn, cur, D);
    b->attn_ln_w = take(cur, D);
    b->attn_ln_b = take(cur, D);
    if (is_decoder) {
        load_attn(&b->cross, cur, D);
        b->cross_ln_w = take(cur, D);
        b->cross_ln_b = take(cur, D);
    }
    b->fc1_w = take(cur, 4LL * D * D);
    b->fc1_b = take(cur, 4LL * D);
    b->fc2_w = take(cur, 4LL * D * D);
    b->fc2_b = take(cur, D);
    b->mlp_ln_w = take(cur, D);
    b->mlp_ln_b = take(cur, D);
}

/* whisper.mojo:60-69,122-128.  `weights` must stay alive while the model is used.
 * Returns NULL when n_floats is not exactly the size the config implies. */
wo_model *wo_create(const wo_config *cfg, const float *weights, long long n_floats) {
    if (n_floats != wo_weight_count(cfg)) return NULL;
    if (cfg->d_model % cfg->n_heads || (cfg->d_model / cfg->n_heads) != 64 || cfg->n_mels % 8) return NULL;
    wo_model *m = (wo_model *)calloc(1, sizeof *m);
    m->cfg = *cfg;
    const int D = cfg->d_model, L = cfg->n_layers;
    const float *cur = weights;
    const float *c1 = take(&cur, (long long)D * cfg->n_mels * 3);
    m->conv1_b = take(&cur, D);
    const float *c2 = take(&cur, (long long)D * D * 3);
    m->conv2_b = take(&cur, D);
    m->conv1_wT = (float *)malloc(sizeof(float) * (size_t)D * cfg->n_mels * 3);
    m->conv2_wT = (float *)malloc(sizeof(float) * (size_t)D * D * 3);
    wo_transpose_conv_weights(m->conv1_wT, c1, D, cfg->n_mels, 3);
    wo_transpose_conv_weights(m->conv2_wT, c2, D, D, 3);
    m->enc_pos = take(&cur, (long long)cfg->n_audio_ctx * D);
    m->enc_blocks = (wo_block *)calloc(L, sizeof(wo_block));
    for (int i = 0; i < L; i++) load_block(&m->enc_blocks[i], &cur, D, 0);
    m->enc_ln_w = take(&cur, D);
    m->enc_ln_b = take(&cur, D);
    m->token_emb = take(&cur, (long long)cfg->vocab * D);
    m->dec_pos = take(&cur, (long long)cfg->n_text_ctx * D);
    m->dec_blocks = (wo_block *)calloc(L, sizeof(wo_block));
    for (int i = 0; i < L; i++) load_block(&m->dec_blocks[i], &cur, D, 1);
    m->dec_ln_w = take(&cur, D);
    m->dec_ln_b = take(&cur, D);
    return m;
}

void wo_destroy(wo_model *m) {
    if (!m) return;
    free(m->conv1_wT);
    free(m->conv2_wT);
    free(m->enc_blocks);
    free(m->dec_blocks);
    free(m);
}

wo_kvcache *wo_kvcache_create(const wo_model *m, int max_len) { /* layers.mojo:30-36,58-63 */
    wo_kvcache *c = (wo_kvcache *)calloc(1, sizeof *c);
    c->n_layers = m->cfg.n_layers;
    c->d_model = m->cfg.d_model;
    c->max_len = max_len;
    c->n_audio_ctx = m->cfg.n_audio_ctx;
    c->layers = (wo_layer_cache *)calloc(c->n_layers, sizeof(wo_layer_cache));
    for (int i = 0; i < c->n_layers; i++) {
        wo_layer_cache *l = &c->layers[i];
        l->self_k = (float *)calloc((size_t)max_len * c->d_model, sizeof(float));
        l->self_v = (float *)calloc((size_t)max_len * c->d_model, sizeof(float));
        l->cross_k = (float *)calloc((size_t)c->n_audio_ctx * c->d_model, sizeof(float));
        l->cross_v = (float *)calloc((size_t)c->n_audio_ctx * c->d_model, sizeof(float));
    }
    return c;
}
void wo_kvcache_destroy(wo_kvcache *c) {
    if (!c) return;
    for (int i = 0; i < c->n_layers; i++) {
        free(c->layers[i].self_k);
        free(c->layers[i].self_v);
        free(c->layers[i].cross_k);
        free(c->layers[i].cross_v);
    }
    free(c->layers);
    free(c);
}
int wo_kvcache_len(const wo_kvcache *c) { return c->layers[0].current_len; }
const float *wo_kvcache_ptr(const wo_kvcache *c, int layer, int which) {
    const wo_layer_cache *l = &c->layers[layer];
    return which == 0 ? l->self_k : which == 1 ? l->self_v : which == 2 ? l->cross_k : l->cross_v;
}

/* ------------------------------------------------------------------------------------------- */
/* L3: attention + block (layers.mojo)                                                         */
/* ------------------------------------------------------------------------------------------- */

/* layers.mojo:186-272  single-query path for one head */
static void mha_decode_head(float *out, const float *q, const float *k, const float *v, int h, int d_model,
                            int head_dim, int k_len, int mask, int mask_limit) {
    const float scale = 1.0f / sqrtf((float)head_dim);
    float *scores = (float *)malloc(sizeof(float) * (size_t)(k_len > 1500 ? k_len : 1500));
    float max_score = -1e10f;
    v8f qv[8];
    for (int c = 0; c < 8; c++) qv[c] = ld8(q + h * head_dim + 8 * c);
    for (int j = 0; j < k_len; j++) {
        const float *kp = k + (size_t)j * d_model + h * head_dim;
        v8f dot = qv[0] * ld8(kp);
        for (int c = 1; c < 8; c++) dot += qv[c] * ld8(kp + 8 * c);
        float score = hsum8(dot) * scale;
        if (mask && j > mask_limit) score = -1e10f;
        if (score > max_score) max_score = score;
        scores[j] = score;
    }
    v8f se = splat8(0.f);
    const int rounded = (k_len / 8) * 8;
    for (int j = 0; j < rounded; j += 8) {
        v8f e;
        for (int l = 0; l < 8; l++) e[l] = expf(scores[j + l] - max_score);
        st8(scores + j, e);
        se += e;
    }
    float sum_exp = hsum8(se);
    for (int j = rounded; j < k_len; j++) {
        float e = expf(scores[j] - max_score);
        scores[j] = e;
        sum_exp += e;
    }
    float inv = 1.0f / sum_exp;
    for (int j = 0; j < k_len; j++) scores[j] *= inv;
    v8f o[8];
    for (int c = 0; c < 8; c++) o[c] = splat8(0.f);
    for (int j = 0; j < k_len; j++) {
        v8f s = splat8(scores[j]);
        const float *vp = v + (size_t)j * d_model + h * head_dim;
        for (int c = 0; c < 8; c++) o[c] += s * ld8(vp + 8 * c);
    }
    for (int c = 0; c < 8; c++) st8(out + h * head_dim + 8 * c, o[c]);
    free(scores);
}

/* layers.mojo:273-342  block path for one head (gather, QK^T, scale+mask, softmax, V^T, PV, scatter) */
static void mha_block_head(float *out, const float *q, const float *k, const float *v, int h, int d_model,
                           int head_dim, int q_len, int k_len, int mask, int mask_base) {
    const float scale = 1.0f / sqrtf((float)head_dim);
    float *q_h = (float *)malloc(sizeof(float) * (size_t)q_len * head_dim);
    float *k_h = (float *)malloc(sizeof(float) * (size_t)k_len * head_dim);
    float *v_hT = (float *)malloc(sizeof(float) * (size_t)k_len * head_dim);
    float *scores = (float *)malloc(sizeof(float) * (size_t)q_len * k_len);
    float *out_h = (float *)malloc(sizeof(float) * (size_t)q_len * head_dim);
    for (int i = 0; i < q_len; i++)
        memcpy(q_h + (size_t)i * head_dim, q + (size_t)i * d_model + h * head_dim, sizeof(float) * head_dim);
    for (int i = 0; i < k_len; i++) {
        memcpy(k_h + (size_t)i * head_dim, k + (size_t)i * d_model + h * head_dim, sizeof(float) * head_dim);
        for (int j = 0; j < head_dim; j++) /* :324-327 */
            v_hT[(size_t)j * k_len + i] = v[(size_t)i * d_model + h * head_dim + j];
    }
    wo_matmul(scores, q_h, k_h, NULL, q_len, k_len, head_dim);
    for (int i = 0; i < q_len; i++) { /* :304-320; threshold = current_len - q_len + i or i */
        float *r = scores + (size_t)i * k_len;
        for (int j = 0; j < k_len; j++) {
            float s = r[j] * scale;
            if (mask && j > mask_base + i) s = -1e10f;
            r[j] = s;
        }
    }
    wo_softmax(scores, q_len, k_len);
    wo_matmul(out_h, scores, v_hT, NULL, q_len, head_dim, k_len);
    for (int i = 0; i < q_len; i++)
        memcpy(out + (size_t)i * d_model + h * head_dim, out_h + (size_t)i * head_dim, sizeof(float) * head_dim);
    free(q_h);
    free(k_h);
    free(v_hT);
    free(scores);
    free(out_h);
}

/* layers.mojo:105-359.  `cache` may be NULL when use_cache == 0. */
static void mha_forward(float *final_out, const wo_attn *w, const float *query, const float *key, const float *value,
                        int q_len, int k_len_in, int d_model, int n_heads, int mask, wo_layer_cache *cache,
                        int is_self_attn, int use_cache) {
    const int head_dim = d_model / n_heads;
    float *q = (float *)malloc(sizeof(float) * (size_t)q_len * d_model);
    wo_matmul(q, query, w->q_w, w->q_b, q_len, d_model, d_model);
    const float *k, *v;
    float *own_k = NULL, *own_v = NULL;
    int k_len;
    if (use_cache) {
        if (is_self_attn) { /* :131-147 */
            float *new_k = (float *)malloc(sizeof(float) * (size_t)q_len * d_model);
            float *new_v = (float *)malloc(sizeof(float) * (size_t)q_len * d_model);
            wo_matmul(new_k, key, w->k_w, NULL, q_len, d_model, d_model);
            wo_matmul(new_v, value, w->v_w, w->v_b, q_len, d_model, d_model);
            size_t off = (size_t)cache->current_len * d_model;
            memcpy(cache->self_k + off, new_k, sizeof(float) * (size_t)q_len * d_model);
            memcpy(cache->self_v + off, new_v, sizeof(float) * (size_t)q_len * d_model);
            cache->current_len += q_len;
            free(new_k);
            free(new_v);
            k = cache->self_k;
            v = cache->self_v;
            k_len = cache->current_len;
        } else { /* :148-157 */
            if (!cache->has_cross) {
                wo_matmul(cache->cross_k, key, w->k_w, NULL, k_len_in, d_model, d_model);
                wo_matmul(cache->cross_v, value, w->v_w, w->v_b, k_len_in, d_model, d_model);
                cache->has_cross = 1;
            }
            k = cache->cross_k;
            v = cache->cross_v;
            k_len = k_len_in;
        }
    } else { /* :158-176 */
        own_k = (float *)malloc(sizeof(float) * (size_t)k_len_in * d_model);
        own_v = (float *)malloc(sizeof(float) * (size_t)k_len_in * d_model);
        wo_matmul(own_k, key, w->k_w, NULL, k_len_in, d_model, d_model);
        wo_matmul(own_v, value, w->v_w, w->v_b, k_len_in, d_model, d_model);
        k = own_k;
        v = own_v;
        k_len = k_len_in;
    }
    float *out = (float *)calloc((size_t)q_len * d_model, sizeof(float));
    const int cached_self = use_cache && is_self_attn;
    if (q_len == 1) { /* :344-346 heads serial */
        int limit = cached_self ? cache->current_len - 1 : 0; /* :213 */
        for (int h = 0; h < n_heads; h++) mha_decode_head(out, q, k, v, h, d_model, head_dim, k_len, mask, limit);
    } else {
        int base = cached_self ? cache->current_len - q_len : 0; /* :311,317 */
        for (int h = 0; h < n_heads; h++)
            mha_block_head(out, q, k, v, h, d_model, head_dim, q_len, k_len, mask, base);
    }
    wo_matmul(final_out, out, w->o_w, w->o_b, q_len, d_model, d_model);
    free(q);
    free(out);
    free(own_k);
    free(own_v);
}

static void add_rows(float *dst, const float *a, const float *b, size_t n) { /* layers.mojo:455-461 etc. */
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) dst[i] = a[i] + b[i];
}

/* layers.mojo:435-519.  x is [rows, D] in, result written to `y` (may not alias x). */
static void block_forward(float *y, const wo_block *b, const float *x, int rows, const float *enc_out,
                          int enc_rows, wo_layer_cache *cache, int use_cache, int D, int H) {
    const size_t n = (size_t)rows * D;
    float *x_norm = (float *)malloc(sizeof(float) * n);
    float *attn_out = (float *)malloc(sizeof(float) * n);
    float *cur = (float *)malloc(sizeof(float) * n);
    wo_layer_norm(x_norm, x, b->attn_ln_w, b->attn_ln_b, rows, D, 1e-5f);
    mha_forward(attn_out, &b->attn, x_norm, x_norm, x_norm, rows, rows, D, H, b->is_decoder, cache, 1, use_cache);
    add_rows(cur, x, attn_out, n);
    if (b->is_decoder && enc_out && enc_rows > 0) { /* :463-488 */
        wo_layer_norm(x_norm, cur, b->cross_ln_w, b->cross_ln_b, rows, D, 1e-5f);
        mha_forward(attn_out, &b->cross, x_norm, enc_out, enc_out, rows, enc_rows, D, H, 0, cache, 0, use_cache);
        add_rows(cur, cur, attn_out, n);
    }
    wo_layer_norm(x_norm, cur, b->mlp_ln_w, b->mlp_ln_b, rows, D, 1e-5f);
    float *hidden = (float *)malloc(sizeof(float) * (size_t)rows * 4 * D);
    wo_matmul(hidden, x_norm, b->fc1_w, b->fc1_b, rows, 4 * D, D);
    wo_gelu(hidden, (size_t)rows * 4 * D);
    wo_matmul(attn_out, hidden, b->fc2_w, b->fc2_b, rows, D, 4 * D);
    add_rows(y, cur, attn_out, n);
    free(hidden);
    free(x_norm);
    free(attn_out);
    free(cur);
}

/* ------------------------------------------------------------------------------------------- */
/* L4: encoder, decoder, greedy loop (whisper.mojo)                                            */
/* ------------------------------------------------------------------------------------------- */

/* whisper.mojo:71-99.  mel [n_mels, 2*n_audio_ctx] -> enc_out [n_audio_ctx, D].
 * Optional taps (may be NULL): conv1_out [D, 2*ctx] after GELU, conv2_out [ctx, D] after GELU. */
void wo_encode_taps(const wo_model *m, const float *mel, float *enc_out, float *conv1_tap, float *conv2_tap) {
    const int D = m->cfg.d_model, S = m->cfg.n_audio_ctx, Lin = 2 * S;
    float *x1 = (float *)malloc(sizeof(float) * (size_t)D * Lin);
    float *x2 = (float *)malloc(sizeof(float) * (size_t)S * D);
    float *x = (float *)malloc(sizeof(float) * (size_t)S * D);
    wo_conv1d(x1, mel, m->conv1_wT, m->conv1_b, m->cfg.n_mels, Lin, D, 1, 1, 0);
    wo_gelu(x1, (size_t)D * Lin);
    if (conv1_tap) memcpy(conv1_tap, x1, sizeof(float) * (size_t)D * Lin);
    wo_conv1d(x2, x1, m->conv2_wT, m->conv2_b, D, Lin, D, 2, 1, 1);
    wo_gelu(x2, (size_t)S * D);
    if (conv2_tap) memcpy(conv2_tap, x2, sizeof(float) * (size_t)S * D);
    add_rows(x, x2, m->enc_pos, (size_t)S * D);
    for (int i = 0; i < m->cfg.n_layers; i++) {
        block_forward(x2, &m->enc_blocks[i], x, S, NULL, 0, NULL, 0, D, m->cfg.n_heads);
        float *t = x;
        x = x2;
        x2 = t;
    }
    wo_layer_norm(enc_out, x, m->enc_ln_w, m->enc_ln_b, S, D, 1e-5f);
    free(x1);
    free(x2);
    free(x);
}
void wo_encode(const wo_model *m, const float *mel, float *enc_out) { wo_encode_taps(m, mel, enc_out, NULL, NULL); }

/* whisper.mojo:130-167.  logits [vocab] for the LAST position; optional hidden tap [D]. */
void wo_decoder_forward(const wo_model *m, wo_kvcache *cache, const int *tokens, int n_tokens, const float *enc_out,
                        int use_cache, int start_pos, float *logits, float *hidden_tap) {
    const int D = m->cfg.d_model;
    float *x = (float *)malloc(sizeof(float) * (size_t)n_tokens * D);
    float *y = (float *)malloc(sizeof(float) * (size_t)n_tokens * D);
    for (int i = 0; i < n_tokens; i++) { /* :138-149 */
        const float *te = m->token_emb + (size_t)tokens[i] * D;
        const float *pe = m->dec_pos + (size_t)(start_pos + i) * D;
        for (int j = 0; j < D; j++) x[(size_t)i * D + j] = te[j] + pe[j];
    }
    for (int i = 0; i < m->cfg.n_layers; i++) {
        block_forward(y, &m->dec_blocks[i], x, n_tokens, enc_out, m->cfg.n_audio_ctx,
                      cache ? &cache->layers[i] : NULL, use_cache, D, m->cfg.n_heads);
        float *t = x;
        x = y;
        y = t;
    }
    wo_layer_norm(y, x, m->dec_ln_w, m->dec_ln_b, n_tokens, D, 1e-5f);
    const float *last = y + (size_t)(n_tokens - 1) * D;
    if (hidden_tap) memcpy(hidden_tap, last, sizeof(float) * D);
    wo_matmul(logits, last, m->token_emb, NULL, 1, m->cfg.vocab, D); /* :162-166 tied embedding */
    free(x);
    free(y);
}

/* Prompt and stop ids: whisper.mojo:187-191,206. */
static const int WO_PROMPT[4] = {50258, 50259, 50359, 50363};
#define WO_EOT 50257

/* whisper.mojo:184-223 with enc_out supplied.  pos_quirk=1 reproduces start_pos = current_len-1
 * (:217); pos_quirk=0 uses current_len (HF positions).  prompt/eot are parameters so configs with a
 * small vocabulary can run; pass NULL/-1 for the reference's ids.  Returns the sequence length. */
int wo_greedy(const wo_model *m, const float *enc_out, int *out_tokens, int pos_quirk, int max_iters,
              const int *prompt, int eot, float *margins) {
    const int V = m->cfg.vocab;
    if (!prompt) prompt = WO_PROMPT;
    if (eot < 0) eot = WO_EOT;
    wo_kvcache *cache = wo_kvcache_create(m, m->cfg.n_text_ctx);
    float *logits = (float *)malloc(sizeof(float) * V);
    int n = 0;
    for (int i = 0; i < 4; i++) out_tokens[n++] = prompt[i];
    wo_decoder_forward(m, cache, prompt, 4, enc_out, 1, 0, logits, NULL);
    int next = wo_argmax(logits, V);
    int step = 0;
    if (margins) {
        float best2 = -INFINITY;
        for (int i = 0; i < V; i++)
            if (i != next && logits[i] > best2) best2 = logits[i];
        margins[step] = logits[next] - best2;
    }
    out_tokens[n++] = next;
    for (int it = 0; it < max_iters; it++) {
        if (next == eot) break;
        int start_pos = cache->layers[0].current_len - (pos_quirk ? 1 : 0);
        wo_decoder_forward(m, cache, &next, 1, enc_out, 1, start_pos, logits, NULL);
        next = wo_argmax(logits, V);
        step++;
        if (margins) {
            float best2 = -INFINITY;
            for (int i = 0; i < V; i++)
                if (i != next && logits[i] > best2) best2 = logits[i];
            margins[step] = logits[next] - best2;
        }
        out_tokens[n++] = next;
    }
    free(logits);
    wo_kvcache_destroy(cache);
    return n;
}

/* whisper.mojo:184-223 end to end: mel -> tokens (<= 4 + 1 + max_iters). */
int wo_transcribe(const wo_model *m, const float *mel, int *out_tokens, int pos_quirk, int max_iters,
                  const int *prompt, int eot) {
    float *enc = (float *)malloc(sizeof(float) * (size_t)m->cfg.n_audio_ctx * m->cfg.d_model);
    wo_encode(m, mel, enc);
    int n = wo_greedy(m, enc, out_tokens, pos_quirk, max_iters, prompt, eot, NULL);
    free(enc);
    return n;
}

/* Teacher-forced decode (test helper built from the same decoder.forward calls the greedy loop
 * makes): prefill with forced[0:4], then feed forced[4:], using the greedy loop's start_pos rule.
 * logits_out receives n_forced-3 rows of `vocab` logits (row 0 = after prefill). */
void wo_teacher_forced(const wo_model *m, const float *enc_out, const int *forced, int n_forced, int pos_quirk,
                       float *logits_out) {
    const int V = m->cfg.vocab;
    wo_kvcache *cache = wo_kvcache_create(m, m->cfg.n_text_ctx);
    wo_decoder_forward(m, cache, forced, 4, enc_out, 1, 0, logits_out, NULL);
    for (int i = 4; i < n_forced; i++) {
        int start_pos = cache->layers[0].current_len - (pos_quirk ? 1 : 0);
        wo_decoder_forward(m, cache, forced + i, 1, enc_out, 1, start_pos, logits_out + (size_t)(i - 3) * V, NULL);
    }
    wo_kvcache_destroy(cache);
}

int wo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Explicit thread count for the timed CPU arms (bench.py): launchers such as torch.distributed.run export
 * OMP_NUM_THREADS=1 to their workers, which would silently serialise the parallelize() analogues above. */
void wo_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
