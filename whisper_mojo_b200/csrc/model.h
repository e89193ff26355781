// model.h -- model-level state of the batched fast path (model.cu).
#pragma once
#include "dtype.h"
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/whisper_b200.h"
#include "decode_chain.h"
#include "kernels.h"

namespace wb {


struct AttnW {  // offsets (in floats) into the fp32 weight image; layers.mojo:96-103
    int64_t q_w, q_b, k_w, v_w, v_b, o_w, o_b;
};
struct BlockW {  // layers.mojo:418-433
    AttnW attn;
    int64_t attn_ln_w, attn_ln_b;
    AttnW cross;
    int64_t cross_ln_w, cross_ln_b;
    int64_t fc1_w, fc1_b, fc2_w, fc2_b, mlp_ln_w, mlp_ln_b;
};
struct Layout {  // export_weights.py:19-90
    int64_t conv1_w, conv1_b, conv2_w, conv2_b, enc_pos;
    std::vector<BlockW> enc, dec;
    int64_t enc_ln_w, enc_ln_b, tok_emb, dec_pos, dec_ln_w, dec_ln_b;
    int64_t total;
    std::vector<std::pair<int64_t, int64_t>> tensors;  // (offset, count) in file order
};
Layout make_layout(const wm_config &c);

struct LayerDev {
    h16 *wqkv = nullptr, *wo = nullptr, *w1 = nullptr, *w2 = nullptr;
    h16 *cwq = nullptr, *cwo = nullptr;  // decoder cross-attention
    h16 *wqk = nullptr, *wov = nullptr;  // folded cross projections [H*D][D], [D][H*D] (absorbed form)
    float *bqk = nullptr, *bov = nullptr;
    float *bqkv = nullptr;                // [3D]: q bias, zeros (k has no bias), v bias
    const float *bo, *b1, *b2, *cbq, *cbo;
    const float *ln1_g, *ln1_b, *ln2_g, *ln2_b, *ln3_g, *ln3_b;  // attn_ln, cross_ln (dec), mlp_ln
};

// Event pairs around individual decode-step launches (option "profile_attn": 1 = cross-attention only, the
// bench's roofline pass; 2 = every kernel of the step, by category).
enum TimedKernel { TK_CROSS = 0, TK_SELF, TK_QKV, TK_O, TK_CQ, TK_CO, TK_FC1, TK_FC2, TK_LN, TK_LOGITS, TK_MISC,
                   TK_CHAIN_FIRST, TK_CHAIN_B, TK_CHAIN_CA, TK_COUNT };
struct KernelTimer {
    std::vector<cudaEvent_t> ev;
    std::vector<int> cat;  // category of event pair i
    int used = 0;
    float total_ms[TK_COUNT] = {};
    int64_t launches[TK_COUNT] = {};
};

struct Model {
    wm_config cfg;
    int D, H, L, V, S, T, NM, F, n_frames, n_samples;
    cudaStream_t stream = nullptr, stream2 = nullptr;  // stream2: host->device uploads, second decode lane
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool own_stream = false;
    bool pdl = false;  // programmatic dependent launch of the decode-step kernels (option "pdl"; env WB_PDL at create)
    int device = 0;    // the CUDA device the model was created on; entry points make it current (DeviceScope)
    int n_caches = 0;  // live wm_kvcache handles (wm_destroy refuses while any exist)
    int gemm_impl = 1, attn_impl = 1, frontend_impl = 1,  // frontend: 0 = fp32 FMA DFT, 1 = TF32x3 tensor-core DFT
         use_graph = 1, profile_attn = 0, enc_batch = 128, wave_max = 2048, small_batch = 0,
        decode_split_k = 1,  // split-K residual GEMMs + fused residual/LayerNorm in the decode step
        decode_fused = 2,  // 1 = the dense work between two attention kernels runs as one persistent chain kernel
                           // (decode_chain.cu: 4 L + 3 kernels per step instead of 12 L + 4); 0 = round 1's kernel per op;
                           // 2 = chain kernels for waves of <= FUSED_MAX_WAVE chunks (latency bound), kernel per op above
        decode_lanes = 1,  // 2 = two half-batches on two streams (measured: no gain, the HBM-bound kernel fills every SM)
        cross_impl = 1,    // 0 = per-layer cross K/V cache (reference form), 1 = absorbed form over enc_out (D <= 384)
        skip_done = 1,     // 1 = chunks that produced EOT drop out of the attention kernels (live list rebuilt every 16 steps)
        prefill_impl = 1;  // 1 = the 4 prompt ids run as ONE q_len = 4 forward with the causal block path (whisper.mojo:195-197);
                           // 0 = fed one by one through the cached step (same ids, 3 more forwards)
    int *stop_sched = nullptr;  // device [stop_sched_n]: forced lengths per chunk of a transcribe call (wm_set_stop_lengths)
    int stop_sched_n = 0;
    Layout lay;
    float *w32 = nullptr;
    bool loaded = false;
    h16 *conv1_w = nullptr, *conv2_w = nullptr, *tok_emb_h16 = nullptr, *cross_wkv = nullptr;
    float *cross_bkv = nullptr;
    std::vector<LayerDev> enc, dec;
    FrontendTables ft;
    std::vector<void *> owned;
    // encoder workspace (sized for enc_cap chunks)
    int enc_cap = 0;
    h16 *e_melT = nullptr, *e_x1T = nullptr, *e_xn = nullptr, *e_qkv = nullptr, *e_attn = nullptr, *e_h = nullptr,
         *e_enc = nullptr;
    float *e_x = nullptr;
    float timing[5] = {0, 0, 0, 0, 0};
    KernelTimer cross_timer;
    std::vector<cudaEvent_t> ev_plain, ev_timed;  // per-sub-batch upload / phase-timing events of transcribe
    // transcribe workspace kept across calls (allocation of a 19 GB cache costs ~0.2 s per call otherwise)
    struct Cache *tr_cache = nullptr;
    float *tr_mel = nullptr;
    size_t tr_mel_cap = 0;
    int *cmax = nullptr;  // frontend: per-chunk log-mel maxima
    int cmax_cap = 0;
    void *stage_in = nullptr, *stage_out = nullptr;  // host-API staging buffers (api.cu)
    size_t stage_in_cap = 0, stage_out_cap = 0;
};

// The reference's prompt: 4 ids run as one q_len = 4 forward (whisper.mojo:190-197).
static constexpr int PREFILL_LEN = 4;
static constexpr int FUSED_MAX_WAVE = 1280;  // decode_fused = 2: largest wave the chain kernels serve (measured crossover)

// A lane = a contiguous sub-batch of the cache's chunks with its own decode workspace and step state.
// Two lanes run on two streams inside one CUDA graph: one lane's small latency-bound kernels overlap the
// other lane's HBM-bound cross-attention.
struct Lane {
    int B = 0, b_off = 0;
    float *x = nullptr, *part_val = nullptr, *attn_ws = nullptr, *logits = nullptr, *part = nullptr;
    h16 *xn = nullptr, *q = nullptr, *attn = nullptr, *h = nullptr, *qp = nullptr, *ctx = nullptr;
    h16 *pf_k = nullptr, *pf_v = nullptr;  // prefill: k / v rows [PREFILL_LEN * B][D] on their way into the cache
    int *part_idx = nullptr, *next = nullptr;
    int cross_splits = 1;
    int part_splits = 4;  // slices the split-K partial buffer `part` holds
    // fused decode step (decode_chain.cu): plans[0] = embed + LN + qkv_0, plans[1 + 2 l] = o -> LN -> cross-q of
    // layer l, plans[2 + 2 l] = cross-o -> LN -> fc1 -> fc2 -> LN -> qkv_{l+1}; plan_last_nolog = the last layer's
    // chain without the final LayerNorm (prompt steps, whose logits nobody reads)
    std::vector<ChainPlan *> plans;
    ChainPlan *plan_last_nolog = nullptr;
    int *counters = nullptr;
    size_t counter_bytes = 0;
    GreedyState g;  // per-chunk arrays point into the cache-wide arrays at b_off; scalars are per lane
};

struct Cache {
    Model *m = nullptr;
    int B = 0, T = 0;
    int host_len = 0;  // mirror of current_len for the step-wise API
    bool has_cross = false;
    h16 *self_kv = nullptr;   // [L][2][B][T][D]
    h16 *cross_kv = nullptr;  // [L][2][B][S][D]   (cross_impl 0)
    h16 *cross_enc = nullptr; // [B][S][D]        (cross_impl 1: enc_out itself, shared by all layers)
    int cross_impl = 0;
    int *tokens_out = nullptr, *out_len = nullptr, *cur_tok = nullptr, *done = nullptr, *scalars = nullptr;
    int *live = nullptr, *stop_at = nullptr;  // [B]: per lane, the indices still decoding / forced lengths
    std::vector<Lane> lanes;
    std::vector<void *> owned;
    int *pinned_scalars = nullptr;           // two snapshots of the lanes' step scalars (lagged EOT poll)
    cudaEvent_t poll_ev[2] = {nullptr, nullptr};
    cudaGraphExec_t graph_exec = nullptr;
    int graph_kernels = 0;  // kernels one replay of graph_exec launches
    int graph_key = 0;      // launch options graph_exec was captured under (decode_fused, skip_done, pdl)
};

int model_create(const wm_config *cfg, void *stream, Model **out);
void model_destroy(Model *m);
int model_load(Model *m, const float *host, int64_t n_floats);
int model_logmel(Model *m, const float *pcm_dev, int n, float *mel_dev);
int model_encode(Model *m, const float *mel_dev, int n, float *enc_out_dev, Cache *into, int cache_off);
int cache_create(Model *m, int B, int max_len, bool want_logits, int n_lanes, Cache **out);
void cache_destroy(Cache *c);
int cache_reset(Cache *c);
int cache_set_encoder(Cache *c, const float *enc_out_dev);
int decode_step(Cache *c, Lane &ln, cudaStream_t st, bool with_logits, bool store_logits, bool advance, int q_len = 1);
int model_transcribe(Model *m, const float *mel_dev, const float *pcm_dev, int n, int32_t *out_tokens_dev,
                     int32_t *out_len_dev, const float *in_host = nullptr);
int model_teacher_forced(Model *m, const float *enc_out_dev, int n, const int32_t *forced_host, int n_forced,
                         float *logits_host);
int cache_step_api(Cache *c, const int32_t *tokens_host, int start_pos, float *logits_host, int32_t *next_host);

}  // namespace wb
