// decode_chain.h -- the decode step's dense work between two attention kernels as ONE persistent kernel.
//
// The reference runs a decoder block as ~20 separate ops (layers.mojo:435-519); round 1 ran it as 11 kernels per
// layer (GEMM, split-K reduce + LayerNorm, ...), 52 per step, each a few microseconds of tensor work wrapped in
// 10-15 us of launch / fill / drain latency.  A chain kernel runs a LIST OF PHASES -- tcgen05 GEMMs (128 x 128
// tiles, split-K partials) and row phases (partial sums + bias + residual + LayerNorm, or the token / position
// embedding) -- on one persistent cooperative grid.  A phase's tile needs only the rows of its own 128-row tile from
// the phase before, so instead of grid-wide barriers every (phase, row tile) has an arrival counter in global memory:
// producers bump it when their part of the tile is stored, consumers poll it before they load that tile.  Weights
// never wait: the TMA producer streams a tile's weight blocks into the ring first and only then waits for the
// activations, so the next phase's weights are in flight while the current phase finishes.
//
//   step = first (embed + LN + qkv_0) | self-attn | B_l (o -> LN -> cross-q) | cross-attn |
//          CA_l (cross-o -> LN -> fc1+GELU -> fc2 -> LN -> qkv_{l+1}) | ... | logits+argmax | argmax reduce + bookkeeping
//   = 4 L + 3 kernels (19 for Tiny) instead of 12 L + 4 (52).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "gemm.h"

namespace wb {

static constexpr int CHAIN_MAX_PHASES = 6;

// One GEMM phase: out = A[M][K] * W[N][K]^T (+ bias), routed like GemmDesc (segments, KV-cache offset), or with
// split_k > 1 as fp32 partial products [split_k][M][N] that the following row phase sums.
struct ChainGemm {
    const h16 *A = nullptr;
    int K = 0;
    const h16 *W = nullptr;
    int N = 0;
    const float *bias = nullptr;
    int epi = EPI_STORE_H16;  // EPI_STORE_H16, EPI_GELU_H16 or EPI_STORE_F32 (split-K partials, no bias)
    int split_k = 1;
    void *out[3] = {nullptr, nullptr, nullptr};
    int64_t out_ld[3] = {0, 0, 0};
    int seg_cols = 0, n_seg_ptrs = 1;
    const int *dyn_off = nullptr;
    int64_t dyn_mult[3] = {0, 0, 0};
};
// One row phase: x[row] (+)= ..., xn[row] = LayerNorm(x[row]) in 16 bits (gamma == nullptr: no LayerNorm output).
//   embed: x[row] = tok_emb[cur_tok[row]] + pos_emb[*pos]                (whisper.mojo:138-149)
//   else : x[row] += bias + part[0][row] + ... + part[n_split-1][row]     (fixed order; layers.mojo:456-461,486-491,515-517)
struct ChainRows {
    float *x = nullptr;
    const float *part = nullptr;
    int n_split = 0;
    const float *bias = nullptr, *gamma = nullptr, *beta = nullptr;
    h16 *xn = nullptr;
    bool embed = false;
    const float *tok_emb = nullptr, *pos_emb = nullptr;
    const int *cur_tok = nullptr, *pos_dev = nullptr;
    int vocab = 0, n_pos = 0;
};

struct ChainPlan;  // device parameter block + launch geometry, built once per (cache, lane, kernel)
ChainPlan *chain_plan_create(int M, int D, int *counters);
void chain_plan_destroy(ChainPlan *p);
int chain_plan_add_gemm(ChainPlan *p, const ChainGemm &g);
int chain_plan_add_rows(ChainPlan *p, const ChainRows &r);
int chain_plan_phases(const ChainPlan *p);
// Number of ints of counter storage one plan needs (the caller zeroes them before every launch of the plan).
size_t chain_counter_ints(int M);
// Fixes the launch geometry (call once, after the last phase was added).
int chain_plan_finalize(ChainPlan *p);
int chain_launch(cudaStream_t st, ChainPlan *p);
// Split-K factor for a residual GEMM with reduction length K (a property of the model, never of the batch size).
int chain_split_k(int K);
// Development aid: with WB_CHAIN_DBG set in the environment, prints CTA 0's per-phase timestamps of the last launch.
void chain_debug_dump();
void chain_debug_init();  // allocates the timestamp buffer (call outside stream capture, e.g. at model creation)

}  // namespace wb
