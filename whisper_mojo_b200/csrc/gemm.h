// gemm.h -- bf16 tensor-core GEMM C = A * W^T (+bias) with fused epilogues, sm_100a.
//
// One kernel family serves every dense contraction of the path: the conv stem as an implicit GEMM
// over 3 shifted taps (whisper_tensor.mojo:367-428), the encoder / decoder projections and MLPs
// (the reference's matmul / MAX wrappers, whisper_tensor.mojo:74-246), the all-layer cross-K/V
// projection (layers.mojo:148-157) and the tied-embedding logit projection fused with argmax
// (whisper.mojo:159-166 + whisper_tensor.mojo:431-439).
#pragma once
#include "dtype.h"
#include <cuda_runtime.h>
#include <stdint.h>

namespace wb {

enum GemmEpi {
    EPI_STORE_H16 = 0,    // out_h16 = acc + bias
    EPI_GELU_H16 = 1,     // out_h16 = gelu(acc + bias)
    EPI_RESID_F32 = 2,     // out_f32 += acc + bias               (residual add in place)
    EPI_STORE_F32 = 3,     // out_f32 = acc + bias
    EPI_ARGMAX = 4,        // per-row (max, first index) over each 128-column tile -> partials; optional f32 logits
    EPI_GELU_POS_F32 = 5,  // out_f32 = gelu(acc + bias) + pos[row_in_batch][n]   (conv2 + positional add)
};

// REF: CUDA-core bring-up kernel.  TC: tcgen05 kernels, CTA-pair (cta_group::2) variant for the large GEMMs and the
// single-CTA one otherwise.  TC_SINGLE / TC_PAIR force one variant (tests, A/B measurements).
enum GemmImpl { GEMM_IMPL_REF = 0, GEMM_IMPL_TC = 1, GEMM_IMPL_TC_SINGLE = 2, GEMM_IMPL_TC_PAIR = 3 };

// Host-side description of one GEMM.
//   A(b, m, tap*Cin + ci) = src[b*a_batch_stride + (m*conv_stride + tap - pad)*lda + ci]
//   (zero when the source row is outside [0, src_rows)); plain GEMM: taps=1, conv_stride=1, pad=0.
//   Output row index = b*rows_per_batch + m.
struct GemmDesc {
    const h16 *A = nullptr;
    int64_t a_batch_stride = 0;
    int lda = 0, src_rows = 0, conv_stride = 1, pad = 0, taps = 1, Cin = 0;
    int batches = 1, rows_per_batch = 0;
    const h16 *W = nullptr;  // [N][taps*Cin]
    int N = 0;
    const float *bias = nullptr;
    int epi = EPI_STORE_H16;
    // Output routing.  Columns are split in segments of seg_cols (a multiple of 128, or N).
    // n_seg_ptrs > 0: segment s writes to out[s] with row stride out_ld[s] (+ dyn offset).
    // n_seg_ptrs == 0: segment s writes to out[0] + s*seg_stride with row stride out_ld[0].
    void *out[3] = {nullptr, nullptr, nullptr};
    int64_t out_ld[3] = {0, 0, 0};
    int64_t seg_stride = 0;
    int seg_cols = 0;  // 0 -> N
    int n_seg_ptrs = 1;
    const int *dyn_off = nullptr;  // device int; element offset added = (*dyn_off) * dyn_mult[s]
    int64_t dyn_mult[3] = {0, 0, 0};
    const float *pos = nullptr;  // EPI_GELU_POS_F32: [rows_per_batch][N]
    float *part_val = nullptr;   // EPI_ARGMAX: [M][gemm_tiles_n(N)]
    int *part_idx = nullptr;
    float *logits = nullptr;  // EPI_ARGMAX: optional full logits [M][N]
    // Split-K (CTA-pair kernel, plain GEMM, EPI_STORE_F32 without bias): the K = taps * Cin reduction is cut into
    // split_k slices of equal length (a multiple of 64); slice s writes its fp32 partial product to rows
    // [s * rows_per_batch, (s + 1) * rows_per_batch) of the output.  The slices run as extra tiles, so a small-M GEMM
    // with a long K (decode: M = 2048, K = 2304) spreads over the whole chip; the caller sums the slices in a
    // fixed order (resid_ln).
    int split_k = 1;
};

// Number of argmax partial slots per row: two per 128-column tile (one per epilogue column half).
static inline int gemm_tiles_n(int N) { return 2 * ((N + 127) / 128); }

int gemm_run(cudaStream_t st, const GemmDesc &d, int impl);

// Reduce the EPI_ARGMAX partials: next[b] = first index of the row maximum.
int argmax_partials(cudaStream_t st, const float *part_val, const int *part_idx, int M, int tiles_n, int *next_dev);
// Bring-up helper: build the same partials from full logits [M][N].
int argmax_partials_from_logits(cudaStream_t st, const float *logits, int M, int N, float *part_val, int *part_idx);

}  // namespace wb
