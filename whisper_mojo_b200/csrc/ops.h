// ops.h -- launchers of the op-level fp32 kernels (ops.cu).  All pointers are device pointers.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace wb {
int op_matmul(cudaStream_t st, float *C, const float *A, const float *B, const float *bias, int M, int N, int K);
int op_layer_norm(cudaStream_t st, float *out, const float *inp, const float *gamma, const float *beta, int rows,
                  int cols, float eps);
int op_gelu(cudaStream_t st, float *t, size_t n);
int op_softmax(cudaStream_t st, float *t, int rows, int cols);
int op_transpose_conv_weights(cudaStream_t st, float *nw, const float *w, int C_out, int C_in, int K);
int op_conv1d(cudaStream_t st, float *out, const float *inp, const float *wT, const float *bias, int C_in, int L_in,
              int C_out, int stride, int padding, int out_T);
int op_argmax(cudaStream_t st, const float *t, int64_t n, long long *out_dev);
int op_add(cudaStream_t st, float *out, const float *a, const float *b, size_t n);
int op_scale_mask(cudaStream_t st, float *s, int rows, int cols, float scale, int mask, long long base);
int op_embed(cudaStream_t st, float *out, const float *tok_emb, const float *pos_emb, const int *tokens_dev, int n,
             int D, int start_pos);
int op_transpose(cudaStream_t st, float *out, const float *in, int rows, int cols);
}  // namespace wb
