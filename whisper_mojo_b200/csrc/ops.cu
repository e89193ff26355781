// ops.cu -- op-level fp32 kernels, 1:1 with the reference's L2 routines (whisper_tensor.mojo).
// These back the wt_* C ABI so layers.mojo / whisper.mojo can keep orchestrating op by op; the
// batched fast path (model.cu) uses the fused bf16 tensor-core kernels instead.
#include "common.cuh"
#include "ops.h"

namespace wb {

// ---- matmul: C[M,N] = A[M,K] * B[N,K]^T (+bias)  (whisper_tensor.mojo:151-246) ---------------

// Small-M path (the reference's M <= 4 "vector" path): one warp per output column n, lanes stride
// over K with float4 loads when K % 4 == 0.
template <int MAXM>
__global__ void matmul_smallm_kernel(float *__restrict__ C, const float *__restrict__ A, const float *__restrict__ B,
                                     const float *__restrict__ bias, int M, int N, int K) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= N) return;
    const float *b = B + (size_t)warp * K;
    float acc[MAXM];
#pragma unroll
    for (int m = 0; m < MAXM; m++) acc[m] = 0.f;
    if ((K & 3) == 0) {
        for (int k = lane * 4; k < K; k += 128) {
            float4 bv = *reinterpret_cast<const float4 *>(b + k);
#pragma unroll
            for (int m = 0; m < MAXM; m++)
                if (m < M) {
                    float4 av = *reinterpret_cast<const float4 *>(A + (size_t)m * K + k);
                    acc[m] += av.x * bv.x + av.y * bv.y + av.z * bv.z + av.w * bv.w;
                }
        }
    } else {
        for (int k = lane; k < K; k += 32) {
            float bv = b[k];
#pragma unroll
            for (int m = 0; m < MAXM; m++)
                if (m < M) acc[m] += A[(size_t)m * K + k] * bv;
        }
    }
#pragma unroll
    for (int m = 0; m < MAXM; m++) {
        float s = warp_sum(acc[m]);
        if (lane == 0 && m < M) C[(size_t)m * N + warp] = s + (bias ? bias[warp] : 0.f);
    }
}

// General path: 64x64 tile, BK = 16, 256 threads, 4x4 micro-tile, fp32 FMA.
__global__ void __launch_bounds__(256) matmul_tiled_kernel(float *__restrict__ C, const float *__restrict__ A,
                                                           const float *__restrict__ B,
                                                           const float *__restrict__ bias, int M, int N, int K) {
    __shared__ float As[16][64 + 4], Bs[16][64 + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += 16) {
        for (int i = threadIdx.x; i < 64 * 16; i += 256) {
            int r = i >> 4, c = i & 15;
            int gm = m0 + r, gn = n0 + r, gk = k0 + c;
            As[c][r] = (gm < M && gk < K) ? A[(size_t)gm * K + gk] : 0.f;
            Bs[c][r] = (gn < N && gk < K) ? B[(size_t)gn * K + gk] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; kk++) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = As[kk][ty * 4 + i], b[i] = Bs[kk][tx * 4 + i];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] += a[i] * b[j];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int gm = m0 + ty * 4 + i, gn = n0 + tx * 4 + j;
            if (gm < M && gn < N) C[(size_t)gm * N + gn] = acc[i][j] + (bias ? bias[gn] : 0.f);
        }
}

int op_matmul(cudaStream_t st, float *C, const float *A, const float *B, const float *bias, int M, int N, int K) {
    if (M <= 0 || N <= 0) return WB_OK;
    if (M <= 4) {
        int warps_per_block = 8;
        matmul_smallm_kernel<4><<<cdiv(N, warps_per_block), warps_per_block * 32, 0, st>>>(C, A, B, bias, M, N, K);
    } else {
        dim3 grid(cdiv(N, 64), cdiv(M, 64));
        matmul_tiled_kernel<<<grid, 256, 0, st>>>(C, A, B, bias, M, N, K);
    }
    WB_LAUNCHED();
    return WB_OK;
}

// ---- layer_norm (whisper_tensor.mojo:249-285): one warp per row, one-pass variance ----------

__global__ void layer_norm_kernel(float *__restrict__ out, const float *__restrict__ inp,
                                  const float *__restrict__ gamma, const float *__restrict__ beta, int rows, int cols,
                                  float eps) {
    int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float *x = inp + (size_t)row * cols;
    float s = 0.f, q = 0.f;
    for (int j = lane; j < cols; j += 32) {
        float v = x[j];
        s += v;
        q += v * v;
    }
    s = warp_sum(s);
    q = warp_sum(q);
    float mean = s / (float)cols;
    float var = q / (float)cols - mean * mean;
    float inv_std = 1.0f / sqrtf(var + eps);
    for (int j = lane; j < cols; j += 32) out[(size_t)row * cols + j] = (x[j] - mean) * inv_std * gamma[j] + beta[j];
}

int op_layer_norm(cudaStream_t st, float *out, const float *inp, const float *gamma, const float *beta, int rows,
                  int cols, float eps) {
    if (rows <= 0) return WB_OK;
    layer_norm_kernel<<<cdiv(rows, 8), 256, 0, st>>>(out, inp, gamma, beta, rows, cols, eps);
    WB_LAUNCHED();
    return WB_OK;
}

// ---- gelu (whisper_tensor.mojo:288-308) -----------------------------------------------------

__global__ void gelu_kernel(float *__restrict__ t, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) t[i] = gelu_ref(t[i]);
}
int op_gelu(cudaStream_t st, float *t, size_t n) {
    if (!n) return WB_OK;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    gelu_kernel<<<blocks, 256, 0, st>>>(t, n);
    WB_LAUNCHED();
    return WB_OK;
}

// ---- softmax (whisper_tensor.mojo:311-355): one block per row --------------------------------

__global__ void __launch_bounds__(256) softmax_kernel(float *__restrict__ t, int cols) {
    __shared__ float red[8];
    float *r = t + (size_t)blockIdx.x * cols;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float m = -INFINITY;
    for (int j = threadIdx.x; j < cols; j += 256) m = fmaxf(m, r[j]);
    m = warp_max(m);
    if (lane == 0) red[w] = m;
    __syncthreads();
    m = red[0];
#pragma unroll
    for (int i = 1; i < 8; i++) m = fmaxf(m, red[i]);
    __syncthreads();
    float s = 0.f;
    for (int j = threadIdx.x; j < cols; j += 256) {
        float e = expf(r[j] - m);
        r[j] = e;
        s += e;
    }
    s = warp_sum(s);
    if (lane == 0) red[w] = s;
    __syncthreads();
    s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) s += red[i];
    for (int j = threadIdx.x; j < cols; j += 256) r[j] = r[j] / s;
}
int op_softmax(cudaStream_t st, float *t, int rows, int cols) {
    if (rows <= 0 || cols <= 0) return WB_OK;
    softmax_kernel<<<rows, 256, 0, st>>>(t, cols);
    WB_LAUNCHED();
    return WB_OK;
}

// ---- conv weights transpose + conv1d (whisper_tensor.mojo:358-428) ---------------------------

__global__ void transpose_conv_weights_kernel(float *__restrict__ nw, const float *__restrict__ w, int C_out,
                                              int C_in, int K) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t n = (size_t)C_out * C_in * K;
    if (i >= n) return;
    int k = (int)(i % K);
    int ci = (int)((i / K) % C_in);
    int co = (int)(i / ((size_t)K * C_in));
    nw[((size_t)co * K + k) * C_in + ci] = w[i];
}
int op_transpose_conv_weights(cudaStream_t st, float *nw, const float *w, int C_out, int C_in, int K) {
    size_t n = (size_t)C_out * C_in * K;
    transpose_conv_weights_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(nw, w, C_out, C_in, K);
    WB_LAUNCHED();
    return WB_OK;
}

// One thread per output element; lo is the fastest index so input reads coalesce.
__global__ void conv1d_kernel(float *__restrict__ out, const float *__restrict__ inp, const float *__restrict__ wT,
                              const float *__restrict__ bias, int C_in, int L_in, int C_out, int L_out, int stride,
                              int padding, int out_T) {
    int lo = blockIdx.x * blockDim.x + threadIdx.x;
    int co = blockIdx.y;
    if (lo >= L_out) return;
    float acc = 0.f;
    int start = lo * stride - padding;
    for (int k = 0; k < 3; k++) {
        int li = start + k;
        if (li < 0 || li >= L_in) continue;
        const float *wp = wT + ((size_t)co * 3 + k) * C_in;
        for (int ci = 0; ci < C_in; ci++) acc += inp[(size_t)ci * L_in + li] * wp[ci];
    }
    acc += bias[co];
    if (out_T) out[(size_t)lo * C_out + co] = acc;
    else out[(size_t)co * L_out + lo] = acc;
}
int op_conv1d(cudaStream_t st, float *out, const float *inp, const float *wT, const float *bias, int C_in, int L_in,
              int C_out, int stride, int padding, int out_T) {
    int L_out = (L_in + 2 * padding - 3) / stride + 1;
    dim3 grid(cdiv(L_out, 128), C_out);
    conv1d_kernel<<<grid, 128, 0, st>>>(out, inp, wT, bias, C_in, L_in, C_out, L_out, stride, padding, out_T);
    WB_LAUNCHED();
    return WB_OK;
}

// ---- argmax (whisper_tensor.mojo:431-439): first maximum wins ---------------------------------

__global__ void __launch_bounds__(1024) argmax_kernel(const float *__restrict__ t, int64_t n,
                                                      long long *__restrict__ out) {
    __shared__ float sv[32];
    __shared__ long long si[32];
    float best = -INFINITY;
    long long bi = 0x7fffffffffffffffLL;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        float v = t[i];
        if (v > best || (v == best && i < bi)) best = v, bi = i;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, best, o);
        long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) best = ov, bi = oi;
    }
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sv[w] = best, si[w] = bi;
    __syncthreads();
    if (w == 0) {
        best = lane < (int)(blockDim.x >> 5) ? sv[lane] : -INFINITY;
        bi = lane < (int)(blockDim.x >> 5) ? si[lane] : 0x7fffffffffffffffLL;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, best, o);
            long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) best = ov, bi = oi;
        }
        if (lane == 0) *out = (bi == 0x7fffffffffffffffLL) ? 0 : bi;  // all -inf/NaN: index 0 like the reference
    }
}
int op_argmax(cudaStream_t st, const float *t, int64_t n, long long *out_dev) {
    argmax_kernel<<<1, 1024, 0, st>>>(t, n, out_dev);
    WB_LAUNCHED();
    return WB_OK;
}

// ---- small elementwise helpers -------------------------------------------------------------

__global__ void add_kernel(float *__restrict__ out, const float *__restrict__ a, const float *__restrict__ b,
                           size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = a[i] + b[i];
}
int op_add(cudaStream_t st, float *out, const float *a, const float *b, size_t n) {
    if (!n) return WB_OK;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    add_kernel<<<blocks, 256, 0, st>>>(out, a, b, n);
    WB_LAUNCHED();
    return WB_OK;
}

__global__ void scale_mask_kernel(float *__restrict__ s, int rows, int cols, float scale, int mask, long long base) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)rows * cols) return;
    int r = (int)(i / cols), c = (int)(i % cols);
    float v = s[i] * scale;
    if (mask && (long long)c > base + r) v = -1e10f;
    s[i] = v;
}
int op_scale_mask(cudaStream_t st, float *s, int rows, int cols, float scale, int mask, long long base) {
    size_t n = (size_t)rows * cols;
    if (!n) return WB_OK;
    scale_mask_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(s, rows, cols, scale, mask, base);
    WB_LAUNCHED();
    return WB_OK;
}

__global__ void embed_kernel(float *__restrict__ out, const float *__restrict__ tok_emb,
                             const float *__restrict__ pos_emb, const int *__restrict__ tokens, int n, int D,
                             int start_pos) {
    int i = blockIdx.x;
    const float *te = tok_emb + (size_t)tokens[i] * D, *pe = pos_emb + (size_t)(start_pos + i) * D;
    for (int j = threadIdx.x; j < D; j += blockDim.x) out[(size_t)i * D + j] = te[j] + pe[j];
}
int op_embed(cudaStream_t st, float *out, const float *tok_emb, const float *pos_emb, const int *tokens_dev, int n,
             int D, int start_pos) {
    if (n <= 0) return WB_OK;
    embed_kernel<<<n, 128, 0, st>>>(out, tok_emb, pos_emb, tokens_dev, n, D, start_pos);
    WB_LAUNCHED();
    return WB_OK;
}

__global__ void transpose_kernel(float *__restrict__ out, const float *__restrict__ in, int rows, int cols) {
    __shared__ float tile[32][33];
    int c = blockIdx.x * 32 + threadIdx.x, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8)
        if (r0 + i < rows && c < cols) tile[i][threadIdx.x] = in[(size_t)(r0 + i) * cols + c];
    __syncthreads();
    int r = r0 + threadIdx.x, c0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += 8)
        if (c0 + i < cols && r < rows) out[(size_t)(c0 + i) * rows + r] = tile[threadIdx.x][i];
}
int op_transpose(cudaStream_t st, float *out, const float *in, int rows, int cols) {
    if (rows <= 0 || cols <= 0) return WB_OK;
    dim3 grid(cdiv(cols, 32), cdiv(rows, 32)), block(32, 8);
    transpose_kernel<<<grid, block, 0, st>>>(out, in, rows, cols);
    WB_LAUNCHED();
    return WB_OK;
}

}  // namespace wb
