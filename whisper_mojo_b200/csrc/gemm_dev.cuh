// gemm_dev.cuh -- device-side pieces shared by the tensor-core GEMM kernels (gemm.cu) and the fused decode-step
// chain kernel (decode_chain.cu): tile constants, the device parameter block, output routing and the epilogue.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "gemm.h"
#include "sm100.cuh"

namespace wb {

static constexpr int BM = 128, BN = 128, BK = 64, STAGES = 6;
static constexpr int A_STAGE_BYTES = BM * BK * 2, B_STAGE_BYTES = BN * BK * 2;
static constexpr int TC_THREADS = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
static constexpr int PART_PER_TILE = 2;  // argmax partials per 128-column tile (one per epilogue column half)
static constexpr int TC_SMEM_BYTES = 1024 + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 256 + 8 * 128 * 4;

// Device-visible parameters (shared by both implementations).
struct GemmDev {
    int rows_per_batch, batches, N, K;
    int tiles_m_per_batch, tiles_n, num_kb, kb_per_tap, taps;
    int a_row_off[3];
    int split_koff;  // > 0: split-K, "batch" b covers K columns [b * split_koff, (b + 1) * split_koff) of A and W
    const float *bias;
    int epi;
    void *out[3];
    long long out_ld[3];
    long long seg_stride;
    int seg_cols, n_seg_ptrs;
    const int *dyn_off;
    long long dyn_mult[3];
    const float *pos;
    float *part_val;
    int *part_idx;
    float *logits;
    // reference-kernel operand addressing
    const h16 *A;
    long long a_batch_stride;
    int lda, src_rows, conv_stride, pad, Cin;
    const h16 *W;
};


// GELU for bf16 outputs: same formula as gelu_ref with the hardware tanh (MUFU.TANH, rel. error ~2^-11,
// far below the bf16 rounding of the result).  fp32 outputs keep the exact tanhf.
__device__ __forceinline__ float gelu_fast(float x) {
    const float k0 = 0.79788456f, k1 = 0.044715f;
    float inner = k0 * (x + k1 * x * x * x), t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(inner));
    return 0.5f * x * (1.0f + t);
}
// Two values at once with the packed f32x2 pipe ops of sm_100 (the GELU epilogue is FMA-issue bound: 7 scalar
// FMA-pipe instructions per element next to one MUFU.TANH): 0.5 x (1 + tanh(k0 x (1 + k1 x^2))).
__device__ __forceinline__ float2 gelu_fast2(float2 x) {
    const float2 k0 = make_float2(0.79788456f, 0.79788456f), k1 = make_float2(0.044715f, 0.044715f);
    const float2 one = make_float2(1.f, 1.f), half = make_float2(0.5f, 0.5f);
    const float2 x2 = __fmul2_rn(x, x);
    const float2 inner = __fmul2_rn(__fmul2_rn(x, k0), __ffma2_rn(k1, x2, one));
    float2 t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(inner.x));
    asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(inner.y));
    const float2 hx = __fmul2_rn(x, half);
    return __ffma2_rn(hx, t, hx);
}

// Resolve the output location of (global row, column n): returns element offset and segment.
__device__ __forceinline__ void out_location(const GemmDev &p, long long grow, int n, int &seg, long long &off) {
    seg = n / p.seg_cols;
    int col = n - seg * p.seg_cols;
    if (p.n_seg_ptrs == 0) {
        off = (long long)seg * p.seg_stride + grow * p.out_ld[0] + col;
        if (p.dyn_off) off += (long long)(*p.dyn_off) * p.dyn_mult[0];
        seg = 0;
    } else {
        off = grow * p.out_ld[seg] + col;
        if (p.dyn_off) off += (long long)(*p.dyn_off) * p.dyn_mult[seg];
    }
}

// Epilogue for one thread = one output row, 32 consecutive columns starting at n (n % 32 == 0), in two
// steps so that no global-memory latency sits between the TMEM read and the stores: `epi_prefetch` issues
// every load the chunk needs (bias, and the old residual / positional values) as independent 128-bit
// loads -- the kernels call it for chunk c+1 before they finish chunk c -- and `epi_finish` adds, applies
// the activation and stores.  The bias slice of the tile is staged in shared memory once per tile by
// `stage_bias` (one coalesced load per warp instead of 8 serial L2-latency loads per chunk).
// Warp-cooperative: sbias[0, ncols) = bias[n0 .. n0+ncols) (zero past N or without a bias); ncols <= 128.
__device__ __forceinline__ void stage_bias(const GemmDev &p, int n0, int ncols, float *sbias, int lane) {
    __syncwarp();  // every lane is done with the previous tile's values
#pragma unroll
    for (int t = 0; t < 4; t++) {
        const int j = lane + 32 * t;
        if (j < ncols) sbias[j] = (p.bias && n0 + j < p.N) ? __ldg(p.bias + n0 + j) : 0.f;
    }
    __syncwarp();
}

template <int EPI>
struct EpiChunk {
    float extra[32];  // EPI_RESID_F32: the values being added to; EPI_GELU_POS_F32: positional embedding
    void *dst;
    bool active, full, vec;
};

template <int EPI>
__device__ __forceinline__ void epi_prefetch(const GemmDev &p, int b, int m, int n, EpiChunk<EPI> &e) {
    e.active = (m < p.rows_per_batch && n < p.N);
    e.dst = nullptr, e.full = e.vec = false;
    if (!e.active) return;
    e.full = (n + 32 <= p.N);
    if (EPI == EPI_ARGMAX) return;
    const long long grow = (long long)b * p.rows_per_batch + m;
    int seg;
    long long off;
    out_location(p, grow, n, seg, off);
    if (EPI == EPI_STORE_H16 || EPI == EPI_GELU_H16) {
        e.dst = reinterpret_cast<h16 *>(p.out[seg]) + off;
        e.vec = e.full && ((reinterpret_cast<uintptr_t>(e.dst) & 15) == 0);
        return;
    }
    float *dst = reinterpret_cast<float *>(p.out[seg]) + off;
    e.dst = dst;
    e.vec = e.full && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
    if (EPI == EPI_RESID_F32) {
        if (e.vec) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 o = *reinterpret_cast<const float4 *>(dst + j);
                e.extra[j] = o.x, e.extra[j + 1] = o.y, e.extra[j + 2] = o.z, e.extra[j + 3] = o.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; j++) e.extra[j] = (n + j < p.N) ? dst[j] : 0.f;
        }
    } else if (EPI == EPI_GELU_POS_F32) {
        const float *pos = p.pos + (long long)m * p.N + n;
        if (e.full && ((reinterpret_cast<uintptr_t>(pos) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 o = __ldg(reinterpret_cast<const float4 *>(pos + j));
                e.extra[j] = o.x, e.extra[j + 1] = o.y, e.extra[j + 2] = o.z, e.extra[j + 3] = o.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; j++) e.extra[j] = (n + j < p.N) ? __ldg(pos + j) : 0.f;
        }
    }
}

template <int EPI>
__device__ __forceinline__ void epi_finish(const GemmDev &p, int b, int m, int n, const uint32_t *vraw,
                                           const EpiChunk<EPI> &e, const float *sbias, float &best, int &best_idx) {
    if (!e.active) return;
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; j += 4) {  // sbias: this chunk's 32 bias values in shared memory (broadcast reads)
        const float4 bv = *reinterpret_cast<const float4 *>(sbias + j);
        v[j] = __uint_as_float(vraw[j]) + bv.x, v[j + 1] = __uint_as_float(vraw[j + 1]) + bv.y;
        v[j + 2] = __uint_as_float(vraw[j + 2]) + bv.z, v[j + 3] = __uint_as_float(vraw[j + 3]) + bv.w;
    }
    if (EPI == EPI_ARGMAX) {
#pragma unroll
        for (int j = 0; j < 32; j++)
            if (n + j < p.N && v[j] > best) best = v[j], best_idx = n + j;
        if (p.logits) {
            const long long grow = (long long)b * p.rows_per_batch + m;
            float *dst = p.logits + grow * p.N + n;
#pragma unroll
            for (int j = 0; j < 32; j++)
                if (n + j < p.N) dst[j] = v[j];
        }
        return;
    }
    if (EPI == EPI_STORE_H16 || EPI == EPI_GELU_H16) {
        if (EPI == EPI_GELU_H16) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                const float2 g = gelu_fast2(make_float2(v[j], v[j + 1]));
                v[j] = g.x, v[j + 1] = g.y;
            }
        }
        h16 *dst = reinterpret_cast<h16 *>(e.dst);
        if (e.vec) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                uint4 u;
                u.x = pack_h2(v[j], v[j + 1]);
                u.y = pack_h2(v[j + 2], v[j + 3]);
                u.z = pack_h2(v[j + 4], v[j + 5]);
                u.w = pack_h2(v[j + 6], v[j + 7]);
                *reinterpret_cast<uint4 *>(dst + j) = u;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; j++)
                if (n + j < p.N) dst[j] = f2h(v[j]);
        }
        return;
    }
    float *dst = reinterpret_cast<float *>(e.dst);
#pragma unroll
    for (int j = 0; j < 32; j++) {
        if (EPI == EPI_RESID_F32) v[j] = e.extra[j] + v[j];
        else if (EPI == EPI_GELU_POS_F32) v[j] = gelu_ref(v[j]) + e.extra[j];
    }
    if (e.vec) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4 *>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
        for (int j = 0; j < 32; j++)
            if (n + j < p.N) dst[j] = v[j];
    }
}

}  // namespace wb
