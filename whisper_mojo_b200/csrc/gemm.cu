// gemm.cu -- C = A * W^T (+bias) in bf16 with fp32 accumulation and fused epilogues (see gemm.h).
//
// GEMM_IMPL_TC : persistent, warp-specialised sm_100a kernel.  One CTA per SM loops over 128x128
//   output tiles; warp 0 feeds a 6-stage shared-memory ring with TMA (128-byte swizzle), warp 1
//   issues tcgen05.mma (UMMA 128x128x16, bf16 -> fp32) into a double-buffered TMEM accumulator,
//   warps 2..9 drain TMEM with tcgen05.ld (two warps per 32-lane quarter, 64 columns each) and run
//   the epilogue while the next tile's MMAs issue.
//   The conv stem runs through the same kernel as an implicit GEMM: the K loop walks 3 taps, each a
//   TMA box shifted by one source row (zero fill outside the chunk = the conv's zero padding).
// GEMM_IMPL_REF: plain CUDA-core kernel with the same operand / epilogue contract, used to
//   validate the tensor-core path on the GPU and for bring-up.
#include "gemm.h"

#include <cuda.h>

#include <mutex>

#include "common.cuh"
#include "sm100.cuh"

namespace wb {

static constexpr int BM = 128, BN = 128, BK = 64, STAGES = 6;
static constexpr int A_STAGE_BYTES = BM * BK * 2, B_STAGE_BYTES = BN * BK * 2;
static constexpr int TC_THREADS = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
static constexpr int PART_PER_TILE = 2;  // argmax partials per 128-column tile (one per epilogue column half)
static constexpr int TC_SMEM_BYTES = 1024 + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 256;

// Device-visible parameters (shared by both implementations).
struct GemmDev {
    int rows_per_batch, batches, N, K;
    int tiles_m_per_batch, tiles_n, num_kb, kb_per_tap, taps;
    int a_row_off[3];
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
    const __nv_bfloat16 *A;
    long long a_batch_stride;
    int lda, src_rows, conv_stride, pad, Cin;
    const __nv_bfloat16 *W;
};

struct GemmTcParams {
    CUtensorMap a_map[3];
    CUtensorMap b_map;
    GemmDev d;
};

// GELU for bf16 outputs: same formula as gelu_ref with the hardware tanh (MUFU.TANH, rel. error ~2^-11,
// far below the bf16 rounding of the result).  fp32 outputs keep the exact tanhf.
__device__ __forceinline__ float gelu_fast(float x) {
    const float k0 = 0.79788456f, k1 = 0.044715f;
    float inner = k0 * (x + k1 * x * x * x), t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(inner));
    return 0.5f * x * (1.0f + t);
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

// Scalar epilogue (reference kernel; also the contract the vectorised tensor-core epilogue follows).
__device__ __forceinline__ void epilogue_scalar(const GemmDev &p, int b, int m, int n, float acc) {
    long long grow = (long long)b * p.rows_per_batch + m;
    float v = acc + (p.bias ? p.bias[n] : 0.f);
    if (p.epi == EPI_ARGMAX) {  // reference kernel only materialises logits; partials come from a helper
        p.logits[grow * p.N + n] = v;
        return;
    }
    int seg;
    long long off;
    out_location(p, grow, n, seg, off);
    switch (p.epi) {
        case EPI_STORE_BF16: reinterpret_cast<__nv_bfloat16 *>(p.out[seg])[off] = __float2bfloat16(v); break;
        case EPI_GELU_BF16: reinterpret_cast<__nv_bfloat16 *>(p.out[seg])[off] = __float2bfloat16(gelu_ref(v)); break;
        case EPI_RESID_F32: reinterpret_cast<float *>(p.out[seg])[off] += v; break;
        case EPI_STORE_F32: reinterpret_cast<float *>(p.out[seg])[off] = v; break;
        case EPI_GELU_POS_F32:
            reinterpret_cast<float *>(p.out[seg])[off] = gelu_ref(v) + p.pos[(long long)m * p.N + n];
            break;
    }
}

// ---------------------------------------------------------------------------------------------
// Reference kernel: 64x64 tile, 256 threads, 4x4 per thread, operands staged in fp32.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gemm_ref_kernel(const GemmDev p) {
    __shared__ float As[16][64 + 4], Bs[16][64 + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int tiles_m = (p.rows_per_batch + 63) / 64;
    const int b = blockIdx.y / tiles_m, m0 = (blockIdx.y % tiles_m) * 64, n0 = blockIdx.x * 64;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < p.K; k0 += 16) {
        for (int i = threadIdx.x; i < 64 * 16; i += 256) {
            int r = i >> 4, c = i & 15;
            int kk = k0 + c;
            int tap = kk / p.Cin, ci = kk - tap * p.Cin;
            int m = m0 + r;
            int srow = m * p.conv_stride + tap - p.pad;
            float av = 0.f;
            if (m < p.rows_per_batch && kk < p.K && srow >= 0 && srow < p.src_rows)
                av = __bfloat162float(p.A[(long long)b * p.a_batch_stride + (long long)srow * p.lda + ci]);
            As[c][r] = av;
            int n = n0 + r;
            Bs[c][r] = (n < p.N && kk < p.K) ? __bfloat162float(p.W[(long long)n * p.K + kk]) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; kk++) {
            float a[4], bb[4];
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = As[kk][ty * 4 + i], bb[i] = Bs[kk][tx * 4 + i];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] += a[i] * bb[j];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
            if (m < p.rows_per_batch && n < p.N) epilogue_scalar(p, b, m, n, acc[i][j]);
        }
}

// ---------------------------------------------------------------------------------------------
// Tensor-core kernel
// ---------------------------------------------------------------------------------------------

// Epilogue for one thread = one output row, 32 consecutive columns starting at n (n % 32 == 0).
template <int EPI>
__device__ __forceinline__ void epilogue_row32(const GemmDev &p, int b, int m, int n, const uint32_t *vraw,
                                               float &best, int &best_idx) {
    const long long grow = (long long)b * p.rows_per_batch + m;
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; j++) v[j] = __uint_as_float(vraw[j]);
    if (p.bias) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            if (n + j + 3 < p.N) {
                float4 bv = __ldg(reinterpret_cast<const float4 *>(p.bias + n + j));
                v[j] += bv.x, v[j + 1] += bv.y, v[j + 2] += bv.z, v[j + 3] += bv.w;
            } else {
#pragma unroll
                for (int t = 0; t < 4; t++)
                    if (n + j + t < p.N) v[j + t] += __ldg(p.bias + n + j + t);
            }
        }
    }
    if (EPI == EPI_ARGMAX) {
#pragma unroll
        for (int j = 0; j < 32; j++)
            if (n + j < p.N && v[j] > best) best = v[j], best_idx = n + j;
        if (p.logits) {
            float *dst = p.logits + grow * p.N + n;
#pragma unroll
            for (int j = 0; j < 32; j++)
                if (n + j < p.N) dst[j] = v[j];
        }
        return;
    }
    const bool full = (n + 32 <= p.N);
    int seg;
    long long off;
    out_location(p, grow, n, seg, off);
    if (EPI == EPI_STORE_BF16 || EPI == EPI_GELU_BF16) {
        if (EPI == EPI_GELU_BF16) {
#pragma unroll
            for (int j = 0; j < 32; j++) v[j] = gelu_fast(v[j]);
        }
        __nv_bfloat16 *dst = reinterpret_cast<__nv_bfloat16 *>(p.out[seg]) + off;
        if (full && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                uint4 u;
                u.x = pack_bf16x2(v[j], v[j + 1]);
                u.y = pack_bf16x2(v[j + 2], v[j + 3]);
                u.z = pack_bf16x2(v[j + 4], v[j + 5]);
                u.w = pack_bf16x2(v[j + 6], v[j + 7]);
                *reinterpret_cast<uint4 *>(dst + j) = u;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; j++)
                if (n + j < p.N) dst[j] = __float2bfloat16(v[j]);
        }
    } else {
        float *dst = reinterpret_cast<float *>(p.out[seg]) + off;
        const float *pos = (EPI == EPI_GELU_POS_F32) ? p.pos + (long long)m * p.N + n : nullptr;
        if (full && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                float4 o;
                if (EPI == EPI_RESID_F32) {
                    o = *reinterpret_cast<const float4 *>(dst + j);
                    o.x += v[j], o.y += v[j + 1], o.z += v[j + 2], o.w += v[j + 3];
                } else if (EPI == EPI_GELU_POS_F32) {
                    float4 pv = __ldg(reinterpret_cast<const float4 *>(pos + j));
                    o.x = gelu_ref(v[j]) + pv.x, o.y = gelu_ref(v[j + 1]) + pv.y;
                    o.z = gelu_ref(v[j + 2]) + pv.z, o.w = gelu_ref(v[j + 3]) + pv.w;
                } else {
                    o.x = v[j], o.y = v[j + 1], o.z = v[j + 2], o.w = v[j + 3];
                }
                *reinterpret_cast<float4 *>(dst + j) = o;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; j++)
                if (n + j < p.N) {
                    if (EPI == EPI_RESID_F32) dst[j] += v[j];
                    else if (EPI == EPI_GELU_POS_F32) dst[j] = gelu_ref(v[j]) + pos[j];
                    else dst[j] = v[j];
                }
        }
    }
}

template <int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tc_kernel(const __grid_constant__ GemmTcParams P) {
    extern __shared__ uint8_t smem_raw[];
    const GemmDev &p = P.d;
    // carve shared memory: [1024-aligned tiles][barriers]
    uint8_t *tiles = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *smem_a = tiles;
    uint8_t *smem_b = tiles + STAGES * A_STAGE_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(tiles + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES));
    uint64_t *full_bar = bars, *empty_bar = bars + STAGES;
    uint64_t *tmem_full = bars + 2 * STAGES, *tmem_empty = bars + 2 * STAGES + 2;
    uint32_t *tmem_holder = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = p.batches * p.tiles_m_per_batch * p.tiles_n;

    if (warp == 0 && lane == 0) {
        for (int t = 0; t < p.taps; t++) ptx::prefetch_tmap(&P.a_map[t]);
        ptx::prefetch_tmap(&P.b_map);
        for (int s = 0; s < STAGES; s++) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; s++) {
            ptx::mbar_init(&tmem_full[s], 1);
            ptx::mbar_init(&tmem_empty[s], 8);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_holder, 256);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int nt = tile % p.tiles_n, mt = tile / p.tiles_n;
                const int b = mt / p.tiles_m_per_batch, m0 = (mt % p.tiles_m_per_batch) * BM;
                for (int kb = 0; kb < p.num_kb; kb++) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    ptx::mbar_expect_tx(&full_bar[stage], A_STAGE_BYTES + B_STAGE_BYTES);
                    const int tap = kb / p.kb_per_tap, c0 = (kb - tap * p.kb_per_tap) * BK;
                    ptx::tma_load_3d(smem_a + stage * A_STAGE_BYTES, &P.a_map[tap], &full_bar[stage], c0,
                                     m0 + p.a_row_off[tap], b);
                    ptx::tma_load_2d(smem_b + stage * B_STAGE_BYTES, &P.b_map, &full_bar[stage], kb * BK, nt * BN);
                    if (++stage == STAGES) stage = 0, phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::umma_idesc_bf16(BM, BN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it++) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < p.num_kb; kb++) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint64_t a_desc = ptx::umma_desc_sw128(ptx::smem_u32(smem_a + stage * A_STAGE_BYTES), 1, 64);
                    const uint64_t b_desc = ptx::umma_desc_sw128(ptx::smem_u32(smem_b + stage * B_STAGE_BYTES), 1, 64);
#pragma unroll
                    for (int k = 0; k < BK / 16; k++)  // +32 bytes (2 x 16 B units) per UMMA_K step inside the atom
                        ptx::mma_bf16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
                    ptx::mma_commit(&empty_bar[stage]);
                    if (++stage == STAGES) stage = 0, phase ^= 1;
                }
                ptx::mma_commit(&tmem_full[acc]);
            }
        }
    } else {
        // ===== epilogue warps: TMEM lanes 32*(warp%4) .. +31, columns 64*half .. +63 =====
        const int q = warp & 3, half = (warp - 2) >> 2;
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it++) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int nt = tile % p.tiles_n, mt = tile / p.tiles_n;
            const int b = mt / p.tiles_m_per_batch, m0 = (mt % p.tiles_m_per_batch) * BM;
            const int m = m0 + q * 32 + lane;
            ptx::mbar_wait(&tmem_full[acc], acc_phase);
            ptx::tc_fence_after();
            float best = -INFINITY;
            int best_idx = 0x7fffffff;
#pragma unroll 1
            for (int c = half * 2; c < half * 2 + 2; c++) {
                uint32_t v[32];
                ptx::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + c * 32, v);
                ptx::tmem_ld_wait();
                const int n = nt * BN + c * 32;
                if (m < p.rows_per_batch && n < p.N) epilogue_row32<EPI>(p, b, m, n, v, best, best_idx);
            }
            if (EPI == EPI_ARGMAX && m < p.rows_per_batch) {
                long long grow = (long long)b * p.rows_per_batch + m;
                p.part_val[(grow * p.tiles_n + nt) * PART_PER_TILE + half] = best;
                p.part_idx[(grow * p.tiles_n + nt) * PART_PER_TILE + half] = best_idx;
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, 256);
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// bf16 tensor map with a {64, box_rows, 1} box and 128-byte swizzle over a [d2][d1][d0] view.
int make_tmap_bf16(CUtensorMap *map, const void *base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_elems,
                   uint64_t stride2_elems, uint32_t box_rows, int rank) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
        return WB_ERR_CUDA;
    }
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {stride1_elems * 2, stride2_elems * 2};
    cuuint32_t box[3] = {64, box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void *>(base), dims, strides,
                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed: %d (dims %llu,%llu,%llu strides %llu,%llu)", (int)r,
                  (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2,
                  (unsigned long long)strides[0], (unsigned long long)strides[1]);
        return WB_ERR_CUDA;
    }
    return WB_OK;
}

static int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

template <int EPI>
static int launch_tc(cudaStream_t st, const GemmTcParams &P, int grid) {
    static bool attr_set = false;
    if (!attr_set) {
        WB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
        attr_set = true;
    }
    gemm_tc_kernel<EPI><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(P);
    WB_LAUNCHED();
    return WB_OK;
}

int gemm_run(cudaStream_t st, const GemmDesc &d, int impl) {
    WB_ARG(d.A && d.W && d.N > 0 && d.Cin > 0 && d.taps >= 1 && d.taps <= 3, "gemm: bad operands");
    if (d.batches <= 0 || d.rows_per_batch <= 0) return WB_OK;
    GemmTcParams P;
    GemmDev &p = P.d;
    p.rows_per_batch = d.rows_per_batch;
    p.batches = d.batches;
    p.N = d.N;
    p.K = d.taps * d.Cin;
    p.tiles_m_per_batch = cdiv(d.rows_per_batch, BM);
    p.tiles_n = cdiv(d.N, BN);
    p.taps = d.taps;
    p.kb_per_tap = d.Cin / BK;
    p.num_kb = d.taps * p.kb_per_tap;
    p.bias = d.bias;
    p.epi = d.epi;
    for (int i = 0; i < 3; i++) p.out[i] = d.out[i], p.out_ld[i] = d.out_ld[i], p.dyn_mult[i] = d.dyn_mult[i];
    p.seg_stride = d.seg_stride;
    p.seg_cols = d.seg_cols > 0 ? d.seg_cols : d.N;
    p.n_seg_ptrs = d.n_seg_ptrs;
    p.dyn_off = d.dyn_off;
    p.pos = d.pos;
    p.part_val = d.part_val;
    p.part_idx = d.part_idx;
    p.logits = d.logits;
    p.A = d.A;
    p.a_batch_stride = d.a_batch_stride;
    p.lda = d.lda;
    p.src_rows = d.src_rows;
    p.conv_stride = d.conv_stride;
    p.pad = d.pad;
    p.Cin = d.Cin;
    p.W = d.W;
    WB_ARG(d.epi == EPI_ARGMAX || d.out[0], "gemm: missing output");
    WB_ARG(d.epi != EPI_ARGMAX || (d.logits || impl == GEMM_IMPL_TC), "gemm: reference argmax needs a logits buffer");
    WB_ARG(p.seg_cols == d.N || p.seg_cols % BN == 0, "gemm: seg_cols must be a multiple of 128");

    if (impl == GEMM_IMPL_REF) {
        dim3 grid(cdiv(d.N, 64), d.batches * cdiv(d.rows_per_batch, 64));
        gemm_ref_kernel<<<grid, 256, 0, st>>>(p);
        WB_LAUNCHED();
        if (d.epi == EPI_ARGMAX)
            WB_CHECK(argmax_partials_from_logits(st, d.logits, d.batches * d.rows_per_batch, d.N, d.part_val,
                                                 d.part_idx));
        return WB_OK;
    }

    WB_ARG(d.Cin % BK == 0, "gemm(tc): Cin=%d must be a multiple of 64", d.Cin);
    WB_ARG(d.lda % 8 == 0 && (d.a_batch_stride % 8) == 0, "gemm(tc): lda / batch stride must be multiples of 8");
    // A maps: one per tap.  source row = m*cs + tap - pad = cs*(m + q) + par with par in [0, cs).
    for (int t = 0; t < d.taps; t++) {
        int delta = t - d.pad;
        int q = floordiv(delta, d.conv_stride), par = delta - q * d.conv_stride;
        p.a_row_off[t] = q;
        uint64_t nrows = d.src_rows > par ? (uint64_t)(d.src_rows - par + d.conv_stride - 1) / d.conv_stride : 0;
        WB_ARG(nrows > 0, "gemm(tc): empty tap view");
        WB_CHECK(make_tmap_bf16(&P.a_map[t], d.A + (size_t)par * d.lda, (uint64_t)d.Cin, nrows, (uint64_t)d.batches,
                                (uint64_t)d.conv_stride * d.lda,
                                d.batches > 1 ? (uint64_t)d.a_batch_stride : (uint64_t)d.conv_stride * d.lda * nrows, BM,
                                3));
    }
    for (int t = d.taps; t < 3; t++) P.a_map[t] = P.a_map[0], p.a_row_off[t] = 0;
    WB_CHECK(make_tmap_bf16(&P.b_map, d.W, (uint64_t)p.K, (uint64_t)d.N, 1, (uint64_t)p.K, 0, BN, 2));

    int total_tiles = p.batches * p.tiles_m_per_batch * p.tiles_n;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int grid = total_tiles < sms ? total_tiles : sms;
    switch (d.epi) {
        case EPI_STORE_BF16: return launch_tc<EPI_STORE_BF16>(st, P, grid);
        case EPI_GELU_BF16: return launch_tc<EPI_GELU_BF16>(st, P, grid);
        case EPI_RESID_F32: return launch_tc<EPI_RESID_F32>(st, P, grid);
        case EPI_STORE_F32: return launch_tc<EPI_STORE_F32>(st, P, grid);
        case EPI_ARGMAX: return launch_tc<EPI_ARGMAX>(st, P, grid);
        case EPI_GELU_POS_F32: return launch_tc<EPI_GELU_POS_F32>(st, P, grid);
    }
    set_error("gemm: unknown epilogue %d", d.epi);
    return WB_ERR_ARG;
}

// ---- argmax over partials --------------------------------------------------------------------

__global__ void argmax_partials_kernel(const float *__restrict__ part_val, const int *__restrict__ part_idx, int M,
                                       int tiles_n, int *__restrict__ next) {
    int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= M) return;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int t = lane; t < tiles_n; t += 32) {
        float v = part_val[(size_t)row * tiles_n + t];
        int i = part_idx[(size_t)row * tiles_n + t];
        if (v > best || (v == best && i < bi)) best = v, bi = i;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) best = ov, bi = oi;
    }
    if (lane == 0) next[row] = (bi == 0x7fffffff) ? 0 : bi;
}
int argmax_partials(cudaStream_t st, const float *part_val, const int *part_idx, int M, int tiles_n, int *next_dev) {
    if (M <= 0) return WB_OK;
    argmax_partials_kernel<<<cdiv(M, 8), 256, 0, st>>>(part_val, part_idx, M, tiles_n, next_dev);
    WB_LAUNCHED();
    return WB_OK;
}

__global__ void argmax_from_logits_kernel(const float *__restrict__ logits, int M, int N, int tiles_n,
                                          float *__restrict__ part_val, int *__restrict__ part_idx) {
    int row = blockIdx.y, t = blockIdx.x, lane = threadIdx.x;  // t = 64-column slot (PART_PER_TILE per 128-col tile)
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int j = lane; j < 64; j += 32) {
        int n = t * 64 + j;
        if (n < N) {
            float v = logits[(size_t)row * N + n];
            if (v > best) best = v, bi = n;  // ascending n per lane: first max kept
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) best = ov, bi = oi;
    }
    if (lane == 0) part_val[(size_t)row * tiles_n + t] = best, part_idx[(size_t)row * tiles_n + t] = bi;
}
int argmax_partials_from_logits(cudaStream_t st, const float *logits, int M, int N, float *part_val, int *part_idx) {
    if (M <= 0) return WB_OK;
    int slots = gemm_tiles_n(N);  // = PART_PER_TILE * ceil(N / 128)
    dim3 grid(slots, M);
    argmax_from_logits_kernel<<<grid, 32, 0, st>>>(logits, M, N, slots, part_val, part_idx);
    WB_LAUNCHED();
    return WB_OK;
}

}  // namespace wb
