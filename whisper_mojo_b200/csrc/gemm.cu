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

#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "gemm_dev.cuh"
#include "sm100.cuh"

namespace wb {

struct GemmTcParams {
    CUtensorMap a_map[3];
    CUtensorMap b_map;
    GemmDev d;
};

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
        case EPI_STORE_H16: reinterpret_cast<h16 *>(p.out[seg])[off] = f2h(v); break;
        case EPI_GELU_H16: reinterpret_cast<h16 *>(p.out[seg])[off] = f2h(gelu_ref(v)); break;
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
                av = h2f(p.A[(long long)b * p.a_batch_stride + (long long)srow * p.lda + ci]);
            As[c][r] = av;
            int n = n0 + r;
            Bs[c][r] = (n < p.N && kk < p.K) ? h2f(p.W[(long long)n * p.K + kk]) : 0.f;
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
    float *bias_smem = reinterpret_cast<float *>(tiles + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 256);  // [8][128]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = p.batches * p.tiles_m_per_batch * p.tiles_n;
    pdl_launch_dependents();

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
    pdl_wait();  // everything above overlapped the previous kernel; operands / outputs are touched from here on

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
            constexpr uint32_t idesc = ptx::umma_idesc_h16(BM, BN, 0, 0);
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
                        ptx::mma_h16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
                    ptx::mma_commit(&empty_bar[stage]);
                    if (++stage == STAGES) stage = 0, phase ^= 1;
                }
                ptx::mma_commit(&tmem_full[acc]);
            }
        }
    } else {
        // ===== epilogue warps: TMEM lanes 32*(warp%4) .. +31, columns 64*half .. +63 =====
        const int q = warp & 3, half = (warp - 2) >> 2;
        float *sbias = bias_smem + (warp - 2) * 128;
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
            const int n_first = nt * BN + half * 64;
            EpiChunk<EPI> e0, e1;
            uint32_t v0[32], v1[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + half * 64;
            ptx::tmem_ld_32x32b_x32(taddr, v0);
            ptx::tmem_ld_32x32b_x32(taddr + 32, v1);
            epi_prefetch<EPI>(p, b, m, n_first, e0);
            epi_prefetch<EPI>(p, b, m, n_first + 32, e1);
            stage_bias(p, n_first, 64, sbias, lane);
            ptx::tmem_ld_wait();
            // the accumulator is in registers: hand the TMEM buffer back before the stores
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_relaxed(&tmem_empty[acc]);
            epi_finish<EPI>(p, b, m, n_first, v0, e0, sbias, best, best_idx);
            epi_finish<EPI>(p, b, m, n_first + 32, v1, e1, sbias + 32, best, best_idx);
            if (EPI == EPI_ARGMAX && m < p.rows_per_batch) {
                long long grow = (long long)b * p.rows_per_batch + m;
                p.part_val[(grow * p.tiles_n + nt) * PART_PER_TILE + half] = best;
                p.part_idx[(grow * p.tiles_n + nt) * PART_PER_TILE + half] = best_idx;
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, 256);
}

// ---------------------------------------------------------------------------------------------
// CTA-pair tensor-core kernel (cta_group::2).  The single-CTA kernel above re-reads 192 KB of operands
// from L2 for every 128x128x384 tile, ~3x what the L2->SM fabric delivers at tensor-pipe speed, so the
// large encoder GEMMs are L2-bound.  Here two CTAs on one TPC compute a 256 x BN2 tile with one UMMA
// (M = 256): each CTA stages its own 128 rows of A and only HALF of the B tile (BN2/2 weight rows), the
// tensor cores read the other half from the peer's shared memory -> half the operand traffic per flop.
//   * both producers (one per CTA) feed the same ring position; their TMA loads complete on the LEADER's
//     "full" barrier, the leader's MMA thread issues for the pair and its tcgen05.commit multicasts the
//     "empty" / "accumulator full" arrivals to both CTAs;
//   * each CTA's eight epilogue warps drain their own TMEM (128 lanes x BN2 columns) and arrive on the
//     leader's "accumulator empty" barrier (remote arrive for the peer).
// ---------------------------------------------------------------------------------------------
static constexpr int PAIR_STAGE_BYTES = 32768;  // A 16 KB + B up to 16 KB per CTA
// 6 ring stages; the TMA-output variant trades two of them for the epilogue staging tiles (8 warps x 2 x 4 KB)
static constexpr int pair_stages(bool tma_out) { return tma_out ? 4 : 6; }
static constexpr int pair_smem_bytes(bool tma_out) {
    return 1024 + pair_stages(tma_out) * PAIR_STAGE_BYTES + 256 + 8 * 128 * 4 + (tma_out ? 8 * 2 * 4096 : 0);
}

struct GemmPairParams {
    CUtensorMap a_map[3];
    CUtensorMap b_map;  // box {64, BN2/2}
    GemmDev d;
    int pair_tiles_m_per_batch, pair_tiles_n;
    CUtensorMap out_map;  // TMA_OUT: fp32 output as {N, rows_per_batch, batches}, box {32, 32, 1}, 128B swizzle
};

// TMA_OUT (plain [rows][N] outputs): each epilogue warp stages its 32 x 32 tile in swizzled shared memory and
// issues one bulk tensor store per tile -- full-row writes instead of 32 scattered 16-byte stores per
// instruction.  For EPI_RESID_F32 the store is a cp.reduce.async.bulk.tensor (.add.f32): the residual add is
// done by the L2, the SM never reads the old values, so no HBM latency sits in the epilogue.  One add per
// element: bit-identical to x + (acc + bias).
template <int EPI, int BN2, bool TMA_OUT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
    gemm_pair_kernel(const __grid_constant__ GemmPairParams P) {
    constexpr int PAIR_STAGES = pair_stages(TMA_OUT);
    extern __shared__ uint8_t smem_raw[];
    const GemmDev &p = P.d;
    constexpr int B_ROWS = BN2 / 2, B_BYTES = B_ROWS * BK * 2;
    uint8_t *tiles = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    // [ring][TMA_OUT: 8 warps x 2 staging tiles of 4 KB, 1024-byte aligned as the 128B swizzle requires][barriers][bias]
    constexpr int RING_BYTES = PAIR_STAGES * PAIR_STAGE_BYTES + (TMA_OUT ? 8 * 2 * 4096 : 0);
    uint8_t *stage_out = tiles + PAIR_STAGES * PAIR_STAGE_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(tiles + RING_BYTES);
    uint64_t *full_bar = bars, *empty_bar = bars + PAIR_STAGES;
    uint64_t *tmem_full = bars + 2 * PAIR_STAGES, *tmem_empty = bars + 2 * PAIR_STAGES + 2;
    uint32_t *tmem_holder = reinterpret_cast<uint32_t *>(bars + 2 * PAIR_STAGES + 4);
    float *bias_smem = reinterpret_cast<float *>(tiles + RING_BYTES + 256);  // [8][128]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const int total_tiles = p.batches * P.pair_tiles_m_per_batch * P.pair_tiles_n;
    const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    pdl_launch_dependents();

    if (warp == 0 && lane == 0) {
        for (int t = 0; t < p.taps; t++) ptx::prefetch_tmap(&P.a_map[t]);
        ptx::prefetch_tmap(&P.b_map);
        if (TMA_OUT) ptx::prefetch_tmap(&P.out_map);
        for (int s = 0; s < PAIR_STAGES; s++) {
            ptx::mbar_init(&full_bar[s], 1);   // used on the leader only
            ptx::mbar_init(&empty_bar[s], 1);  // one multicast commit per use
        }
        for (int s = 0; s < 2; s++) {
            ptx::mbar_init(&tmem_full[s], 1);
            ptx::mbar_init(&tmem_empty[s], 16);  // leader only: 8 epilogue warps of each CTA
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc_pair(tmem_holder, 512);
        ptx::tmem_relinquish_pair();
    }
    ptx::tc_fence_before();
    ptx::cluster_sync();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    pdl_wait();

    if (warp == 0) {
        // ===== TMA producer (one per CTA) =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = pair_id; tile < total_tiles; tile += n_pairs) {
                const int nt = tile % P.pair_tiles_n, mt = tile / P.pair_tiles_n;
                const int b = mt / P.pair_tiles_m_per_batch;
                const int m0 = (mt % P.pair_tiles_m_per_batch) * 256 + (int)rank * BM;
                const int n0 = nt * BN2 + (int)rank * B_ROWS;
                for (int kb = 0; kb < p.num_kb; kb++) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (rank == 0) ptx::mbar_expect_tx(&full_bar[stage], 2 * (A_STAGE_BYTES + B_BYTES));
                    const uint32_t bar = ptx::mapa_u32(&full_bar[stage], 0);
                    uint8_t *sa = tiles + stage * PAIR_STAGE_BYTES, *sb = sa + A_STAGE_BYTES;
                    const int tap = kb / p.kb_per_tap, c0 = (kb - tap * p.kb_per_tap) * BK;
                    const int koff = b * p.split_koff;  // split-K: the "batch" index selects a K slice, not rows
                    ptx::tma_load_3d_pair(sa, &P.a_map[tap], bar, c0 + koff, m0 + p.a_row_off[tap], p.split_koff ? 0 : b);
                    ptx::tma_load_2d_pair(sb, &P.b_map, bar, kb * BK + koff, n0);
                    if (++stage == PAIR_STAGES) stage = 0, phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA only) =====
        if (rank == 0 && lane == 0) {
            constexpr uint32_t idesc = ptx::umma_idesc_h16(256, BN2, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = pair_id; tile < total_tiles; tile += n_pairs, it++) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN2;
                for (int kb = 0; kb < p.num_kb; kb++) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    uint8_t *sa = tiles + stage * PAIR_STAGE_BYTES, *sb = sa + A_STAGE_BYTES;
                    const uint64_t a_desc = ptx::umma_desc_sw128(ptx::smem_u32(sa), 1, 64);
                    const uint64_t b_desc = ptx::umma_desc_sw128(ptx::smem_u32(sb), 1, 64);
#pragma unroll
                    for (int k = 0; k < BK / 16; k++)
                        ptx::mma_h16_ss_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
                    ptx::mma_commit_pair(&empty_bar[stage], 3);
                    if (++stage == PAIR_STAGES) stage = 0, phase ^= 1;
                }
                ptx::mma_commit_pair(&tmem_full[acc], 3);
            }
        }
    } else {
        // ===== epilogue warps: TMEM lanes 32*(warp%4) .. +31, columns [half*BN2/2, +BN2/2) =====
        const int q = warp & 3, half = (warp - 2) >> 2;
        constexpr int CH = BN2 / 64;  // 32-column chunks per half
        float *sbias = bias_smem + (warp - 2) * 128;
        int it = 0;
        uint32_t out_ctr = 0;  // staging-buffer parity (TMA_OUT)
        for (int tile = pair_id; tile < total_tiles; tile += n_pairs, it++) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int nt = tile % P.pair_tiles_n, mt = tile / P.pair_tiles_n;
            const int b = mt / P.pair_tiles_m_per_batch;
            const int m = (mt % P.pair_tiles_m_per_batch) * 256 + (int)rank * BM + q * 32 + lane;
            ptx::mbar_wait(&tmem_full[acc], acc_phase);
            ptx::tc_fence_after();
            float best = -INFINITY;
            int best_idx = 0x7fffffff;
            const int n_first = nt * BN2 + half * (BN2 / 2);
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN2 + half * (BN2 / 2);
            // software pipeline over the CH chunks: TMEM read of chunk c and every global load of chunk c+1
            // are in flight while chunk c-1 is converted and stored
            EpiChunk<EPI> e[2];
            uint32_t v[2][32];
            ptx::tmem_ld_32x32b_x32(taddr, v[0]);
            if (!TMA_OUT) epi_prefetch<EPI>(p, b, m, n_first, e[0]);
            stage_bias(p, n_first, BN2 / 2, sbias, lane);
#pragma unroll
            for (int c = 0; c < CH; c++) {
                ptx::tmem_ld_wait();  // chunk c is in v[c & 1]
                if (c + 1 < CH) {
                    ptx::tmem_ld_32x32b_x32(taddr + (c + 1) * 32, v[(c + 1) & 1]);
                    if (!TMA_OUT) epi_prefetch<EPI>(p, b, m, n_first + (c + 1) * 32, e[(c + 1) & 1]);
                } else {
                    // the whole accumulator slice has left TMEM: hand the buffer back before the last stores
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive_cluster_relaxed(ptx::mapa_u32(&tmem_empty[acc], 0));
                }
                if (TMA_OUT) {
                    // 32 rows x 32 values -> this warp's staging tile, 16-byte chunks XOR-swizzled exactly as the
                    // output tensor map expects (fp32: 128 B rows / 128B swizzle, bf16: 64 B rows / 64B swizzle),
                    // then one bulk tensor store (bf16) or reduce-add (fp32 residual) per tile
                    uint8_t *buf = stage_out + ((warp - 2) * 2 + (out_ctr & 1)) * 4096;
                    if (lane == 0) ptx::bulk_wait_group_read<1>();  // the store that last used this buffer has read it
                    __syncwarp();
                    float o[32];
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const float4 bv = *reinterpret_cast<const float4 *>(sbias + c * 32 + 4 * j);
                        o[4 * j] = __uint_as_float(v[c & 1][4 * j]) + bv.x, o[4 * j + 1] = __uint_as_float(v[c & 1][4 * j + 1]) + bv.y;
                        o[4 * j + 2] = __uint_as_float(v[c & 1][4 * j + 2]) + bv.z, o[4 * j + 3] = __uint_as_float(v[c & 1][4 * j + 3]) + bv.w;
                    }
                    if (EPI == EPI_RESID_F32 || EPI == EPI_STORE_F32) {
#pragma unroll
                        for (int j = 0; j < 8; j++)
                            *reinterpret_cast<float4 *>(buf + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                                make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
                    } else {
                        if (EPI == EPI_GELU_H16) {
#pragma unroll
                            for (int j = 0; j < 32; j += 2) {
                                const float2 g = gelu_fast2(make_float2(o[j], o[j + 1]));
                                o[j] = g.x, o[j + 1] = g.y;
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            *reinterpret_cast<uint4 *>(buf + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) =
                                make_uint4(pack_h2(o[8 * j], o[8 * j + 1]), pack_h2(o[8 * j + 2], o[8 * j + 3]),
                                           pack_h2(o[8 * j + 4], o[8 * j + 5]), pack_h2(o[8 * j + 6], o[8 * j + 7]));
                    }
                    ptx::fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        const int m_w = (mt % P.pair_tiles_m_per_batch) * 256 + (int)rank * BM + q * 32;
                        if (EPI == EPI_RESID_F32) ptx::tma_reduce_add_3d(&P.out_map, buf, n_first + c * 32, m_w, b);
                        else ptx::tma_store_3d(&P.out_map, buf, n_first + c * 32, m_w, b);
                        ptx::bulk_commit_group();
                    }
                    out_ctr++;
                } else {
                    epi_finish<EPI>(p, b, m, n_first + c * 32, v[c & 1], e[c & 1], sbias + c * 32, best, best_idx);
                    if (EPI == EPI_ARGMAX && (c & 1)) {
                        // one (max, first index) partial per 64 columns, the slot layout of the single-CTA kernel
                        if (m < p.rows_per_batch) {
                            const long long grow = (long long)b * p.rows_per_batch + m;
                            const long long slot = grow * p.tiles_n * PART_PER_TILE + (n_first + (c - 1) * 32) / 64;
                            p.part_val[slot] = best;
                            p.part_idx[slot] = best_idx;
                        }
                        best = -INFINITY, best_idx = 0x7fffffff;
                    }
                }
            }
        }
    }
    if (TMA_OUT && warp >= 2 && lane == 0) ptx::bulk_wait_group<0>();  // staged tiles fully written before exit
    ptx::tc_fence_before();
    ptx::cluster_sync();  // the peer may still signal this CTA's barriers / read its shared memory
    if (warp == 1) ptx::tmem_dealloc_pair(tmem_base, 512);
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
int make_tmap_h16(CUtensorMap *map, const void *base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_elems,
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
    CUresult r = fn(map, H16_TMAP_DTYPE, (cuuint32_t)rank, const_cast<void *>(base), dims, strides,
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

// Generic tiled tensor map; strides in bytes for dims 1..rank-1.  swizzle_bytes in {0, 32, 64, 128}.
static int make_tmap_any(CUtensorMap *map, CUtensorMapDataType dtype, int swizzle_bytes, const void *base, int rank,
                         const uint64_t *dims, const uint64_t *strides_bytes, const uint32_t *box) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
        return WB_ERR_CUDA;
    }
    cuuint64_t d[3] = {1, 1, 1}, s[2] = {0, 0};
    cuuint32_t b[3] = {1, 1, 1}, estr[3] = {1, 1, 1};
    for (int i = 0; i < rank; i++) d[i] = dims[i], b[i] = box[i];
    for (int i = 0; i + 1 < rank; i++) s[i] = strides_bytes[i];
    const CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                  : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                        : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = fn(map, dtype, (cuuint32_t)rank, const_cast<void *>(base), d, s, b, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed: %d (dims %llu,%llu,%llu strides %llu,%llu)", (int)r,
                  (unsigned long long)d[0], (unsigned long long)d[1], (unsigned long long)d[2], (unsigned long long)s[0],
                  (unsigned long long)s[1]);
        return WB_ERR_CUDA;
    }
    return WB_OK;
}
int make_tmap_any_pub(CUtensorMap *map, CUtensorMapDataType dtype, int swizzle_bytes, const void *base, int rank,
                      const uint64_t *dims, const uint64_t *strides_bytes, const uint32_t *box) {
    return make_tmap_any(map, dtype, swizzle_bytes, base, rank, dims, strides_bytes, box);
}
// fp32 tensor map, 128-byte swizzle (box[0] = 32 elements).
int make_tmap_f32(CUtensorMap *map, const void *base, int rank, const uint64_t *dims, const uint64_t *strides_bytes,
                  const uint32_t *box) {
    return make_tmap_any(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 128, base, rank, dims, strides_bytes, box);
}

static int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

template <int EPI>
static int launch_tc(cudaStream_t st, const GemmTcParams &P, int grid) {
    WB_CUDA(ensure_dyn_smem(gemm_tc_kernel<EPI>, TC_SMEM_BYTES));
    WB_CUDA(launch_pdl(gemm_tc_kernel<EPI>, dim3(grid), dim3(TC_THREADS), TC_SMEM_BYTES, st, P));
    WB_LAUNCHED();
    return WB_OK;
}

template <int EPI, int BN2, bool TMA_OUT>
static int launch_pair(cudaStream_t st, const GemmPairParams &P, int grid) {
    constexpr int smem = pair_smem_bytes(TMA_OUT);
    WB_CUDA(ensure_dyn_smem(gemm_pair_kernel<EPI, BN2, TMA_OUT>, smem));
    WB_CUDA(launch_pdl(gemm_pair_kernel<EPI, BN2, TMA_OUT>, dim3(grid), dim3(TC_THREADS), smem, st, P));
    WB_LAUNCHED();
    return WB_OK;
}
template <int EPI, bool TMA_OUT = false>
static int launch_pair_bn(cudaStream_t st, const GemmPairParams &P, int grid, int bn2) {
    switch (bn2) {
        case 256: return launch_pair<EPI, 256, TMA_OUT>(st, P, grid);
        case 192: return launch_pair<EPI, 192, TMA_OUT>(st, P, grid);
        default: return launch_pair<EPI, 128, TMA_OUT>(st, P, grid);
    }
}

bool g_tma_out = true;  // A/B switch (WB_TMA_OUT=0 in the environment restores the per-thread store epilogues)

// Pair-tile width: the widest of 256 / 192 / 128 that wastes the fewest padded columns.
static int pick_pair_bn(int N) {
    int best = 128, best_cols = cdiv(N, 128) * 128;
    for (int bn : {192, 256}) {
        int cols = cdiv(N, bn) * bn;
        if (cols <= best_cols) best = bn, best_cols = cols;
    }
    return best;
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
    p.split_koff = 0;
    if (d.split_k > 1) {
        WB_ARG(d.taps == 1 && d.batches == 1 && d.epi == EPI_STORE_F32 && !d.bias && d.n_seg_ptrs == 1 && !d.dyn_off &&
                   impl != GEMM_IMPL_REF && d.Cin % (d.split_k * BK) == 0,
               "gemm: split_k needs a plain bias-free EPI_STORE_F32 GEMM with K %% (64 * split_k) == 0 on the tcgen05 path");
        p.split_koff = d.Cin / d.split_k;
        p.batches = d.split_k;  // K slices ride on the batch index of the tile scheduler and of the output rows
        p.num_kb = p.split_koff / BK;
        p.kb_per_tap = p.num_kb;
    }
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
    WB_ARG(d.epi != EPI_ARGMAX || (d.logits || impl != GEMM_IMPL_REF), "gemm: reference argmax needs a logits buffer");
    WB_ARG(p.seg_cols == d.N || p.seg_cols % 32 == 0, "gemm: seg_cols must be a multiple of 32");
    WB_ARG(impl >= GEMM_IMPL_REF && impl <= GEMM_IMPL_TC_PAIR, "gemm: unknown implementation %d", impl);

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
        WB_CHECK(make_tmap_h16(&P.a_map[t], d.A + (size_t)par * d.lda, (uint64_t)d.Cin, nrows, (uint64_t)d.batches,
                                (uint64_t)d.conv_stride * d.lda,
                                d.batches > 1 ? (uint64_t)d.a_batch_stride : (uint64_t)d.conv_stride * d.lda * nrows, BM,
                                3));
    }
    for (int t = d.taps; t < 3; t++) P.a_map[t] = P.a_map[0], p.a_row_off[t] = 0;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);

    static const bool env_once = [] {
        const char *e = getenv("WB_TMA_OUT");
        if (e && e[0] == '0') g_tma_out = false;
        return true;
    }();
    (void)env_once;
    const int pair_tiles_m = cdiv(d.rows_per_batch, 256);
    const bool pair_ok = true;
    // CTA-pair kernel: always for a full wave of pair tiles; for the decode-step GEMMs (M = batch <= 2048) already
    // from 12 pair tiles on (measured: 10-15 % faster than 48..288 single-CTA tiles) unless the output goes through
    // the device-side KV-cache offset (per-thread stores, where the single-CTA kernel wins).
    const int64_t pair_tiles = (int64_t)p.batches * pair_tiles_m * cdiv(d.N, 256);
    const bool want_pair = impl == GEMM_IMPL_TC_PAIR || d.split_k > 1 ||
                           (impl == GEMM_IMPL_TC && (pair_tiles >= sms / 2 || (pair_tiles >= 12 && !d.dyn_off)));
    if (pair_ok && want_pair) {
        GemmPairParams Q;
        // output routing is per 32-column chunk, so any tile width works; the argmax partials need 64-column pairs
        const int bn2 = d.epi == EPI_ARGMAX ? 256 : pick_pair_bn(d.N);
        for (int t = 0; t < 3; t++) Q.a_map[t] = P.a_map[t];
        WB_CHECK(make_tmap_h16(&Q.b_map, d.W, (uint64_t)p.K, (uint64_t)d.N, 1, (uint64_t)p.K, 0, bn2 / 2, 2));
        Q.d = p;
        Q.pair_tiles_m_per_batch = pair_tiles_m;
        Q.pair_tiles_n = cdiv(d.N, bn2);
        const int total = p.batches * pair_tiles_m * Q.pair_tiles_n;
        const int grid = 2 * std::min(total, sms / 2);
        switch (d.epi) {
            case EPI_STORE_H16:
            case EPI_GELU_H16:
            case EPI_RESID_F32:
            case EPI_STORE_F32: {
                // plain [rows][N] output -> coalesced TMA stores / residual add by TMA reduce (fp32: the SM never
                // reads the old values)
                const int esz = (d.epi == EPI_RESID_F32 || d.epi == EPI_STORE_F32) ? 4 : 2;
                const bool plain = d.n_seg_ptrs == 1 && !d.dyn_off && p.seg_cols == d.N && (d.out_ld[0] * esz) % 16 == 0 &&
                                   (reinterpret_cast<uintptr_t>(d.out[0]) & 15) == 0 && g_tma_out;
                if (plain) {
                    const uint64_t dims[3] = {(uint64_t)d.N, (uint64_t)d.rows_per_batch, (uint64_t)p.batches};
                    const uint64_t str[2] = {(uint64_t)d.out_ld[0] * esz, (uint64_t)d.rows_per_batch * d.out_ld[0] * esz};
                    const uint32_t box[3] = {32, 32, 1};
                    WB_CHECK(make_tmap_any(&Q.out_map, esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : H16_TMAP_DTYPE,
                                           esz == 4 ? 128 : 64, d.out[0], 3, dims, str, box));
                    if (d.epi == EPI_STORE_H16) return launch_pair_bn<EPI_STORE_H16, true>(st, Q, grid, bn2);
                    if (d.epi == EPI_GELU_H16) return launch_pair_bn<EPI_GELU_H16, true>(st, Q, grid, bn2);
                    if (d.epi == EPI_STORE_F32) return launch_pair_bn<EPI_STORE_F32, true>(st, Q, grid, bn2);
                    return launch_pair_bn<EPI_RESID_F32, true>(st, Q, grid, bn2);
                }
                if (d.epi == EPI_STORE_H16) return launch_pair_bn<EPI_STORE_H16>(st, Q, grid, bn2);
                if (d.epi == EPI_GELU_H16) return launch_pair_bn<EPI_GELU_H16>(st, Q, grid, bn2);
                if (d.epi == EPI_STORE_F32) return launch_pair_bn<EPI_STORE_F32>(st, Q, grid, bn2);
                return launch_pair_bn<EPI_RESID_F32>(st, Q, grid, bn2);
            }
            case EPI_ARGMAX: return launch_pair<EPI_ARGMAX, 256, false>(st, Q, grid);
            case EPI_GELU_POS_F32: return launch_pair_bn<EPI_GELU_POS_F32>(st, Q, grid, bn2);
        }
    }
    WB_CHECK(make_tmap_h16(&P.b_map, d.W, (uint64_t)p.K, (uint64_t)d.N, 1, (uint64_t)p.K, 0, BN, 2));

    int total_tiles = p.batches * p.tiles_m_per_batch * p.tiles_n;
    int grid = total_tiles < sms ? total_tiles : sms;
    switch (d.epi) {
        case EPI_STORE_H16: return launch_tc<EPI_STORE_H16>(st, P, grid);
        case EPI_GELU_H16: return launch_tc<EPI_GELU_H16>(st, P, grid);
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
    pdl_launch_dependents();
    pdl_wait();
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
    WB_CUDA(launch_pdl(argmax_partials_kernel, dim3(cdiv(M, 8)), dim3(256), 0, st, part_val, part_idx, M, tiles_n, next_dev));
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
