// attn_tc.cu -- encoder self-attention on tcgen05 tensor cores (sm_100a), flash-attention style.
//
// Replaces the reference's per-head gather -> QK^T -> scale -> softmax -> V^T -> PV -> scatter
// (layers.mojo:273-342, no mask for the encoder) without ever materialising the 1500 x 1500 score
// matrix.  One CTA = 128 queries of one (chunk, head); it walks the keys in blocks of 128:
//
//   warp 4  TMA producer : Q tile once, then K_j (2-deep ring) and V_j straight out of the fused
//                          [tokens][3D] qkv activation (one 3-D tensor map; rows past the chunk's
//                          1500 tokens are zero-filled by TMA, never read from the next chunk)
//   warp 5  MMA issuer   : S = Q K_j^T  (UMMA 128x128x16, K-major x K-major) into TMEM cols [0,128)
//                          O += P_j V_j (UMMA 128x64x16, P K-major from smem, V MN-major from smem)
//                          into TMEM cols [128,192); S_{j+1} is issued before PV_j so it overlaps
//                          the softmax of block j
//   warps 0-3 softmax    : thread = query row; the 128 scores of the block are pulled into registers
//                          with one TMEM read (S is released to the MMA warp right away), then max,
//                          ex2.approx -> bf16 P into 128B-swizzled smem; the online-softmax shift is
//                          lazy (moves only when the maximum grows by > 2^8), so the O rescale through
//                          tcgen05.ld/st is rare.  Bound: MUFU.EX2, 16/clk/SM = 1024 clk per 128x128 block
//
// 96 KB of shared memory and 256 TMEM columns per CTA -> two CTAs per SM, so one CTA's MMAs run
// under the other's exponentials.  Output: O / l in bf16, one 128-byte row segment per thread.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "sm100.cuh"

namespace wb {

int make_tmap_h16(CUtensorMap *map, const void *base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_elems,
                   uint64_t stride2_elems, uint32_t box_rows, int rank);  // gemm.cu

static constexpr int ATT_THREADS = 192;
static constexpr int TILE_BYTES = 128 * 64 * 2;  // 128 rows x 64 bf16
// KB = keys per block.  128: 96 KB of shared memory, 256 TMEM columns, 2 CTAs per SM.  64: 58 KB, 128 columns,
// 3 CTAs per SM -- three softmax warps per SM sub-partition instead of two keep the MUFU pipe busier while
// the other warps sit in their TMEM-load / max / P-store / barrier phases.
template <int KB>
struct AttCfg {
    static constexpr int KV_BYTES = KB * 64 * 2;           // K or V tile
    static constexpr int P_BYTES = (KB / 64) * TILE_BYTES;  // P: KB / 64 sub-tiles of [128 queries x 64 keys]
    static constexpr int SMEM = 1024 + TILE_BYTES + 3 * KV_BYTES + P_BYTES + 256;  // Q, K0, K1, V, P + barriers
    static constexpr int TMEM_COLS = 2 * KB;                // S: KB columns, O: 64 columns
    static constexpr int MIN_CTAS = KB == 128 ? 2 : 3;
};

struct AttnTcParams {
    CUtensorMap qkv_map;  // dims (3D, S, B), box (64, 128, 1): Q tiles
    CUtensorMap kv_map;   // same tensor, box (64, KB, 1): K / V tiles
    h16 *out;   // [B*S][D]
    int S, D, H, n_kblocks;
    unsigned long long *dbg;  // optional SM-clock timestamp dump (development aid): [2 CTAs][2 roles][16 blocks][8 events]
    int dbg_cta1;             // linear id of the second traced CTA (the first one is CTA 0)
};

#define EA_STAMP(role, blk, ev)                                                                               \
    do {                                                                                                      \
        if (DBG && dbg_slot >= 0 && (blk) < 16) P.dbg[((dbg_slot * 2 + (role)) * 16 + (blk)) * 8 + (ev)] = (unsigned long long)clock64(); \
    } while (0)

template <int KB, bool DBG>  // DBG: timestamp dump build (costs ~12 %, only launched by the WB_EA_DBG debug hook)
__global__ void __launch_bounds__(ATT_THREADS, AttCfg<KB>::MIN_CTAS)
    encoder_attn_tc_kernel(const __grid_constant__ AttnTcParams P) {
    using C = AttCfg<KB>;
    constexpr int KV_BYTES = C::KV_BYTES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *tiles = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sQ = tiles, *sK = tiles + TILE_BYTES, *sV = sK + 2 * KV_BYTES, *sP = sV + KV_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sP + C::P_BYTES);
    uint64_t *q_full = bars, *k_full = bars + 1, *k_empty = bars + 3, *v_full = bars + 5, *s_full = bars + 6,
             *s_empty = bars + 7, *p_full = bars + 8, *pv_done = bars + 9;
    uint32_t *tmem_holder = reinterpret_cast<uint32_t *>(bars + 10);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
    const int nblk = P.n_kblocks;
    const int lin_cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    const int dbg_slot = !DBG ? -1 : (lin_cta == 0 ? 0 : (lin_cta == P.dbg_cta1 ? 1 : -1));
    if (DBG && threadIdx.x == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        if (dbg_slot >= 0) P.dbg[2 * 2 * 16 * 8 + dbg_slot] = smid;
        if (lin_cta < 512) P.dbg[2 * 2 * 16 * 8 + 2 + lin_cta] = smid + 1;  // which early CTAs share an SM
    }

    if (warp == 4 && lane == 0) {
        ptx::prefetch_tmap(&P.qkv_map);
        ptx::prefetch_tmap(&P.kv_map);
        ptx::mbar_init(q_full, 1);
        for (int i = 0; i < 2; i++) ptx::mbar_init(&k_full[i], 1), ptx::mbar_init(&k_empty[i], 1);
        ptx::mbar_init(v_full, 1);
        ptx::mbar_init(s_full, 1);
        ptx::mbar_init(s_empty, 4);
        ptx::mbar_init(p_full, 4);
        ptx::mbar_init(pv_done, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 5) {
        ptx::tmem_alloc(tmem_holder, C::TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    const uint32_t tS = tmem_base, tO = tmem_base + KB;

    if (warp == 4) {
        // ===== TMA producer =====
        if (lane == 0) {
            ptx::mbar_expect_tx(q_full, TILE_BYTES);
            ptx::tma_load_3d(sQ, &P.qkv_map, q_full, h * 64, q0, b);
            auto load_k = [&](int j) {
                const int s = j & 1;
                ptx::mbar_wait(&k_empty[s], ((j >> 1) & 1) ^ 1);
                ptx::mbar_expect_tx(&k_full[s], KV_BYTES);
                ptx::tma_load_3d(sK + s * KV_BYTES, &P.kv_map, &k_full[s], P.D + h * 64, j * KB, b);
            };
            load_k(0);
            for (int j = 0; j < nblk; j++) {
                if (j + 1 < nblk) load_k(j + 1);  // K runs one block ahead of V
                if (j > 0) ptx::mbar_wait(pv_done, (j - 1) & 1);  // V buffer is free once PV_{j-1} has completed
                ptx::mbar_expect_tx(v_full, KV_BYTES);
                ptx::tma_load_3d(sV, &P.kv_map, v_full, 2 * P.D + h * 64, j * KB, b);
            }
        }
    } else if (warp == 5) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc_s = ptx::umma_idesc_h16(128, KB, 0, 0);  // Q (K-major) x K (K-major)
            constexpr uint32_t idesc_o = ptx::umma_idesc_h16(128, 64, 0, 1);   // P (K-major) x V (MN-major)
            const uint64_t q_desc = ptx::umma_desc_sw128(ptx::smem_u32(sQ), 1, 64);
            auto issue_s = [&](int j) {
                const int s = j & 1;
                ptx::mbar_wait(&k_full[s], (j >> 1) & 1);
                ptx::mbar_wait(s_empty, (j & 1) ^ 1);  // softmax has drained S_{j-1} from TMEM
                ptx::tc_fence_after();
                const uint64_t k_desc = ptx::umma_desc_sw128(ptx::smem_u32(sK + s * KV_BYTES), 1, 64);
                EA_STAMP(1, j, 0);
#pragma unroll
                for (int k = 0; k < 4; k++) ptx::mma_h16_ss(tS, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0);
                ptx::mma_commit(&k_empty[s]);
                ptx::mma_commit(s_full);
            };
            ptx::mbar_wait(q_full, 0);
            issue_s(0);
            for (int j = 0; j < nblk; j++) {
                if (j + 1 < nblk) issue_s(j + 1);
                ptx::mbar_wait(p_full, j & 1);
                EA_STAMP(1, j, 1);
                ptx::mbar_wait(v_full, j & 1);
                EA_STAMP(1, j, 2);
                ptx::tc_fence_after();
#pragma unroll
                for (int k = 0; k < KB / 16; k++) {
                    // P: 64-key sub-tiles, 32 B per 16-key step inside a sub-tile.
                    const uint64_t p_desc =
                        ptx::umma_desc_sw128(ptx::smem_u32(sP + (k >> 2) * TILE_BYTES), 1, 64) + 2 * (k & 3);
                    // V (MN-major): 16 keys = two 8-row groups of 1024 B.
                    const uint64_t v_desc = ptx::umma_desc_sw128(ptx::smem_u32(sV + k * 2048), 1, 64);
                    ptx::mma_h16_ss(tO, p_desc, v_desc, idesc_o, (j | k) != 0);
                }
                ptx::mma_commit(pv_done);
            }
        }
    } else {
        // ===== softmax warps: thread = query row (TMEM lane 32*warp + lane) =====
        const int row = warp * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
        const float c = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
        // Lazy rescale: the running shift m_use only moves when a block's maximum exceeds it by more than
        // TAU (in score units: 2^8 after scaling), so p <= 256 and the O/l rescale is rare after block 0.
        const float TAU = 8.0f / c;
        float m_use = -INFINITY, l_run = 0.f;
        uint8_t *p_row = sP + row * 128;
        const int sw = row & 7;
        for (int j = 0; j < nblk; j++) {
            const int kvalid = min(KB, P.S - j * KB);  // keys of this block inside the chunk
            ptx::mbar_wait(s_full, j & 1);
            if (threadIdx.x == 0) EA_STAMP(0, j, 0);
            ptx::tc_fence_after();
            // the whole score row of the block into registers with one wait, then hand S back to the MMA warp
            uint32_t v[KB];
#pragma unroll
            for (int cc = 0; cc < KB / 32; cc++) ptx::tmem_ld_32x32b_x32(tS + lane_addr + cc * 32, v + cc * 32);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int t = 0; t < KB; t++) asm volatile("" : "+r"(v[t]));  // no use of v[] may move above the wait
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(s_empty);  // S_{j+1} = Q K_{j+1}^T runs under this block's exponentials
            if (threadIdx.x == 0) EA_STAMP(0, j, 1);
            if (kvalid < KB) {
#pragma unroll
                for (int t = 0; t < KB; t++)
                    if (t >= kvalid) v[t] = 0xff800000u;  // -inf: exp2 -> 0, never the maximum
            }
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
            for (int t = 0; t < KB; t += 4) {
                mx0 = fmaxf(mx0, __uint_as_float(v[t])), mx1 = fmaxf(mx1, __uint_as_float(v[t + 1]));
                mx2 = fmaxf(mx2, __uint_as_float(v[t + 2])), mx3 = fmaxf(mx3, __uint_as_float(v[t + 3]));
            }
            const float m_blk = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
            float alpha = 1.0f;
            if (j == 0) {
                m_use = m_blk;
            } else if (m_blk > m_use + TAU) {
                alpha = ptx::ex2((m_use - m_blk) * c);
                m_use = m_blk;
            }
            const float nmc = -m_use * c;
            if (threadIdx.x == 0) EA_STAMP(0, j, 2);
            // p = exp2(s*c - m*c) -> bf16 pairs, in place over the score registers; row sum in fp32
            // (packed f32x2 FMA / ADD: same per-lane IEEE results, half the issue slots next to the MUFU stream;
            // measured 776 -> 751 us per launch)
            float2 l01 = make_float2(0.f, 0.f);
            const float2 c2 = make_float2(c, c), nmc2 = make_float2(nmc, nmc);
#pragma unroll
            for (int t = 0; t < KB; t += 2) {
                const float2 a0 = __ffma2_rn(make_float2(__uint_as_float(v[t]), __uint_as_float(v[t + 1])), c2, nmc2);
                const float2 p01 = make_float2(ptx::ex2(a0.x), ptx::ex2(a0.y));
                l01 = __fadd2_rn(l01, p01);
                v[t >> 1] = pack_h2(p01.x, p01.y);
            }
            l_run = l_run * alpha + (l01.x + l01.y);
            if (threadIdx.x == 0) EA_STAMP(0, j, 3);
            if (j > 0) ptx::mbar_wait(pv_done, (j - 1) & 1);  // P buffer free, O holds blocks < j
            if (threadIdx.x == 0) EA_STAMP(0, j, 4);
            // P -> swizzled smem (K-major A operand): 64-key sub-tiles x 8 chunks of 16 B
#pragma unroll
            for (int i = 0; i < KB / 8; i++) {
                uint8_t *dst = p_row + (i >> 3) * TILE_BYTES;
                *reinterpret_cast<uint4 *>(dst + (((i & 7) ^ sw) << 4)) =
                    make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
            // rescale the running output when some row of this warp moved its shift
            if (j > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {
                ptx::tc_fence_after();
#pragma unroll 1
                for (int cc = 0; cc < 2; cc++) {
                    uint32_t o[32];
                    ptx::tmem_ld_32x32b_x32(tO + lane_addr + cc * 32, o);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int t = 0; t < 32; t++) o[t] = __float_as_uint(__uint_as_float(o[t]) * alpha);
                    ptx::tmem_st_32x32b_x32(tO + lane_addr + cc * 32, o);
                }
                ptx::tmem_st_wait();
            }
            ptx::fence_proxy_async_smem();  // P (generic-proxy stores) -> visible to the MMA's async proxy
            if (threadIdx.x == 0) EA_STAMP(0, j, 5);
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(p_full);
            if (threadIdx.x == 0) EA_STAMP(0, j, 6);
        }
        // epilogue: O / l -> bf16 -> global
        ptx::mbar_wait(pv_done, (nblk - 1) & 1);
        ptx::tc_fence_after();
        const float inv = 1.0f / l_run;
        const int q = q0 + row;
        h16 *dst = P.out + ((size_t)b * P.S + q) * P.D + h * 64;
#pragma unroll 1
        for (int cc = 0; cc < 2; cc++) {
            uint32_t o[32];
            ptx::tmem_ld_32x32b_x32(tO + lane_addr + cc * 32, o);
            ptx::tmem_ld_wait();
            if (q < P.S) {
#pragma unroll
                for (int t = 0; t < 32; t += 8) {
                    uint4 u;
                    u.x = pack_h2(__uint_as_float(o[t]) * inv, __uint_as_float(o[t + 1]) * inv);
                    u.y = pack_h2(__uint_as_float(o[t + 2]) * inv, __uint_as_float(o[t + 3]) * inv);
                    u.z = pack_h2(__uint_as_float(o[t + 4]) * inv, __uint_as_float(o[t + 5]) * inv);
                    u.w = pack_h2(__uint_as_float(o[t + 6]) * inv, __uint_as_float(o[t + 7]) * inv);
                    *reinterpret_cast<uint4 *>(dst + cc * 32 + t) = u;
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 5) ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
}

template <int KB, bool DBG>
static int launch_attn(cudaStream_t st, AttnTcParams &P, const h16 *qkv, int B, int S, int H, int D) {
    WB_CHECK(make_tmap_h16(&P.kv_map, qkv, (uint64_t)3 * D, (uint64_t)S, (uint64_t)B, (uint64_t)3 * D, (uint64_t)S * 3 * D,
                            KB, 3));
    P.n_kblocks = cdiv(S, KB);
    WB_CUDA(ensure_dyn_smem(encoder_attn_tc_kernel<KB, DBG>, AttCfg<KB>::SMEM));
    dim3 grid(cdiv(S, 128), H, B);
    encoder_attn_tc_kernel<KB, DBG><<<grid, ATT_THREADS, AttCfg<KB>::SMEM, st>>>(P);
    WB_LAUNCHED();
    return WB_OK;
}

unsigned long long *g_ea_dbg = nullptr;  // set by the debug hook (WB_EA_DBG) to collect timestamps
int g_ea_dbg_cta1 = 148;
int g_attn_kb = 128;  // keys per block of the encoder attention kernel (WB_ATTN_KB=64: 3-CTA/SM variant, measured equal)

int encoder_attention_tc(cudaStream_t st, const h16 *qkv, h16 *out, int B, int S, int H, int D) {
    if (B <= 0) return WB_OK;
    WB_ARG(D == H * 64, "encoder_attention_tc: head_dim must be 64");
    AttnTcParams P;
    WB_CHECK(make_tmap_h16(&P.qkv_map, qkv, (uint64_t)3 * D, (uint64_t)S, (uint64_t)B, (uint64_t)3 * D,
                            (uint64_t)S * 3 * D, 128, 3));
    P.out = out, P.S = S, P.D = D, P.H = H;
    P.dbg = g_ea_dbg, P.dbg_cta1 = g_ea_dbg_cta1;
    static const bool env_once = [] {
        const char *e = getenv("WB_ATTN_KB");
        if (e) g_attn_kb = atoi(e) == 128 ? 128 : 64;
        return true;
    }();
    (void)env_once;
    if (P.dbg) return launch_attn<128, true>(st, P, qkv, B, S, H, D);
    return g_attn_kb == 128 ? launch_attn<128, false>(st, P, qkv, B, S, H, D) : launch_attn<64, false>(st, P, qkv, B, S, H, D);
}

}  // namespace wb
