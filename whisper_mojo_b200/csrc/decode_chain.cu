// decode_chain.cu -- persistent phase-list kernel for the dense part of one decode step (see decode_chain.h).
//
// Grid: one CTA per SM (cooperative launch: every CTA is resident, which the inter-CTA arrival counters rely on).
// Roles per CTA, as in gemm_tc_kernel: warp 0 = TMA producer (6-stage ring, 128-byte swizzle), warp 1 = tcgen05.mma
// issuer (UMMA 128 x 128 x 16, fp32 accumulators double-buffered in TMEM), warps 2..9 = epilogue AND row phases.
// Every role walks the phase list in order with its own loop; they meet only through the mbarrier pipeline.
//
// Dependencies.  Phase p consumes, for a row tile mt (128 chunks), what phase p-1 produced for the same rows.
// counters[p][mt] counts arrivals of phase p's producers for that tile (8 epilogue warps per GEMM tile; 8 warps per
// 16-row unit of a row phase).  Consumers poll with ld.acquire.gpu; producers store, fence (gpu scope + the
// generic->async proxy fence, since the next reader is a TMA load), then red.release.gpu.  Tiles are walked row-tile
// major in every phase, so row tile 0 of phase p+1 starts while the later row tiles of phase p are still running.
// No cycles: a wait only ever points at an EARLIER phase, and every role finishes its phase-p work before p+1.
#include "decode_chain.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "gemm_dev.cuh"

namespace wb {

int make_tmap_h16(CUtensorMap *map, const void *base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_elems,
                  uint64_t stride2_elems, uint32_t box_rows, int rank);  // gemm.cu

enum { PH_GEMM = 0, PH_ROWS = 1 };
// 4 ring stages of 32 KB + one 4 KB output staging tile per epilogue warp and 32-column chunk (TMA stores)
static constexpr int CH_STAGES = 4;
static constexpr int CH_STAGE_OUT_BYTES = 8 * 2 * 4096;
static constexpr int CH_SMEM_BYTES = 1024 + CH_STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + CH_STAGE_OUT_BYTES + 256 + 8 * 128 * 4;
static constexpr int UNIT_ROWS = 8;  // rows per unit of a row phase: one per epilogue warp

struct ChainPhase {
    CUtensorMap a_map;  // A [M][K], box {64, 128}
    CUtensorMap b_map;  // W [N][K], box {64, 128}
    CUtensorMap o_map;  // tma_out: output as {N, M, splits}, box {32, 32, 1} (fp32: 128B swizzle, 16-bit: 64B swizzle)
    int tma_out;        // plain [rows][N] output: coalesced bulk tensor stores instead of per-thread row stores
    const void *w_base;  // the phase's weight matrix, for the L2 prefetch at kernel start
    unsigned long long w_bytes;
    unsigned long long w_per_cta;  // bytes of it each CTA prefetches (multiple of 4 KB; set at launch-geometry time)
    GemmDev d;          // rows_per_batch = M, batches = split_k ("batch" b = K slice), num_kb per slice
    int type, tiles_n, splits;
    // row phase
    float *x;
    const float *part;
    int n_split;
    const float *rbias, *gamma, *beta;
    h16 *xn;
    int embed;
    const float *tok_emb, *pos_emb;
    const int *cur_tok, *pos_dev;
    int vocab, n_pos;
};

struct ChainParams {
    int n_phases, M, D, tiles_m;
    int pf_mode;    // L2 prefetch of the chain's weights at kernel start: 0 = none, 1 = bulk prefetches (TMA unit), 2 = per line (LSU)
    int *counters;  // [CHAIN_MAX_PHASES][tiles_m]
    unsigned long long *dbg;  // development aid (WB_CHAIN_DBG): CTA 0's timestamps [phase][role][8]
    ChainPhase ph[CHAIN_MAX_PHASES];
};

struct ChainPlan {
    ChainParams P;
    int grid = 1;
    int grid_final = 0;  // CTAs of the launch once the device is known (chain_grid)
};

#define CH_STAMP(p, role, ev)                                                                          \
    do {                                                                                               \
        if (P.dbg && blockIdx.x == 0) P.dbg[((p)*3 + (role)) * 8 + (ev)] = ptx::globaltimer_ns();      \
    } while (0)

__device__ __forceinline__ void chain_wait(const int *ctr, int target) {
    uint64_t t0 = 0;
    for (uint32_t spins = 1;; ++spins) {
        int v;
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        if (v >= target) return;
        if ((spins & 0x3ffu) == 0) {  // bounded: a protocol bug traps instead of hanging the GPU
            uint64_t t = ptx::globaltimer_ns();
            if (t0 == 0) t0 = t;
            else if (t - t0 > 10000000000ull) __trap();
        }
    }
}
__device__ __forceinline__ void chain_signal(int *ctr) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(ctr), "r"(1) : "memory");
}
// After cp.async.bulk.wait_group 0 the bulk stores ARE performed: nothing of this thread is in flight, so the arrival
// needs no release fence (a releasing reduction costs a MEMBAR.GPU, 1.5-3 us under load: timestamps in profiles/)
__device__ __forceinline__ void chain_signal_relaxed(int *ctr) {
    asm volatile("red.relaxed.gpu.global.add.s32 [%0], %1;" ::"l"(ctr), "r"(1) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// arrivals that complete row tile `mt` of phase `q`
__device__ __forceinline__ int chain_target(const ChainParams &P, int q, int mt) {
    const ChainPhase &ph = P.ph[q];
    if (ph.type == PH_GEMM) return ph.tiles_n * ph.splits * 8;
    const int rows = min(BM, P.M - mt * BM);
    return ((rows + UNIT_ROWS - 1) / UNIT_ROWS) * 8;
}

// R rows of a row phase handled by one warp (D % 128 == 0, D <= 1024), all their loads in flight together: see
// ChainRows.  g / b: this lane's slices of gamma / beta, loaded by the caller BEFORE it waited for the producers.
// NV = D / 128 float4 per lane and row (compile time: the arrays must stay in registers).
template <int R, int NV>
__device__ __forceinline__ void chain_rows(const ChainPhase &ph, int row0, int M, int lane) {
    constexpr int nvec = NV, D = NV * 128;
    float4 v[R][NV];
    float4 g[NV], b[NV];  // gamma / beta slices: constants, their loads are independent of everything below
    if (ph.gamma) {
#pragma unroll
        for (int i = 0; i < NV; i++) {
            g[i] = __ldg(reinterpret_cast<const float4 *>(ph.gamma) + i * 32 + lane);
            b[i] = __ldg(reinterpret_cast<const float4 *>(ph.beta) + i * 32 + lane);
        }
    }
    if (ph.embed) {
        int pos = *ph.pos_dev;
        pos = pos < 0 ? 0 : (pos >= ph.n_pos ? ph.n_pos - 1 : pos);
        const float4 *pe = reinterpret_cast<const float4 *>(ph.pos_emb + (size_t)pos * D);
#pragma unroll
        for (int r = 0; r < R; r++) {
            if (row0 + r >= M) continue;
            int tok = ph.cur_tok[row0 + r];
            tok = tok < 0 ? 0 : (tok >= ph.vocab ? ph.vocab - 1 : tok);
            const float4 *te = reinterpret_cast<const float4 *>(ph.tok_emb + (size_t)tok * D);
#pragma unroll
            for (int i = 0; i < NV; i++)
                if (i < nvec) {
                    const float4 a = te[i * 32 + lane], c = pe[i * 32 + lane];
                    v[r][i] = make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w);
                    reinterpret_cast<float4 *>(ph.x + (size_t)(row0 + r) * D)[i * 32 + lane] = v[r][i];
                }
        }
    } else {
        const long long split_stride = (long long)M * D;
        // every load of both rows is issued before the first use: one L2 round trip, not one per slice
        // slices whose loads are in flight together (one L2 round trip per SB slices).  Six at once (the longest
        // reduction of d_model <= 384 in one round trip) spills at the kernel's 168-register cap and measured slower.
        constexpr int SB = NV <= 3 ? 4 : 2;
        float4 pp[SB][R][NV];
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int i = 0; i < NV; i++)
                if (i < nvec && row0 + r < M) v[r][i] = reinterpret_cast<const float4 *>(ph.x + (size_t)(row0 + r) * D)[i * 32 + lane];
        for (int s0 = 0; s0 < ph.n_split; s0 += SB) {
#pragma unroll
            for (int k = 0; k < SB; k++)
#pragma unroll
                for (int r = 0; r < R; r++)
#pragma unroll
                    for (int i = 0; i < NV; i++)
                        if (s0 + k < ph.n_split && row0 + r < M)
                            // written by other CTAs of this launch, and rewritten by a later phase of it: ld.global.cg, so
                            // a stale L1 line from an earlier read of these addresses on this SM cannot be hit
                            pp[k][r][i] = __ldcg(reinterpret_cast<const float4 *>(ph.part + (size_t)(s0 + k) * split_stride + (size_t)(row0 + r) * D) + i * 32 + lane);
#pragma unroll
            for (int k = 0; k < SB; k++)  // fixed order: x + bias, then the slices in order (as resid_ln): deterministic
#pragma unroll
                for (int r = 0; r < R; r++)
#pragma unroll
                    for (int i = 0; i < NV; i++)
                        if (s0 + k < ph.n_split && row0 + r < M) {
                            float4 a = v[r][i];
                            if (s0 + k == 0 && ph.rbias) {
                                const float4 c = __ldg(reinterpret_cast<const float4 *>(ph.rbias) + i * 32 + lane);
                                a.x += c.x, a.y += c.y, a.z += c.z, a.w += c.w;
                            }
                            a.x += pp[k][r][i].x, a.y += pp[k][r][i].y, a.z += pp[k][r][i].z, a.w += pp[k][r][i].w;
                            v[r][i] = a;
                        }
        }
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int i = 0; i < NV; i++)
                if (i < nvec && row0 + r < M) {
                    if (ph.n_split == 0 && ph.rbias) {
                        const float4 c = __ldg(reinterpret_cast<const float4 *>(ph.rbias) + i * 32 + lane);
                        v[r][i].x += c.x, v[r][i].y += c.y, v[r][i].z += c.z, v[r][i].w += c.w;
                    }
                    reinterpret_cast<float4 *>(ph.x + (size_t)(row0 + r) * D)[i * 32 + lane] = v[r][i];
                }
    }
    if (!ph.gamma) return;
    float s[R], q[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        s[r] = q[r] = 0.f;
#pragma unroll
        for (int i = 0; i < NV; i++)
            if (i < nvec && row0 + r < M) {
                s[r] += v[r][i].x + v[r][i].y + v[r][i].z + v[r][i].w;
                q[r] += v[r][i].x * v[r][i].x + v[r][i].y * v[r][i].y + v[r][i].z * v[r][i].z + v[r][i].w * v[r][i].w;
            }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int r = 0; r < R; r++) {
            s[r] += __shfl_xor_sync(0xffffffffu, s[r], o);
            q[r] += __shfl_xor_sync(0xffffffffu, q[r], o);
        }
#pragma unroll
    for (int r = 0; r < R; r++) {
        if (row0 + r >= M) continue;
        const float mean = s[r] / (float)D;  // one-pass variance, whisper_tensor.mojo:249-285
        const float var = q[r] / (float)D - mean * mean;
        const float inv_std = 1.0f / sqrtf(var + 1e-5f);
#pragma unroll
        for (int i = 0; i < NV; i++)
            if (i < nvec) {
                uint2 pk;
                pk.x = pack_h2((v[r][i].x - mean) * inv_std * g[i].x + b[i].x, (v[r][i].y - mean) * inv_std * g[i].y + b[i].y);
                pk.y = pack_h2((v[r][i].z - mean) * inv_std * g[i].z + b[i].z, (v[r][i].w - mean) * inv_std * g[i].w + b[i].w);
                reinterpret_cast<uint2 *>(ph.xn + (size_t)(row0 + r) * D)[i * 32 + lane] = pk;
            }
    }
}

template <int EPI>
__device__ __forceinline__ void chain_epilogue(const GemmDev &p, int b, int m, int n_first, const uint32_t *v0,
                                               const uint32_t *v1, const float *sbias) {
    // (the accumulator is already in registers; global loads this needs are none for the three epilogues used here)
    EpiChunk<EPI> e0, e1;
    epi_prefetch<EPI>(p, b, m, n_first, e0);
    epi_prefetch<EPI>(p, b, m, n_first + 32, e1);
    float best = 0.f;
    int best_idx = 0;
    epi_finish<EPI>(p, b, m, n_first, v0, e0, sbias, best, best_idx);
    epi_finish<EPI>(p, b, m, n_first + 32, v1, e1, sbias + 32, best, best_idx);
}

// NV = d_model / 128 (the row phases keep a row in registers).
// The parameter block stays a __grid_constant__ kernel parameter: a copy in global memory that all threads pull into
// shared memory at kernel start was tried (to avoid first-touch constant-cache misses per phase and role) and measured
// slower -- 68.5 -> 72.2 ms at batch 1, 114.5 -> 117.6 at 256 chunks: the extra global round trip in front of the CTA
// barrier and tensor maps fetched from global memory cost more than the misses.
template <int NV>
__global__ void __launch_bounds__(TC_THREADS, 1) decode_chain_kernel(const __grid_constant__ ChainParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *tiles = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *smem_a = tiles;
    constexpr int STAGES = CH_STAGES;
    constexpr int RING_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES);
    uint8_t *smem_b = tiles + STAGES * A_STAGE_BYTES;
    uint8_t *stage_out = tiles + RING_BYTES;  // [8 warps][2 chunks][4 KB], 1024-byte aligned (swizzled TMA source)
    uint64_t *bars = reinterpret_cast<uint64_t *>(tiles + RING_BYTES + CH_STAGE_OUT_BYTES);
    uint64_t *full_bar = bars, *empty_bar = bars + STAGES;
    uint64_t *tmem_full = bars + 2 * STAGES, *tmem_empty = bars + 2 * STAGES + 2;
    uint32_t *tmem_holder = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 4);
    float *bias_smem = reinterpret_cast<float *>(tiles + RING_BYTES + CH_STAGE_OUT_BYTES + 256);  // [8][128]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = (int)gridDim.x;
    if (P.dbg && threadIdx.x == 0 && blockIdx.x < 248) P.dbg[CHAIN_MAX_PHASES * 3 * 8 + blockIdx.x] = ptx::globaltimer_ns();


#define CH_PRO(k)                                                                                       \
    do {                                                                                                \
        if (P.dbg && blockIdx.x == 0 && lane == 0) P.dbg[CHAIN_MAX_PHASES * 3 * 8 + 504 + (k)] = ptx::globaltimer_ns(); \
    } while (0)
    // Prologue, spread over the warps so that the CTA barrier falls ~0.9 us after the kernel starts (one thread walking
    // every phase's tensor maps and one warp computing the L2 prefetch ranges in front of the barrier took 2.2-2.7 us
    // per launch, nine launches per decode step: prologue timestamps in profiles/)
    if (warp == 3 && lane < P.n_phases && P.ph[lane].type == PH_GEMM) {  // lane p: the tensor maps of phase p
        ptx::prefetch_tmap(&P.ph[lane].a_map);
        ptx::prefetch_tmap(&P.ph[lane].b_map);
        if (P.ph[lane].tma_out) ptx::prefetch_tmap(&P.ph[lane].o_map);
    }
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; s++) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; s++) {
            ptx::mbar_init(&tmem_full[s], 1);
            ptx::mbar_init(&tmem_empty[s], 8);
        }
        ptx::fence_barrier_init();
        CH_PRO(1);  // tensor maps prefetched, barriers initialised
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_holder, 256);
        ptx::tmem_relinquish();
        CH_PRO(2);  // TMEM allocated
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    if (warp == 3) CH_PRO(3);  // past the CTA barrier
    if (warp >= 2) {
        // The attention kernel that ran before this one streamed gigabytes through the L2, so every weight matrix of
        // this chain is cold; its first reader would pay the HBM latency per ring refill.  Each CTA asks the L2 for
        // its 1/G slice of every matrix of the chain right away (4 KB pieces, one per epilogue thread: a handful of
        // instructions each, issued while these warps would otherwise wait for the first accumulator), so the later
        // phases' weights arrive while the first phases run.
        const unsigned tid = (unsigned)(warp - 2) * 32u + (unsigned)lane;
        for (int p = 0; p < P.n_phases; p++) {
            const ChainPhase &ph = P.ph[p];
            if (ph.type != PH_GEMM || !ph.w_bytes) continue;
            const unsigned long long lo = ph.w_per_cta * blockIdx.x, hi = lo + ph.w_per_cta < ph.w_bytes ? lo + ph.w_per_cta : ph.w_bytes;
            if (P.pf_mode == 1) {
                for (unsigned long long off = lo + 4096ull * tid; off < hi; off += 4096ull * 256) {
                    const unsigned n = (unsigned)(hi - off < 4096ull ? hi - off : 4096ull) & ~15u;
                    if (n) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const char *>(ph.w_base) + off), "r"(n) : "memory");
                }
            } else if (P.pf_mode == 2) {  // one 128-byte line per instruction through the load/store unit (the TMA unit stays free)
                for (unsigned long long off = lo + 128ull * tid; off < hi; off += 128ull * 256)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(ph.w_base) + off) : "memory");
            }
        }
        if (warp == 2) CH_PRO(0);  // L2 prefetches issued
    }

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int p = 0; p < P.n_phases; p++) {
                const ChainPhase &ph = P.ph[p];
                if (ph.type != PH_GEMM) continue;
                const int per_mt = ph.tiles_n * ph.splits, total = P.tiles_m * per_mt, nkb = ph.d.num_kb;
                for (int t = blockIdx.x; t < total; t += G) {
                    const int mt = t / per_mt, r = t - mt * per_mt, sp = r / ph.tiles_n, nt = r - sp * ph.tiles_n;
                    const int koff = sp * ph.d.split_koff;
                    bool ready = (p == 0);
                    if (t == (int)blockIdx.x) CH_STAMP(p, 0, 0);
                    for (int kb0 = 0; kb0 < nkb; kb0 += STAGES) {
                        const int n = min(STAGES, nkb - kb0);
                        // weights first: they depend on nothing, so they stream while the previous phase finishes
                        int s = stage;
                        uint32_t phs = phase;
                        for (int i = 0; i < n; i++) {
                            ptx::mbar_wait(&empty_bar[s], phs ^ 1);
                            ptx::mbar_expect_tx(&full_bar[s], A_STAGE_BYTES + B_STAGE_BYTES);
                            ptx::tma_load_2d(smem_b + s * B_STAGE_BYTES, &ph.b_map, &full_bar[s], (kb0 + i) * BK + koff, nt * BN);
                            if (++s == STAGES) s = 0, phs ^= 1;
                        }
                        if (t == (int)blockIdx.x && kb0 == 0) CH_STAMP(p, 0, 1);
                        if (!ready) {  // the activations of this row tile: produced by phase p-1 of this launch
                            chain_wait(P.counters + (p - 1) * P.tiles_m + mt, chain_target(P, p - 1, mt));
                            fence_proxy_async_all();
                            ready = true;
                        }
                        if (t == (int)blockIdx.x && kb0 == 0) CH_STAMP(p, 0, 2);
                        s = stage;
                        for (int i = 0; i < n; i++) {
                            ptx::tma_load_2d(smem_a + s * A_STAGE_BYTES, &ph.a_map, &full_bar[s], (kb0 + i) * BK + koff, mt * BM);
                            if (++s == STAGES) s = 0;
                        }
                        stage = s, phase = phs;
                        if (t == (int)blockIdx.x && kb0 == 0) CH_STAMP(p, 0, 3);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::umma_idesc_h16(BM, BN, 0, 0);
            int stage = 0, it = 0;
            uint32_t phase = 0;
            for (int p = 0; p < P.n_phases; p++) {
                const ChainPhase &ph = P.ph[p];
                if (ph.type != PH_GEMM) continue;
                const int total = P.tiles_m * ph.tiles_n * ph.splits, nkb = ph.d.num_kb;
                for (int t = blockIdx.x; t < total; t += G, it++) {
                    const int acc = it & 1;
                    ptx::mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);
                    ptx::tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * BN;
                    for (int kb = 0; kb < nkb; kb++) {
                        ptx::mbar_wait(&full_bar[stage], phase);
                        if (t == (int)blockIdx.x && kb == 0) CH_STAMP(p, 1, 0);
                        if (t == (int)blockIdx.x && kb >= 1 && kb <= 6) CH_STAMP(p, 1, 1 + kb);  // k-blocks 1..6 have landed
                        ptx::tc_fence_after();
                        const uint64_t a_desc = ptx::umma_desc_sw128(ptx::smem_u32(smem_a + stage * A_STAGE_BYTES), 1, 64);
                        const uint64_t b_desc = ptx::umma_desc_sw128(ptx::smem_u32(smem_b + stage * B_STAGE_BYTES), 1, 64);
#pragma unroll
                        for (int k = 0; k < BK / 16; k++)
                            ptx::mma_h16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
                        ptx::mma_commit(&empty_bar[stage]);
                        if (++stage == STAGES) stage = 0, phase ^= 1;
                    }
                    ptx::mma_commit(&tmem_full[acc]);
                    if (t == (int)blockIdx.x) CH_STAMP(p, 1, 1);
                }
            }
        }
    } else {
        // ===== epilogue warps (TMEM lanes 32 * (warp % 4) .. +31, columns 64 * half .. +63) and row phases =====
        const int q = warp & 3, half = (warp - 2) >> 2, ew = warp - 2;
        float *sbias = bias_smem + ew * 128;
        int it = 0;
        for (int p = 0; p < P.n_phases; p++) {
            const ChainPhase &ph = P.ph[p];
            int *ctr = P.counters + p * P.tiles_m;
            const bool signal = p + 1 < P.n_phases;  // the last phase is followed by the kernel boundary
            if (ph.type == PH_GEMM) {
                const GemmDev &d = ph.d;
                const int per_mt = ph.tiles_n * ph.splits, total = P.tiles_m * per_mt;
                for (int t = blockIdx.x; t < total; t += G, it++) {
                    const int mt = t / per_mt, r = t - mt * per_mt, sp = r / ph.tiles_n, nt = r - sp * ph.tiles_n;
                    const int acc = it & 1;
                    const int m = mt * BM + q * 32 + lane;
                    const int n_first = nt * BN + half * 64;
                    stage_bias(d, n_first, 64, sbias, lane);  // before the wait: its global loads overlap the MMAs
                    ptx::mbar_wait(&tmem_full[acc], (it >> 1) & 1);
                    if (t == (int)blockIdx.x && ew == 0 && lane == 0) CH_STAMP(p, 2, 0);
                    ptx::tc_fence_after();
                    uint32_t v0[32], v1[32];
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + half * 64;
                    ptx::tmem_ld_32x32b_x32(taddr, v0);
                    ptx::tmem_ld_32x32b_x32(taddr + 32, v1);
                    ptx::tmem_ld_wait();
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive_relaxed(&tmem_empty[acc]);  // accumulator is in registers
                    if (ph.tma_out) {
                        // 32 rows x 32 values per chunk -> this warp's staging tile, 16-byte pieces XOR-swizzled as the
                        // output tensor map expects, then ONE bulk tensor store per chunk: full-line writes instead of
                        // 32 scattered row stores per instruction (those took 4 us per tile and another 2-4 us until
                        // the releasing signal had drained them; timestamps in profiles/).  Rows / columns outside
                        // [M, N) are clipped by the tensor map.
                        const int m_w = mt * BM + q * 32;
#pragma unroll
                        for (int c = 0; c < 2; c++) {
                            const uint32_t *vv = c ? v1 : v0;
                            uint8_t *buf = stage_out + (ew * 2 + c) * 4096;
                            float o[32];
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                const float4 bv = *reinterpret_cast<const float4 *>(sbias + c * 32 + 4 * j);
                                o[4 * j] = __uint_as_float(vv[4 * j]) + bv.x, o[4 * j + 1] = __uint_as_float(vv[4 * j + 1]) + bv.y;
                                o[4 * j + 2] = __uint_as_float(vv[4 * j + 2]) + bv.z, o[4 * j + 3] = __uint_as_float(vv[4 * j + 3]) + bv.w;
                            }
                            if (d.epi == EPI_STORE_F32) {
#pragma unroll
                                for (int j = 0; j < 8; j++)
                                    *reinterpret_cast<float4 *>(buf + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                                        make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
                            } else {
                                if (d.epi == EPI_GELU_H16) {
#pragma unroll
                                    for (int j = 0; j < 32; j += 2) {
                                        const float2 gl = gelu_fast2(make_float2(o[j], o[j + 1]));
                                        o[j] = gl.x, o[j + 1] = gl.y;
                                    }
                                }
#pragma unroll
                                for (int j = 0; j < 4; j++)
                                    *reinterpret_cast<uint4 *>(buf + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) =
                                        make_uint4(pack_h2(o[8 * j], o[8 * j + 1]), pack_h2(o[8 * j + 2], o[8 * j + 3]),
                                                   pack_h2(o[8 * j + 4], o[8 * j + 5]), pack_h2(o[8 * j + 6], o[8 * j + 7]));
                            }
                        }
                        ptx::fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            ptx::tma_store_3d(&ph.o_map, stage_out + (ew * 2) * 4096, n_first, m_w, sp);
                            ptx::tma_store_3d(&ph.o_map, stage_out + (ew * 2 + 1) * 4096, n_first + 32, m_w, sp);
                            ptx::bulk_commit_group();
                            ptx::bulk_wait_group<0>();  // written (not just read): the staging tiles are free and the data can be signalled
                        }
                        if (t == (int)blockIdx.x && ew == 0 && lane == 0) CH_STAMP(p, 2, 1);
                        if (signal) {
                            if (lane == 0) chain_signal_relaxed(ctr + mt);
                        }
                        __syncwarp();
                    } else {
                        switch (d.epi) {
                            case EPI_GELU_H16: chain_epilogue<EPI_GELU_H16>(d, sp, m, n_first, v0, v1, sbias); break;
                            case EPI_STORE_F32: chain_epilogue<EPI_STORE_F32>(d, sp, m, n_first, v0, v1, sbias); break;
                            default: chain_epilogue<EPI_STORE_H16>(d, sp, m, n_first, v0, v1, sbias); break;
                        }
                        if (t == (int)blockIdx.x && ew == 0 && lane == 0) CH_STAMP(p, 2, 1);
                        if (signal) {
                            // every lane's stores -> (generic -> async proxy fence) -> warp barrier -> one releasing
                            // reduction at gpu scope: the release is cumulative over the lanes the barrier ordered before it
                            fence_proxy_async_all();
                            __syncwarp();
                            if (lane == 0) chain_signal(ctr + mt);
                        }
                    }
                    if (t == (int)blockIdx.x && ew == 0 && lane == 0) CH_STAMP(p, 2, 2);
                }
            } else {
                const int n_units = (P.M + UNIT_ROWS - 1) / UNIT_ROWS;
                for (int u = blockIdx.x; u < n_units; u += G) {
                    const int mt = (u * UNIT_ROWS) / BM;
                    if (p > 0) {
                        // ONE poller per CTA: eight warps spinning on the same counter line from every CTA kept its L2
                        // slice saturated with polls, which is also where the producers' arrivals have to land
                        if (ew == 0 && lane == 0) chain_wait(P.counters + (p - 1) * P.tiles_m + mt, chain_target(P, p - 1, mt));
                        asm volatile("bar.sync 1, 256;" ::: "memory");  // the eight epilogue warps
                    }
                    if (u == (int)blockIdx.x && ew == 0 && lane == 0) CH_STAMP(p, 2, 0);
                    const int row0 = u * UNIT_ROWS + ew * (UNIT_ROWS / 8);
                    chain_rows<UNIT_ROWS / 8, NV>(ph, row0, P.M, lane);
                    if (u == (int)blockIdx.x && ew == 0 && lane == 0) CH_STAMP(p, 2, 1);
                    if (signal) {
                        fence_proxy_async_all();
                        __syncwarp();
                        if (lane == 0) chain_signal(ctr + mt);
                    }
                    if (u == (int)blockIdx.x && ew == 0 && lane == 0) CH_STAMP(p, 2, 2);
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (P.dbg && threadIdx.x == 0 && blockIdx.x < 248) P.dbg[CHAIN_MAX_PHASES * 3 * 8 + 256 + blockIdx.x] = ptx::globaltimer_ns();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, 256);
}

int make_tmap_any_pub(CUtensorMap *map, CUtensorMapDataType dtype, int swizzle_bytes, const void *base, int rank,
                      const uint64_t *dims, const uint64_t *strides_bytes, const uint32_t *box);  // gemm.cu

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
size_t chain_counter_ints(int M) { return (size_t)CHAIN_MAX_PHASES * cdiv(M, BM); }

int chain_split_k(int K) {
    // ~6 k-blocks of 64 per tile: one SM streams operands at ~65-75 GB/s whatever the batch (4 x 32 KB in flight per
    // ring, measured), so a tile's K extent, not its flops, sets the phase time -- 12 k-blocks took 5-6 us, 6 take
    // ~3.  The slices must tile K exactly, and the factor depends on K only: never on the batch size.
    const int kb = K / BK;
    int s = std::max(1, (kb + 5) / 6);
    while (kb % s) s++;
    return s;
}

static unsigned long long *g_chain_dbg = nullptr;
void chain_debug_init() {  // development aid: the timestamp buffer must exist before any stream capture
    if (getenv("WB_CHAIN_DBG") && !g_chain_dbg &&
        cudaMalloc((void **)&g_chain_dbg, ((size_t)CHAIN_MAX_PHASES * 3 * 8 + 512) * 8) != cudaSuccess)
        g_chain_dbg = nullptr;
}
ChainPlan *chain_plan_create(int M, int D, int *counters) {
    ChainPlan *p = new ChainPlan();
    p->P.n_phases = 0, p->P.M = M, p->P.D = D, p->P.tiles_m = cdiv(M, BM), p->P.counters = counters, p->P.dbg = nullptr;
    return p;
}
void chain_plan_destroy(ChainPlan *p) { delete p; }
int chain_plan_phases(const ChainPlan *p) { return p->P.n_phases; }

int chain_plan_add_gemm(ChainPlan *pl, const ChainGemm &g) {
    ChainParams &P = pl->P;
    WB_ARG(P.n_phases < CHAIN_MAX_PHASES, "chain: too many phases");
    WB_ARG(g.A && g.W && g.K > 0 && g.K % BK == 0 && g.N > 0 && g.split_k >= 1 && (g.K / BK) % g.split_k == 0,
           "chain: bad GEMM phase (K=%d N=%d split=%d)", g.K, g.N, g.split_k);
    WB_ARG(g.epi == EPI_STORE_H16 || g.epi == EPI_GELU_H16 || g.epi == EPI_STORE_F32, "chain: unsupported epilogue %d", g.epi);
    WB_ARG(g.split_k == 1 || (g.epi == EPI_STORE_F32 && !g.bias && g.n_seg_ptrs == 1 && !g.dyn_off),
           "chain: split-K needs a plain bias-free fp32 partial output");
    ChainPhase &ph = P.ph[P.n_phases];
    memset(&ph, 0, sizeof ph);
    ph.type = PH_GEMM;
    ph.tiles_n = cdiv(g.N, BN);
    ph.splits = g.split_k;
    ph.w_base = g.W, ph.w_bytes = (unsigned long long)g.N * g.K * sizeof(h16);
    WB_CHECK(make_tmap_h16(&ph.a_map, g.A, (uint64_t)g.K, (uint64_t)P.M, 1, (uint64_t)g.K, 0, BM, 2));
    WB_CHECK(make_tmap_h16(&ph.b_map, g.W, (uint64_t)g.K, (uint64_t)g.N, 1, (uint64_t)g.K, 0, BN, 2));
    GemmDev &d = ph.d;
    d.rows_per_batch = P.M, d.batches = g.split_k, d.N = g.N, d.K = g.K;
    d.tiles_m_per_batch = P.tiles_m, d.tiles_n = ph.tiles_n;
    d.split_koff = g.K / g.split_k;
    d.num_kb = d.split_koff / BK, d.kb_per_tap = d.num_kb, d.taps = 1;
    d.bias = g.bias, d.epi = g.epi;
    for (int i = 0; i < 3; i++) d.out[i] = g.out[i], d.out_ld[i] = g.out_ld[i], d.dyn_mult[i] = g.dyn_mult[i];
    d.seg_cols = g.seg_cols > 0 ? g.seg_cols : g.N;
    d.n_seg_ptrs = g.n_seg_ptrs;
    d.dyn_off = g.dyn_off;
    WB_ARG(d.out[0] && (d.seg_cols == g.N || d.seg_cols % 32 == 0), "chain: bad output routing");
    // plain [rows][N] outputs go through TMA stores (fp32 partials, q', h); segmented / cache-offset outputs keep the
    // per-thread stores
    const int esz = g.epi == EPI_STORE_F32 ? 4 : 2;
    ph.tma_out = g.n_seg_ptrs == 1 && !g.dyn_off && d.seg_cols == g.N && (g.out_ld[0] * esz) % 16 == 0 &&
                 (reinterpret_cast<uintptr_t>(g.out[0]) & 15) == 0;
    if (ph.tma_out) {
        const uint64_t dims[3] = {(uint64_t)g.N, (uint64_t)P.M, (uint64_t)g.split_k};
        const uint64_t str[2] = {(uint64_t)g.out_ld[0] * esz, (uint64_t)P.M * g.out_ld[0] * esz};
        const uint32_t box[3] = {32, 32, 1};
        WB_CHECK(make_tmap_any_pub(&ph.o_map, esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : H16_TMAP_DTYPE, esz == 4 ? 128 : 64,
                                   g.out[0], 3, dims, str, box));
    }
    P.n_phases++;
    pl->grid = std::max(pl->grid, P.tiles_m * ph.tiles_n * ph.splits);
    return WB_OK;
}

int chain_plan_add_rows(ChainPlan *pl, const ChainRows &r) {
    ChainParams &P = pl->P;
    WB_ARG(P.n_phases < CHAIN_MAX_PHASES, "chain: too many phases");
    WB_ARG(P.D % 128 == 0 && P.D <= 768, "chain: D=%d must be a multiple of 128 and <= 768", P.D);
    WB_ARG(r.x && (!r.gamma || (r.beta && r.xn)) && (r.embed ? (r.tok_emb && r.pos_emb && r.cur_tok && r.pos_dev) : (r.n_split == 0 || r.part)),
           "chain: bad row phase");
    ChainPhase &ph = P.ph[P.n_phases];
    memset(&ph, 0, sizeof ph);
    ph.type = PH_ROWS;
    ph.x = r.x, ph.part = r.part, ph.n_split = r.n_split, ph.rbias = r.bias, ph.gamma = r.gamma, ph.beta = r.beta, ph.xn = r.xn;
    ph.embed = r.embed ? 1 : 0;
    ph.tok_emb = r.tok_emb, ph.pos_emb = r.pos_emb, ph.cur_tok = r.cur_tok, ph.pos_dev = r.pos_dev;
    ph.vocab = r.vocab, ph.n_pos = r.n_pos;
    P.n_phases++;
    pl->grid = std::max(pl->grid, cdiv(P.M, UNIT_ROWS));
    return WB_OK;
}

// WB_CHAIN_DBG=1: every launch records CTA 0's timestamps; chain_debug_dump prints those of the LAST launch
static int g_chain_dbg_phases = 0;
void chain_debug_dump() {
    if (!g_chain_dbg) return;
    cudaDeviceSynchronize();
    const size_t n = (size_t)CHAIN_MAX_PHASES * 3 * 8 + 512;
    std::vector<unsigned long long> h(n);
    if (cudaMemcpy(h.data(), g_chain_dbg, n * 8, cudaMemcpyDeviceToHost) != cudaSuccess) return;
    unsigned long long t0 = ~0ull, s_max = 0, e_min = ~0ull, e_max = 0;
    const size_t np_ = (size_t)CHAIN_MAX_PHASES * 3 * 8;
    int n_cta = 0;
    for (size_t i = 0; i < 248; i++)
        if (h[np_ + i]) {
            n_cta++;
            t0 = std::min(t0, h[np_ + i]), s_max = std::max(s_max, h[np_ + i]);
            e_min = std::min(e_min, h[np_ + 256 + i]), e_max = std::max(e_max, h[np_ + 256 + i]);
        }
    fprintf(stderr, "%d CTAs: started within %.2f us, ended between %.2f and %.2f us after the first start (CTA 0: %.2f .. %.2f)\n", n_cta,
            (double)(s_max - t0) / 1e3, (double)(e_min - t0) / 1e3, (double)(e_max - t0) / 1e3, (double)(h[np_] - t0) / 1e3,
            (double)(h[np_ + 256] - t0) / 1e3);
    fprintf(stderr, "prologue of CTA 0 (us): L2 prefetches issued %.2f, tensor maps + barriers %.2f, TMEM allocated %.2f, past the CTA barrier %.2f\n",
            (double)(h[np_ + 504] - t0) / 1e3, (double)(h[np_ + 505] - t0) / 1e3, (double)(h[np_ + 506] - t0) / 1e3, (double)(h[np_ + 507] - t0) / 1e3);
    fprintf(stderr, "chain timestamps of the last launch, CTA 0, first tile / unit of each phase (us since the first stamp)\n"
                    "phase | producer: start B_issued dep_ready A_issued | mma: first_full commit kb1..kb6 landed | epi/rows: ready stored signalled\n");
    for (int p = 0; p < g_chain_dbg_phases; p++) {
        fprintf(stderr, "%5d |", p);
        for (int r = 0; r < 3; r++) {
            const int ne = r == 0 ? 4 : (r == 1 ? 8 : 3);  // mma: first_full, commit, then the arrival of k-blocks 1..6
            for (int e = 0; e < ne; e++) {
                const unsigned long long v = h[((size_t)p * 3 + r) * 8 + e];
                if (v) fprintf(stderr, " %8.2f", (double)(v - t0) / 1e3);
                else fprintf(stderr, "        -");
            }
            fprintf(stderr, " |");
        }
        fprintf(stderr, "\n");
    }
}

// Fix the launch geometry (CTAs of the plan on this device, each phase's per-CTA L2 prefetch slice), once all phases
// are added.
int chain_plan_finalize(ChainPlan *pl) {
    WB_ARG(pl && pl->P.n_phases > 0, "chain: empty plan");
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int G = std::min(pl->grid, sms);
    for (int p = 0; p < pl->P.n_phases; p++) {
        ChainPhase &ph = pl->P.ph[p];
        if (ph.type == PH_GEMM) ph.w_per_cta = ((ph.w_bytes / (unsigned long long)G) + 4095ull) & ~4095ull;
    }
    pl->grid_final = G;
    pl->P.dbg = g_chain_dbg;  // WB_CHAIN_DBG: every launch records CTA 0's timestamps
    static const int pf_env = getenv("WB_CHAIN_PF") ? atoi(getenv("WB_CHAIN_PF")) : 1;
    pl->P.pf_mode = pf_env;
    return WB_OK;
}

int chain_launch(cudaStream_t st, ChainPlan *pl) {
    WB_ARG(pl && pl->grid_final > 0, "chain: plan was not finalized");
    const int G = pl->grid_final;
    if (pl->P.dbg) {
        cudaMemsetAsync(pl->P.dbg, 0, ((size_t)CHAIN_MAX_PHASES * 3 * 8 + 512) * 8, st);
        g_chain_dbg_phases = pl->P.n_phases;
    }
    void (*kernel)(const ChainParams) = nullptr;
    switch (pl->P.D >> 7) {
        case 1: kernel = decode_chain_kernel<1>; break;
        case 2: kernel = decode_chain_kernel<2>; break;
        case 3: kernel = decode_chain_kernel<3>; break;
        case 4: kernel = decode_chain_kernel<4>; break;
        case 5: kernel = decode_chain_kernel<5>; break;
        case 6: kernel = decode_chain_kernel<6>; break;
        default: set_error("chain: d_model %d not supported", pl->P.D); return WB_ERR_ARG;
    }
    WB_CUDA(ensure_dyn_smem(kernel, CH_SMEM_BYTES));
    // Cooperative launch: the whole grid is resident at once (one CTA per SM), which the arrival counters need --
    // also when another stream's kernels compete for the SMs.
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(G), cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = CH_SMEM_BYTES, cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    WB_CUDA(cudaLaunchKernelEx(&cfg, kernel, pl->P));
    WB_LAUNCHED();
    // Timing experiment (results are garbage): the same kernel again with warm instruction / constant / L2 caches.
    // Measured at 256 chunks: 30.9 -> 27.2 us for the last layer's chain, all of it in the first phase and the row
    // phases; the steady 4.5-5.5 us per phase are memory round trips of the protocol, not cold misses.
    static const bool twice = getenv("WB_CHAIN_TWICE") != nullptr;
    if (twice) {
        cudaMemsetAsync(pl->P.counters, 0, (size_t)CHAIN_MAX_PHASES * pl->P.tiles_m * sizeof(int), st);
        if (pl->P.dbg) cudaMemsetAsync(pl->P.dbg, 0, ((size_t)CHAIN_MAX_PHASES * 3 * 8 + 512) * 8, st);
        WB_CUDA(cudaLaunchKernelEx(&cfg, kernel, pl->P));
    }
    return WB_OK;
}

}  // namespace wb
