// cross_attn_tc.cu -- decode-step cross-attention over the encoder output itself ("absorbed" form).
//
// The reference keeps per-layer cross K = enc*Wk^T and V = enc*Wv^T + bv (layers.mojo:148-157) and the
// decode step reads both (2 x 1500 x D values per chunk per layer, layers.mojo:186-272).  Since
//     q_h . K_j,h  = (Wk_h^T q_h) . enc_j          and     sum_j p_j V_j,h = Wv_h (sum_j p_j enc_j) + bv_h
// every head can attend over enc_out directly: Wk is folded into the query projection and Wv into
// the output projection at load time (model.cu), and this kernel reads each chunk's enc_out ONCE per
// layer -- half the HBM bytes of the K/V form, and the cache shrinks from L*2 tensors to 1 per chunk.
//
// One persistent CTA per SM walks chunks; per chunk it streams enc_out in blocks of 128 keys:
//   warp 4  TMA     : Q' tile [16 x D] (heads padded to 16 rows by TMA zero fill) and the
//                     [128 keys x D] block as D/64 swizzled [128 x 64] atoms, 2-stage ring
//   warp 5  MMA (scores) : S[128 keys x 16 heads]   = enc_blk (A, K-major)   x Q'^T (B, K-major)    N = 16
//   warp 6  MMA (context): C[D x 16 heads]         += enc_blk^T (A, MN-major) x P^T (B, K-major)     N = 16
//                     (keys are the UMMA M dimension, so the 6 heads cost N = 16, not M = 128; the 48
//                     small MMAs per block are issue bound, hence two issuing threads in parallel)
//   warps 0-3 softmax: thread = key; per-head max across the 128 threads (CREDUX.MAX.F32 per warp + smem),
//                     online-softmax rescale of C through tcgen05.ld/st when a head's max moved,
//                     P^T written to one of two swizzled smem buffers; at the end C / l -> bf16 ctx [B][H*D]
// Scores arrive already multiplied by log2(e)/8 (folded into Wqk), so p = exp2(s - m).
//
// d_model 512 / 768 run as CTA PAIRS (cross_attn_absorbed_pair_kernel, cluster of 2): each CTA streams half of the
// channels of the chunk in the same 128-key blocks, the two swap their partial scores per block through distributed
// shared memory (st.async + mbarrier complete_tx) and each accumulates the context of its own channels; see xa_body.
// Small-N UMMAs cost ~40-50 clocks each whatever M and N are (tools/ubench_mma.cu), so what this kernel pays per key
// is the NUMBER of UMMAs: 48 per 128-key x 384-channel stage here, 72 per 64-key x 768-channel stage in the
// single-CTA form that d_model 640 still uses.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "sm100.cuh"

namespace wb {

int make_tmap_h16(CUtensorMap *map, const void *base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_elems,
                   uint64_t stride2_elems, uint32_t box_rows, int rank);  // gemm.cu

static constexpr int XA_THREADS = 224;
static constexpr int QATOM_BYTES = 16 * 64 * 2;  // [16 heads x 64 channels]
// Keys per block: 128 for D <= 384 (96 KB stages); 64 for D up to 768 -- the stage stays 96 KB, the score UMMAs run
// with M = 64 (accumulator row i in TMEM lane 32 (i / 16) + i % 16, measured with tools/ubench_tmem.cu), so only the
// first 16 lanes of every softmax warp own a key.
static constexpr int xa_keys(int atoms) { return atoms <= 6 ? 128 : 64; }

struct CrossAttnParams {
    CUtensorMap enc_map;  // dims (D, S, B), box (64, KEYS, 1)
    CUtensorMap q_map;    // dims (D, q_rows * H, B), box (64, 16, 1): rows >= q_rows * H are zero filled
    h16 *ctx;   // [B][q_rows][H*D]
    int B, S, D, H, n_blocks, atoms;
    // Prefill (whisper.mojo:195-197): q_rows query rows per chunk in q' / ctx; this launch attends for the NQ rows
    // q0 .. q0 + NQ - 1 of every chunk (their heads are the score columns q * H + h).  A decode step: q_rows = 1, q0 = 0.
    int q_rows, q0;
    // Finished chunks are skipped (whisper.mojo:206-207 `if next_token == 50257: break`, batched): when `live` is set the
    // kernel walks live[0 .. *n_live) -- the indices of the chunks still decoding -- instead of 0 .. B.
    const int *live, *n_live;
    unsigned long long *dbg;  // optional timestamp dump (CTA 0): [role][block][event]
};

__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t *v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t *v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

#define XA_STAMP(role, blk, ev)                                                              \
    do {                                                                                     \
        if (P.dbg && blockIdx.x == 0 && (blk) < 64) P.dbg[((role)*64 + (blk)) * 8 + (ev)] = ptx::globaltimer_ns(); \
    } while (0)

// 16-byte store into the peer CTA's shared memory that completes 16 transaction bytes on the peer's mbarrier: data and
// signal travel together, no cluster-scope release fence (a releasing remote arrive cost ~2 us per block, measured)
__device__ __forceinline__ void st_async_v4(uint32_t cluster_addr, uint32_t cluster_bar, float a, float b, float c, float d) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(cluster_addr),
                 "f"(a), "f"(b), "f"(c), "f"(d), "r"(cluster_bar)
                 : "memory");
}
// wait on a barrier of this CTA whose arrivals come from the peer CTA of the cluster (acquire at cluster scope, so the
// peer's shared::cluster stores issued before its releasing arrive are visible afterwards); bounded like mbar_wait
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {
    uint64_t t0 = 0;
    for (uint32_t spins = 1;; ++spins) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(ptx::smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return;
        if ((spins & 0xfffu) == 0) {
            uint64_t t = ptx::globaltimer_ns();
            if (t0 == 0) t0 = t;
            else if (t - t0 > 10000000000ull) __trap();
        }
    }
}

// ATOMS: 64-channel atoms THIS CTA streams.  CL = 1: one CTA per chunk, D = 64 ATOMS.  CL = 2 (D = 512 / 768): a
// cluster of two CTAs per chunk, each streaming its half of the channels as 128-key blocks (the per-CTA shape of
// the D = 384 kernel: 96 KB stages, M = 128 score UMMAs, half the UMMA count per key of the single-CTA 64-key
// variant, whose 72 small UMMAs per 96 KB at ~40-50 clocks each were the limit: tools/ubench_mma.cu).  The score of a
// key is the sum over all channels, so per block the two CTAs swap their partial scores [128 keys x H] through
// distributed shared memory, compute the same softmax, and each accumulates the context of its own channels.
// NQ: query rows per chunk served by one pass over the chunk's encoder output (1 = decode step; 2 = prefill with
// 2 H <= 16 score columns: the N = 16 UMMAs cost the same whatever number of their columns is in use).
template <int ATOMS, int CL, int NQ = 1>
__device__ __forceinline__ void xa_body(const CrossAttnParams &P) {
    constexpr int KEYS = xa_keys(ATOMS);
    constexpr int ATOM_BYTES = KEYS * 64 * 2;  // [KEYS x 64 channels] bf16
    constexpr int XA_MAX_ATOMS = ATOMS;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int atoms = ATOMS;
    constexpr int stage_bytes = atoms * ATOM_BYTES;
    uint8_t *sEnc = base;                                  // [2][atoms][KEYS x 64]
    uint8_t *sQ = base + 2 * stage_bytes;                  // [atoms][16 x 64]
    uint8_t *sP = sQ + XA_MAX_ATOMS * QATOM_BYTES;         // [2][KEYS / 64][16 x 64]   P^T, keys 0-63 | 64-127
    constexpr int P_BYTES = (KEYS / 64) * QATOM_BYTES;     // one P^T buffer; block g uses buffer g & 1
    uint64_t *bars = reinterpret_cast<uint64_t *>(sP + 2 * P_BYTES);
    // enc_full[stage][atom]: one barrier per 64-channel atom, so the score MMAs of an atom issue as soon as it has
    // landed instead of after the whole 96 KB stage (the S MMAs are issue bound: ~1 us per block otherwise sits
    // between "stage arrived" and "scores ready", and with only two stages that latency caps the HBM stream)
    // enc_empty[stage][atom pair]: the context MMAs walk the atom pairs in order and release each pair as soon as
    // its 8 MMAs are done, so the refill of the stage starts ~0.7 us before the block's last MMA retires
    uint64_t *enc_full = bars, *enc_empty = bars + 2 * XA_MAX_ATOMS, *q_full = enc_empty + XA_MAX_ATOMS, *q_empty = q_full + 1,
             *s_full = q_empty + 1, *s_empty = s_full + 2, *p_full = s_empty + 2, *c_done = p_full + 1,  // c_done[2]: block g commits to c_done[g & 1]
             *c_empty = c_done + 2,  // c_empty[2]: the context accumulator is double buffered over chunks
             *x_full = c_empty + 2;  // x_full[2] (CL = 2): the peer's partial scores of block parity g & 1 have arrived
    uint32_t *tmem_holder = reinterpret_cast<uint32_t *>(x_full + 2);
    float *s_red = reinterpret_cast<float *>(x_full + 3);  // [4 warps][16] cross-warp reduction scratch
    // [2][KEYS][H] partial scores written by the peer CTA (CL = 2)
    float *xchg = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(s_red + 64) + 15) & ~uintptr_t(15));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nblk = P.n_blocks, D = P.D;
    pdl_launch_dependents();
    constexpr int HEADS = ATOMS * CL;  // head_dim = 64, so the model has D / 64 heads
    constexpr int H = HEADS * NQ;      // score / context columns in use: (query row, head) pairs
    static_assert(H <= 16, "at most 16 score columns");
    static_assert(NQ == 1 || CL == 1, "several query rows per pass: single-CTA form only");
    constexpr int n_acc = ATOMS / 2;  // C accumulators of [128 channels x 16 heads]
    static_assert(CL == 1 || KEYS == 128, "the cluster form streams 128-key blocks");
    const uint32_t rank = CL > 1 ? ptx::cluster_ctarank() : 0u;
    const int b_first = CL > 1 ? (int)(blockIdx.x / CL) : (int)blockIdx.x;  // chunks walk over clusters
    const int b_step = (int)gridDim.x / CL;
    const int ch0 = (int)rank * ATOMS * 64;  // first channel of this CTA

    if (warp == 4 && lane == 0) {
        ptx::prefetch_tmap(&P.enc_map);
        ptx::prefetch_tmap(&P.q_map);
        for (int i = 0; i < 2 * XA_MAX_ATOMS; i++) ptx::mbar_init(&enc_full[i], 1);
        for (int i = 0; i < XA_MAX_ATOMS; i++) ptx::mbar_init(&enc_empty[i], 1);
        for (int i = 0; i < 2; i++) {
            ptx::mbar_init(&s_full[i], 1);
            ptx::mbar_init(&s_empty[i], 4);
        }
        ptx::mbar_init(q_full, 1);
        ptx::mbar_init(q_empty, 1);
        ptx::mbar_init(p_full, 4);
        ptx::mbar_init(&c_done[0], 1);
        ptx::mbar_init(&c_done[1], 1);
        ptx::mbar_init(&c_empty[0], 4);
        ptx::mbar_init(&c_empty[1], 4);
        ptx::mbar_init(&x_full[0], 1);
        ptx::mbar_init(&x_full[1], 1);
        ptx::fence_barrier_init();
    }
    if (warp == 5) {
        ptx::tmem_alloc(tmem_holder, 512);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    if constexpr (CL > 1) ptx::cluster_sync();  // the peer's barriers are initialised before anything arrives on them
    else __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    // TMEM columns: score partials S[stage][part] @ 16 * (stage * n_acc + part), then the context accumulators C_m of
    // chunk parity p @ 32 n_acc + 16 (p n_acc + m) (two context buffers: the epilogue of chunk i runs under the first
    // block of chunk i+1); 64 n_acc columns in all (192 for D = 384, 384 for D = 768).
    // Back-to-back MMAs into one accumulator serialise (~45 ns each), so the K = D reduction of the scores is
    // split into one partial accumulator per atom pair and the issue order interleaves accumulators.
    const uint32_t tS0 = tmem_base, tC0 = tmem_base + 32 * n_acc;
    constexpr int C_BUF = 16 * n_acc;  // columns of one context buffer
    pdl_wait();  // the prologue above overlapped the previous kernel (programmatic dependent launch)
    const int n_work = P.live ? *P.n_live : P.B;  // chunks this launch attends for (rebuilt by an earlier kernel)

    if (warp == 4) {
        // ===== TMA producer =====
        if (lane == 0) {
            int g = 0, ci = 0;  // g: global block counter of this CTA, ci: chunk counter
            for (int wi = b_first; wi < n_work; wi += b_step, ci++) {
                const int b = P.live ? P.live[wi] : wi;
                ptx::mbar_wait(q_empty, (ci & 1) ^ 1);  // score MMAs of the previous chunk are done with sQ
                ptx::mbar_expect_tx(q_full, atoms * QATOM_BYTES);
                for (int a = 0; a < atoms; a++) ptx::tma_load_3d(sQ + a * QATOM_BYTES, &P.q_map, q_full, ch0 + a * 64, P.q0 * HEADS, b);
                for (int j = 0; j < nblk; j++, g++) {
                    const int s = g & 1;
                    for (int a = 0; a < atoms; a++) {
                        if ((a & 1) == 0) ptx::mbar_wait(&enc_empty[s * (XA_MAX_ATOMS / 2) + (a >> 1)], ((g >> 1) & 1) ^ 1);
                        if (a == 0) XA_STAMP(0, g, 0);
                        uint64_t *bar = &enc_full[s * XA_MAX_ATOMS + a];
                        ptx::mbar_expect_tx(bar, ATOM_BYTES);
                        ptx::tma_load_3d(sEnc + s * stage_bytes + a * ATOM_BYTES, &P.enc_map, bar, ch0 + a * 64, j * KEYS, b);
                    }
                }
            }
        }
    } else if (warp == 5) {
        // ===== MMA issuer 1: scores =====
        if (lane == 0) {
            constexpr uint32_t idesc_s = ptx::umma_idesc_h16(KEYS, 16, 0, 0);  // enc (K-major) x Q' (K-major)
            uint64_t a_desc0[2], b_desc0 = ptx::umma_desc_sw128(ptx::smem_u32(sQ), 1, 64);
            for (int s = 0; s < 2; s++) a_desc0[s] = ptx::umma_desc_sw128(ptx::smem_u32(sEnc + s * stage_bytes), 1, 64);
            int g = 0, ci = 0;
            for (int wi = b_first; wi < n_work; wi += b_step, ci++) {
                const int b = P.live ? P.live[wi] : wi;
                ptx::mbar_wait(q_full, ci & 1);
                for (int j = 0; j < nblk; j++, g++) {
                    const int s = g & 1;
                    ptx::mbar_wait(&s_empty[s], ((g >> 1) & 1) ^ 1);
                    XA_STAMP(1, g, 1);
#pragma unroll
                    for (int a = 0; a < atoms; a++) {  // atoms in arrival order; atom pair `part` shares an accumulator
                        ptx::mbar_wait(&enc_full[s * XA_MAX_ATOMS + a], (g >> 1) & 1);
                        if (a == 0) XA_STAMP(1, g, 0);
                        ptx::tc_fence_after();
                        const int part = a >> 1;
#pragma unroll
                        for (int k = 0; k < 4; k++)  // 4 K-steps of 16 channels; descriptors advance in 16-byte units
                            ptx::mma_h16_ss(tS0 + 16 * (s * n_acc + part), a_desc0[s] + a * (ATOM_BYTES >> 4) + 2 * k,
                                             b_desc0 + a * (QATOM_BYTES >> 4) + 2 * k, idesc_s, ((a & 1) | k) != 0);
                    }
                    ptx::mma_commit(&s_full[s]);
                    if (j + 1 == nblk) ptx::mma_commit(q_empty);
                }
            }
        }
    } else if (warp == 6) {
        // ===== MMA issuer 2: context =====
        if (lane == 0) {
            constexpr uint32_t idesc_c = ptx::umma_idesc_h16(128, 16, 1, 0);  // enc^T (MN-major) x P^T (K-major)
            uint64_t a_desc0[2];
            // A = enc_blk^T: channels [128 m, 128 m + 128) = atoms 2m, 2m+1 (LBO = one atom),
            // K = keys: 8-row groups 1024 B apart, 16 keys per MMA = 2048 B
            for (int s = 0; s < 2; s++)
                a_desc0[s] = ptx::umma_desc_sw128(ptx::smem_u32(sEnc + s * stage_bytes), ATOM_BYTES >> 4, 64);
            const uint64_t p_desc0 = ptx::umma_desc_sw128(ptx::smem_u32(sP), 1, 64);
            int g = 0, ci = 0;
            for (int wi = b_first; wi < n_work; wi += b_step, ci++) {
                const int b = P.live ? P.live[wi] : wi;
                ptx::mbar_wait(&c_empty[ci & 1], ((ci >> 1) & 1) ^ 1);  // epilogue of chunk ci - 2 has drained this buffer
                const uint32_t tC = tC0 + (ci & 1) * C_BUF;
                for (int j = 0; j < nblk; j++, g++) {
                    const int s = g & 1;
                    ptx::mbar_wait(p_full, g & 1);  // softmax(g) done => scores(g) done reading the stage too
                    XA_STAMP(1, g, 2);
                    ptx::tc_fence_after();
#pragma unroll
                    for (int m = 0; m < n_acc; m++) {  // atom pair m = channels [128 m, 128 m + 128)
#pragma unroll
                        for (int k = 0; k < KEYS / 16; k++)
                            ptx::mma_h16_ss(tC + 16 * m, a_desc0[s] + (2 * m) * (ATOM_BYTES >> 4) + k * (2048 >> 4),
                                             p_desc0 + (g & 1) * (P_BYTES >> 4) + (k >> 2) * (QATOM_BYTES >> 4) + 2 * (k & 3),
                                             idesc_c, (j | k) != 0);
                        ptx::mma_commit(&enc_empty[s * (XA_MAX_ATOMS / 2) + m]);
                    }
                    ptx::mma_commit(&c_done[g & 1]);
                    XA_STAMP(1, g, 3);
                }
            }
        }
    } else {
        // ===== softmax / correction / epilogue warps =====
        const int row = warp * 32 + lane;  // TMEM lane: channel lane for C, and (KEYS = 128) the key within the block
        const bool owns_key = KEYS == 128 || lane < 16;            // M = 64 accumulators: 16 rows per warp
        const int key = KEYS == 128 ? row : warp * 16 + lane;      // key within the block of this thread
        const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
        for (int i = row; i < 2 * P_BYTES / 16; i += 128) reinterpret_cast<uint4 *>(sP)[i] = make_uint4(0, 0, 0, 0);
        ptx::fence_proxy_async_smem();
        named_bar_sync(2, 128);
        // chunk epilogue: l[h] = sum over the 128 threads; ctx = C / l.  Deferred: it runs after the softmax of the
        // NEXT chunk's first block (its context MMAs are long done by then), so the softmax warps never wait on c_done
        auto epilogue = [&](int b_out, float *l_sum, uint32_t tC) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int h = 0; h < H; h++) l_sum[h] += __shfl_xor_sync(0xffffffffu, l_sum[h], o);
            }
            if (lane < H) {
                float v = l_sum[0];
#pragma unroll
                for (int h = 1; h < H; h++)
                    if (lane == h) v = l_sum[h];
                s_red[warp * 16 + lane] = v;
            }
            named_bar_sync(1, 128);
            float inv[H];
#pragma unroll
            for (int h = 0; h < H; h++) inv[h] = 1.0f / (s_red[h] + s_red[16 + h] + s_red[32 + h] + s_red[48 + h]);
            ptx::tc_fence_after();
            // column h = (query row q0 + h / HEADS, head h % HEADS): consecutive [D] slices of ctx [B][q_rows][HEADS * D]
            h16 *dst = P.ctx + ((size_t)b_out * P.q_rows + P.q0) * HEADS * D;
            for (int m = 0; m < n_acc; m++) {
                uint32_t cv[16];
                tmem_ld_32x32b_x16(tC + 16 * m + lane_addr, cv);
                ptx::tmem_ld_wait();
                const int c = ch0 + m * 128 + row;
#pragma unroll
                for (int h = 0; h < H; h++) dst[(size_t)h * D + c] = f2h(__uint_as_float(cv[h]) * inv[h]);
            }
            ptx::tc_fence_before();
            named_bar_sync(2, 128);  // s_red reads are done; C buffer drained by all four warps
        };
        int g = 0, ci = 0, b_prev = -1;
        float l_prev[H];
        for (int wi = b_first; wi < n_work; wi += b_step, ci++) {
                const int b = P.live ? P.live[wi] : wi;
            float m_run[H], l_part[H];
            const uint32_t tC = tC0 + (ci & 1) * C_BUF;
#pragma unroll
            for (int h = 0; h < H; h++) m_run[h] = -INFINITY, l_part[h] = 0.f;
            for (int j = 0; j < nblk; j++, g++) {
                const int s = g & 1;
                const bool valid = owns_key && j * KEYS + key < P.S;
                ptx::mbar_wait(&s_full[s], (g >> 1) & 1);
                if (threadIdx.x == 0) XA_STAMP(2, g, 0);
                ptx::tc_fence_after();
                uint32_t sv[16], svp[n_acc > 1 ? n_acc - 1 : 1][16];  // the n_acc partial score accumulators
                tmem_ld_32x32b_x16(tS0 + 16 * (s * n_acc) + lane_addr, sv);
#pragma unroll
                for (int pa = 1; pa < n_acc; pa++) tmem_ld_32x32b_x16(tS0 + 16 * (s * n_acc + pa) + lane_addr, svp[pa - 1]);
                ptx::tmem_ld_wait();
                if (threadIdx.x == 0) XA_STAMP(2, g, 4);
#pragma unroll
                for (int h = 0; h < H; h++) {
                    float v = __uint_as_float(sv[h]);
#pragma unroll
                    for (int pa = 1; pa < n_acc; pa++) v += __uint_as_float(svp[pa - 1][h]);
                    sv[h] = __float_as_uint(v);
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&s_empty[s]);
                if constexpr (CL > 1) {
                    // swap partial scores with the peer: my [key][H] row goes into the peer's buffer of parity g & 1 as
                    // st.async stores that complete transaction bytes on the peer's barrier; wait for the peer's rows.
                    // (buffer g & 1 is rewritten at block g + 2, which the peer reaches only after my arrive of
                    // block g + 1, i.e. after I have read block g's row)
                    float *mine = xchg + ((g & 1) * KEYS + key) * H;
                    const uint32_t dst = ptx::mapa_u32(mine, rank ^ 1u), dst_bar = ptx::mapa_u32(&x_full[g & 1], rank ^ 1u);
                    if (threadIdx.x == 0) ptx::mbar_expect_tx(&x_full[g & 1], KEYS * H * 4);  // what the peer will send me
#pragma unroll
                    for (int h = 0; h < H; h += 4)
                        st_async_v4(dst + h * 4, dst_bar, __uint_as_float(sv[h]), __uint_as_float(sv[h + 1]),
                                    __uint_as_float(sv[h + 2]), __uint_as_float(sv[h + 3]));
                    if (threadIdx.x == 0) XA_STAMP(1, g, 4);
                    if (threadIdx.x == 0) XA_STAMP(1, g, 5);
                    mbar_wait_cluster(&x_full[g & 1], (g >> 1) & 1);
                    if (threadIdx.x == 0) XA_STAMP(1, g, 6);
#pragma unroll
                    for (int h = 0; h < H; h += 4) {
                        const float4 o = *reinterpret_cast<const float4 *>(mine + h);
                        sv[h] = __float_as_uint(__uint_as_float(sv[h]) + o.x);
                        sv[h + 1] = __float_as_uint(__uint_as_float(sv[h + 1]) + o.y);
                        sv[h + 2] = __float_as_uint(__uint_as_float(sv[h + 2]) + o.z);
                        sv[h + 3] = __float_as_uint(__uint_as_float(sv[h + 3]) + o.w);
                    }
                }
                // per-head maximum over the 128 keys of the block: butterfly steps outermost so the H independent
                // shuffles of a step pipeline (head-outermost compiles to 5 x H serially dependent SHFLs)
                float sc[H], mx[H];
#pragma unroll
                for (int h = 0; h < H; h++) mx[h] = sc[h] = valid ? __uint_as_float(sv[h]) : -INFINITY;
                // warp maximum per head: one CREDUX.MAX.F32 each (sm_100a) instead of a 5-step shuffle butterfly
#pragma unroll
                for (int h = 0; h < H; h++) asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(mx[h]) : "f"(mx[h]));
                if (lane < H) {
                    float v = mx[0];
#pragma unroll
                    for (int h = 1; h < H; h++)
                        if (lane == h) v = mx[h];
                    s_red[warp * 16 + lane] = v;
                }
                if (threadIdx.x == 0) XA_STAMP(2, g, 5);
                named_bar_sync(1, 128);
                if (threadIdx.x == 0) XA_STAMP(2, g, 6);
                float alpha[H];
                bool moved = false;
#pragma unroll
                for (int h = 0; h < H; h++) {
                    float mb = fmaxf(fmaxf(s_red[h], s_red[16 + h]), fmaxf(s_red[32 + h], s_red[48 + h]));
                    float mn = fmaxf(m_run[h], mb);
                    alpha[h] = exp2f(m_run[h] - mn);  // 0 on the first block
                    moved |= (alpha[h] != 1.0f);
                    m_run[h] = mn;
                }
                // P^T[h][key] = exp2(s - m) in bf16, 128B-swizzled rows of 64 keys
                if (threadIdx.x == 0) XA_STAMP(2, g, 1);
                // P^T is double buffered: block g - 2's context MMAs must be done with this buffer (they long are)
                if (g > 1) ptx::mbar_wait(&c_done[g & 1], ((g - 2) >> 1) & 1);
                if (threadIdx.x == 0) XA_STAMP(2, g, 2);
                {
                    uint8_t *atom = sP + (g & 1) * P_BYTES + (key >> 6) * QATOM_BYTES;
                    const int kk = key & 63;
#pragma unroll
                    for (int h = 0; h < H; h++) {  // rows of the padded heads (h >= H) were zeroed once at kernel start
                        float p = valid ? exp2f(sc[h] - m_run[h]) : 0.f;
                        l_part[h] = l_part[h] * alpha[h] + p;
                        if (owns_key)
                            *reinterpret_cast<h16 *>(atom + h * 128 + (((kk >> 3) ^ (h & 7)) << 4) + (kk & 7) * 2) =
                                f2h(p);
                    }
                }
                // rescale the running context when a head's maximum moved (uniform across the CTA)
                if (j > 0 && moved) {
                    ptx::mbar_wait(&c_done[(g - 1) & 1], ((g - 1) >> 1) & 1);  // C holds every block < g
                    ptx::tc_fence_after();
                    for (int m = 0; m < n_acc; m++) {
                        uint32_t cv[16];
                        tmem_ld_32x32b_x16(tC + 16 * m + lane_addr, cv);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int h = 0; h < H; h++) cv[h] = __float_as_uint(__uint_as_float(cv[h]) * alpha[h]);
                        tmem_st_32x32b_x16(tC + 16 * m + lane_addr, cv);
                    }
                    ptx::tmem_st_wait();
                }
                if (threadIdx.x == 0) XA_STAMP(2, g, 7);
                ptx::fence_proxy_async_smem();
                ptx::tc_fence_before();
                named_bar_sync(2, 128);  // all four warps are past their s_red reads before the next block writes it
                if (lane == 0) ptx::mbar_arrive(p_full);
                if (threadIdx.x == 0) XA_STAMP(2, g, 3);
                if (j == 0 && b_prev >= 0) {  // previous chunk: its last context MMAs (block g - 1) are long done by now
                    ptx::mbar_wait(&c_done[(g - 1) & 1], ((g - 1) >> 1) & 1);
                    epilogue(b_prev, l_prev, tC0 + ((ci - 1) & 1) * C_BUF);
                    if (lane == 0) ptx::mbar_arrive(&c_empty[(ci - 1) & 1]);
                    b_prev = -1;
                }
            }
#pragma unroll
            for (int h = 0; h < H; h++) l_prev[h] = l_part[h];
            b_prev = b;
        }
        if (b_prev >= 0) {  // last chunk of this CTA
            ptx::mbar_wait(&c_done[(g - 1) & 1], ((g - 1) >> 1) & 1);
            epilogue(b_prev, l_prev, tC0 + ((ci - 1) & 1) * C_BUF);
            if (lane == 0) ptx::mbar_arrive(&c_empty[(ci - 1) & 1]);
        }
    }
    ptx::tc_fence_before();
    if constexpr (CL > 1) ptx::cluster_sync();  // the peer may still write this CTA's exchange buffer / barriers
    else __syncthreads();
    if (warp == 5) ptx::tmem_dealloc(tmem_base, 512);
}

template <int ATOMS, int NQ = 1>
__global__ void __launch_bounds__(XA_THREADS, 1) cross_attn_absorbed_kernel(const __grid_constant__ CrossAttnParams P) {
    xa_body<ATOMS, 1, NQ>(P);
}
template <int ATOMS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(XA_THREADS, 1)
    cross_attn_absorbed_pair_kernel(const __grid_constant__ CrossAttnParams P) {
    xa_body<ATOMS, 2>(P);
}

// D = 512 / 768 run as CTA pairs (two CTAs per chunk, half the channels each) unless WB_XA_NO_PAIR is set
static bool xa_use_pair(int atoms) {
    static const bool off = getenv("WB_XA_NO_PAIR") != nullptr;
    return !off && (atoms == 8 || atoms == 12);
}

size_t cross_attn_absorbed_smem(int D) {
    const bool pair = xa_use_pair(D / 64);
    const int atoms = pair ? D / 128 : D / 64;  // atoms per CTA
    const int keys = xa_keys(atoms);
    return 1024 + (size_t)2 * atoms * keys * 64 * 2 + atoms * QATOM_BYTES + 2 * (keys / 64) * QATOM_BYTES +
           (3 * atoms + 16) * 8 + 64 * 4 + 64 + (pair ? 2 * 128 * (D / 64) * 4 + 16 : 0);
}

bool cross_attn_absorbed_supported(int D, int H) { return D % 128 == 0 && D <= 768 && H <= 16; }

// q' bf16 [B][H*D] (already scaled by log2(e)/8 through the folded weights), enc bf16 [B][S][D]
// -> ctx bf16 [B][H*D] with ctx[b][h] = softmax_j(q'_h . enc_j) weighted sum of enc_j.
unsigned long long *g_xa_dbg = nullptr;  // set by the debug hook to collect timestamps

int cross_attention_absorbed(cudaStream_t st, const h16 *qp, const h16 *enc, h16 *ctx,
                             int B, int S, int D, int H, const int *live, const int *n_live, int q_rows) {
    if (B <= 0) return WB_OK;
    WB_ARG(cross_attn_absorbed_supported(D, H) && H * 64 == D,
           "absorbed cross-attention needs head_dim 64, D %% 128 == 0, D <= 768 (D=%d H=%d)", D, H);
    WB_ARG(q_rows >= 1 && q_rows <= 4, "absorbed cross-attention: q_rows=%d", q_rows);
    CrossAttnParams P;
    const bool pair = xa_use_pair(D / 64);
    const int keys = xa_keys(pair ? D / 128 : D / 64);
    WB_CHECK(make_tmap_h16(&P.enc_map, enc, (uint64_t)D, (uint64_t)S, (uint64_t)B, (uint64_t)D, (uint64_t)S * D, keys, 3));
    WB_CHECK(make_tmap_h16(&P.q_map, qp, (uint64_t)D, (uint64_t)q_rows * H, (uint64_t)B, (uint64_t)D, (uint64_t)q_rows * H * D, 16, 3));
    P.ctx = ctx, P.B = B, P.S = S, P.D = D, P.H = H, P.n_blocks = cdiv(S, keys), P.atoms = D / 64;
    P.q_rows = q_rows, P.q0 = 0;
    P.dbg = g_xa_dbg;
    P.live = (live && n_live) ? live : nullptr, P.n_live = n_live;
    const size_t smem = cross_attn_absorbed_smem(D);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = pair ? 2 * std::min(B, sms / 2) : (B < sms ? B : sms);
    auto launch = [&](auto kernel, int) {
        WB_CUDA(ensure_dyn_smem(kernel, smem));
        WB_CUDA(launch_pdl(kernel, dim3(grid), dim3(XA_THREADS), smem, st, P));
        WB_LAUNCHED();
        return WB_OK;
    };
    auto one_pass = [&]() -> int {
        if (pair) return P.atoms == 8 ? launch(cross_attn_absorbed_pair_kernel<4>, 6) : launch(cross_attn_absorbed_pair_kernel<6>, 7);
        switch (P.atoms) {
            case 2: return launch(cross_attn_absorbed_kernel<2>, 0);
            case 4: return launch(cross_attn_absorbed_kernel<4>, 1);
            case 6: return launch(cross_attn_absorbed_kernel<6>, 2);
            case 8: return launch(cross_attn_absorbed_kernel<8>, 3);
            case 10: return launch(cross_attn_absorbed_kernel<10>, 4);
            default: return launch(cross_attn_absorbed_kernel<12>, 5);
        }
    };
    if (q_rows == 1) return one_pass();
    // Prefill: the q_rows query rows of a chunk attend over the same encoder output.  With 2 H <= 16 two rows share
    // one pass (their heads are 2 H of the 16 score columns the N = 16 UMMAs compute anyway), so the 4-id prompt
    // costs two reads of enc_out instead of four; otherwise one pass per row.
    const bool two = !pair && 2 * H <= 16 && q_rows % 2 == 0;
    for (int q0 = 0; q0 < q_rows; q0 += two ? 2 : 1) {
        P.q0 = q0;
        if (!two) {
            WB_CHECK(one_pass());
            continue;
        }
        switch (P.atoms) {
            case 2: WB_CHECK(launch(cross_attn_absorbed_kernel<2, 2>, 8)); break;
            case 4: WB_CHECK(launch(cross_attn_absorbed_kernel<4, 2>, 9)); break;
            case 6: WB_CHECK(launch(cross_attn_absorbed_kernel<6, 2>, 10)); break;
            default: WB_CHECK(launch(cross_attn_absorbed_kernel<8, 2>, 11)); break;
        }
    }
    return WB_OK;
}

}  // namespace wb
