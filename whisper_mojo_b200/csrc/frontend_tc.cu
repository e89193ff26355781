// frontend_tc.cu -- log-mel frontend with the STFT as a tensor-core GEMM (sm_100a, tcgen05 kind::tf32).
//
// Same arithmetic contract as frontend.cu (HF WhisperFeatureExtractor, export_weights.py:116), but the
// 400-point windowed DFT of every frame is one row of
//      X[frame, col] = sum_n y[160*frame + n] * T[col, n]      T = Hann window folded into cos / sin
// computed on the tensor cores.  TF32 alone (10-bit mantissa) is ~100x too coarse for the 1e-4 log-mel
// tolerance, so both operands are split y = y_hi + y_lo, T = T_hi + T_lo (each part exactly
// representable in TF32) and three products hi*hi + lo*hi + hi*lo accumulate into the same fp32 TMEM
// accumulator (error ~2^-21 per term; measured 1.7e-5 max-abs on the log-mel, fp32 FMA kernel: 9e-6).
//
//   pre-pass      reflect-pads every chunk by 200 samples and writes y_hi / y_lo  (padded_split_kernel)
//   A operand     the frames are OVERLAPPING rows of the padded signal: a 3-D tensor map with a row
//                 stride of 160 samples and 416-sample rows (400 + zero-weighted tail) -- no im2col copy
//   B operand     T_hi / T_lo [416 cols][416 n]: rows 0..200 = w*cos (bins 0..200), rows 208..407 = w*sin
//                 (bins 0..199), everything else zero; 1.4 MB, L2 resident
//   kernel        CTA pair (cta_group::2): 256 frames x 416 columns per pair tile as two UMMAs of
//                 N = 208 (re | im halves), K = 8 per instruction, 13 k-blocks x 4 k-steps x 3 products;
//                 each CTA stages its 128 frames and half of the twiddle rows (two 84 KB stages)
//   epilogue      thread = frame: power = re^2 + im^2 from TMEM, sparse slaney mel filterbank as a
//                 running pair of accumulators (every bin feeds at most two adjacent filters), log10,
//                 coalesced stores of the [mel][frame] tile, per-chunk maximum by ordered-int atomicMax
#include <cuda.h>
#include <math.h>

#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "sm100.cuh"

namespace wb {

int make_tmap_f32(CUtensorMap *map, const void *base, int rank, const uint64_t *dims, const uint64_t *strides_bytes,
                  const uint32_t *box);  // gemm.cu

static constexpr int FT_KB = 13;                       // k-blocks of 32 samples (416 = 400 + 16 zero-weighted)
static constexpr int FT_NH = 208, FT_BOX = 104;        // columns per half (re | im), twiddle rows per CTA and half
static constexpr int FT_IM_COL = 256;                  // TMEM column of the im accumulator (re at 0)
static constexpr int FT_A_BYTES = 128 * 32 * 4;        // 128 frames x 32 samples fp32
static constexpr int FT_B_BYTES = FT_BOX * 32 * 4;     // 104 twiddle rows x 32 samples
static constexpr int FT_STAGE_BYTES = 2 * FT_A_BYTES + 4 * FT_B_BYTES;  // A_hi A_lo | B_hi re,im | B_lo re,im
static constexpr int FT_STAGES = 2;
static constexpr int FT_THREADS = 320;
static constexpr int FT_MAX_MELS = 80;
static constexpr int FT_SMEM = 1024 + FT_STAGES * FT_STAGE_BYTES + FT_MAX_MELS * 128 * 4 + 3 * FT_NH * 4 + 256;

struct FrontendTcParams {
    CUtensorMap a_hi, a_lo;  // dims {416, n_frames, B}, strides {160 samples, padded chunk}
    CUtensorMap b_hi, b_lo;  // dims {416 n, 416 cols}, box {32, 104}
    const int *bin_j;        // [208] first mel filter bin k feeds (non-decreasing)
    const float *bin_wa, *bin_wb;  // [208] weights into filters bin_j[k] and bin_j[k] + 1
    float *mel_raw;
    int *chunk_max;
    int n_frames, n_mels, tiles_per_chunk, total_tiles;
};

__device__ __forceinline__ int ft_float_to_ordered(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}

// y = reflect_pad(x, 200) (torch.stft center=True), split into TF32-exact hi / lo parts.
__global__ void padded_split_kernel(const float *__restrict__ pcm, int n_samples, int pad_len, float *__restrict__ y_hi,
                                    float *__restrict__ y_lo) {
    const int b = blockIdx.y;
    const float *x = pcm + (size_t)b * n_samples;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < pad_len; i += gridDim.x * blockDim.x) {
        int s = i - 200;
        if (s < 0) s = -s;
        if (s >= n_samples) s = 2 * (n_samples - 1) - s;
        const float v = (s >= 0 && s < n_samples) ? __ldg(x + s) : 0.f;
        const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        const float lo = __uint_as_float(__float_as_uint(v - hi) & 0xFFFFE000u);
        y_hi[(size_t)b * pad_len + i] = hi;
        y_lo[(size_t)b * pad_len + i] = lo;
    }
}

__device__ __forceinline__ void mma_tf32_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                 uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// kind::tf32 instruction descriptor: fp32 accumulate, TF32 x TF32, both operands K-major.
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FT_THREADS, 1)
    logmel_tc_kernel(const __grid_constant__ FrontendTcParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *tiles = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float *s_mel = reinterpret_cast<float *>(tiles + FT_STAGES * FT_STAGE_BYTES);  // [n_mels][128 frames]
    int *s_j = reinterpret_cast<int *>(s_mel + FT_MAX_MELS * 128);
    float *s_wa = reinterpret_cast<float *>(s_j + FT_NH), *s_wb = s_wa + FT_NH;
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_wb + FT_NH);
    uint64_t *full_bar = bars, *empty_bar = bars + FT_STAGES, *tmem_full = bars + 2 * FT_STAGES,
             *tmem_empty = bars + 2 * FT_STAGES + 1;
    uint32_t *tmem_holder = reinterpret_cast<uint32_t *>(bars + 2 * FT_STAGES + 2);
    float *s_red = reinterpret_cast<float *>(bars + 2 * FT_STAGES + 3);  // [8]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

    for (int i = threadIdx.x; i < P.n_mels * 128; i += FT_THREADS) s_mel[i] = 0.f;
    for (int i = threadIdx.x; i < FT_NH; i += FT_THREADS) s_j[i] = P.bin_j[i], s_wa[i] = P.bin_wa[i], s_wb[i] = P.bin_wb[i];
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&P.a_hi), ptx::prefetch_tmap(&P.a_lo), ptx::prefetch_tmap(&P.b_hi), ptx::prefetch_tmap(&P.b_lo);
        for (int s = 0; s < FT_STAGES; s++) ptx::mbar_init(&full_bar[s], 1), ptx::mbar_init(&empty_bar[s], 1);
        ptx::mbar_init(tmem_full, 1);
        ptx::mbar_init(tmem_empty, 16);  // leader only: 8 epilogue warps of each CTA
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc_pair(tmem_holder, 512);
        ptx::tmem_relinquish_pair();
    }
    __syncthreads();
    ptx::tc_fence_before();
    ptx::cluster_sync();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        // ===== TMA producer (one per CTA): its 128 frames, its half of every twiddle block =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = pair_id; tile < P.total_tiles; tile += n_pairs) {
                const int b = tile / P.tiles_per_chunk;
                const int f0 = (tile % P.tiles_per_chunk) * 256 + (int)rank * 128;
                for (int kb = 0; kb < FT_KB; kb++) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (rank == 0) ptx::mbar_expect_tx(&full_bar[stage], 2 * FT_STAGE_BYTES);
                    const uint32_t bar = ptx::mapa_u32(&full_bar[stage], 0);
                    uint8_t *sa = tiles + stage * FT_STAGE_BYTES, *sb = sa + 2 * FT_A_BYTES;
                    ptx::tma_load_3d_pair(sa, &P.a_hi, bar, kb * 32, f0, b);
                    ptx::tma_load_3d_pair(sa + FT_A_BYTES, &P.a_lo, bar, kb * 32, f0, b);
                    ptx::tma_load_2d_pair(sb, &P.b_hi, bar, kb * 32, (int)rank * FT_BOX);
                    ptx::tma_load_2d_pair(sb + FT_B_BYTES, &P.b_hi, bar, kb * 32, FT_NH + (int)rank * FT_BOX);
                    ptx::tma_load_2d_pair(sb + 2 * FT_B_BYTES, &P.b_lo, bar, kb * 32, (int)rank * FT_BOX);
                    ptx::tma_load_2d_pair(sb + 3 * FT_B_BYTES, &P.b_lo, bar, kb * 32, FT_NH + (int)rank * FT_BOX);
                    if (++stage == FT_STAGES) stage = 0, phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA): re -> TMEM cols [0,208), im -> [256,464) =====
        if (rank == 0 && lane == 0) {
            constexpr uint32_t idesc = umma_idesc_tf32(256, FT_NH);
            const uint32_t d_re = tmem_base, d_im = tmem_base + FT_IM_COL;
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = pair_id; tile < P.total_tiles; tile += n_pairs, it++) {
                ptx::mbar_wait(tmem_empty, (it & 1) ^ 1);
                ptx::tc_fence_after();
                for (int kb = 0; kb < FT_KB; kb++) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = ptx::smem_u32(tiles + stage * FT_STAGE_BYTES), sb = sa + 2 * FT_A_BYTES;
                    const uint64_t a_hi = ptx::umma_desc_sw128(sa, 1, 64), a_lo = ptx::umma_desc_sw128(sa + FT_A_BYTES, 1, 64);
                    const uint64_t bh_re = ptx::umma_desc_sw128(sb, 1, 64), bh_im = ptx::umma_desc_sw128(sb + FT_B_BYTES, 1, 64);
                    const uint64_t bl_re = ptx::umma_desc_sw128(sb + 2 * FT_B_BYTES, 1, 64),
                                   bl_im = ptx::umma_desc_sw128(sb + 3 * FT_B_BYTES, 1, 64);
#pragma unroll
                    for (int k = 0; k < 4; k++) {  // 8 samples (32 bytes = 2 x 16 B units) per UMMA
                        const uint32_t acc = (kb | k) != 0;
                        mma_tf32_ss_pair(d_re, a_hi + 2 * k, bh_re + 2 * k, idesc, acc);
                        mma_tf32_ss_pair(d_im, a_hi + 2 * k, bh_im + 2 * k, idesc, acc);
                        mma_tf32_ss_pair(d_re, a_lo + 2 * k, bh_re + 2 * k, idesc, 1);
                        mma_tf32_ss_pair(d_im, a_lo + 2 * k, bh_im + 2 * k, idesc, 1);
                        mma_tf32_ss_pair(d_re, a_hi + 2 * k, bl_re + 2 * k, idesc, 1);
                        mma_tf32_ss_pair(d_im, a_hi + 2 * k, bl_im + 2 * k, idesc, 1);
                    }
                    ptx::mma_commit_pair(&empty_bar[stage], 3);
                    if (++stage == FT_STAGES) stage = 0, phase ^= 1;
                }
                ptx::mma_commit_pair(tmem_full, 3);
            }
        }
    } else {
        // ===== epilogue warps: thread = frame (TMEM lane 32*(warp%4) + lane); the two warps of a lane
        // quarter split the bins: half 0 -> bins 0..95, half 1 -> bins 96..200 =====
        const int q = warp & 3, half = (warp - 2) >> 2, row = q * 32 + lane, tid = threadIdx.x - 64;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        const int c_begin = half ? 3 : 0, c_end = half ? 7 : 3;
        int it = 0;
        for (int tile = pair_id; tile < P.total_tiles; tile += n_pairs, it++) {
            const int b = tile / P.tiles_per_chunk;
            const int f_tile = (tile % P.tiles_per_chunk) * 256 + (int)rank * 128;
            ptx::mbar_wait(tmem_full, it & 1);
            ptx::tc_fence_after();
            int cur_j = s_j[c_begin * 32];
            float acc_a = 0.f, acc_b = 0.f;
            for (int c = c_begin; c < c_end; c++) {
                uint32_t re[32], im[32];
                ptx::tmem_ld_32x32b_x32(lane_addr + c * 32, re);
                ptx::tmem_ld_32x32b_x32(lane_addr + FT_IM_COL + c * 32, im);
                ptx::tmem_ld_wait();
                if (c == c_end - 1) {  // this warp's slice of the accumulator is in registers
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive_cluster_relaxed(ptx::mapa_u32(tmem_empty, 0));
                }
#pragma unroll
                for (int i = 0; i < 32; i++) {
                    const int k = c * 32 + i;
                    if (k >= FT_NH) break;  // warp-uniform; bins 201..207 carry zero weights
                    const float r = __uint_as_float(re[i]), s = __uint_as_float(im[i]);
                    const float pw = r * r + s * s;
                    const int jk = s_j[k];  // warp-uniform
                    while (cur_j < jk) {    // filter cur_j is complete: no later bin feeds it
                        atomicAdd(&s_mel[cur_j * 128 + row], acc_a);
                        acc_a = acc_b, acc_b = 0.f, cur_j++;
                    }
                    acc_a = fmaf(s_wa[k], pw, acc_a);
                    acc_b = fmaf(s_wb[k], pw, acc_b);
                }
            }
            atomicAdd(&s_mel[cur_j * 128 + row], acc_a);
            if (cur_j + 1 < P.n_mels) atomicAdd(&s_mel[(cur_j + 1) * 128 + row], acc_b);
            epi_bar_sync();
            // log10 + coalesced store of the [mel][128 frames] tile; the tile is zeroed for the next pass
            float lmax = -INFINITY;
            for (int i = tid; i < P.n_mels * 128; i += 256) {
                const int m = i >> 7, fg = f_tile + (i & 127);
                const float acc = s_mel[i];
                s_mel[i] = 0.f;
                if (fg < P.n_frames) {
                    const float v = log10f(fmaxf(acc, 1e-10f));
                    P.mel_raw[((size_t)b * P.n_mels + m) * P.n_frames + fg] = v;
                    lmax = fmaxf(lmax, v);
                }
            }
            lmax = warp_max(lmax);
            if (lane == 0) s_red[warp - 2] = lmax;
            epi_bar_sync();
            if (tid == 0) {
                float mx = s_red[0];
                for (int i = 1; i < 8; i++) mx = fmaxf(mx, s_red[i]);
                if (mx > -INFINITY) atomicMax(&P.chunk_max[b], ft_float_to_ordered(mx));
            }
        }
    }
    ptx::tc_fence_before();
    ptx::cluster_sync();
    if (warp == 1) ptx::tmem_dealloc_pair(tmem_base, 512);
}

// ---- host side ---------------------------------------------------------------------------------

static float tf32_rn(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    u = (u + 0x1000u) & 0xFFFFE000u;
    memcpy(&f, &u, 4);
    return f;
}

// Twiddle tables and the per-bin view of the mel filterbank.  `rows` = the dense [n_mels][201] filterbank.
int frontend_tc_tables_create(FrontendTables *t, const std::vector<std::vector<float>> &rows) {
    t->tc_ok = false;
    const int n_mels = (int)rows.size();
    if (n_mels > FT_MAX_MELS) return WB_OK;  // the fp32 FMA kernel serves other filterbank sizes
    // every bin must feed at most two ADJACENT filters (true for triangular banks with one centre per filter)
    std::vector<int> bj(FT_NH, 0);
    std::vector<float> wa(FT_NH, 0.f), wb(FT_NH, 0.f);
    int last_j = 0;
    for (int k = 0; k <= 200; k++) {
        int first = -1, cnt = 0, last = -1;
        for (int m = 0; m < n_mels; m++)
            if (rows[m][k] != 0.f) {
                if (first < 0) first = m;
                last = m, cnt++;
            }
        if (cnt == 0) {
            bj[k] = last_j;
            continue;
        }
        if (cnt > 2 || last - first > 1 || first < last_j) return WB_OK;
        bj[k] = first, wa[k] = rows[first][k], wb[k] = (last > first) ? rows[last][k] : 0.f;
        last_j = first;
    }
    for (int k = 201; k < FT_NH; k++) bj[k] = last_j;
    const double PI = 3.14159265358979323846;
    const int K = 32 * FT_KB, N = 2 * FT_NH;
    std::vector<float> hi((size_t)N * K, 0.f), lo((size_t)N * K, 0.f);
    for (int col = 0; col < N; col++) {
        const bool is_sin = col >= FT_NH;
        const int k = is_sin ? col - FT_NH : col;
        if ((!is_sin && k > 200) || (is_sin && k > 199)) continue;
        for (int n = 0; n < 400; n++) {
            const double w = 0.5 - 0.5 * cos(2.0 * PI * n / 400.0);
            const int idx = (int)(((long long)k * n) % 400);  // exact argument reduction
            const double v = w * (is_sin ? sin(2.0 * PI * idx / 400.0) : cos(2.0 * PI * idx / 400.0));
            const float h = tf32_rn((float)v);
            hi[(size_t)col * K + n] = h;
            lo[(size_t)col * K + n] = tf32_rn((float)(v - (double)h));
        }
    }
    auto up = [](auto **dst, const auto &v) {
        if (cudaMalloc((void **)dst, v.size() * sizeof(v[0])) != cudaSuccess) return false;
        return cudaMemcpy(*dst, v.data(), v.size() * sizeof(v[0]), cudaMemcpyHostToDevice) == cudaSuccess;
    };
    if (!up(&t->tc_b_hi, hi) || !up(&t->tc_b_lo, lo) || !up(&t->tc_bin_j, bj) || !up(&t->tc_bin_wa, wa) ||
        !up(&t->tc_bin_wb, wb)) {
        set_error("frontend tables: %s", cudaGetErrorString(cudaGetLastError()));
        return WB_ERR_CUDA;
    }
    t->tc_ok = true;
    return WB_OK;
}

void frontend_tc_tables_destroy(FrontendTables *t) {
    cudaFree(t->tc_b_hi), cudaFree(t->tc_b_lo), cudaFree(t->tc_bin_j), cudaFree(t->tc_bin_wa), cudaFree(t->tc_bin_wb);
    cudaFree(t->tc_ws);
}

int logmel_raw_tc(cudaStream_t st, FrontendTables &t, const float *pcm, int B, int n_frames, float *mel_raw,
                  int *chunk_max_enc) {
    WB_ARG(t.tc_ok, "tensor-core frontend tables not available");
    const int n_samples = n_frames * 160;
    // padded length: 200 + n_samples + 200, + 16 zero-weighted tail samples, rounded so the chunk stride is a
    // multiple of the 160-sample row stride (tensor-map strides must nest)
    const int pad_len = cdiv(n_samples + 416, 160) * 160;
    const size_t need = (size_t)2 * B * pad_len;
    if (t.tc_ws_cap < need) {
        cudaFree(t.tc_ws);
        t.tc_ws = nullptr, t.tc_ws_cap = 0;
        WB_CUDA(cudaMalloc((void **)&t.tc_ws, need * sizeof(float)));
        t.tc_ws_cap = need;
    }
    float *y_hi = t.tc_ws, *y_lo = t.tc_ws + (size_t)B * pad_len;
    dim3 g1(std::min(cdiv(pad_len, 256), 592), B);
    padded_split_kernel<<<g1, 256, 0, st>>>(pcm, n_samples, pad_len, y_hi, y_lo);
    WB_LAUNCHED();

    FrontendTcParams P;
    const uint64_t a_dims[3] = {416, (uint64_t)n_frames, (uint64_t)B};
    const uint64_t a_str[2] = {160 * 4, (uint64_t)pad_len * 4};
    const uint32_t a_box[3] = {32, 128, 1};
    WB_CHECK(make_tmap_f32(&P.a_hi, y_hi, 3, a_dims, a_str, a_box));
    WB_CHECK(make_tmap_f32(&P.a_lo, y_lo, 3, a_dims, a_str, a_box));
    const uint64_t b_dims[2] = {416, 416};
    const uint64_t b_str[1] = {416 * 4};
    const uint32_t b_box[2] = {32, FT_BOX};
    WB_CHECK(make_tmap_f32(&P.b_hi, t.tc_b_hi, 2, b_dims, b_str, b_box));
    WB_CHECK(make_tmap_f32(&P.b_lo, t.tc_b_lo, 2, b_dims, b_str, b_box));
    P.bin_j = t.tc_bin_j, P.bin_wa = t.tc_bin_wa, P.bin_wb = t.tc_bin_wb;
    P.mel_raw = mel_raw, P.chunk_max = chunk_max_enc;
    P.n_frames = n_frames, P.n_mels = t.n_mels;
    P.tiles_per_chunk = cdiv(n_frames, 256);
    P.total_tiles = B * P.tiles_per_chunk;
    WB_CUDA(ensure_dyn_smem(logmel_tc_kernel, FT_SMEM));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = 2 * std::min(P.total_tiles, sms / 2);
    logmel_tc_kernel<<<grid, FT_THREADS, FT_SMEM, st>>>(P);
    WB_LAUNCHED();
    return WB_OK;
}

}  // namespace wb
