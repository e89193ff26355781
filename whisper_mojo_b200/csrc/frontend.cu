// frontend.cu -- log-mel frontend on the GPU.
//
// Replaces the HF WhisperFeatureExtractor call the reference makes offline (export_weights.py:116):
// reflect-pad 200, 400-sample periodic-Hann frames at hop 160, |rDFT|^2 (201 bins), slaney mel
// filterbank (80 x 201), log10(max(., 1e-10)), clamp to (chunk max - 8), (x + 4) / 4.
//
// The DFT is done in fp32 FMA (bf16 tensor cores are not accurate enough for the 1e-4 log-mel
// tolerance) as a register-tiled contraction with two symmetry folds that cut the work 4x:
//   window symmetry  w[n] = w[400-n]:  a[n] = w[n](x[n] + x[400-n]),  d[n] = w[n](x[n] - x[400-n]), n = 1..199
//   bin symmetry     cos(2pi(200-k)n/400) = (-1)^n cos(2pi k n/400):  even / odd n accumulate separately and
//                    give bins k and 200-k at once, so only k = 0..100 is contracted.
// A CTA handles 32 frames of one chunk: samples -> shared, folded a/d -> shared ([n][frame]),
// 26 x 8 threads each own 4 bins x 4 frames x {Ce, Co, Se, So}; the power tile goes back through
// shared memory into the sparse mel filterbank and log10; the per-chunk maximum is an integer
// atomicMax on an order-preserving encoding (exact, order independent).
#include <math.h>

#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace wb {

static constexpr int FR = 32;        // frames per CTA
static constexpr int KPAD = 104;     // 101 bins padded to a multiple of 4
static constexpr int NFOLD = 200;    // n = 0..199 (row 0 unused: w[0] = 0)
static constexpr int SEG = FR * 160 + 240;  // samples a CTA touches

__device__ __forceinline__ int float_to_ordered(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void logmel_init_max_kernel(int *chunk_max, int B) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) chunk_max[i] = float_to_ordered(-INFINITY);
}

__global__ void __launch_bounds__(256) logmel_raw_kernel(const float *__restrict__ pcm, int n_frames,
                                                         const float *__restrict__ tw_cos,
                                                         const float *__restrict__ tw_sin,
                                                         const float *__restrict__ window,
                                                         const float *__restrict__ mel_w,
                                                         const int *__restrict__ mel_start,
                                                         const int *__restrict__ mel_len,
                                                         const int *__restrict__ mel_off, int n_mels,
                                                         float *__restrict__ mel_raw, int *__restrict__ chunk_max) {
    extern __shared__ float sm[];
    float *s_x = sm;                    // [SEG] samples
    float *s_a = sm + SEG;              // [200][FR]; s_a..s_d later reused as power[FR][204]
    float *s_d = s_a + NFOLD * FR;      // [200][FR]
    float *s_mid = s_d + NFOLD * FR;    // [FR] x[200] term
    __shared__ float s_red[8];

    const int b = blockIdx.y, f0 = blockIdx.x * FR;
    const int n_samples = n_frames * 160;
    const float *x = pcm + (size_t)b * n_samples;
    // samples f0*160-200 .. f0*160-200+SEG-1, reflect padded (torch.stft center=True, pad_mode="reflect")
    for (int i = threadIdx.x; i < SEG; i += 256) {
        int s = f0 * 160 - 200 + i;
        if (s < 0) s = -s;
        if (s >= n_samples) s = 2 * (n_samples - 1) - s;
        s_x[i] = (s >= 0 && s < n_samples) ? x[s] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NFOLD * FR; i += 256) {
        int n = i / FR, f = i % FR;
        if (n == 0) {  // w[0] = 0: row 0 is never read; the slot carries the x[200] term instead (w[200] = 1)
            s_a[i] = 0.f, s_d[i] = 0.f;
            s_mid[f] = s_x[f * 160 + 200];
            continue;
        }
        float lo = s_x[f * 160 + n], hi = s_x[f * 160 + 400 - n];
        float w = window[n];
        s_a[i] = w * (lo + hi);
        s_d[i] = w * (lo - hi);
    }
    __syncthreads();

    const int kq = threadIdx.x >> 3, fq = threadIdx.x & 7;  // 26 bin-quads x 8 frame-quads
    float ce[4][4] = {}, co[4][4] = {}, se[4][4] = {}, so[4][4] = {};
    if (kq < 26) {
        for (int n = 1; n < NFOLD; n += 2) {
            // odd n
            {
                float4 c4 = __ldg(reinterpret_cast<const float4 *>(tw_cos + n * KPAD) + kq);
                float4 s4 = __ldg(reinterpret_cast<const float4 *>(tw_sin + n * KPAD) + kq);
                float4 a4 = *reinterpret_cast<const float4 *>(s_a + n * FR + fq * 4);
                float4 d4 = *reinterpret_cast<const float4 *>(s_d + n * FR + fq * 4);
                const float cc[4] = {c4.x, c4.y, c4.z, c4.w}, ss[4] = {s4.x, s4.y, s4.z, s4.w};
                const float aa[4] = {a4.x, a4.y, a4.z, a4.w}, dd[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        co[i][j] += cc[i] * aa[j];
                        so[i][j] += ss[i] * dd[j];
                    }
            }
            // even n
            if (n + 1 < NFOLD) {
                const int ne = n + 1;
                float4 c4 = __ldg(reinterpret_cast<const float4 *>(tw_cos + ne * KPAD) + kq);
                float4 s4 = __ldg(reinterpret_cast<const float4 *>(tw_sin + ne * KPAD) + kq);
                float4 a4 = *reinterpret_cast<const float4 *>(s_a + ne * FR + fq * 4);
                float4 d4 = *reinterpret_cast<const float4 *>(s_d + ne * FR + fq * 4);
                const float cc[4] = {c4.x, c4.y, c4.z, c4.w}, ss[4] = {s4.x, s4.y, s4.z, s4.w};
                const float aa[4] = {a4.x, a4.y, a4.z, a4.w}, dd[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        ce[i][j] += cc[i] * aa[j];
                        se[i][j] += ss[i] * dd[j];
                    }
            }
        }
    }
    __syncthreads();  // the contraction is done with s_a / s_d; reuse them as power[FR][204]
    float *s_pow = s_a;
    if (kq < 26) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int k = kq * 4 + i;
            if (k > 100) continue;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int f = fq * 4 + j;
                const float mid = s_mid[f];
                const float sgn = (k & 1) ? -mid : mid;  // (-1)^k x[200]; (-1)^(200-k) is the same sign
                float re0 = sgn + ce[i][j] + co[i][j];
                float im0 = -(se[i][j] + so[i][j]);
                float re1 = sgn + ce[i][j] - co[i][j];
                float im1 = se[i][j] - so[i][j];
                s_pow[f * 204 + k] = re0 * re0 + im0 * im0;
                if (k < 100) s_pow[f * 204 + 200 - k] = re1 * re1 + im1 * im1;
            }
        }
    }
    __syncthreads();
    // mel filterbank (sparse rows) + log10; thread -> (mel m, frame f) with f fastest for coalesced stores
    float lmax = -INFINITY;
    for (int i = threadIdx.x; i < n_mels * FR; i += 256) {
        const int m = i / FR, f = i % FR;
        if (f0 + f >= n_frames) continue;
        const int st = mel_start[m], ln = mel_len[m];
        const float *w = mel_w + mel_off[m];
        float acc = 0.f;
        for (int t = 0; t < ln; t++) acc += w[t] * s_pow[f * 204 + st + t];
        float v = log10f(fmaxf(acc, 1e-10f));
        mel_raw[((size_t)b * n_mels + m) * n_frames + f0 + f] = v;
        lmax = fmaxf(lmax, v);
    }
    lmax = warp_max(lmax);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = lmax;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = s_red[0];
        for (int i = 1; i < 8; i++) m = fmaxf(m, s_red[i]);
        atomicMax(&chunk_max[b], float_to_ordered(m));
    }
}

__global__ void logmel_finalize_kernel(float *__restrict__ mel, const int *__restrict__ chunk_max, size_t per_chunk,
                                       size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        float cm = ordered_to_float(chunk_max[i / per_chunk]);
        float v = fmaxf(mel[i], cm - 8.0f);
        mel[i] = (v + 4.0f) / 4.0f;
    }
}

// [n_mels][n_frames] f32 -> [n_frames][128] bf16 through a 32x33 shared tile.
__global__ void mel_to_h16_T_kernel(const float *__restrict__ mel, h16 *__restrict__ out, int n_mels,
                                     int n_frames) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, f0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
    const float *src = mel + (size_t)b * n_mels * n_frames;
    for (int i = threadIdx.y; i < 32; i += 8) {
        int m = m0 + i, f = f0 + threadIdx.x;
        tile[i][threadIdx.x] = (m < n_mels && f < n_frames) ? src[(size_t)m * n_frames + f] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        int f = f0 + i, m = m0 + threadIdx.x;
        if (f < n_frames) out[((size_t)b * n_frames + f) * 128 + m] = f2h(tile[threadIdx.x][i]);
    }
}

// ---- host side ---------------------------------------------------------------------------------

static double hz_to_mel(double f) {
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_mel + log(f / min_log_hz) / logstep : f / f_sp;
}
static double mel_to_hz(double m) {
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * exp(logstep * (m - min_log_mel)) : f_sp * m;
}

template <typename T>
static int upload(T **dst, const std::vector<T> &v) {
    WB_CUDA(cudaMalloc((void **)dst, v.size() * sizeof(T)));
    WB_CUDA(cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return WB_OK;
}

int frontend_tc_tables_create(FrontendTables *t, const std::vector<std::vector<float>> &rows);  // frontend_tc.cu
void frontend_tc_tables_destroy(FrontendTables *t);

int frontend_tables_create(FrontendTables *t, int n_mels) {
    t->n_mels = n_mels;
    const double PI = 3.14159265358979323846;
    std::vector<float> c((size_t)NFOLD * KPAD, 0.f), s((size_t)NFOLD * KPAD, 0.f), w(400);
    for (int n = 0; n < NFOLD; n++)
        for (int k = 0; k <= 100; k++) {
            int idx = (int)(((long long)k * n) % 400);  // exact argument reduction
            c[(size_t)n * KPAD + k] = (float)cos(2.0 * PI * idx / 400.0);
            s[(size_t)n * KPAD + k] = (float)sin(2.0 * PI * idx / 400.0);
        }
    for (int n = 0; n < 400; n++) w[n] = (float)(0.5 - 0.5 * cos(2.0 * PI * n / 400.0));
    // slaney mel filterbank, float64 then cast (same construction as HF mel_filter_bank(norm="slaney",
    // mel_scale="slaney"); oracle/logmel_oracle.py holds the test-side restatement)
    const int n_freqs = 201;
    std::vector<double> f_pts(n_mels + 2);
    const double m_lo = hz_to_mel(0.0), m_hi = hz_to_mel(8000.0);
    for (int i = 0; i < n_mels + 2; i++) f_pts[i] = mel_to_hz(m_lo + (m_hi - m_lo) * i / (n_mels + 1));
    std::vector<float> mw;
    std::vector<int> mstart(n_mels), mlen(n_mels), moff(n_mels);
    std::vector<std::vector<float>> dense;
    for (int m = 0; m < n_mels; m++) {
        const double enorm = 2.0 / (f_pts[m + 2] - f_pts[m]);
        int first = -1, last = -1;
        std::vector<float> row(n_freqs, 0.f);
        for (int k = 0; k < n_freqs; k++) {
            const double fr = 8000.0 * k / (n_freqs - 1);
            const double lower = (fr - f_pts[m]) / (f_pts[m + 1] - f_pts[m]);
            const double upper = (f_pts[m + 2] - fr) / (f_pts[m + 2] - f_pts[m + 1]);
            const double v = fmax(0.0, fmin(lower, upper)) * enorm;
            row[k] = (float)v;
            if (row[k] != 0.f) {
                if (first < 0) first = k;
                last = k;
            }
        }
        dense.push_back(row);
        if (first < 0) first = 0, last = -1;
        mstart[m] = first;
        mlen[m] = last - first + 1;
        moff[m] = (int)mw.size();
        for (int k = first; k <= last; k++) mw.push_back(row[k]);
    }
    if (mw.empty()) mw.push_back(0.f);
    WB_CHECK(upload(&t->tw_cos, c));
    WB_CHECK(upload(&t->tw_sin, s));
    WB_CHECK(upload(&t->window, w));
    WB_CHECK(upload(&t->mel_w, mw));
    WB_CHECK(upload(&t->mel_start, mstart));
    WB_CHECK(upload(&t->mel_len, mlen));
    WB_CHECK(upload(&t->mel_off, moff));
    return frontend_tc_tables_create(t, dense);
}

void frontend_tables_destroy(FrontendTables *t) {
    frontend_tc_tables_destroy(t);
    cudaFree(t->tw_cos);
    cudaFree(t->tw_sin);
    cudaFree(t->window);
    cudaFree(t->mel_w);
    cudaFree(t->mel_start);
    cudaFree(t->mel_len);
    cudaFree(t->mel_off);
    *t = FrontendTables();
}

int logmel_raw(cudaStream_t st, FrontendTables &t, const float *pcm, int B, int n_frames, float *mel_raw,
               int *chunk_max_enc, int impl) {
    if (B <= 0) return WB_OK;
    WB_ARG(t.tw_cos && t.n_mels > 0, "frontend tables not initialised");
    logmel_init_max_kernel<<<cdiv(B, 256), 256, 0, st>>>(chunk_max_enc, B);
    WB_LAUNCHED();
    if (impl == 1 && t.tc_ok) return logmel_raw_tc(st, t, pcm, B, n_frames, mel_raw, chunk_max_enc);
    const size_t smem = (size_t)(SEG + 2 * NFOLD * FR + FR) * sizeof(float);
    WB_CUDA(ensure_dyn_smem(logmel_raw_kernel, smem));
    static_assert(2 * NFOLD * FR >= FR * 204, "power tile must fit in the folded-sample buffers");
    dim3 grid(cdiv(n_frames, FR), B);
    logmel_raw_kernel<<<grid, 256, smem, st>>>(pcm, n_frames, t.tw_cos, t.tw_sin, t.window, t.mel_w, t.mel_start,
                                               t.mel_len, t.mel_off, t.n_mels, mel_raw, chunk_max_enc);
    WB_LAUNCHED();
    return WB_OK;
}

int logmel_finalize(cudaStream_t st, float *mel, const int *chunk_max_enc, int B, int n_mels, int n_frames) {
    if (B <= 0) return WB_OK;
    size_t per = (size_t)n_mels * n_frames, total = per * B;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    logmel_finalize_kernel<<<blocks, 256, 0, st>>>(mel, chunk_max_enc, per, total);
    WB_LAUNCHED();
    return WB_OK;
}

int mel_to_h16_T(cudaStream_t st, const float *mel, h16 *out, int B, int n_mels, int n_frames) {
    if (B <= 0) return WB_OK;
    dim3 grid(cdiv(n_frames, 32), 4, B), block(32, 8);  // 4 x 32 = 128 channels (>= n_mels zero)
    mel_to_h16_T_kernel<<<grid, block, 0, st>>>(mel, out, n_mels, n_frames);
    WB_LAUNCHED();
    return WB_OK;
}

}  // namespace wb
