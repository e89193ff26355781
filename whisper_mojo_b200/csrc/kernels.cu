// kernels.cu -- LayerNorm, embedding, single-query (decode) attention, bring-up encoder attention and
// greedy-loop bookkeeping for the batched fast path.  All HBM-bound: coalesced 128-bit accesses,
// warp-shuffle reductions, no atomics in value-producing reductions (results do not depend on batch
// size or position in the batch).
#include "kernels.h"

#include <math.h>

#include <algorithm>

#include "common.cuh"

namespace wb {

// ---------------------------------------------------------------------------------------------
// LayerNorm fp32 -> bf16: one warp per row.
// ---------------------------------------------------------------------------------------------
enum { LN_PLAIN = 0, LN_EMBED = 1, LN_RESID = 2 };
// MODE LN_RESID: x[row] += bias + sum_s part[s][row] (fixed order: deterministic) before the LayerNorm -- the second
// half of a split-K residual GEMM fused with the LayerNorm that follows it; gamma == nullptr skips the LN output.
// NVT: d_model / 128 at compile time (0 = read it from D): the LN_RESID form keeps a whole row and four slices of it in
// registers, and sized for the general case (8 vectors) that is 224 registers = one CTA per SM.
template <int EMBED, int NVT = 0>
__global__ void __launch_bounds__(256) ln_h16_kernel(const float *__restrict__ x_in, const float *__restrict__ gamma,
                                                      const float *__restrict__ beta, int rows, int D,
                                                      h16 *__restrict__ out_h16,
                                                      float *__restrict__ out_f32,
                                                      // EMBED only:
                                                      const float *__restrict__ tok_emb,
                                                      const float *__restrict__ pos_emb,
                                                      const int *__restrict__ cur_tok, const int *__restrict__ pos_dev,
                                                      int vocab, int n_pos, float *__restrict__ x_out,
                                                      // LN_EMBED, prefill (q_len > 1): row r holds prompt token r % q_len at
                                                      // position r % q_len; *set_len = q_len - 1 (the bookkeeping kernel
                                                      // after the logits adds the last 1)
                                                      int q_len, int4 prompt, int *__restrict__ set_len,
                                                      // LN_RESID only:
                                                      const float *__restrict__ part, int n_split,
                                                      long long split_stride, const float *__restrict__ bias) {
    pdl_launch_dependents();
    pdl_wait();
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= rows) return;
    // D is a multiple of 128 for every supported config (heads * 64 with an even head count);
    // up to 8 float4 per lane (D <= 1024) are kept in registers.
    constexpr int NVM = NVT ? NVT : 8;
    float4 v[NVM];
    const int nvec = NVT ? NVT : D >> 7;
    float s = 0.f, q = 0.f;
    if (EMBED == LN_RESID) {
        // Every load of the row is issued before the first add: the row of x, the bias and the slices four at a time
        // (two L2 round trips for six slices).  Vector by vector -- x, bias, slices, store, next vector -- this was
        // ~6 serial round trips per row for d_model = 384.  The order of the adds is unchanged: x + bias + part[0] + ...
        float4 *xr = reinterpret_cast<float4 *>(x_out + (size_t)row * D);
        float4 bb[NVM];
#pragma unroll
        for (int i = 0; i < NVM; i++)
            if (i < nvec) {
                v[i] = xr[i * 32 + lane];
                if (bias) bb[i] = reinterpret_cast<const float4 *>(bias)[i * 32 + lane];
            }
        for (int s0 = 0; s0 < n_split; s0 += 4) {
            float4 pp[4][NVM];
#pragma unroll
            for (int k = 0; k < 4; k++)
#pragma unroll
                for (int i = 0; i < NVM; i++)
                    if (i < nvec && s0 + k < n_split)
                        pp[k][i] = reinterpret_cast<const float4 *>(part + (size_t)(s0 + k) * split_stride + (size_t)row * D)[i * 32 + lane];
            if (s0 == 0 && bias) {
#pragma unroll
                for (int i = 0; i < NVM; i++)
                    if (i < nvec) v[i].x += bb[i].x, v[i].y += bb[i].y, v[i].z += bb[i].z, v[i].w += bb[i].w;
            }
#pragma unroll
            for (int k = 0; k < 4; k++)
#pragma unroll
                for (int i = 0; i < NVM; i++)
                    if (i < nvec && s0 + k < n_split)
                        v[i].x += pp[k][i].x, v[i].y += pp[k][i].y, v[i].z += pp[k][i].z, v[i].w += pp[k][i].w;
        }
        if (n_split == 0 && bias) {
#pragma unroll
            for (int i = 0; i < NVM; i++)
                if (i < nvec) v[i].x += bb[i].x, v[i].y += bb[i].y, v[i].z += bb[i].z, v[i].w += bb[i].w;
        }
#pragma unroll
        for (int i = 0; i < NVM; i++)
            if (i < nvec) xr[i * 32 + lane] = v[i];
        if (!gamma) return;
    } else if (EMBED == LN_EMBED) {
        int tok, pos;
        if (q_len > 1) {  // whisper.mojo:195-197: the prompt ids at positions 0 .. q_len-1, every chunk alike
            pos = row % q_len;
            tok = pos == 0 ? prompt.x : (pos == 1 ? prompt.y : (pos == 2 ? prompt.z : prompt.w));
            if (row == 0 && lane == 0 && set_len) *set_len = q_len - 1;
        } else {
            tok = cur_tok[row];
            pos = *pos_dev;
        }
        tok = tok < 0 ? 0 : (tok >= vocab ? vocab - 1 : tok);
        pos = pos < 0 ? 0 : (pos >= n_pos ? n_pos - 1 : pos);
        const float4 *te = reinterpret_cast<const float4 *>(tok_emb + (size_t)tok * D);
        const float4 *pe = reinterpret_cast<const float4 *>(pos_emb + (size_t)pos * D);
#pragma unroll
        for (int i = 0; i < NVM; i++)
            if (i < nvec) {
                float4 a = te[i * 32 + lane], b = pe[i * 32 + lane];
                v[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
                reinterpret_cast<float4 *>(x_out + (size_t)row * D)[i * 32 + lane] = v[i];
            }
    } else {
        const float4 *xr = reinterpret_cast<const float4 *>(x_in + (size_t)row * D);
#pragma unroll
        for (int i = 0; i < NVM; i++)
            if (i < nvec) v[i] = xr[i * 32 + lane];
    }
#pragma unroll
    for (int i = 0; i < NVM; i++)
        if (i < nvec) {
            s += v[i].x + v[i].y + v[i].z + v[i].w;
            q += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
        }
    s = warp_sum(s);
    q = warp_sum(q);
    const float mean = s / (float)D;
    const float var = q / (float)D - mean * mean;
    const float inv_std = 1.0f / sqrtf(var + 1e-5f);
#pragma unroll
    for (int i = 0; i < NVM; i++)
        if (i < nvec) {
            float4 g = reinterpret_cast<const float4 *>(gamma)[i * 32 + lane];
            float4 b = reinterpret_cast<const float4 *>(beta)[i * 32 + lane];
            float4 r;
            r.x = (v[i].x - mean) * inv_std * g.x + b.x;
            r.y = (v[i].y - mean) * inv_std * g.y + b.y;
            r.z = (v[i].z - mean) * inv_std * g.z + b.z;
            r.w = (v[i].w - mean) * inv_std * g.w + b.w;
            uint2 pk;
            pk.x = pack_h2(r.x, r.y);
            pk.y = pack_h2(r.z, r.w);
            reinterpret_cast<uint2 *>(out_h16 + (size_t)row * D)[i * 32 + lane] = pk;
            if (out_f32) reinterpret_cast<float4 *>(out_f32 + (size_t)row * D)[i * 32 + lane] = r;
        }
}

int ln_h16(cudaStream_t st, const float *x, const float *gamma, const float *beta, int rows, int D,
            h16 *out_h16, float *out_f32) {
    WB_ARG(D % 128 == 0 && D <= 1024, "ln_h16: D=%d must be a multiple of 128 and <= 1024", D);
    if (rows <= 0) return WB_OK;
    WB_CUDA(launch_pdl(ln_h16_kernel<LN_PLAIN>, dim3(cdiv(rows, 8)), dim3(256), 0, st, x, gamma, beta, rows, D, out_h16,
                       out_f32, nullptr, nullptr, nullptr, nullptr, 0, 0, nullptr, 1, make_int4(0, 0, 0, 0), nullptr, nullptr, 0, 0, nullptr));
    WB_LAUNCHED();
    return WB_OK;
}

int resid_ln(cudaStream_t st, float *x, const float *part, int n_split, const float *bias, const float *gamma,
             const float *beta, int rows, int D, h16 *out_h16) {
    WB_ARG(D % 128 == 0 && D <= 1024, "resid_ln: D=%d must be a multiple of 128 and <= 1024", D);
    WB_ARG(x && part && n_split >= 1 && (!gamma || (beta && out_h16)), "resid_ln: bad arguments");
    if (rows <= 0) return WB_OK;
    auto go = [&](auto kernel) {
        return launch_pdl(kernel, dim3(cdiv(rows, 8)), dim3(256), 0, st, (const float *)nullptr, gamma, beta, rows, D, out_h16,
                          (float *)nullptr, (const float *)nullptr, (const float *)nullptr, (const int *)nullptr, (const int *)nullptr, 0, 0,
                          x, 1, make_int4(0, 0, 0, 0), (int *)nullptr, part, n_split, (long long)rows * D, bias);
    };
    switch (D >> 7) {  // row + slices in registers: sized to d_model at compile time
        case 1: WB_CUDA(go(ln_h16_kernel<LN_RESID, 1>)); break;
        case 2: WB_CUDA(go(ln_h16_kernel<LN_RESID, 2>)); break;
        case 3: WB_CUDA(go(ln_h16_kernel<LN_RESID, 3>)); break;
        case 4: WB_CUDA(go(ln_h16_kernel<LN_RESID, 4>)); break;
        case 5: WB_CUDA(go(ln_h16_kernel<LN_RESID, 5>)); break;
        case 6: WB_CUDA(go(ln_h16_kernel<LN_RESID, 6>)); break;
        default: WB_CUDA(go(ln_h16_kernel<LN_RESID, 0>)); break;
    }
    WB_LAUNCHED();
    return WB_OK;
}

int embed_ln(cudaStream_t st, const float *tok_emb, const float *pos_emb, const int *cur_tok, const int *pos_dev,
             int B, int D, int vocab, int n_pos, const float *gamma, const float *beta, float *x,
             h16 *xn) {
    WB_ARG(D % 128 == 0 && D <= 1024, "embed_ln: D=%d must be a multiple of 128 and <= 1024", D);
    if (B <= 0) return WB_OK;
    WB_CUDA(launch_pdl(ln_h16_kernel<LN_EMBED>, dim3(cdiv(B, 8)), dim3(256), 0, st, nullptr, gamma, beta, B, D, xn, nullptr,
                       tok_emb, pos_emb, cur_tok, pos_dev, vocab, n_pos, x, 1, make_int4(0, 0, 0, 0), nullptr, nullptr, 0, 0, nullptr));
    WB_LAUNCHED();
    return WB_OK;
}

int embed_ln_prefill(cudaStream_t st, const float *tok_emb, const float *pos_emb, const int *prompt4_host, int q_len, int B,
                     int D, int vocab, int n_pos, const float *gamma, const float *beta, float *x, h16 *xn, int *set_len) {
    WB_ARG(D % 128 == 0 && D <= 1024, "embed_ln_prefill: D=%d must be a multiple of 128 and <= 1024", D);
    WB_ARG(q_len >= 2 && q_len <= 4, "embed_ln_prefill: q_len=%d (the reference's prompt has 4 ids)", q_len);
    if (B <= 0) return WB_OK;
    const int4 pr = make_int4(prompt4_host[0], prompt4_host[1], prompt4_host[2], prompt4_host[3]);
    WB_CUDA(launch_pdl(ln_h16_kernel<LN_EMBED>, dim3(cdiv(B * q_len, 8)), dim3(256), 0, st, nullptr, gamma, beta, B * q_len, D, xn,
                       nullptr, tok_emb, pos_emb, nullptr, nullptr, vocab, n_pos, x, q_len, pr, set_len, nullptr, 0, 0, nullptr));
    WB_LAUNCHED();
    return WB_OK;
}

// k / v rows of a q_len-token forward, [B * q_len][D] each (row b * q_len + p), into rows 0 .. q_len-1 of the self K/V
// cache (layers.mojo:131-143 with start_pos = 0): one 16-byte piece per thread.
__global__ void kv_scatter_kernel(const h16 *__restrict__ k, const h16 *__restrict__ v, h16 *__restrict__ Kc,
                                  h16 *__restrict__ Vc, int rows, int q_len, int D, long long kv_batch_stride) {
    pdl_launch_dependents();
    pdl_wait();
    const int per_row = D >> 3;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)rows * per_row) return;
    const int r = (int)(i / per_row), c = (int)(i - (long long)r * per_row);
    const long long dst = (long long)(r / q_len) * kv_batch_stride + (long long)(r % q_len) * D + c * 8;
    *reinterpret_cast<uint4 *>(Kc + dst) = *reinterpret_cast<const uint4 *>(k + (size_t)r * D + c * 8);
    *reinterpret_cast<uint4 *>(Vc + dst) = *reinterpret_cast<const uint4 *>(v + (size_t)r * D + c * 8);
}
int kv_scatter(cudaStream_t st, const h16 *k, const h16 *v, h16 *Kc, h16 *Vc, int B, int q_len, int D, int64_t kv_batch_stride) {
    WB_ARG(D % 8 == 0 && q_len >= 1, "kv_scatter: bad shape");
    if (B <= 0) return WB_OK;
    const long long n = (long long)B * q_len * (D >> 3);
    WB_CUDA(launch_pdl(kv_scatter_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, k, v, Kc, Vc, B * q_len, q_len, D,
                       (long long)kv_batch_stride));
    WB_LAUNCHED();
    return WB_OK;
}

// ---------------------------------------------------------------------------------------------
// Decode attention: grid (B, splits), one warp per head.  Lane layout inside a warp: 4 key rows x
// 8 lanes, each lane owning 8 of the 64 head dims (one 128-bit load); a key row's head slice is one
// full 128-byte line, and the H warps of a CTA walk the same rows, so a CTA streams whole K / V rows.
// Three phases as in the reference (scores -> softmax -> weighted sum); scores live in shared memory.
// ---------------------------------------------------------------------------------------------
struct DecodeAttnDev {
    const h16 *q, *K, *V;
    h16 *out;
    long long kv_batch_stride;
    int H, D, len_const, len_add, splits, smem_len;
    int q_len;  // > 1: grid row vb is query vb % q_len of chunk vb / q_len; without len_const it attends keys 0 .. vb % q_len
    const int *len_dev;
    float *ws;
    const int *done;
};

// K / V rows travel global -> shared memory as 16-byte cp.async copies (L2 only), DA_STAGES iterations of 16 rows
// ahead of their use: loads in flight cost no registers, so a warp keeps 3 x 4 x 16 B per lane outstanding instead of
// 4 x 16 B (the kernel was bound by bytes in flight, not by HBM: ~73 KB per SM against the ~85 KB that 43 GB/s per SM
// at ~2 us loaded latency need).  Every lane reads back only the slots it filled itself, so no barrier is involved,
// and the arithmetic (lane -> row / dims mapping, order of the sums) is unchanged: same bits as the register version.
static constexpr int DA_STAGES = 3;
static constexpr int DA_STAGE_BYTES_PER_WARP = DA_STAGES * 4 * 32 * 16;
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ uint4 ld_stream(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

__global__ void __launch_bounds__(384) decode_attn_kernel(const DecodeAttnDev p) {
    extern __shared__ float s_scores[];  // [H][smem_len]
    pdl_launch_dependents();
    pdl_wait();
    // b: row of q / out / ws; kb: the chunk whose K / V it attends over.  q_len > 1 is the causal block path of the
    // reference (layers.mojo:304-320: keys j > i are filled with -1e10, i.e. they get weight exp(-1e10 - max) = 0
    // exactly) done as q_len independent single-query problems of lengths 1 .. q_len -- the arithmetic of a row is the
    // one a cached single-token step at that position runs, bit for bit.
    const int b = blockIdx.x, split = blockIdx.y;
    const int kb = p.q_len > 1 ? b / p.q_len : b;
    if (p.done && p.done[kb]) return;  // finished chunk (whisper.mojo:206-207): nothing reads its output any more
    const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = lane & 7, sub = lane >> 3;
    const int len = p.len_dev ? (*p.len_dev + p.len_add) : (p.len_const > 0 ? p.len_const : b % p.q_len + 1);
    int chunk = (len + p.splits - 1) / p.splits;
    chunk = (chunk + 3) & ~3;
    const int j0 = split * chunk;
    const int j1 = min(len, j0 + chunk);
    const int n = max(j1 - j0, 0);
    float *sc = s_scores + (size_t)h * p.smem_len;
    // this warp's staging slots [stage][u][lane] of 16 bytes, behind the H score rows (16-byte aligned: smem_len % 4 == 0)
    uint4 *stg = reinterpret_cast<uint4 *>(s_scores + (size_t)p.H * p.smem_len) + (size_t)h * (DA_STAGES * 4 * 32) + lane;
    const int n_it = (n + 15) >> 4;
    // queue the copies of iteration `it` (rows it*16 + sub + 4u, u < 4) of the matrix at `base`, one commit group each
    auto issue = [&](const h16 *base, int it) {
        if (it < n_it) {
            const int i = it * 16 + sub;
            uint4 *dst = stg + (it % DA_STAGES) * (4 * 32);
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (i + 4 * u < n) cp_async16(dst + u * 32, base + (size_t)(j0 + i + 4 * u) * p.D);
        }
        cp_async_commit();
    };

    float qf[8];
    h8_to_float(*reinterpret_cast<const uint4 *>(p.q + (size_t)b * p.D + h * 64 + c * 8), qf);
    const float scale = 0.125f;  // 1/sqrt(head_dim), head_dim = 64 (layers.mojo:184)
    const h16 *Kb = p.K + (size_t)kb * p.kv_batch_stride + h * 64 + c * 8;
    const h16 *Vb = p.V + (size_t)kb * p.kv_batch_stride + h * 64 + c * 8;

    // phase 1: scores
    float m = -1e10f;  // layers.mojo:188
#pragma unroll
    for (int s = 0; s < DA_STAGES - 1; s++) issue(Kb, s);
    for (int i0 = 0, it = 0; i0 < n; i0 += 16, it++) {  // warp-uniform trip count: the shuffles below need all 32 lanes
        const int i = i0 + sub;
        issue(Kb, it + DA_STAGES - 1);  // refills the slot read in the previous iteration
        cp_async_wait<DA_STAGES - 1>();
        uint4 kv[4];
        const uint4 *src = stg + (it % DA_STAGES) * (4 * 32);
#pragma unroll
        for (int u = 0; u < 4; u++) {
            int jj = i + 4 * u;
            kv[u] = (jj < n) ? src[u * 32] : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            float kf[8];
            h8_to_float(kv[u], kf);
            float d = 0.f;
#pragma unroll
            for (int t = 0; t < 8; t++) d += qf[t] * kf[t];
            d += __shfl_xor_sync(0xffffffffu, d, 1);
            d += __shfl_xor_sync(0xffffffffu, d, 2);
            d += __shfl_xor_sync(0xffffffffu, d, 4);
            int jj = i + 4 * u;
            if (jj < n) {
                float s = d * scale;
                m = fmaxf(m, s);
                if (c == 0) sc[jj] = s;
            }
        }
    }
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
    cp_async_wait<0>();  // (only empty groups are left) every slot is free again:
#pragma unroll
    for (int s = 0; s < DA_STAGES - 1; s++) issue(Vb, s);  // the first V rows travel under the softmax
    __syncwarp();
    // phase 2: exp and sum
    float l = 0.f;
    for (int j = lane; j < n; j += 32) {
        float e = __expf(sc[j] - m);
        sc[j] = e;
        l += e;
    }
    l = warp_sum(l);
    __syncwarp();
    // phase 3: weighted sum of V
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int i0 = 0, it = 0; i0 < n; i0 += 16, it++) {
        const int i = i0 + sub;
        issue(Vb, it + DA_STAGES - 1);
        cp_async_wait<DA_STAGES - 1>();
        uint4 vv[4];
        const uint4 *src = stg + (it % DA_STAGES) * (4 * 32);
#pragma unroll
        for (int u = 0; u < 4; u++) {
            int jj = i + 4 * u;
            vv[u] = (jj < n) ? src[u * 32] : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            int jj = i + 4 * u;
            float pj = (jj < n) ? sc[jj] : 0.f;
            float vf[8];
            h8_to_float(vv[u], vf);
#pragma unroll
            for (int t = 0; t < 8; t++) acc[t] += pj * vf[t];
        }
    }
#pragma unroll
    for (int t = 0; t < 8; t++) {
        acc[t] += __shfl_xor_sync(0xffffffffu, acc[t], 8);
        acc[t] += __shfl_xor_sync(0xffffffffu, acc[t], 16);
    }
    if (p.splits == 1) {
        if (sub == 0) {
            const float inv = 1.0f / l;
            uint4 o;
            o.x = pack_h2(acc[0] * inv, acc[1] * inv);
            o.y = pack_h2(acc[2] * inv, acc[3] * inv);
            o.z = pack_h2(acc[4] * inv, acc[5] * inv);
            o.w = pack_h2(acc[6] * inv, acc[7] * inv);
            *reinterpret_cast<uint4 *>(p.out + (size_t)b * p.D + h * 64 + c * 8) = o;
        }
    } else {
        float *w = p.ws + (((size_t)b * p.splits + split) * p.H + h) * 66;
        if (sub == 0) {
#pragma unroll
            for (int t = 0; t < 8; t++) w[c * 8 + t] = acc[t];
            if (c == 0) w[64] = m, w[65] = l;
        }
    }
}

// Merge split-K partials: one warp per (b, h); lane owns 2 head dims.
__global__ void decode_attn_combine_kernel(const float *__restrict__ ws, h16 *__restrict__ out, int B, int H,
                                           int D, int splits) {
    int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= B * H) return;
    int b = w / H, h = w % H;
    float m = -1e10f;
    for (int s = 0; s < splits; s++) m = fmaxf(m, ws[(((size_t)b * splits + s) * H + h) * 66 + 64]);
    float l = 0.f, o0 = 0.f, o1 = 0.f;
    for (int s = 0; s < splits; s++) {
        const float *p = ws + (((size_t)b * splits + s) * H + h) * 66;
        float f = __expf(p[64] - m);
        l += p[65] * f;
        o0 += p[2 * lane] * f;
        o1 += p[2 * lane + 1] * f;
    }
    float inv = 1.0f / l;
    *reinterpret_cast<uint32_t *>(out + (size_t)b * D + h * 64 + 2 * lane) = pack_h2(o0 * inv, o1 * inv);
}

// Split-K factor for a constant key length.  Every CTA streams the same number of K/V bytes, so the
// kernel's efficiency is (CTAs / resident slots) / ceil(CTAs / resident slots): with one CTA per chunk
// 2048 chunks over ~1000 slots is 2.05 waves -> 3 rounds.  Pick the smallest split count whose last
// wave is >= 92 % full (or the best one), keeping at least 96 keys per split.
int decode_attention_splits(int B, int len, int H) {
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int max_splits = len / 96;
    if (max_splits < 1) max_splits = 1;
    if (max_splits > 32) max_splits = 32;
    int best = 1;
    double best_eff = 0.0;
    for (int s = 1; s <= max_splits; s++) {
        int chunk = ((len + s - 1) / s + 3) & ~3;
        // resident CTAs per SM for this shared-memory footprint, from the occupancy calculator
        size_t smem = (size_t)H * (chunk + 4) * sizeof(float) + (size_t)H * DA_STAGE_BYTES_PER_WARP;
        ensure_dyn_smem(decode_attn_kernel, smem);
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_attn_kernel, H * 32, smem) != cudaSuccess)
            per_sm = 4, cudaGetLastError();
        if (per_sm < 1) per_sm = 1;
        double waves = (double)B * s / ((double)sms * per_sm);
        double eff = waves / ceil(waves);
        if (waves < 1.0) eff = waves;  // not even one full wave: more splits = more parallelism
        eff *= 1.0 - 0.004 * (s - 1);  // fixed per-CTA cost (query load, softmax, partial write)
        if (eff > best_eff + 1e-9) best_eff = eff, best = s;
        if (eff >= 0.92) return s;
    }
    return best;
}

int decode_attention(cudaStream_t st, const DecodeAttnArgs &a) {
    WB_ARG(a.H >= 1 && a.H <= 12 && a.D == a.H * 64, "decode_attention: H=%d D=%d unsupported", a.H, a.D);
    WB_ARG(a.splits >= 1 && (a.splits == 1 || a.ws), "decode_attention: splits need a workspace");
    if (a.B <= 0) return WB_OK;
    DecodeAttnDev p;
    p.q = a.q, p.K = a.K, p.V = a.V, p.out = a.out;
    p.kv_batch_stride = a.kv_batch_stride;
    p.H = a.H, p.D = a.D, p.len_const = a.len_const, p.len_add = a.len_add, p.splits = a.splits;
    p.len_dev = a.len_dev, p.ws = a.ws, p.done = a.done;
    p.q_len = a.q_len < 1 ? 1 : a.q_len;
    WB_ARG(p.q_len == 1 || (!a.len_dev && a.B % p.q_len == 0), "decode_attention: q_len=%d needs B %% q_len == 0 and no device length", p.q_len);
    WB_ARG(a.len_dev || a.len_const > 0 || p.q_len > 1, "decode_attention: no key length");
    int chunk = (a.max_len + a.splits - 1) / a.splits;
    p.smem_len = ((chunk + 3) & ~3) + 4;
    size_t smem = (size_t)a.H * p.smem_len * sizeof(float) + (size_t)a.H * DA_STAGE_BYTES_PER_WARP;
    WB_CUDA(ensure_dyn_smem(decode_attn_kernel, smem));
    dim3 grid(a.B, a.splits);
    WB_CUDA(launch_pdl(decode_attn_kernel, grid, dim3(a.H * 32), smem, st, p));
    WB_LAUNCHED();
    if (a.splits > 1) {
        decode_attn_combine_kernel<<<cdiv(a.B * a.H, 8), 256, 0, st>>>(a.ws, a.out, a.B, a.H, a.D, a.splits);
        WB_LAUNCHED();
    }
    return WB_OK;
}

// ---------------------------------------------------------------------------------------------
// Encoder attention, bring-up version: one warp per query row, scores in shared memory.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) encoder_attn_ref_kernel(const h16 *__restrict__ qkv,
                                                               h16 *__restrict__ out, int S, int H, int D) {
    extern __shared__ float s_sc[];  // [8][S]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qi = blockIdx.x * 8 + warp, h = blockIdx.y, b = blockIdx.z;
    if (qi >= S) return;
    float *sc = s_sc + (size_t)warp * S;
    const size_t ld = 3 * (size_t)D;
    const h16 *base = qkv + (size_t)b * S * ld;
    float2 qv = h22f2(*reinterpret_cast<const h16x2 *>(base + (size_t)qi * ld + h * 64 + 2 * lane));
    float m = -INFINITY;
    for (int j = 0; j < S; j++) {
        float2 kv = h22f2(
            *reinterpret_cast<const h16x2 *>(base + (size_t)j * ld + D + h * 64 + 2 * lane));
        float d = warp_sum(qv.x * kv.x + qv.y * kv.y) * 0.125f;
        m = fmaxf(m, d);
        if (lane == 0) sc[j] = d;
    }
    __syncwarp();
    float l = 0.f;
    for (int j = lane; j < S; j += 32) {
        float e = expf(sc[j] - m);
        sc[j] = e;
        l += e;
    }
    l = warp_sum(l);
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < S; j++) {
        float2 vv = h22f2(
            *reinterpret_cast<const h16x2 *>(base + (size_t)j * ld + 2 * D + h * 64 + 2 * lane));
        float pj = sc[j];
        o0 += pj * vv.x;
        o1 += pj * vv.y;
    }
    float inv = 1.0f / l;
    *reinterpret_cast<uint32_t *>(out + ((size_t)b * S + qi) * D + h * 64 + 2 * lane) = pack_h2(o0 * inv, o1 * inv);
}

int encoder_attention_ref(cudaStream_t st, const h16 *qkv, h16 *out, int B, int S, int H, int D) {
    if (B <= 0) return WB_OK;
    size_t smem = (size_t)8 * S * sizeof(float);
    WB_CUDA(ensure_dyn_smem(encoder_attn_ref_kernel, smem));
    dim3 grid(cdiv(S, 8), H, B);
    encoder_attn_ref_kernel<<<grid, 256, smem, st>>>(qkv, out, S, H, D);
    WB_LAUNCHED();
    return WB_OK;
}

// ---------------------------------------------------------------------------------------------
// Greedy-loop bookkeeping
// ---------------------------------------------------------------------------------------------
__global__ void greedy_init_kernel(GreedyState g, int B, int p0, int p1, int p2, int p3) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) g.scalars[0] = 0, g.scalars[1] = 0, g.scalars[2] = 0, g.scalars[3] = B;
    if (i >= B) return;
    if (g.live) g.live[i] = i;
    int *row = g.tokens_out + (size_t)i * g.T_out;
    for (int t = 0; t < g.T_out; t++) row[t] = -1;
    row[0] = p0, row[1] = p1, row[2] = p2, row[3] = p3;
    g.out_len[i] = 4;
    g.cur_tok[i] = p0;
    g.done[i] = 0;
}
// Ordered compaction of the not-done chunk indices: one CTA of 1024 threads, each owning a contiguous run of chunks;
// an exclusive scan of the per-thread counts places the runs.
__global__ void __launch_bounds__(1024) greedy_rebuild_live_kernel(GreedyState g, int B) {
    __shared__ int warp_tot[32];
    const int tid = threadIdx.x, per = (B + 1023) / 1024, lo = min(B, tid * per), hi = min(B, lo + per);
    int cnt = 0;
    for (int i = lo; i < hi; i++) cnt += g.done[i] == 0;
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((tid & 31) >= o) incl += v;
    }
    if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32) {
        int w = warp_tot[tid], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, wi, o);
            if (tid >= o) wi += v;
        }
        warp_tot[tid] = wi - w;  // exclusive prefix of the warp totals
        if (tid == 31) g.scalars[3] = wi;
    }
    __syncthreads();
    int pos = warp_tot[tid >> 5] + incl - cnt;
    for (int i = lo; i < hi; i++)
        if (g.done[i] == 0) g.live[pos++] = i;
}
int greedy_rebuild_live(cudaStream_t st, const GreedyState &g, int B) {
    WB_ARG(g.live && g.done, "rebuild_live: no live list");
    greedy_rebuild_live_kernel<<<1, 1024, 0, st>>>(g, B);
    WB_LAUNCHED();
    return WB_OK;
}

int greedy_init(cudaStream_t st, const GreedyState &g, int B, const int *prompt) {
    greedy_init_kernel<<<cdiv(B, 256), 256, 0, st>>>(g, B, prompt[0], prompt[1], prompt[2], prompt[3]);
    WB_LAUNCHED();
    return WB_OK;
}

// Scalars are read by every thread before any thread of block 0 rewrites them at the end of the
// NEXT launch; within this launch only thread 0 of block 0 writes them, after computing from the
// values all threads would read -- other threads use only per-row state, so there is no race.
__global__ void greedy_advance_kernel(GreedyState g, int B, int mode, int next_prompt_token,
                                      const int *__restrict__ next) {
    pdl_launch_dependents();
    pdl_wait();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) {
        if (mode == 0) {
            g.cur_tok[i] = next_prompt_token;
        } else {
            int tok = next[i];
            if (!g.done[i]) {
                int n = g.out_len[i];
                if (g.stop_at && n + 1 >= g.stop_at[i]) tok = g.eot;  // forced length (test / bench hook)
                if (n < g.T_out) {
                    g.tokens_out[(size_t)i * g.T_out + n] = tok;
                    g.out_len[i] = n + 1;
                }
                if (tok == g.eot) {
                    g.done[i] = 1;
                    atomicAdd(&g.scalars[2], 1);
                }
            }
            g.cur_tok[i] = tok;
        }
    }
    if (i == 0) {
        int cl = g.scalars[0] + 1;
        g.scalars[0] = cl;
        // prefill tokens sit at positions 0..3; generated tokens use current_len - quirk (whisper.mojo:217)
        g.scalars[1] = (mode == 0) ? cl : cl - g.pos_quirk;
    }
}
// argmax over the logits GEMM's per-tile partials (first index of the row maximum, whisper_tensor.mojo:431-439) fused
// with the greedy bookkeeping of the row: one kernel instead of argmax_partials + greedy_advance.  Warp per row.
__global__ void greedy_argmax_advance_kernel(GreedyState g, int B, const float *__restrict__ part_val,
                                             const int *__restrict__ part_idx, int tiles_n, int *__restrict__ next) {
    pdl_launch_dependents();
    pdl_wait();
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row < B) {
        float best = -INFINITY;
        int bi = 0x7fffffff;
        for (int t = lane; t < tiles_n; t += 32) {
            const float v = part_val[(size_t)row * tiles_n + t];
            const int i = part_idx[(size_t)row * tiles_n + t];
            if (v > best || (v == best && i < bi)) best = v, bi = i;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) best = ov, bi = oi;
        }
        if (lane == 0) {
            int tok = (bi == 0x7fffffff) ? 0 : bi;
            next[row] = tok;
            if (!g.done[row]) {
                const int n = g.out_len[row];
                if (g.stop_at && n + 1 >= g.stop_at[row]) tok = g.eot;  // forced length (test / bench hook)
                if (n < g.T_out) {
                    g.tokens_out[(size_t)row * g.T_out + n] = tok;
                    g.out_len[row] = n + 1;
                }
                if (tok == g.eot) {
                    g.done[row] = 1;
                    atomicAdd(&g.scalars[2], 1);
                }
            }
            g.cur_tok[row] = tok;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {  // no thread of this launch reads the scalars (see greedy_advance_kernel)
        const int cl = g.scalars[0] + 1;
        g.scalars[0] = cl;
        g.scalars[1] = cl - g.pos_quirk;  // generated tokens use current_len - quirk (whisper.mojo:217)
    }
}
int greedy_argmax_advance(cudaStream_t st, const GreedyState &g, int B, const float *part_val, const int *part_idx,
                          int tiles_n, int *next) {
    if (B <= 0) return WB_OK;
    WB_CUDA(launch_pdl(greedy_argmax_advance_kernel, dim3(cdiv(B, 8)), dim3(256), 0, st, g, B, part_val, part_idx, tiles_n, next));
    WB_LAUNCHED();
    return WB_OK;
}

int greedy_advance(cudaStream_t st, const GreedyState &g, int B, int mode, int next_prompt_token, const int *next) {
    WB_CUDA(launch_pdl(greedy_advance_kernel, dim3(cdiv(B, 256)), dim3(256), 0, st, g, B, mode, next_prompt_token, next));
    WB_LAUNCHED();
    return WB_OK;
}

// ---------------------------------------------------------------------------------------------
// Weight conversion
// ---------------------------------------------------------------------------------------------
__global__ void convert_f32_h16_kernel(const float *__restrict__ src, h16 *__restrict__ dst, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = f2h(src[i]);
}
int convert_f32_h16(cudaStream_t st, const float *src, h16 *dst, size_t n) {
    if (!n) return WB_OK;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    convert_f32_h16_kernel<<<blocks, 256, 0, st>>>(src, dst, n);
    WB_LAUNCHED();
    return WB_OK;
}

__global__ void convert_conv_weight_kernel(const float *__restrict__ w, h16 *__restrict__ dst, int C_out,
                                           int C_in, int C_in_pad) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t n = (size_t)C_out * 3 * C_in_pad;
    if (i >= n) return;
    int ci = (int)(i % C_in_pad);
    int k = (int)((i / C_in_pad) % 3);
    int co = (int)(i / ((size_t)3 * C_in_pad));
    float v = ci < C_in ? w[((size_t)co * C_in + ci) * 3 + k] : 0.f;
    dst[i] = f2h(v);
}
int convert_conv_weight(cudaStream_t st, const float *w, h16 *dst, int C_out, int C_in, int C_in_pad) {
    size_t n = (size_t)C_out * 3 * C_in_pad;
    convert_conv_weight_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(w, dst, C_out, C_in, C_in_pad);
    WB_LAUNCHED();
    return WB_OK;
}

// ---------------------------------------------------------------------------------------------
// Folding of the cross-attention projections (see kernels.h); one thread per output element.
// ---------------------------------------------------------------------------------------------
__global__ void fold_qk_kernel(const float *__restrict__ Wq, const float *__restrict__ bq,
                               const float *__restrict__ Wk, int D, int H, h16 *__restrict__ Wqk,
                               float *__restrict__ bqk) {
    const float alpha = 0.125f * 1.4426950408889634f;
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t n = (size_t)H * D * (D + 1);  // last column of each row = bias
    if (idx >= n) return;
    int i = (int)(idx % (D + 1));
    int hc = (int)(idx / (D + 1));
    int h = hc / D, c = hc % D;
    float acc = 0.f;
    for (int d = 0; d < 64; d++) {
        float wk = Wk[(size_t)(h * 64 + d) * D + c];
        acc += wk * (i < D ? Wq[(size_t)(h * 64 + d) * D + i] : bq[h * 64 + d]);
    }
    if (i < D) Wqk[(size_t)hc * D + i] = f2h(acc * alpha);
    else bqk[hc] = acc * alpha;
}
__global__ void fold_ov_kernel(const float *__restrict__ Wv, const float *__restrict__ bv,
                               const float *__restrict__ Wo, const float *__restrict__ bo, int D, int H,
                               h16 *__restrict__ Wov, float *__restrict__ bov) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t n = (size_t)D * (H * D + 1);
    if (idx >= n) return;
    int col = (int)(idx % ((size_t)H * D + 1));
    int row = (int)(idx / ((size_t)H * D + 1));
    if (col < H * D) {
        int h = col / D, c = col % D;
        float acc = 0.f;
        for (int d = 0; d < 64; d++) acc += Wo[(size_t)row * D + h * 64 + d] * Wv[(size_t)(h * 64 + d) * D + c];
        Wov[(size_t)row * H * D + col] = f2h(acc);
    } else {
        float acc = bo[row];
        for (int j = 0; j < D; j++) acc += Wo[(size_t)row * D + j] * bv[j];
        bov[row] = acc;
    }
}
int fold_cross_weights(cudaStream_t st, const float *Wq, const float *bq, const float *Wk, const float *Wv,
                       const float *bv, const float *Wo, const float *bo, int D, int H, h16 *Wqk,
                       float *bqk, h16 *Wov, float *bov) {
    size_t n1 = (size_t)H * D * (D + 1), n2 = (size_t)D * ((size_t)H * D + 1);
    fold_qk_kernel<<<(unsigned)((n1 + 255) / 256), 256, 0, st>>>(Wq, bq, Wk, D, H, Wqk, bqk);
    WB_LAUNCHED();
    fold_ov_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(Wv, bv, Wo, bo, D, H, Wov, bov);
    WB_LAUNCHED();
    return WB_OK;
}

}  // namespace wb
