// ubench_tmem.cu -- microbenchmarks that size the softmax / epilogue warps of the tcgen05 kernels:
//   (1) tcgen05.ld 32x32b.x32 throughput with 1..4 warps of a CTA (one per TMEM lane quarter) and
//       with 1 or 2 CTAs per SM,
//   (2) MUFU.EX2 throughput per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_tmem tools/ubench_tmem.cu
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../whisper_mojo_b200/csrc/sm100.cuh"

using namespace wb;

__global__ void __launch_bounds__(128) ldtm_kernel(int iters, int active_warps, long long *out, float *sink) {
    __shared__ uint32_t holder;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        ptx::tmem_alloc(&holder, 256);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t base = holder;
    float acc = 0.f;
    long long t0 = 0, t1 = 0;
    if (warp < active_warps) {
        const uint32_t addr = base + ((uint32_t)(warp * 32) << 16);
        t0 = clock64();
        for (int i = 0; i < iters; i++) {
            uint32_t v[128];
            ptx::tmem_ld_32x32b_x32(addr, v);
            ptx::tmem_ld_32x32b_x32(addr + 32, v + 32);
            ptx::tmem_ld_32x32b_x32(addr + 64, v + 64);
            ptx::tmem_ld_32x32b_x32(addr + 96, v + 96);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int t = 0; t < 128; t += 16) acc += __uint_as_float(v[t]);
        }
        t1 = clock64();
    }
    __syncthreads();
    if (lane == 0 && warp < active_warps) out[blockIdx.x * 4 + warp] = t1 - t0;
    if (acc == 12345.678f) sink[0] = acc;
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc(base, 256);
}

__global__ void __launch_bounds__(512) mufu_kernel(int iters, long long *out, float *sink) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = -1e-3f * (threadIdx.x + i);
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = ptx::ex2(x[i]) - 1.0f;
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i++) s += x[i];
    if (s == 12345.678f) sink[0] = s;
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

// the exponential phase of the attention softmax in isolation: 128 scores per thread -> ex2(fma) -> row sum -> bf16x2
// pack -> 16 x st.shared.v4; variant 1 drops the pack, variant 2 drops the sum, variant 3 only fma + ex2
template <int VARIANT>
__global__ void __launch_bounds__(256) exp_phase_kernel(int iters, const float *in, long long *out, float *sink) {
    __shared__ uint4 sm[256 * 4];
    float v[128];
#pragma unroll
    for (int t = 0; t < 128; t++) v[t] = in[(threadIdx.x * 128 + t) & 1023];
    float l0 = 0.f, l1 = 0.f;
    unsigned acc = 0;
    const float c = 0.18033688f, nmc = -0.5f;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        uint32_t pk[64];
#pragma unroll
        for (int t = 0; t < 128; t += 2) {
            const float p0 = ptx::ex2(fmaf(v[t], c, nmc));
            const float p1 = ptx::ex2(fmaf(v[t + 1], c, nmc));
            if (VARIANT != 2 && VARIANT != 3) l0 += p0, l1 += p1;
            if (VARIANT == 0 || VARIANT == 2) {
                __nv_bfloat162 b = __floats2bfloat162_rn(p0, p1);
                pk[t >> 1] = *reinterpret_cast<uint32_t *>(&b);
            } else {
                pk[t >> 1] = __float_as_uint(p0) ^ __float_as_uint(p1);
            }
            v[t] = p0 - 1.0f, v[t + 1] = p1 - 1.0f;  // feed back so iterations depend on each other
        }
        if (VARIANT == 0) {
#pragma unroll
            for (int i = 0; i < 4; i++) sm[threadIdx.x * 4 + i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
        }
#pragma unroll
        for (int i = 0; i < 64; i++) acc ^= pk[i];
    }
    long long t1 = clock64();
    if (l0 + l1 == 12345.678f || acc == 0x12345u) sink[0] = l0 + l1 + sm[threadIdx.x].x;
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

// Where do the rows of an M = 64 accumulator live in TMEM?  D[i][n] = i + 1 for a 64 x 16 product; every lane dumps
// column 0.  (Both operands K-major, no swizzle would need other descriptors: the 128B-swizzle atoms of the kernels
// are reused -- A is [64 rows][64 k] with only k = 0 set, B is [16 rows][64 k] with k = 0 = 1.)
__global__ void __launch_bounds__(128) m64_layout_kernel(float *out) {
    __shared__ __align__(1024) uint8_t sA[64 * 128];
    __shared__ __align__(1024) uint8_t sB[16 * 128];
    __shared__ uint64_t bar;
    __shared__ uint32_t holder;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 64 * 128 / 2; i += 128) reinterpret_cast<__nv_bfloat16 *>(sA)[i] = __float2bfloat16(0.f);
    for (int i = threadIdx.x; i < 16 * 128 / 2; i += 128) reinterpret_cast<__nv_bfloat16 *>(sB)[i] = __float2bfloat16(0.f);
    __syncthreads();
    // element (row r, k = 0): 16-byte chunk 0 of row r sits at chunk (0 ^ (r & 7)) under the 128B swizzle
    if (threadIdx.x < 64) reinterpret_cast<__nv_bfloat16 *>(sA + threadIdx.x * 128 + ((0 ^ (threadIdx.x & 7)) << 4))[0] = __float2bfloat16((float)(threadIdx.x + 1));
    if (threadIdx.x < 16) reinterpret_cast<__nv_bfloat16 *>(sB + threadIdx.x * 128 + ((0 ^ (threadIdx.x & 7)) << 4))[0] = __float2bfloat16(1.f);
    if (threadIdx.x == 0) {
        ptx::mbar_init(&bar, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 0) {
        ptx::tmem_alloc(&holder, 32);
        ptx::tmem_relinquish();
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t base = holder;
    // zero the 32 columns first so untouched lanes read 0
    {
        uint32_t z[32];
#pragma unroll
        for (int i = 0; i < 32; i++) z[i] = 0;
        ptx::tmem_st_32x32b_x32(base + ((uint32_t)(warp * 32) << 16), z);
        ptx::tmem_st_wait();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (threadIdx.x == 0) {
        const uint32_t idesc = ptx::umma_idesc_bf16(64, 16, 0, 0);
        ptx::mma_bf16_ss(base, ptx::umma_desc_sw128(ptx::smem_u32(sA), 1, 64), ptx::umma_desc_sw128(ptx::smem_u32(sB), 1, 64),
                         idesc, 0);
        ptx::mma_commit(&bar);
    }
    ptx::mbar_wait(&bar, 0);
    ptx::tc_fence_after();
    uint32_t v[32];
    ptx::tmem_ld_32x32b_x32(base + ((uint32_t)(warp * 32) << 16), v);
    ptx::tmem_ld_wait();
    out[threadIdx.x * 2] = __uint_as_float(v[0]);
    out[threadIdx.x * 2 + 1] = __uint_as_float(v[15]);
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc(base, 32);
}

#include <cuda_bf16.h>

int main() {
    {
        float *o, ho[256];
        cudaMalloc(&o, 256 * 4);
        cudaMemset(o, 0, 256 * 4);
        m64_layout_kernel<<<1, 128>>>(o);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(ho, o, 256 * 4, cudaMemcpyDeviceToHost);
        printf("M=64 accumulator layout (%s): lane -> row+1 (col 0 | col 15)\n", cudaGetErrorString(e));
        for (int l = 0; l < 128; l++) printf("%s%3.0f|%3.0f", (l % 16 == 0) ? "\n  " : " ", ho[2 * l], ho[2 * l + 1]);
        printf("\n");
    }
    long long *out;
    float *sink;
    cudaMalloc(&out, 4096 * sizeof(long long));
    cudaMalloc(&sink, 16);
    long long h[4096];
    const int iters = 2000;
    for (int ctas_per_sm = 1; ctas_per_sm <= 2; ctas_per_sm++)
        for (int w = 1; w <= 4; w++) {
            const int grid = 148 * ctas_per_sm;
            ldtm_kernel<<<grid, 128>>>(iters, w, out, sink);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
                printf("ldtm error %s\n", cudaGetErrorString(e));
                return 1;
            }
            cudaMemcpy(h, out, grid * 4 * sizeof(long long), cudaMemcpyDeviceToHost);
            double clk = (double)h[0] / iters;
            // bytes per warp per iteration: 32 lanes x 128 columns x 4 B = 16 KiB
            printf("ldtm: %d CTA/SM, %d warps/CTA: %.1f clk per 16 KiB/warp -> %.1f B/clk/warp, %.1f B/clk/SM\n",
                   ctas_per_sm, w, clk, 16384.0 / clk, 16384.0 * w * ctas_per_sm / clk);
        }
    for (int threads = 128; threads <= 512; threads *= 2) {
        mufu_kernel<<<148, threads>>>(iters, out, sink);
        cudaDeviceSynchronize();
        cudaMemcpy(h, out, sizeof(long long), cudaMemcpyDeviceToHost);
        double clk = (double)h[0] / iters;
        printf("mufu.ex2 (+fadd): %d threads/SM: %.1f clk per 16 ex2/thread -> %.2f ex2/clk/SM\n", threads, clk,
               16.0 * threads / clk);
    }
    {
        float *in;
        cudaMalloc(&in, 1024 * 4);
        cudaMemset(in, 0, 1024 * 4);
        auto run = [&](auto kern, const char *name) {
            for (int threads : {128, 256}) {
                kern<<<148, threads>>>(200, in, out, sink);
                cudaDeviceSynchronize();
                cudaMemcpy(h, out, sizeof(long long), cudaMemcpyDeviceToHost);
                printf("exp phase [%s], %d warps/SMSP: %.0f clk per 128 scores/thread (MUFU bound %d)\n", name, threads / 128,
                       (double)h[0] / 200, 1024 * threads / 128);
            }
        };
        run(exp_phase_kernel<0>, "fma+ex2+sum+pack+sts");
        run(exp_phase_kernel<1>, "fma+ex2+sum");
        run(exp_phase_kernel<2>, "fma+ex2+pack");
        run(exp_phase_kernel<3>, "fma+ex2");
    }
    return 0;
}
