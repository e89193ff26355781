// ubench_mma.cu -- cost of the small-N tcgen05.mma instructions the absorbed cross-attention is built from:
//   scores   S[keys x 16]  += enc[keys x 16ch] (A, K-major, M = 128 or 64) x Q'^T (B, K-major, N = 16)
//   context  C[128ch x 16] += enc^T (A, MN-major, M = 128)                 x P^T  (B, K-major, N = 16)
// One thread issues `n` MMAs round-robin over `acc` independent TMEM accumulators (operands in zeroed smem,
// the same descriptors the kernel uses), commits, and the clocks from first issue to the commit's arrival are
// divided by n.  Variants: accumulator count (dependent chains), M, N, A major-ness, and two issuing threads.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/ubench_mma tools/ubench_mma.cu
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../whisper_mojo_b200/csrc/sm100.cuh"

using namespace wb;

struct Args {
    int n, acc, M, N, a_mn, two_threads, a_stride16;  // a_stride16: descriptor advance between consecutive MMAs (16 B units)
};

template <int ACC, int STRIDE16>
__global__ void __launch_bounds__(128) mma_kernel(Args a, long long *out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint32_t holder;
    __shared__ uint64_t bar[2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 96 * 1024 / 16; i += 128) reinterpret_cast<uint4 *>(base)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        ptx::mbar_init(&bar[0], 1), ptx::mbar_init(&bar[1], 1);
        ptx::fence_barrier_init();
    }
    if (warp == 0) {
        ptx::tmem_alloc(&holder, 512);
        ptx::tmem_relinquish();
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tm = holder;
    const int issuers = a.two_threads ? 2 : 1;
    if (warp < issuers && lane == 0) {
        const uint32_t idesc = a.M == 128 ? (a.a_mn ? (a.N == 16 ? ptx::umma_idesc_bf16(128, 16, 1, 0) : ptx::umma_idesc_bf16(128, 64, 1, 0))
                                                     : (a.N == 16 ? ptx::umma_idesc_bf16(128, 16, 0, 0)
                                                                  : (a.N == 64 ? ptx::umma_idesc_bf16(128, 64, 0, 0) : ptx::umma_idesc_bf16(128, 128, 0, 0))))
                                          : ptx::umma_idesc_bf16(64, 16, 0, 0);
        // A: K-major [M x 64] atom (SBO = 1024 B) or MN-major (LBO = one 16 KB atom); B: K-major [N x 64] atom at +64 KB
        const uint64_t a_desc = a.a_mn ? ptx::umma_desc_sw128(ptx::smem_u32(base), 16384 >> 4, 64) : ptx::umma_desc_sw128(ptx::smem_u32(base), 1, 64);
        const uint64_t b_desc = ptx::umma_desc_sw128(ptx::smem_u32(base + 65536 + warp * 8192), 1, 64);
        const long long t0 = clock64();
        const uint32_t d0 = tm + warp * 256, nn = a.N;
        for (int i = 0; i < a.n; i += 24) {  // unrolled: the issue loop itself must not be the bottleneck
#pragma unroll
            for (int k = 0; k < 24; k++)
                ptx::mma_bf16_ss(d0 + (k % ACC) * nn, a_desc + (uint64_t)((k % 16) * STRIDE16), b_desc + 2 * (k & 3), idesc, 1);
        }
        ptx::mma_commit(&bar[warp]);
        const long long t1 = clock64();
        ptx::mbar_wait(&bar[warp], 0);
        const long long t2 = clock64();
        out[warp * 2] = t1 - t0, out[warp * 2 + 1] = t2 - t0;
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc(tm, 512);
}

int main() {
    long long *d, h[4];
    cudaMalloc(&d, 64);
    cudaFuncSetAttribute(mma_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(mma_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(mma_kernel<3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(mma_kernel<6, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(mma_kernel<1, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(mma_kernel<3, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int n = 960;
    struct V { const char *name; Args a; } vs[] = {
        {"scores  M128 N16 K-major A, 1 accumulator ", {n, 1, 128, 16, 0, 0, 2}},
        {"scores  M128 N16 K-major A, 3 accumulators", {n, 3, 128, 16, 0, 0, 2}},
        {"scores  M128 N16 K-major A, 6 accumulators", {n, 6, 128, 16, 0, 0, 2}},
        {"scores  M64  N16 K-major A, 1 accumulator ", {n, 1, 64, 16, 0, 0, 2}},
        {"scores  M64  N16 K-major A, 6 accumulators", {n, 6, 64, 16, 0, 0, 2}},
        {"context M128 N16 MN-major A, 1 accumulator ", {n, 1, 128, 16, 1, 0, 128}},
        {"context M128 N16 MN-major A, 3 accumulators", {n, 3, 128, 16, 1, 0, 128}},
        {"         M128 N64  K-major A, 3 accumulators", {n, 3, 128, 64, 0, 0, 2}},
        {"         M128 N128 K-major A, 2 accumulators", {n, 2, 128, 128, 0, 0, 2}},
        {"scores  M128 N16, 3 acc, TWO issuing threads ", {n, 3, 128, 16, 0, 1, 2}},
        {"context M128 N16 MN, 3 acc, TWO issuing threads", {n, 3, 128, 16, 1, 1, 128}},
    };
    for (auto &v : vs) {
        for (int rep = 0; rep < 2; rep++) {
            if (v.a.a_stride16 == 2) {
                if (v.a.acc == 1) mma_kernel<1, 2><<<1, 128, 100 * 1024>>>(v.a, d);
                else if (v.a.acc == 2) mma_kernel<2, 2><<<1, 128, 100 * 1024>>>(v.a, d);
                else if (v.a.acc == 3) mma_kernel<3, 2><<<1, 128, 100 * 1024>>>(v.a, d);
                else mma_kernel<6, 2><<<1, 128, 100 * 1024>>>(v.a, d);
            } else {
                if (v.a.acc == 1) mma_kernel<1, 128><<<1, 128, 100 * 1024>>>(v.a, d);
                else mma_kernel<3, 128><<<1, 128, 100 * 1024>>>(v.a, d);
            }
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s: %s\n", v.name, cudaGetErrorString(e)); return 1; }
        }
        cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
        printf("%s: issue %.1f clk/MMA, complete %.1f clk/MMA", v.name, (double)h[0] / n, (double)h[1] / n);
        if (v.a.two_threads) printf("  | thread 2: issue %.1f complete %.1f (per thread; both run concurrently)", (double)h[2] / n, (double)h[3] / n);
        printf("\n");
    }
    return 0;
}
