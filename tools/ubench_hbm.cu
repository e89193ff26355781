// ubench_hbm.cu -- read-only HBM streaming bandwidth (what an ideal cross-attention pass could reach):
//   (1) grid-stride 128-bit ld.global.nc loads, (2) per-CTA contiguous 1.15 MB regions streamed with
//   cp.async.bulk into a shared-memory ring (the access pattern of cross_attn_absorbed_kernel).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/ubench_hbm tools/ubench_hbm.cu
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../whisper_mojo_b200/csrc/sm100.cuh"
using namespace wb;

__global__ void __launch_bounds__(512) read_ldg(const uint4 *__restrict__ p, size_t n, unsigned *sink) {
    unsigned acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + i));
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

// each CTA streams `per_cta` bytes (contiguous) through a ring of STAGES x STAGE_BYTES with bulk copies
template <int STAGES, int STAGE_BYTES>
__global__ void __launch_bounds__(64) read_bulk(const uint8_t *__restrict__ p, size_t per_cta, int n_regions, unsigned *sink) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *buf = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    uint64_t *full = reinterpret_cast<uint64_t *>(buf + STAGES * STAGE_BYTES), *empty = full + STAGES;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) ptx::mbar_init(&full[s], 1), ptx::mbar_init(&empty[s], 1);
        ptx::fence_barrier_init();
    }
    __syncthreads();
    const int nblk = (int)(per_cta / STAGE_BYTES);
    if (threadIdx.x == 0) {  // producer
        int g = 0;
        for (int r = blockIdx.x; r < n_regions; r += gridDim.x)
            for (int j = 0; j < nblk; j++, g++) {
                const int s = g % STAGES;
                ptx::mbar_wait(&empty[s], ((g / STAGES) & 1) ^ 1);
                ptx::mbar_expect_tx(&full[s], STAGE_BYTES);
                ptx::bulk_load(buf + s * STAGE_BYTES, p + (size_t)r * per_cta + (size_t)j * STAGE_BYTES, STAGE_BYTES, &full[s]);
            }
    } else if (threadIdx.x == 32) {  // consumer: release immediately
        int g = 0;
        unsigned acc = 0;
        for (int r = blockIdx.x; r < n_regions; r += gridDim.x)
            for (int j = 0; j < nblk; j++, g++) {
                const int s = g % STAGES;
                ptx::mbar_wait(&full[s], (g / STAGES) & 1);
                acc ^= *reinterpret_cast<volatile unsigned *>(buf + s * STAGE_BYTES);
                ptx::mbar_arrive(&empty[s]);
            }
        if (acc == 0x12345678u) sink[0] = acc;
    }
}

// the cross-attention's own pattern: per stage six 3-D tensor loads of [128 keys x 64 channels] boxes (128B swizzle)
// out of a [B][1500][384] bf16 tensor, 2 stages of 96 KB
__global__ void __launch_bounds__(64) read_tensor(const __grid_constant__ CUtensorMap map, int B, int nblk, unsigned *sink) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *buf = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int SB = 98304;
    uint64_t *full = reinterpret_cast<uint64_t *>(buf + 2 * SB), *empty = full + 2;
    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; s++) ptx::mbar_init(&full[s], 1), ptx::mbar_init(&empty[s], 1);
        ptx::fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int g = 0;
        for (int b = blockIdx.x; b < B; b += gridDim.x)
            for (int j = 0; j < nblk; j++, g++) {
                const int s = g & 1;
                ptx::mbar_wait(&empty[s], ((g >> 1) & 1) ^ 1);
                ptx::mbar_expect_tx(&full[s], SB);
                for (int a = 0; a < 6; a++) ptx::tma_load_3d(buf + s * SB + a * 16384, &map, &full[s], a * 64, j * 128, b);
            }
    } else if (threadIdx.x == 32) {
        int g = 0;
        unsigned acc = 0;
        for (int b = blockIdx.x; b < B; b += gridDim.x)
            for (int j = 0; j < nblk; j++, g++) {
                const int s = g & 1;
                ptx::mbar_wait(&full[s], (g >> 1) & 1);
                acc ^= *reinterpret_cast<volatile unsigned *>(buf + s * SB);
                ptx::mbar_arrive(&empty[s]);
            }
        if (acc == 0x12345678u) sink[0] = acc;
    }
}

template <typename F>
static float time_ms(F f, int reps) {
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < reps; i++) f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main() {
    const size_t per = 1500ull * 384 * 2;  // one chunk's enc_out
    const int regions = 2048;
    const size_t bytes = per * regions;
    uint8_t *p;
    unsigned *sink;
    cudaMalloc(&p, bytes), cudaMalloc(&sink, 16);
    cudaMemset(p, 1, bytes);
    for (int g : {148 * 2, 148 * 4, 148 * 8}) {
        float ms = time_ms([&] { read_ldg<<<g, 512>>>((const uint4 *)p, bytes / 16, sink); }, 5);
        printf("ld.global.nc v4, grid %4d x 512: %.3f ms  %.0f GB/s\n", g, ms, bytes / ms / 1e6);
    }
    {
        constexpr int SB = 49152, ST = 4;
        cudaFuncSetAttribute(read_bulk<ST, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, ST * SB + 512);
        // 1152000 bytes per region is not a multiple of 49152: stream 23 blocks (1130496 B) of each region
        float ms = time_ms([&] { read_bulk<ST, SB><<<148, 64, ST * SB + 512>>>(p, per, regions, sink); }, 5);
        printf("bulk ring 4 x 48 KB, 148 CTAs: %.3f ms  %.0f GB/s\n", ms, (double)(per / SB) * SB * regions / ms / 1e6);
    }
    {
        constexpr int SB = 98304, ST = 2;
        cudaFuncSetAttribute(read_bulk<ST, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, ST * SB + 512);
        float ms = time_ms([&] { read_bulk<ST, SB><<<148, 64, ST * SB + 512>>>(p, per, regions, sink); }, 5);
        printf("bulk ring 2 x 96 KB, 148 CTAs: %.3f ms  %.0f GB/s\n", ms, (double)(per / SB) * SB * regions / ms / 1e6);
    }
    {
        constexpr int SB = 16384, ST = 12;
        cudaFuncSetAttribute(read_bulk<ST, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, ST * SB + 512);
        float ms = time_ms([&] { read_bulk<ST, SB><<<148, 64, ST * SB + 512>>>(p, per, regions, sink); }, 5);
        printf("bulk ring 12 x 16 KB, 148 CTAs: %.3f ms  %.0f GB/s\n", ms, (double)(per / SB) * SB * regions / ms / 1e6);
    }
    {
        constexpr int SB = 32768, ST = 3;
        cudaFuncSetAttribute(read_bulk<ST, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, ST * SB + 512);
        float ms = time_ms([&] { read_bulk<ST, SB><<<296, 64, ST * SB + 512>>>(p, per, regions, sink); }, 5);
        printf("bulk ring 3 x 32 KB, 296 CTAs (2/SM): %.3f ms  %.0f GB/s\n", ms, (double)(per / SB) * SB * regions / ms / 1e6);
    }
    {
        typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                     const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void *fp = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
        CUtensorMap map;
        cuuint64_t dims[3] = {384, 1500, (cuuint64_t)regions}, str[2] = {768, 1500 * 768};
        cuuint32_t box[3] = {64, 128, 1}, es[3] = {1, 1, 1};
        for (int promo = 0; promo < 2; promo++) {
            CUresult r = ((EncodeFn)fp)(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, p, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_128B,
                                        promo ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("tensor map encode failed %d\n", (int)r); return 1; }
            cudaFuncSetAttribute(read_tensor, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 98304 + 2048);
            float ms = time_ms([&] { read_tensor<<<148, 64, 2 * 98304 + 2048>>>(map, regions, 12, sink); }, 5);
            printf("tensor loads 6 x [128x64] per 96 KB stage, 2 stages, L2 promotion %s: %.3f ms  %.0f GB/s (valid bytes)\n",
                   promo ? "256B" : "128B", ms, (double)per * regions / ms / 1e6);
        }
    }
    return 0;
}
