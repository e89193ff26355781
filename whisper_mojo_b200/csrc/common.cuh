// common.cuh -- shared helpers for the sm_100a kernels of libwhisper_b200.
#pragma once
#include "dtype.h"
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/whisper_b200.h"

namespace wb {

void set_error(const char *fmt, ...);
extern std::atomic<int64_t> g_launches;

#define WB_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            wb::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));   \
            return WB_ERR_CUDA;                                                                    \
        }                                                                                          \
    } while (0)

#define WB_CHECK(call)                 \
    do {                               \
        int rc__ = (call);             \
        if (rc__ != WB_OK) return rc__; \
    } while (0)

#define WB_ARG(cond, ...)               \
    do {                                \
        if (!(cond)) {                  \
            wb::set_error(__VA_ARGS__); \
            return WB_ERR_ARG;          \
        }                               \
    } while (0)

// Call after every <<<>>> launch: counts the launch and surfaces configuration errors.
#define WB_LAUNCHED()                                                                              \
    do {                                                                                           \
        wb::g_launches.fetch_add(1, std::memory_order_relaxed);                                    \
        cudaError_t e__ = cudaPeekAtLastError();                                                   \
        if (e__ != cudaSuccess) {                                                                  \
            wb::set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return WB_ERR_CUDA;                                                                    \
        }                                                                                          \
    } while (0)

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Opt a kernel into `bytes` of dynamic shared memory on the CURRENT device.  The attribute is per device and per
// kernel, so the cache is keyed by both; it only ever grows (a smaller request never lowers the limit another model
// in the process relies on).  Returns a cudaError_t.
cudaError_t ensure_dyn_smem_impl(const void *func, int bytes);
template <typename K>
static inline cudaError_t ensure_dyn_smem(K kernel, size_t bytes) {
    return ensure_dyn_smem_impl(reinterpret_cast<const void *>(kernel), (int)bytes);
}

// Programmatic dependent launch (option "pdl", default off: measured 432 vs 431 ms of decode, i.e. the captured
// graph already hides launch latency): the kernels of the decode step are launched with programmatic
// stream serialisation, so kernel N+1's CTAs are scheduled -- and run their prologue (barrier init, TMEM
// allocation, tensor-map prefetch) -- while kernel N still executes; every such kernel executes pdl_wait()
// before its first access to memory a predecessor may write.  Without the attribute both instructions are no-ops.
// Per thread, set from the model's own `pdl` field at every model-level entry point (PdlScope): two models driven
// from two host threads never see each other's setting.
extern thread_local bool g_pdl;
struct PdlScope {
    bool saved;
    explicit PdlScope(bool on) : saved(g_pdl) { g_pdl = on; }
    ~PdlScope() { g_pdl = saved; }
};
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr, cfg.numAttrs = g_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- device helpers ------------------------------------------------------------------------

__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// tanh GELU exactly as whisper_tensor.mojo:288-308 writes it (constants truncated as there).
__device__ __forceinline__ float gelu_ref(float x) {
    const float k0 = 0.79788456f, k1 = 0.044715f;
    float x3 = x * x * x;
    float inner = k0 * (x + k1 * x3);
    return 0.5f * x * (1.0f + tanhf(inner));
}

__device__ __forceinline__ void h8_to_float(const uint4 &v, float *f) {
    const h16x2 *p = reinterpret_cast<const h16x2 *>(&v);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float2 t = h22f2(p[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    h16x2 t = f22h2(a, b);
    return *reinterpret_cast<uint32_t *>(&t);
}

}  // namespace wb
