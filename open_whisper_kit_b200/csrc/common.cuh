// Shared device/host helpers for the sm_100a kernels of the batched transcription path.
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <mutex>

namespace wb {

// ---- error handling -------------------------------------------------------------------------------
// No exception crosses the C ABI (reference contract, include/whisper.h: errors are return codes), so
// CUDA failures are recorded, logged once and surfaced by the caller as the API's failure value.
void cuda_fail(cudaError_t e, const char * expr, const char * file, int line);
bool cuda_failed();            // sticky flag of the calling thread (one API call = one thread = one context)
void cuda_clear_failure();

#define WB_CUDA(expr)                                                        \
    do {                                                                     \
        cudaError_t e__ = (expr);                                            \
        if (e__ != cudaSuccess) ::wb::cuda_fail(e__, #expr, __FILE__, __LINE__); \
    } while (0)

// Launch-site helper: runs `configure` once per device, before the first launch of a kernel on that device (kernel function
// attributes such as the dynamic shared-memory limit are per device, and one process may hold contexts on several GPUs that
// are driven from several threads).  The done-bit is published only after `configure` returned, and late-comers wait on the
// mutex, so no thread can launch ahead of the attribute.
struct DeviceOnce {
    std::atomic<unsigned long long> done{0};
    std::mutex mu;
};
template <typename F> static inline void once_per_device(DeviceOnce & o, F && configure) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    const unsigned long long bit = 1ull << (dev & 63);
    if (o.done.load(std::memory_order_acquire) & bit) return;
    std::lock_guard<std::mutex> lock(o.mu);
    if (o.done.load(std::memory_order_relaxed) & bit) return;
    configure();
    o.done.fetch_or(bit, std::memory_order_release);
}

template <typename T> static inline T ceil_div(T a, T b) { return (a + b - 1) / b; }
template <typename T> static inline T round_up(T a, T b) { return ceil_div(a, b) * b; }

// ---- 16-bit operand type of the tensor path (f16 matches the reference's F16 weight files bit for
// bit; bf16 is the alternative named by the spec).  Selected per context at load time. -----------------
enum class DType : int { F16 = 0, BF16 = 1 };

#ifdef __CUDACC__
// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------------
// The decoder step is ~390 small dependent kernels; back to back they cost ~4 us each just in launch latency.  Kernels
// launched through launch_pdl() may be scheduled while their predecessor is still running: they call pdl_trigger() first
// (lets THEIR successor be scheduled early) and pdl_wait() before touching any memory another kernel may write -- the wait
// returns only when every earlier grid has completed and flushed, so ordering is unchanged; only launch latency, prologue
// work and loads of never-written data (weights, cross K/V) overlap the predecessor's tail.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    WB_CUDA(cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...));
}

template <typename T> struct Half16;
template <> struct Half16<__half> {
    static __device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
    static __device__ __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
    static constexpr int kind = 0;
};
template <> struct Half16<__nv_bfloat16> {
    static __device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
    static __device__ __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
    static constexpr int kind = 1;
};

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

// GELU, tanh form, as ggml computes it (reference ggml/src/ggml-cpu/vec.h:  0.5*x*(1+tanh(sqrt(2/pi)*x*(1+0.044715*x*x))))
// 0.5*x*(1+tanh(u)) == x / (1 + exp(-2u)): two MUFU ops (ex2, rcp) instead of a full tanhf, ~1e-6 relative.
__device__ __forceinline__ float gelu_tanh(float x) {
    const float k = -2.0f * 0.79788456080286535587989211986876f * 1.4426950408889634f;   // -2*sqrt(2/pi)*log2(e)
    const float ka = k * 0.044715f;
    const float arg = x * fmaf(ka, x * x, k);
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(arg));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return x * r;
}
#endif

}  // namespace wb
