// Weight-streaming GEMM for the decoder step on the 5th-generation tensor cores: out = epilogue(X[M,K] * W[N,K]^T), M <= 128.
//
// Same contract, grid and reduction as skinny_gemm.cu -- 64-column output tiles, the K range split over the CTAs of a
// thread-block cluster whose partial tiles are added in rank order through DSMEM -- but the inner product is tcgen05:
// legacy mma.sync tops out near 180 TFLOP/s on this part (measured), which makes the three wide GEMMs of a decoder layer
// (QKV, MLP up, MLP down: 13 MB of weights each) instruction-bound at 14-16 us although their HBM time is 2 us.
//   warp 4 (one lane): TMA producer -- weight tile [64 rows][64 k] and activation tile [M rows][64 k] per k-block into a ring
//                      of stages (128-byte swizzle); the weight tiles of the first stages are requested BEFORE the
//                      programmatic-launch dependency resolves (weights are never written on the device);
//   warp 5 (one lane): MMA issuer   -- 4 x tcgen05.mma 128x64x16 per k-block, f32 accumulator in TMEM (rows >= M of the A
//                      tile are zero-filled by TMA or stale shared memory: they only produce accumulator rows nobody reads);
//   warps 0-3:         tcgen05.ld of their 32 accumulator lanes (= rows) into the CTA's f32 tile in shared memory, then the
//                      cluster-wide reduction and the bias / scale / GELU / residual epilogue exactly as skinny_gemm.cu.
// Replaces the same reference operators (ggml_mul_mat of whisper_build_graph_decoder, src/whisper.cpp:2525-2799).
#include "tc_skinny.h"

#include <cooperative_groups.h>
#include <cuda.h>

#include <mutex>
#include <type_traits>
#include <unordered_map>

#include "ptx.cuh"

namespace cg = cooperative_groups;

namespace wb {

namespace {

constexpr int TB = 64;                 // output tile columns and k per stage
constexpr int T_THREADS = 192;
constexpr int T_STAGES = 4;
constexpr int T_PITCH = 68;               // floats per row of the f32 tile in shared memory

struct TcSkinnyParams {
    int M, N, K, KS, rows_pad;         // rows_pad: 64 or 128 (accumulator rows carried through the reduction)
    const float * bias;
    float scale; int scale_cols;
    int gelu, ref_f16_gelu;
    const float * resid; int ldr;
    void * out16; int ldo16;
    float * out32; int ldo32;
};

__device__ __forceinline__ void ts_wait(uint64_t * bar, uint32_t parity) {        // bounded: trap instead of hanging the GPU
    for (unsigned spins = 0; !ptx::mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 26)) __trap();
}

template <typename T16> __device__ __forceinline__ float gelu_ts(float v, int ref_f16) {
    if (ref_f16) {      // the reference evaluates GELU through an F16 table (ggml/src/ggml-cpu/vec.h:996-1009)
        const float x = __half2float(__float2half_rn(v));
        const float y = __half2float(__float2half_rn(gelu_tanh(x)));
        return v <= -10.0f ? 0.0f : (v >= 10.0f ? v : y);
    }
    return gelu_tanh(v);
}

template <typename T16>
__global__ void __launch_bounds__(T_THREADS, 2)
tc_skinny_kernel(const __grid_constant__ TMap tm_x, const __grid_constant__ TMap tm_w, const TcSkinnyParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t b_full[T_STAGES], b_empty[T_STAGES], b_acc;
    __shared__ uint32_t s_tmem;
    uint8_t * smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nt = blockIdx.x, ks = blockIdx.y;
    const int n0 = nt * TB;
    const int kblocks = p.K / TB;
    const int kb0 = (int) ((long long) kblocks * ks / p.KS), kb1 = (int) ((long long) kblocks * (ks + 1) / p.KS);
    const int nkb = kb1 - kb0;
    const uint32_t x_bytes = (uint32_t) p.rows_pad * 128u, stage_bytes = x_bytes + TB * 128u;
    float * tile_sum = reinterpret_cast<float *>(smem + T_STAGES * stage_bytes);          // [rows_pad][T_PITCH] f32

    if (tid == 0) {
        for (int s = 0; s < T_STAGES; ++s) {
            ptx::mbar_init(&b_full[s], 1);
            ptx::mbar_init(&b_empty[s], 1);
        }
        ptx::mbar_init(&b_acc, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 0) {
        ptx::tmem_alloc(&s_tmem, 64);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = s_tmem;
    pdl_trigger();

    if (warp == 4) {
        if (lane == 0) {
            // ===== TMA producer =====
            ptx::prefetch_tensormap(&tm_x);
            ptx::prefetch_tensormap(&tm_w);
            const int npre = nkb < T_STAGES ? nkb : T_STAGES;
            for (int i = 0; i < npre; ++i) {                  // weights first: they do not depend on the predecessor grid
                ptx::mbar_arrive_expect_tx(&b_full[i], stage_bytes);
                ptx::tma_load_2d(smem + i * stage_bytes + x_bytes, &tm_w, &b_full[i], (kb0 + i) * TB, n0);
            }
            pdl_wait();
            for (int i = 0; i < nkb; ++i) {
                const int s = i % T_STAGES;
                if (i >= npre) {
                    ts_wait(&b_empty[s], ((i / T_STAGES) - 1) & 1);
                    ptx::mbar_arrive_expect_tx(&b_full[s], stage_bytes);
                    ptx::tma_load_2d(smem + s * stage_bytes + x_bytes, &tm_w, &b_full[s], (kb0 + i) * TB, n0);
                }
                ptx::tma_load_2d(smem + s * stage_bytes, &tm_x, &b_full[s], (kb0 + i) * TB, 0);
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            // ===== MMA issuer =====
            const uint32_t idesc = ptx::make_idesc_f16(Half16<T16>::kind, 128, TB);
            for (int i = 0; i < nkb; ++i) {
                const int s = i % T_STAGES;
                ts_wait(&b_full[s], (i / T_STAGES) & 1);
                ptx::tc_fence_after();
                const uint32_t sx = ptx::smem_u32(smem + s * stage_bytes);
                const uint64_t da = ptx::make_sw128_kmajor_desc(sx), db = ptx::make_sw128_kmajor_desc(sx + x_bytes);
#pragma unroll
                for (int k = 0; k < 4; ++k) ptx::umma_f16(tmem, da + (uint64_t) (2 * k), db + (uint64_t) (2 * k), idesc, (uint32_t) (i != 0 || k != 0));
                ptx::umma_commit(&b_empty[s]);
            }
            ptx::umma_commit(&b_acc);
        }
    } else {
        // ===== accumulator -> this CTA's f32 tile in shared memory =====
        // (a CTA whose K slice is empty contributes zeros)
        pdl_wait();
        const int row = warp * 32 + lane;
        if (nkb > 0) {
            ts_wait(&b_acc, 0);
            ptx::tc_fence_after();
        }
        if (warp * 32 < p.rows_pad) {
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                uint32_t r[32];
                if (nkb > 0) {
                    ptx::tmem_ld_32x32(tmem + ((uint32_t) (warp * 32) << 16) + (uint32_t) (c * 32), r);
                    ptx::tmem_ld_wait();
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = 0u;
                }
                // lane = row; the 68-float row pitch keeps the eight lanes of a quarter-warp on distinct banks
                if (row < p.rows_pad) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4 *>(tile_sum + row * T_PITCH + c * 32 + j) =
                            make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                }
            }
        }
        ptx::tc_fence_before();
    }
    __syncthreads();
    pdl_wait();          // every thread takes part in the epilogue below (residual reads, output writes)

    // ---- cluster-wide reduction and epilogue: every CTA of the cluster (one per K split) finishes a slice of the tile ----
    const int n_elem = p.rows_pad * TB;
    int e_lo = 0, e_hi = n_elem;
    cg::cluster_group cluster = cg::this_cluster();
    if (p.KS > 1) {
        cluster.sync();
        e_lo = (int) ((long long) n_elem * ks / p.KS) & ~3;
        e_hi = ks == p.KS - 1 ? n_elem : ((int) ((long long) n_elem * (ks + 1) / p.KS) & ~3);
    }
    T16 * out16 = reinterpret_cast<T16 *>(p.out16);
    auto finish = [&](auto ks_tag) {
        constexpr int KSC = decltype(ks_tag)::value;
        const float * peer[KSC];
#pragma unroll
        for (int r = 0; r < KSC; ++r) peer[r] = KSC > 1 ? cluster.map_shared_rank(tile_sum, r) : tile_sum;
        for (int e = e_lo + tid; e < e_hi; e += T_THREADS) {
            float part[KSC];
            const int ea = (e >> 6) * T_PITCH + (e & 63);
#pragma unroll
            for (int r = 0; r < KSC; ++r) part[r] = peer[r][ea];         // all remote loads in flight together
            float x = part[0];
#pragma unroll
            for (int r = 1; r < KSC; ++r) x += part[r];                   // fixed rank order
            const int m = e >> 6, n = n0 + (e & 63);
            if (m >= p.M || n >= p.N) continue;
            if (p.bias) x += __ldg(p.bias + n);
            if (n < p.scale_cols) x *= p.scale;
            if (p.gelu) x = gelu_ts<T16>(x, p.ref_f16_gelu);
            if (p.resid) x += p.resid[(size_t) m * p.ldr + n];
            if (p.out32) p.out32[(size_t) m * p.ldo32 + n] = x;
            if (out16) out16[(size_t) m * p.ldo16 + n] = Half16<T16>::from_f(x);
        }
    };
    switch (p.KS) {
        case 1: finish(std::integral_constant<int, 1>{}); break;
        case 2: finish(std::integral_constant<int, 2>{}); break;
        case 4: finish(std::integral_constant<int, 4>{}); break;
        default: finish(std::integral_constant<int, 8>{}); break;
    }
    if (p.KS > 1) cluster.sync();     // peers may still be reading this CTA's tile
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem, 64);
    }
}

// weight tensor maps never change for a given (pointer, shape): encode once
struct WKey {
    const void * w; int N, K, ldw, dt;
    bool operator==(const WKey & o) const { return w == o.w && N == o.N && K == o.K && ldw == o.ldw && dt == o.dt; }
};
struct WKeyHash {
    size_t operator()(const WKey & k) const { return std::hash<const void *>()(k.w) ^ ((size_t) k.N * 1315423911u) ^ ((size_t) k.K << 20); }
};

}  // namespace

bool tc_skinny_usable(const GemmArgs & g) {
    return g.M > 0 && g.M <= 128 && g.K % TB == 0 && g.lda % 8 == 0 && g.ldw % 8 == 0 && !g.pos &&
           !(reinterpret_cast<uintptr_t>(g.a) & 15) && !(reinterpret_cast<uintptr_t>(g.w) & 15);
}

bool tc_skinny_gemm(const GemmArgs & g, cudaStream_t stream) {
    if (g.M <= 0 || g.N <= 0 || g.K <= 0) return true;
    if (!tc_skinny_usable(g)) return false;
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        WB_CUDA(cudaGetDevice(&dev));
        WB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        if (n_sm <= 0) n_sm = 148;
    }
    const int n_tiles = ceil_div(g.N, TB), kblocks = g.K / TB;
    // K splits = cluster size (<= 8 portable): enough CTAs to keep one wave of SMs streaming, >= 2 k-blocks each
    int want = std::min(ceil_div(n_sm, n_tiles), std::max(1, kblocks / 2));
    int KS = 1;
    while (KS < 8 && KS * 2 <= want) KS *= 2;

    static std::mutex mu;
    static std::unordered_map<WKey, TMap, WKeyHash> wmaps;
    TMap tm_x, tm_w;
    const int rows_pad = g.M <= 64 ? 64 : 128;
    if (!tc_make_tmap(&tm_x, g.a, g.M, g.K, g.lda, rows_pad, g.dtype)) return false;
    {
        std::lock_guard<std::mutex> lock(mu);
        const WKey key = {g.w, g.N, g.K, g.ldw, (int) g.dtype};
        auto it = wmaps.find(key);
        if (it == wmaps.end()) {
            TMap m;
            if (!tc_make_tmap(&m, g.w, g.N, g.K, g.ldw, TB, g.dtype)) return false;
            it = wmaps.emplace(key, m).first;
        }
        tm_w = it->second;
    }
    TcSkinnyParams p;
    p.M = g.M; p.N = g.N; p.K = g.K; p.KS = KS; p.rows_pad = rows_pad;
    p.bias = g.bias; p.scale = g.scale; p.scale_cols = g.scale_cols;
    p.gelu = g.gelu ? 1 : 0; p.ref_f16_gelu = g.dtype == DType::F16 ? 1 : 0;
    p.resid = g.resid; p.ldr = g.ldr; p.out16 = g.out16; p.ldo16 = g.ldo16; p.out32 = g.out32; p.ldo32 = g.ldo32;
    const int smem = T_STAGES * (rows_pad * 128 + TB * 128) + rows_pad * T_PITCH * 4 + 1024;

    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_tiles, KS, 1);
    cfg.blockDim = dim3(T_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = KS;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    constexpr int kMaxSmem = T_STAGES * (128 * 128 + TB * 128) + 128 * T_PITCH * 4 + 1024;
    if (g.dtype == DType::F16) {
        static bool set = false;
        if (!set) {
            WB_CUDA(cudaFuncSetAttribute(tc_skinny_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
            set = true;
        }
        WB_CUDA(cudaLaunchKernelEx(&cfg, tc_skinny_kernel<__half>, tm_x, tm_w, p));
    } else {
        static bool set = false;
        if (!set) {
            WB_CUDA(cudaFuncSetAttribute(tc_skinny_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
            set = true;
        }
        WB_CUDA(cudaLaunchKernelEx(&cfg, tc_skinny_kernel<__nv_bfloat16>, tm_x, tm_w, p));
    }
    return !cuda_failed();
}

}  // namespace wb
