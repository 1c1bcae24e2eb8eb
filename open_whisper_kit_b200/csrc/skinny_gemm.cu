// Weight-streaming GEMM for the decoder step: out = epilogue(X[M,K] * W[N,K]^T) with M = live sequences (<= 128).
//
// At M <= 128 the decoder's linear layers are pure HBM streaming of the weights (1.6 GB per step for large-v3, reference
// src/whisper.cpp:2525-2799; on CUDA the reference uses mul_mat_vec_f / mul_mat_f here, ggml/src/ggml-cuda/mmvf.cu:8,
// mmf.cuh:50).  A 128 x 256 tcgen05 tile grid has only N/256 = 5..20 CTAs for these shapes, far too few to pull 6.5 TB/s,
// so this kernel trades tensor-core peak for memory-level parallelism:
//   * CTA tile 64 (M) x 64 (N) x 64 (K per stage), grid = N/64 tiles x KS K-splits (>= 2 CTAs per SM in flight),
//   * 6-stage cp.async ring (16-byte coalesced loads of W and X into XOR-swizzled shared memory),
//   * the 4 warps of a CTA split the K slice of a stage, so the X fragments are read once per warp and shared memory
//     traffic stays at 2x the weight bytes; mma.sync m16n8k16, f32 accumulation,
//   * the KS CTAs that split one tile's K range form a thread-block cluster: after a cluster barrier every CTA adds the
//     KS partial tiles for its own slice of the 64x64 outputs straight out of the peers' shared memory (DSMEM), in fixed
//     rank order (bit-reproducible, no atomics, no global scratch), then applies bias / scale / GELU / residual.
#include "skinny_gemm.h"

#include <cooperative_groups.h>

#include <type_traits>

namespace cg = cooperative_groups;

namespace wb {

namespace {

constexpr int SB = 64;            // tile edge (M, N and K-per-stage)
constexpr int S_THREADS = 128;
constexpr int S_STAGES = 6;
constexpr int S_STAGE_BYTES = 2 * SB * SB * 2;     // X tile + W tile
constexpr int S_RED_STRIDE = 72;
constexpr int S_SMEM = S_STAGES * S_STAGE_BYTES;   // 96 KB, reused for the cross-warp reduction (4*64*72*4 = 72 KB)

__device__ __forceinline__ void cp16(void * smem, const void * gmem, bool valid) {
    const uint32_t s = (uint32_t) __cvta_generic_to_shared(smem);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t & r0, uint32_t & r1, uint32_t & r2, uint32_t & r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
template <typename T16> __device__ __forceinline__ void mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <> __device__ __forceinline__ void mma<__half>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <> __device__ __forceinline__ void mma<__nv_bfloat16>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t sw(int row, int chunk) { return (uint32_t) (row * 128 + ((chunk ^ (row & 7)) << 4)); }

struct SkinnyParams {
    int M, N, K, KS;
    const void * x; int ldx;
    const void * w; int ldw;
    const float * bias;
    float scale; int scale_cols;
    int gelu, ref_f16_gelu;
    const float * resid; int ldr;
    void * out16; int ldo16;
    float * out32; int ldo32;
};

template <typename T16> __device__ __forceinline__ float gelu_sk(float v, int ref_f16) {
    if (ref_f16) {
        const float x = __half2float(__float2half_rn(v));
        const float y = __half2float(__float2half_rn(gelu_tanh(x)));
        return v <= -10.0f ? 0.0f : (v >= 10.0f ? v : y);
    }
    return gelu_tanh(v);
}

template <typename T16>
__global__ void __launch_bounds__(S_THREADS, 2)
skinny_gemm_kernel(const SkinnyParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nt = blockIdx.x, ks = blockIdx.y, mb = blockIdx.z;
    const int n0 = nt * SB, m0 = mb * SB;
    const int kblocks = p.K / SB;
    const int kb0 = (int) ((long long) kblocks * ks / p.KS), kb1 = (int) ((long long) kblocks * (ks + 1) / p.KS);
    const int nkb = kb1 - kb0;
    const T16 * X = reinterpret_cast<const T16 *>(p.x);
    const T16 * W = reinterpret_cast<const T16 *>(p.w);

    auto load_x = [&](int stage, int kb) {
        uint8_t * sx = smem + stage * S_STAGE_BYTES;
        const int k0 = kb * SB;
#pragma unroll
        for (int i = 0; i < (SB * 8) / S_THREADS; ++i) {
            const int idx = tid + i * S_THREADS;
            const int r = idx >> 3, c = idx & 7;
            const bool okx = (m0 + r) < p.M;
            cp16(sx + sw(r, c), X + (size_t) (okx ? m0 + r : 0) * p.ldx + k0 + c * 8, okx);
        }
    };
    auto load_w = [&](int stage, int kb) {
        uint8_t * swt = smem + stage * S_STAGE_BYTES + SB * SB * 2;
        const int k0 = kb * SB;
#pragma unroll
        for (int i = 0; i < (SB * 8) / S_THREADS; ++i) {
            const int idx = tid + i * S_THREADS;
            const int r = idx >> 3, c = idx & 7;
            const bool okw = (n0 + r) < p.N;
            cp16(swt + sw(r, c), W + (size_t) (okw ? n0 + r : 0) * p.ldw + k0 + c * 8, okw);
        }
    };

    float acc[4][8][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.0f;

    // Weights are never written by a kernel: their first S-1 tiles are requested BEFORE waiting for the predecessor
    // grid (PDL), the activation tiles after it.  cp.async groups retire in order, so waiting for an X group implies the
    // (older) W group.
    pdl_trigger();
#pragma unroll
    for (int s = 0; s < S_STAGES - 1; ++s)
        if (s < nkb) load_w(s, kb0 + s);
    cp_commit();
    // the bias slice of this tile is a weight too: have it in L2 by the time the epilogue asks for it
    if (p.bias && tid == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.bias + n0));
    if (p.bias && tid == 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.bias + n0 + 32));
    pdl_wait();
#pragma unroll
    for (int s = 0; s < S_STAGES - 1; ++s) {
        if (s < nkb) load_x(s, kb0 + s);
        cp_commit();
    }
    for (int it = 0; it < nkb; ++it) {
        cp_wait<S_STAGES - 2>();
        __syncthreads();
        {   // prefetch the stage that was consumed in the previous iteration
            const int nx = it + S_STAGES - 1;
            if (nx < nkb) {
                load_x(nx % S_STAGES, kb0 + nx);
                load_w(nx % S_STAGES, kb0 + nx);
            }
            cp_commit();
        }
        const uint32_t sx = (uint32_t) __cvta_generic_to_shared(smem + (it % S_STAGES) * S_STAGE_BYTES);
        const uint32_t swt = sx + SB * SB * 2;
        // this warp's 16-wide K slice of the stage: chunks 2*warp, 2*warp+1
        uint32_t a[4][4];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            const int r = mt * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
            ldsm4(sx + sw(r, warp * 2 + (lane >> 4)), a[mt][0], a[mt][1], a[mt][2], a[mt][3]);
        }
#pragma unroll
        for (int np = 0; np < 4; ++np) {
            uint32_t b0, b1, b2, b3;
            const int r = np * 16 + (lane & 7) + 8 * (lane >> 4);
            ldsm4(swt + sw(r, warp * 2 + ((lane >> 3) & 1)), b0, b1, b2, b3);
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) {
                mma<T16>(acc[mt][2 * np], a[mt], b0, b1);
                mma<T16>(acc[mt][2 * np + 1], a[mt], b2, b3);
            }
        }
    }
    cp_wait<0>();
    __syncthreads();

    // cross-warp reduction of the four K-slices through shared memory
    float * red = reinterpret_cast<float *>(smem);
    {
        float * my = red + warp * SB * S_RED_STRIDE;
        const int g = lane >> 2, tq = lane & 3;
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int n8 = 0; n8 < 8; ++n8) {
                *reinterpret_cast<float2 *>(my + (mt * 16 + g) * S_RED_STRIDE + n8 * 8 + 2 * tq) = make_float2(acc[mt][n8][0], acc[mt][n8][1]);
                *reinterpret_cast<float2 *>(my + (mt * 16 + g + 8) * S_RED_STRIDE + n8 * 8 + 2 * tq) = make_float2(acc[mt][n8][2], acc[mt][n8][3]);
            }
    }
    __syncthreads();

    // fold the four warps' K-slices into one 64x64 f32 tile at the start of shared memory (row stride 64)
    float * tile_sum = reinterpret_cast<float *>(smem + 4 * SB * S_RED_STRIDE * sizeof(float));   // 16 KB after the 72 KB
    for (int e = tid; e < SB * SB; e += S_THREADS) {
        const int r = e >> 6, c = e & 63;
        const float * q = red + r * S_RED_STRIDE + c;
        tile_sum[e] = (q[0] + q[SB * S_RED_STRIDE]) + (q[2 * SB * S_RED_STRIDE] + q[3 * SB * S_RED_STRIDE]);
    }
    // every CTA of the cluster (one per K split) finishes a slice of the tile: elements [e_lo, e_hi)
    int e_lo = 0, e_hi = SB * SB;
    cg::cluster_group cluster = cg::this_cluster();
    if (p.KS > 1) {
        cluster.sync();
        e_lo = (int) ((long long) SB * SB * ks / p.KS) & ~3;
        e_hi = ks == p.KS - 1 ? SB * SB : ((int) ((long long) SB * SB * (ks + 1) / p.KS) & ~3);
    } else {
        __syncthreads();
    }
    T16 * out16 = reinterpret_cast<T16 *>(p.out16);
    auto finish = [&](auto ks_tag) {
        constexpr int KSC = decltype(ks_tag)::value;
        const float * peer[KSC];
#pragma unroll
        for (int r = 0; r < KSC; ++r) peer[r] = KSC > 1 ? cluster.map_shared_rank(tile_sum, r) : tile_sum;
        for (int e = e_lo + tid; e < e_hi; e += S_THREADS) {
            float part[KSC];
#pragma unroll
            for (int r = 0; r < KSC; ++r) part[r] = peer[r][e];          // all remote loads in flight together
            float x = part[0];
#pragma unroll
            for (int r = 1; r < KSC; ++r) x += part[r];                   // fixed rank order
            const int m = m0 + (e >> 6), n = n0 + (e & 63);
            if (m >= p.M || n >= p.N) continue;
            if (p.bias) x += __ldg(p.bias + n);
            if (n < p.scale_cols) x *= p.scale;
            if (p.gelu) x = gelu_sk<T16>(x, p.ref_f16_gelu);
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
}

}  // namespace

bool skinny_gemm(const GemmArgs & g, SkinnyWorkspace & wsp, cudaStream_t stream) {
    if (g.M <= 0 || g.N <= 0 || g.K <= 0) return true;
    if (g.K % SB != 0 || (g.lda % 8) || (g.ldw % 8) || g.pos) return false;
    if ((reinterpret_cast<uintptr_t>(g.a) & 15) || (reinterpret_cast<uintptr_t>(g.w) & 15)) return false;
    const int n_tiles = ceil_div(g.N, SB), m_blocks = ceil_div(g.M, SB), kblocks = g.K / SB;
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        WB_CUDA(cudaGetDevice(&dev));
        WB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        if (n_sm <= 0) n_sm = 148;
    }
    // K splits = cluster size (<= 8 portable): enough CTAs to keep >= one wave of SMs streaming, >= 2 k-blocks each
    int want = std::min(ceil_div(n_sm, n_tiles * m_blocks), std::max(1, kblocks / 2));
    int KS = 1;
    while (KS < 8 && KS * 2 <= want) KS *= 2;       // 1, 2, 4 or 8
    (void) wsp;
    SkinnyParams p;
    p.M = g.M; p.N = g.N; p.K = g.K; p.KS = KS;
    p.x = g.a; p.ldx = g.lda; p.w = g.w; p.ldw = g.ldw;
    p.bias = g.bias; p.scale = g.scale; p.scale_cols = g.scale_cols;
    p.gelu = g.gelu ? 1 : 0; p.ref_f16_gelu = g.dtype == DType::F16 ? 1 : 0;
    p.resid = g.resid; p.ldr = g.ldr; p.out16 = g.out16; p.ldo16 = g.ldo16; p.out32 = g.out32; p.ldo32 = g.ldo32;

    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_tiles, KS, m_blocks);
    cfg.blockDim = dim3(S_THREADS);
    cfg.dynamicSmemBytes = S_SMEM;
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
    if (g.dtype == DType::F16) {
        static DeviceOnce set;      // function attributes are per device
        once_per_device(set, [&] {
            WB_CUDA(cudaFuncSetAttribute(skinny_gemm_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, S_SMEM));
        });
        WB_CUDA(cudaLaunchKernelEx(&cfg, skinny_gemm_kernel<__half>, p));
    } else {
        static DeviceOnce set;      // function attributes are per device
        once_per_device(set, [&] {
            WB_CUDA(cudaFuncSetAttribute(skinny_gemm_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, S_SMEM));
        });
        WB_CUDA(cudaLaunchKernelEx(&cfg, skinny_gemm_kernel<__nv_bfloat16>, p));
    }
    return !cuda_failed();
}

SkinnyWorkspace::~SkinnyWorkspace() {
    if (partial) cudaFree(partial);
    if (counters) cudaFree(counters);
}

}  // namespace wb
