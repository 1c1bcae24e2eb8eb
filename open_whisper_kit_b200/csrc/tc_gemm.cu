// C[M,N] = epilogue( A[M,K] * W[N,K]^T )  on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM),
// operands staged by TMA into 128B-swizzled shared memory through an mbarrier ring.  Both operands are K-major
// ("TN"): activations are [tokens][features] and every Whisper weight is stored [out][in]
// (reference src/whisper.cpp:1775-1793), so no transposes exist anywhere on the path.
//
// This kernel replaces, for every GEMM of the encoder / cross-K/V / prompt pass, the reference's
// ggml_mul_mat + separate bias add / scale / GELU / residual add / cast launches
// (src/whisper.cpp:2006-2014, 2112-2237, 2300-2339; on CUDA: cublasGemmEx + convert + k_bin_bcast + unary kernels,
// ggml/src/ggml-cuda/ggml-cuda.cu:1228-1382).
//
// Structure (one CTA per SM, persistent over output tiles, 384 threads):
//   warp 0      : TMA producer   (one elected lane)       global -> smem ring, kStages deep
//   warp 1      : MMA issuer     (one elected lane)       tcgen05.mma 128 x BN x 16, accumulates over K in TMEM
//   warp 2      : TMEM allocator (2 accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1)
//   warps 4..11 : epilogue       tcgen05.ld -> registers -> bias / scale / GELU / positional add / residual -> global
//                 (warp % 4 selects the TMEM lane quarter, (warp-4)/4 the column half; two warps per SM sub-partition
//                  so MUFU/convert latency of the GELU epilogue hides behind the other warp)
#include "tc_gemm.h"

#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "ptx.cuh"

namespace wb {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;           // 64 x 16-bit = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int kEpiWarp0 = 4;
constexpr int kEpiWarps = 8;        // two per SM sub-partition: lane quarter = warp % 4, column half = (warp-4)/4
constexpr int kThreads = (kEpiWarp0 + kEpiWarps) * 32;

template <int BN> struct Cfg {
    static constexpr int kStages = BN == 256 ? 4 : 6;
    static constexpr int kABytes = BM * BK * 2;
    static constexpr int kBBytes = BN * BK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kTmemCols = 2 * BN;
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct EpiParams {
    int M, N;
    const float * bias;
    float scale;
    int scale_cols;
    int gelu;          // 0 none, 1 tanh-GELU
    int ref_f16_gelu;  // round GELU input/output through f16 like ggml's lookup table
    const float * pos;
    int pos_rows;
    const float * resid;
    int ldr;
    void * out16;
    int ldo16;
    int hm_T;          // head-major remap of out16 (GemmArgs::head_major_T)
    float * out32;
    int ldo32;
    void * vt;         // V^T scratch of the encoder attention (GemmArgs::vt), or null
    int vt_col0, vt_T, vt_TP, vt_H;
};

template <typename T16> __device__ __forceinline__ float gelu_epi(float v, int ref_f16) {
    if (ref_f16) {
        // ggml_vec_gelu_f32 with GGML_GELU_FP16 (reference ggml/src/ggml-cpu/vec.h:996-1009): table lookup on the
        // f16-rounded input, f16 result; identity / zero outside (-10, 10).  Branch-free.
        const float x = __half2float(__float2half_rn(v));
        const float y = __half2float(__float2half_rn(gelu_tanh(x)));
        return v <= -10.0f ? 0.0f : (v >= 10.0f ? v : y);
    }
    return gelu_tanh(v);
}

// The same GELU for two values at once when the result only ever leaves as a 16-bit number (MLP up, conv 1): the f16-rounded
// inputs go through exactly the operations of gelu_tanh() -- x*x, fma, x*t, ex2, 1+e, rcp, x*r, all in f32 -- but as two-wide
// f32 instructions (mul / fma / add .f32x2, sm_100), and the result is NOT rounded to f16 and back here: the store's own
// rounding produces the same 16-bit number.  The reference's range clauses (0 below -10, v above 10) fall out of the formula
// once the result is a 16-bit number: 1 + 2^-126 rounds to 1, so y = x = f16(v) above 10, and x * rcp(2^126) is a zero below -10.
// With the scalar version the epilogue of a K = 1280 GEMM was longer than its main loop (MLP up: 990 against 1 243 TFLOP/s).
__device__ __forceinline__ void gelu_f16_pair(float & v0, float & v1) {
    uint32_t h;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(v1), "f"(v0));
    const float x0 = __half2float(__ushort_as_half((unsigned short) (h & 0xffffu)));
    const float x1 = __half2float(__ushort_as_half((unsigned short) (h >> 16)));
    const float k = -2.0f * 0.79788456080286535587989211986876f * 1.4426950408889634f;   // as gelu_tanh()
    const float ka = k * 0.044715f;
    asm("{\n\t"
        ".reg .b64 x, t, a, kk, kka, one, e, d, r;\n\t"
        ".reg .f32 a0, a1, e0, e1, d0, d1, r0, r1;\n\t"
        "mov.b64 x, {%2, %3};\n\t"
        "mov.b64 kk, {%4, %4};\n\t"
        "mov.b64 kka, {%5, %5};\n\t"
        "mov.b64 one, {0f3F800000, 0f3F800000};\n\t"
        "mul.rn.f32x2 t, x, x;\n\t"
        "fma.rn.f32x2 t, kka, t, kk;\n\t"
        "mul.rn.f32x2 a, x, t;\n\t"
        "mov.b64 {a0, a1}, a;\n\t"
        "ex2.approx.ftz.f32 e0, a0;\n\t"
        "ex2.approx.ftz.f32 e1, a1;\n\t"
        "mov.b64 e, {e0, e1};\n\t"
        "add.rn.f32x2 d, e, one;\n\t"
        "mov.b64 {d0, d1}, d;\n\t"
        "rcp.approx.ftz.f32 r0, d0;\n\t"
        "rcp.approx.ftz.f32 r1, d1;\n\t"
        "mov.b64 r, {r0, r1};\n\t"
        "mul.rn.f32x2 r, x, r;\n\t"
        "mov.b64 {%0, %1}, r;\n\t"
        "}"
        : "=f"(v0), "=f"(v1)
        : "f"(x0), "f"(x1), "f"(k), "f"(ka));
}

template <int BN, typename T16>
__global__ void __launch_bounds__(kThreads, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, int K,
               const EpiParams ep) {
    using C = Cfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    // 128B swizzle needs 1024-byte aligned tiles
    uint8_t * smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t * smem_a = smem;
    uint8_t * smem_b = smem + C::kStages * C::kABytes;
    uint64_t * bars = reinterpret_cast<uint64_t *>(smem + C::kStages * C::kStageBytes);
    uint64_t * full_bar = bars;                       // [kStages]   TMA -> MMA
    uint64_t * empty_bar = bars + C::kStages;         // [kStages]   MMA -> TMA
    uint64_t * tfull_bar = bars + 2 * C::kStages;     // [2]         MMA -> epilogue
    uint64_t * tempty_bar = tfull_bar + 2;            // [2]         epilogue -> MMA
    uint32_t * tmem_slot = reinterpret_cast<uint32_t *>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int tiles_m = (ep.M + BM - 1) / BM;
    const int tiles_n = (ep.N + BN - 1) / BN;
    const int n_tiles = tiles_m * tiles_n;
    const int k_blocks = (K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmap_a);
        ptx::prefetch_tensormap(&tmap_b);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < C::kStages; ++i) {
            ptx::mbar_init(&full_bar[i], 1);
            ptx::mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&tfull_bar[i], 1);
            ptx::mbar_init(&tempty_bar[i], kEpiWarps);   // one arrive per epilogue warp
        }
        ptx::fence_mbar_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(tmem_slot, C::kTmemCols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int m0 = (tile / tiles_n) * BM;
                const int n0 = (tile % tiles_n) * BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    ptx::mbar_arrive_expect_tx(&full_bar[stage], C::kStageBytes);
                    ptx::tma_load_2d(smem_a + stage * C::kABytes, &tmap_a, &full_bar[stage], kb * BK, m0);
                    ptx::tma_load_2d(smem_b + stage * C::kBBytes, &tmap_b, &full_bar[stage], kb * BK, n0);
                    if (++stage == C::kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_f16(Half16<T16>::kind, BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);       // epilogue drained this accumulator
                ptx::tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t) (acc * BN);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint64_t da = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem_a + stage * C::kABytes));
                    const uint64_t db = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem_b + stage * C::kBBytes));
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        // advance 32 bytes (16 elements) along K inside the swizzle atom: +2 in 16-byte units
                        ptx::umma_f16(tmem_d, da + (uint64_t) (2 * k), db + (uint64_t) (2 * k), idesc,
                                      (uint32_t) ((kb | k) != 0));
                    }
                    ptx::umma_commit(&empty_bar[stage]);                // smem slot free once these MMAs retire
                    if (++stage == C::kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                ptx::umma_commit(&tfull_bar[acc]);                       // accumulator complete
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
    } else if (warp >= kEpiWarp0) {
        // ===== epilogue =====
        const int ew = warp & 3;                       // TMEM lane quarter this warp may read (warp % 4)
        const int chalf = (warp - kEpiWarp0) >> 2;     // which half of the tile's columns
        int acc = 0;
        uint32_t acc_phase = 0;
        T16 * out16 = reinterpret_cast<T16 *>(ep.out16);
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int m0 = (tile / tiles_n) * BM;
            const int n0 = (tile % tiles_n) * BN;
            ptx::mbar_wait(&tfull_bar[acc], acc_phase);
            ptx::tc_fence_after();
            const int row = m0 + ew * 32 + lane;
            const bool row_ok = row < ep.M;
            const float * pos_row = ep.pos ? ep.pos + (size_t) (row % ep.pos_rows) * ep.N : nullptr;
            const float * res_row = ep.resid ? ep.resid + (size_t) row * ep.ldr : nullptr;
#pragma unroll 1
            for (int c = chalf * (BN / 64); c < (chalf + 1) * (BN / 64); ++c) {
                const int col0 = n0 + c * 32;
                if (col0 >= ep.N) break;               // warp-uniform
                uint32_t r[32];
                ptx::tmem_ld_32x32(tmem_base + ((uint32_t) (ew * 32) << 16) + (uint32_t) (acc * BN + c * 32), r);
                ptx::tmem_ld_wait();
                if (!row_ok) continue;
                const bool full = col0 + 32 <= ep.N;
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                if (ep.bias) {
                    if (full) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b = __ldg(reinterpret_cast<const float4 *>(ep.bias + col0 + j));
                            v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (col0 + j < ep.N) v[j] += __ldg(ep.bias + col0 + j);
                    }
                }
                if (col0 < ep.scale_cols) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (col0 + j < ep.scale_cols) v[j] *= ep.scale;
                }
                if (ep.gelu) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = gelu_epi<T16>(v[j], ep.ref_f16_gelu);
                }
                if (pos_row) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (col0 + j < ep.N) v[j] += __ldg(pos_row + col0 + j);
                }
                if (res_row) {
                    if (full) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b = *reinterpret_cast<const float4 *>(res_row + col0 + j);
                            v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (col0 + j < ep.N) v[j] += res_row[col0 + j];
                    }
                }
                if (ep.out32) {
                    float * o = ep.out32 + (size_t) row * ep.ldo32 + col0;
                    if (full) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            *reinterpret_cast<float4 *>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (col0 + j < ep.N) o[j] = v[j];
                    }
                }
                if (ep.vt && col0 >= ep.vt_col0) {
                    // V third of the encoder's QKV projection: straight into the attention kernel's V^T layout
                    const int w = row / ep.vt_T, t = row - w * ep.vt_T, c0 = col0 - ep.vt_col0;
                    T16 * o = reinterpret_cast<T16 *>(ep.vt) + ((size_t) (w * ep.vt_H + (c0 >> 6)) * 80 + (c0 & 63)) * ep.vt_TP + t;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (col0 + j < ep.N) o[(size_t) j * ep.vt_TP] = Half16<T16>::from_f(v[j]);
                } else if (out16) {
                    T16 * o = out16 + (size_t) row * ep.ldo16 + col0;
                    if (ep.hm_T > 0) {          // cross-K/V pool: [window][head][K|V][T][64]; a 32-column chunk stays inside one head
                        const int w = row / ep.hm_T, t = row - w * ep.hm_T, half = ep.N >> 1;
                        const int kv = col0 >= half ? 1 : 0, hc = col0 - kv * half;
                        o = out16 + (size_t) w * ep.hm_T * ep.N + ((size_t) ((hc >> 6) * 2 + kv) * ep.hm_T + t) * 64 + (hc & 63);
                    }
                    if (full) {
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            union {
                                uint4 u;
                                T16 h[8];
                            } pk;
#pragma unroll
                            for (int q = 0; q < 8; ++q) pk.h[q] = Half16<T16>::from_f(v[j + q]);
                            *reinterpret_cast<uint4 *>(o + j) = pk.u;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (col0 + j < ep.N) o[j] = Half16<T16>::from_f(v[j]);
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, C::kTmemCols);
    }
}

// ---- 2-CTA variant: one 256 x 256 output tile per pair of SMs (tcgen05 cta_group::2) -------------------------------------------
// The 1-CTA kernel above loads 48 KB of operands per 128 x 256 x 64 block of MMAs (87 FLOP per byte): at K = 1280 it is bound by the
// L2 -> SM path (~950 TFLOP/s).  Here the two CTAs of a cluster own one 256 x 256 tile: each stages ITS 128 rows of A and ITS 128
// rows of the weight tile (32 KB per k-block and CTA, 131 FLOP per byte), the leader CTA (cluster rank 0) issues
// tcgen05.mma.cta_group::2 (M = 256: the tensor cores of both SMs run, reading A from their own SM and the two halves of B from
// both), and each CTA's TMEM receives the accumulator rows of its own 128 A rows.
//   both CTAs, warp 0 : TMA producer; every load reports its bytes to the LEADER's full barrier (which expects 64 KB per stage)
//   leader,    warp 1 : MMA issuer; tcgen05.commit multicast to both CTAs frees the stage / publishes the accumulator in both
//   both CTAs, warps 4-11: epilogue on the own 128 rows, then an arrive on the leader's "accumulator drained" barrier
struct Cfg2 {
    static constexpr int BN = 256;
    static constexpr int kStages = 6;
    static constexpr int kABytes = BM * BK * 2;              // this CTA's 128 rows of A
    static constexpr int kBBytes = (BN / 2) * BK * 2;        // this CTA's 128 rows of the 256-row weight tile
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kTmemCols = 2 * BN;                 // two accumulator stages: all 512 columns
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

__device__ __forceinline__ void g2_wait(uint64_t * bar, uint32_t parity) {       // bounded: a protocol error must trap, not hang
    for (unsigned spins = 0; !ptx::mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 28)) __trap();
}
__device__ __forceinline__ void g2_wait_cluster(uint64_t * bar, uint32_t parity) {      // arrivals come from the other CTA too
    for (unsigned spins = 0;; ++spins) {
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(ok)
            : "r"(ptx::smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return;
        if (spins > (1u << 28)) __trap();
    }
}
__device__ __forceinline__ uint32_t g2_map_to_cta(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
// 2-D tiled load into THIS CTA's shared memory; completion bytes go to the mbarrier at cluster address `bar` (the leader's)
__device__ __forceinline__ void g2_tma_load_2d(void * smem_dst, const void * tmap, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(ptx::smem_u32(smem_dst)), "l"(tmap), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void g2_umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the mbarrier at this offset in BOTH CTAs of the pair once all MMAs issued so far have completed
__device__ __forceinline__ void g2_commit_both(uint64_t * bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(ptx::smem_u32(bar)),
                 "h"((uint16_t) 3)
                 : "memory");
}

template <typename T16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
tc_gemm2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, int K, const EpiParams ep) {
    using C = Cfg2;
    constexpr int BN = C::BN;
    extern __shared__ uint8_t smem_raw[];
    // both CTAs of the pair must see identical offsets (the leader's descriptors address the peer's memory too): the dynamic
    // window starts at the same shared-memory address in every CTA of a launch, so the same round-up applies
    uint8_t * smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t * smem_a = smem;
    uint8_t * smem_b = smem + C::kStages * C::kABytes;
    uint64_t * bars = reinterpret_cast<uint64_t *>(smem + C::kStages * C::kStageBytes);
    uint64_t * full_bar = bars;                       // [kStages]   both producers -> leader's MMA issuer   (leader's copy is used)
    uint64_t * empty_bar = bars + C::kStages;         // [kStages]   leader's MMA issuer -> producer of each CTA
    uint64_t * tfull_bar = bars + 2 * C::kStages;     // [2]         leader's MMA issuer -> epilogue of each CTA
    uint64_t * tempty_bar = tfull_bar + 2;            // [2]         epilogues of both CTAs -> leader's MMA issuer (leader's copy)
    uint32_t * tmem_slot = reinterpret_cast<uint32_t *>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

    const int tiles_m = (ep.M + 2 * BM - 1) / (2 * BM);
    const int tiles_n = ep.N / BN;
    const int n_tiles = tiles_m * tiles_n;
    const int k_blocks = (K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmap_a);
        ptx::prefetch_tensormap(&tmap_b);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < C::kStages; ++i) {
            ptx::mbar_init(&full_bar[i], 1);
            ptx::mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&tfull_bar[i], 1);
            ptx::mbar_init(&tempty_bar[i], 2 * kEpiWarps);   // one arrive per epilogue warp of both CTAs
        }
        ptx::fence_mbar_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(tmem_slot)), "r"(C::kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    ptx::tc_fence_before();
    __syncthreads();
    // the peer's barriers must exist before a load or a commit signals them
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer (both CTAs) =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = pair; tile < n_tiles; tile += n_pairs) {
                const int m0 = (tile / tiles_n) * 2 * BM + (int) rank * BM;
                const int n0 = (tile % tiles_n) * BN + (int) rank * (BN / 2);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    g2_wait(&empty_bar[stage], phase ^ 1);
                    if (rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * C::kStageBytes);
                    const uint32_t bar = g2_map_to_cta(ptx::smem_u32(&full_bar[stage]), 0);
                    g2_tma_load_2d(smem_a + stage * C::kABytes, &tmap_a, bar, kb * BK, m0);
                    g2_tma_load_2d(smem_b + stage * C::kBBytes, &tmap_b, bar, kb * BK, n0);
                    if (++stage == C::kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA only) =====
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_f16(Half16<T16>::kind, 2 * BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = pair; tile < n_tiles; tile += n_pairs) {
                g2_wait_cluster(&tempty_bar[acc], acc_phase ^ 1);       // both epilogues drained this accumulator
                ptx::tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t) (acc * BN);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    g2_wait_cluster(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint64_t da = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem_a + stage * C::kABytes));
                    const uint64_t db = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem_b + stage * C::kBBytes));
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)
                        g2_umma_f16(tmem_d, da + (uint64_t) (2 * k), db + (uint64_t) (2 * k), idesc, (uint32_t) ((kb | k) != 0));
                    g2_commit_both(&empty_bar[stage]);                  // the stage is free in both CTAs once these MMAs retire
                    if (++stage == C::kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                g2_commit_both(&tfull_bar[acc]);                         // accumulator complete in both CTAs
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
    } else if (warp >= kEpiWarp0) {
        // ===== epilogue (both CTAs, own 128 rows) =====
        const int ew = warp & 3;                       // TMEM lane quarter this warp may read (warp % 4)
        const int chalf = (warp - kEpiWarp0) >> 2;     // which half of the tile's columns
        int acc = 0;
        uint32_t acc_phase = 0;
        T16 * out16 = reinterpret_cast<T16 *>(ep.out16);
        const uint32_t tempty_leader = g2_map_to_cta(ptx::smem_u32(&tempty_bar[0]), 0);
        for (int tile = pair; tile < n_tiles; tile += n_pairs) {
            const int m0 = (tile / tiles_n) * 2 * BM + (int) rank * BM;
            const int n0 = (tile % tiles_n) * BN;
            g2_wait(&tfull_bar[acc], acc_phase);
            ptx::tc_fence_after();
            const int row = m0 + ew * 32 + lane;
            const bool row_ok = row < ep.M;
            const float * pos_row = ep.pos ? ep.pos + (size_t) (row % ep.pos_rows) * ep.N : nullptr;
            const float * res_row = ep.resid ? ep.resid + (size_t) row * ep.ldr : nullptr;
#pragma unroll 1
            for (int c = chalf * (BN / 64); c < (chalf + 1) * (BN / 64); ++c) {
                const int col0 = n0 + c * 32;
                uint32_t r[32];
                ptx::tmem_ld_32x32(tmem_base + ((uint32_t) (ew * 32) << 16) + (uint32_t) (acc * BN + c * 32), r);
                ptx::tmem_ld_wait();
                if (!row_ok) continue;
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                if (ep.bias) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b = __ldg(reinterpret_cast<const float4 *>(ep.bias + col0 + j));
                        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
                    }
                }
                if (col0 < ep.scale_cols) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (col0 + j < ep.scale_cols) v[j] *= ep.scale;
                }
                if (ep.gelu) {
                    if (ep.ref_f16_gelu && !ep.out32 && !pos_row && !res_row) {       // 16-bit output only: two-wide, no range clauses
#pragma unroll
                        for (int j = 0; j < 32; j += 2) gelu_f16_pair(v[j], v[j + 1]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = gelu_epi<T16>(v[j], ep.ref_f16_gelu);
                    }
                }
                if (pos_row) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += __ldg(pos_row + col0 + j);
                }
                if (res_row) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b = *reinterpret_cast<const float4 *>(res_row + col0 + j);
                        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
                    }
                }
                if (ep.out32) {
                    float * o = ep.out32 + (size_t) row * ep.ldo32 + col0;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4 *>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                }
                if (ep.vt && col0 >= ep.vt_col0) {
                    // V third of the encoder's QKV projection: straight into the attention kernel's V^T layout
                    const int w = row / ep.vt_T, t = row - w * ep.vt_T, c0 = col0 - ep.vt_col0;
                    T16 * o = reinterpret_cast<T16 *>(ep.vt) + ((size_t) (w * ep.vt_H + (c0 >> 6)) * 80 + (c0 & 63)) * ep.vt_TP + t;
#pragma unroll
                    for (int j = 0; j < 32; ++j) o[(size_t) j * ep.vt_TP] = Half16<T16>::from_f(v[j]);
                } else if (out16) {
                    T16 * o = out16 + (size_t) row * ep.ldo16 + col0;
                    if (ep.hm_T > 0) {          // cross-K/V pool: [window][head][K|V][T][64]; a 32-column chunk stays inside one head
                        const int w = row / ep.hm_T, t = row - w * ep.hm_T, half = ep.N >> 1;
                        const int kv = col0 >= half ? 1 : 0, hc = col0 - kv * half;
                        o = out16 + (size_t) w * ep.hm_T * ep.N + ((size_t) ((hc >> 6) * 2 + kv) * ep.hm_T + t) * 64 + (hc & 63);
                    }
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        union {
                            uint4 u;
                            T16 h[8];
                        } pk;
#pragma unroll
                        for (int q = 0; q < 8; ++q) pk.h[q] = Half16<T16>::from_f(v[j + q]);
                        *reinterpret_cast<uint4 *>(o + j) = pk.u;
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0)
                asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(tempty_leader + (uint32_t) (acc * 8)) : "memory");
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    // neither CTA may leave while the other still signals its barriers or (the leader's MMAs) reads its shared memory
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 2) {
        ptx::tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::kTmemCols) : "memory");
    }
}

// ---- host: tensor maps ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void * p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess) {
            fn = (EncodeTiledFn) p;
        }
    });
    return fn;
}

bool make_tmap(CUtensorMap * tm, const void * base, int rows, int cols, int ld_elems, int box_rows, DType dt) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {(cuuint64_t) cols, (cuuint64_t) rows};
    cuuint64_t strides[1] = {(cuuint64_t) ld_elems * 2};
    cuuint32_t box[2] = {(cuuint32_t) BK, (cuuint32_t) box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, dt == DType::F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                    const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

template <int BN, typename T16> void launch(const GemmArgs & g, const EpiParams & ep, int n_sm, cudaStream_t stream,
                                            bool & ok) {
    using C = Cfg<BN>;
    CUtensorMap ta, tb;
    if (!make_tmap(&ta, g.a, g.M, g.K, g.lda, BM, g.dtype) || !make_tmap(&tb, g.w, g.N, g.K, g.ldw, BN, g.dtype)) {
        ok = false;
        return;
    }
    static DeviceOnce attr_set;      // function attributes are per device
    once_per_device(attr_set, [&] {
        WB_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<BN, T16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     C::kSmemBytes));
    });
    const int tiles = ceil_div(g.M, BM) * ceil_div(g.N, BN);
    const int grid = tiles < n_sm ? tiles : n_sm;
    tc_gemm_kernel<BN, T16><<<grid, kThreads, C::kSmemBytes, stream>>>(ta, tb, g.K, ep);
    WB_CUDA(cudaGetLastError());
}

template <typename T16> void launch2(const GemmArgs & g, const EpiParams & ep, int n_sm, cudaStream_t stream, bool & ok) {
    using C = Cfg2;
    CUtensorMap ta, tb;
    if (!make_tmap(&ta, g.a, g.M, g.K, g.lda, BM, g.dtype) || !make_tmap(&tb, g.w, g.N, g.K, g.ldw, C::BN / 2, g.dtype)) {
        ok = false;
        return;
    }
    static DeviceOnce attr_set;      // function attributes are per device
    once_per_device(attr_set, [&] {
        WB_CUDA(cudaFuncSetAttribute(tc_gemm2_kernel<T16>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    });
    const int tiles = ceil_div(g.M, 2 * BM) * (g.N / C::BN);
    const int pairs = std::min(tiles, n_sm / 2);
    tc_gemm2_kernel<T16><<<2 * pairs, kThreads, C::kSmemBytes, stream>>>(ta, tb, g.K, ep);
    WB_CUDA(cudaGetLastError());
}

}  // namespace

bool tc_make_tmap(TMap * tm, const void * base, int rows, int cols, int ld_elems, int box_rows, DType dt) {
    static_assert(sizeof(TMap) == sizeof(CUtensorMap), "TMap must mirror CUtensorMap");
    return make_tmap(reinterpret_cast<CUtensorMap *>(tm), base, rows, cols, ld_elems, box_rows, dt);
}

bool tc_make_tmap3d(TMap * tm, const void * base, size_t d0, size_t d1, size_t d2, size_t stride1_bytes, size_t stride2_bytes,
                    int box0, int box1, DType dt) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t dims[3] = {(cuuint64_t) d0, (cuuint64_t) d1, (cuuint64_t) d2};
    cuuint64_t strides[2] = {(cuuint64_t) stride1_bytes, (cuuint64_t) stride2_bytes};
    cuuint32_t box[3] = {(cuuint32_t) box0, (cuuint32_t) box1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(reinterpret_cast<CUtensorMap *>(tm), dt == DType::F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                    3, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

bool tc_gemm(const GemmArgs & g, cudaStream_t stream) {
    if (g.M <= 0 || g.N <= 0 || g.K <= 0) return true;
    // TMA needs 16-byte aligned bases and row pitches; the epilogue's vector stores need aligned leading dims.
    if ((g.lda % 8) || (g.ldw % 8) || (reinterpret_cast<uintptr_t>(g.a) & 15) || (reinterpret_cast<uintptr_t>(g.w) & 15))
        return false;
    if (g.out16 && (g.ldo16 % 8)) return false;
    if (g.head_major_T > 0 && (!g.out16 || g.N % 128 || g.M % g.head_major_T)) return false;
    if (g.out32 && (g.ldo32 % 4)) return false;
    if (g.resid && (g.ldr % 4)) return false;
    EpiParams ep;
    ep.M = g.M;
    ep.N = g.N;
    ep.bias = g.bias;
    ep.scale = g.scale;
    ep.scale_cols = g.scale_cols;
    ep.gelu = g.gelu ? 1 : 0;
    ep.ref_f16_gelu = (g.dtype == DType::F16) ? 1 : 0;
    ep.pos = g.pos;
    ep.pos_rows = g.pos_rows > 0 ? g.pos_rows : 1;
    ep.resid = g.resid;
    ep.ldr = g.ldr;
    ep.out16 = g.out16;
    ep.ldo16 = g.ldo16;
    ep.hm_T = g.head_major_T;
    ep.vt = g.vt; ep.vt_col0 = g.vt_col0; ep.vt_T = g.vt_T; ep.vt_TP = g.vt_TP; ep.vt_H = g.vt_H;
    if (g.vt && (g.vt_col0 % 64 || g.vt_T <= 0 || g.vt_TP < g.vt_T || g.vt_H <= 0 || g.M % g.vt_T || (g.N - g.vt_col0) != g.vt_H * 64)) return false;
    ep.out32 = g.out32;
    ep.ldo32 = g.ldo32;
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        WB_CUDA(cudaGetDevice(&dev));
        WB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        if (n_sm <= 0) n_sm = 148;
    }
    bool ok = true;
    const bool wide = (g.N % 256 == 0) || g.N > 1024;
    // 2-CTA tiles (256 x 256 per SM pair): whole 256-column tiles, vector epilogue, enough tiles to give every pair several
    static const bool two_cta_env = !(getenv("WHISPER_B200_GEMM_2CTA") && atoi(getenv("WHISPER_B200_GEMM_2CTA")) == 0);
    const bool vec_ok = !(reinterpret_cast<uintptr_t>(g.bias) & 15) && !(reinterpret_cast<uintptr_t>(g.resid) & 15) &&
                        !(reinterpret_cast<uintptr_t>(g.out32) & 15) && !(reinterpret_cast<uintptr_t>(g.out16) & 15) &&
                        !(reinterpret_cast<uintptr_t>(g.pos) & 15);
    if (two_cta_env && g.N % 256 == 0 && g.M >= 2048 && vec_ok && (!g.pos || g.N % 4 == 0)) {
        if (g.dtype == DType::F16) launch2<__half>(g, ep, n_sm, stream, ok);
        else launch2<__nv_bfloat16>(g, ep, n_sm, stream, ok);
        return ok && !cuda_failed();
    }
    if (g.dtype == DType::F16) {
        if (wide) launch<256, __half>(g, ep, n_sm, stream, ok);
        else launch<128, __half>(g, ep, n_sm, stream, ok);
    } else {
        if (wide) launch<256, __nv_bfloat16>(g, ep, n_sm, stream, ok);
        else launch<128, __nv_bfloat16>(g, ep, n_sm, stream, ok);
    }
    return ok && !cuda_failed();
}

}  // namespace wb
