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

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>
#include <unordered_map>
#include <vector>

#include "ptx.cuh"

namespace cg = cooperative_groups;

namespace wb {

namespace {

constexpr int TB = 64;                 // output tile columns and k per stage
constexpr int T_THREADS = 192;
// k-blocks in flight per CTA: the loop is bound by load latency (~1.2 us a round trip), so deeper is faster -- but two CTAs
// of 5 x 16 KB still fit the 164 KB shared-memory carve-out; one more stage moves the SM to the 228 KB carve-out, which
// stays in force for the cross-attention kernel that follows and costs it a third of its speed (28 KB of L1 left).
constexpr int T_STAGES = 5;
constexpr int T_CTRL_BYTES = 128 + 512;   // mbarriers + TMEM base address | per-row {mean, rstd} of the LayerNorm-fused A operand
constexpr int T_PITCH = 68;               // floats per row of the f32 tile in shared memory

struct TcSkinnyParams {
    int M, N, K, KS, rows_pad;         // rows_pad: 64 or 128 (accumulator rows carried through the reduction)
    const float * bias;
    float scale; int scale_cols;
    int gelu, ref_f16_gelu;
    int vec_io;                        // bias / residual / outputs are 16-byte (8-byte for 16-bit) addressable per 4 columns
    const float * resid; int ldr;
    void * out16; int ldo16;
    float * out32; int ldo32;
    unsigned long long * trace;        // development aid (WHISPER_B200_TCS_TRACE): 8 time stamps per CTA, or null
    // LayerNorm fusion (GemmArgs): statistics of the output rows for the next GEMM / A operand = LayerNorm(ln_x) built in place
    float2 * ln_part_out;
    const float * ln_x; int ld_lnx;
    const float2 * ln_part_in;
    const float * ln_gamma; const float * ln_beta;
    float ln_eps;
};

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define TS_STAMP(i)                                                                                       \
    do {                                                                                                  \
        if (p.trace && tid == 0) p.trace[((size_t) (blockIdx.y * gridDim.x + blockIdx.x)) * 8 + (i)] = gtime(); \
    } while (0)

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
    // no static shared memory and no alignment slack: the dynamic window starts 1024-byte aligned (checked below)
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nt = blockIdx.x, ks = blockIdx.y;
    const int n0 = nt * TB;
    const int kblocks = p.K / TB;
    const int kb0 = (int) ((long long) kblocks * ks / p.KS), kb1 = (int) ((long long) kblocks * (ks + 1) / p.KS);
    const int nkb = kb1 - kb0;
    const uint32_t x_bytes = (uint32_t) p.rows_pad * 128u, stage_bytes = x_bytes + TB * 128u;
    uint64_t * b_full = reinterpret_cast<uint64_t *>(smem + T_STAGES * stage_bytes);
    uint64_t * b_empty = b_full + T_STAGES;
    uint64_t & b_acc = b_empty[T_STAGES];
    uint32_t & s_tmem = *reinterpret_cast<uint32_t *>(b_empty + T_STAGES + 1);
    static_assert((2 * T_STAGES + 2) * 8 <= 128, "control block");
    float2 * s_stat = reinterpret_cast<float2 *>(reinterpret_cast<uint8_t *>(b_full) + 128);      // [64] rows (LayerNorm-fused A operand)
    const bool ln_in = p.ln_x != nullptr;
    // [rows_pad][T_PITCH] f32; overlays the operand ring, which is dead once the last MMA has completed
    float * tile_sum = reinterpret_cast<float *>(smem);

    // Epilogue work unit = four consecutive columns of one row (float4); this CTA finishes units [v_lo, v_hi) of the tile.
    const int n_vec = p.rows_pad * (TB / 4);
    const int v_lo = p.KS > 1 ? (int) ((long long) n_vec * ks / p.KS) : 0;
    const int v_hi = p.KS > 1 ? (int) ((long long) n_vec * (ks + 1) / p.KS) : n_vec;
    const int U_rt = p.KS <= 2 ? 4 : (p.KS <= 4 ? 2 : 1);         // units a thread keeps in flight (= U of finish<KS>)
    auto load_br = [&](int v, float4 & b, float4 & r) {            // bias and residual of unit v (zeros when absent / outside)
        const int m = v >> 4, n = n0 + ((v & 15) << 2);
        b = r = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (v >= v_hi || m >= p.M || n >= p.N) return;
        if (p.vec_io && n + 3 < p.N) {
            if (p.bias) b = __ldg(reinterpret_cast<const float4 *>(p.bias + n));
            if (p.resid) r = *reinterpret_cast<const float4 *>(p.resid + (size_t) m * p.ldr + n);
        } else {
            float b4[4] = {0.0f, 0.0f, 0.0f, 0.0f}, r4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            for (int i = 0; i < 4; ++i)
                if (n + i < p.N) {
                    if (p.bias) b4[i] = __ldg(p.bias + n + i);
                    if (p.resid) r4[i] = p.resid[(size_t) m * p.ldr + n + i];
                }
            b = make_float4(b4[0], b4[1], b4[2], b4[3]);
            r = make_float4(r4[0], r4[1], r4[2], r4[3]);
        }
    };
    // the first batch's bias / residual are requested as soon as the predecessor grid is complete, under the main loop
    float4 pre_b[4], pre_r[4];
    auto preload = [&]() {
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (u < U_rt) load_br(v_lo + u * T_THREADS + tid, pre_b[u], pre_r[u]);
    };

    TS_STAMP(0);          // CTA start
    if (tid == 0) {
        if (ptx::smem_u32(smem) & 1023u) __trap();
        for (int s = 0; s < T_STAGES; ++s) {
            ptx::mbar_init(&b_full[s], ln_in ? 2 : 1);        // TMA producer (+ the warps that build the normalised A tile)
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
            const uint32_t tx_bytes = ln_in ? TB * 128u : stage_bytes;      // LayerNorm-fused: only the weight half arrives by TMA
            for (int i = 0; i < npre; ++i) {                  // weights first: they do not depend on the predecessor grid
                ptx::mbar_arrive_expect_tx(&b_full[i], tx_bytes);
                ptx::tma_load_2d(smem + i * stage_bytes + x_bytes, &tm_w, &b_full[i], (kb0 + i) * TB, n0);
            }
            pdl_wait();
            for (int i = 0; i < nkb; ++i) {
                const int s = i % T_STAGES;
                if (i >= npre) {
                    ts_wait(&b_empty[s], ((i / T_STAGES) - 1) & 1);
                    ptx::mbar_arrive_expect_tx(&b_full[s], tx_bytes);
                    ptx::tma_load_2d(smem + s * stage_bytes + x_bytes, &tm_w, &b_full[s], (kb0 + i) * TB, n0);
                }
                if (!ln_in) ptx::tma_load_2d(smem + s * stage_bytes, &tm_x, &b_full[s], (kb0 + i) * TB, 0);
            }
        }
        pdl_wait();
        preload();
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
        pdl_wait();
        preload();
    } else {
        // ===== accumulator -> this CTA's f32 tile in shared memory =====
        // (a CTA whose K slice is empty contributes zeros)
        pdl_wait();
        TS_STAMP(1);      // predecessor grid complete
        preload();
        if (ln_in) {
            // ===== A operand = LayerNorm(ln_x), built here instead of by a kernel of its own (rows_pad == 64, checked on the host) =====
            // Everything this needs from the predecessor is requested at once -- the row statistics the producer GEMM left per
            // 64-column tile AND the first k-block of f32 rows -- so the critical path stays ONE round trip to L2, like the TMA
            // load it replaces.  Two threads per row combine the K / 64 partial {mean, centred sum of squares} pairs (Chan et
            // al.; as accurate as a two-pass variance) and merge their halves through a shuffle.
            // Per k-block thread (r8 = tid / 16, c4 = tid % 16) normalises four consecutive columns of rows r8, r8 + 8, ... and
            // stores them as 8 bytes of the 128-byte-swizzled K-major tile the MMA descriptor expects (16-byte chunk index XOR
            // row % 8); the next k-block's rows are in flight meanwhile.
            const int c4 = tid & 15, r8 = tid >> 4;
            const int srow = tid >> 1, shalf = tid & 1, n_part = p.K / TB;          // n_part <= 20 (host check)
            float2 q[10];
#pragma unroll
            for (int t = 0; t < 10; ++t) {
                const int tt = 2 * t + shalf;
                q[t] = (srow < p.M && tt < n_part) ? p.ln_part_in[(size_t) tt * p.M + srow] : make_float2(0.0f, 0.0f);
            }
            // the f32 rows of up to three k-blocks (a CTA's whole K slice in the step's shapes) are requested before anything is
            // consumed: 24 x 16 bytes per thread in flight
            float4 xv[3][8];
            auto load_x = [&](int i0) {
#pragma unroll
                for (int bb = 0; bb < 3; ++bb) {
                    if (i0 + bb >= nkb) break;
                    const float * src = p.ln_x + (size_t) (kb0 + i0 + bb) * TB + 4 * c4;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int r = r8 + 8 * j;
                        xv[bb][j] = r < p.M ? *reinterpret_cast<const float4 *>(src + (size_t) r * p.ld_lnx) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    }
                }
            };
            load_x(0);
            {
                float mean = 0.0f, m2 = 0.0f, cnt = 0.0f;
#pragma unroll
                for (int t = 0; t < 10; ++t) {
                    if (2 * t + shalf < n_part) {
                        const float delta = q[t].x - mean, tot = cnt + (float) TB;
                        mean += delta * ((float) TB / tot);
                        m2 += q[t].y + delta * delta * (cnt * (float) TB / tot);
                        cnt = tot;
                    }
                }
                // merge the two halves of the row in a fixed order (even tiles, then odd tiles)
                const float mean_o = __shfl_xor_sync(0xffffffffu, mean, 1), m2_o = __shfl_xor_sync(0xffffffffu, m2, 1),
                            cnt_o = __shfl_xor_sync(0xffffffffu, cnt, 1);
                const float mA = shalf ? mean_o : mean, qA = shalf ? m2_o : m2, nA = shalf ? cnt_o : cnt;
                const float mB = shalf ? mean : mean_o, qB = shalf ? m2 : m2_o, nB = shalf ? cnt : cnt_o;
                const float tot = nA + nB, delta = mB - mA;
                const float mu = tot > 0.0f ? mA + delta * (nB / tot) : 0.0f;
                const float ss = tot > 0.0f ? qA + qB + delta * delta * (nA * nB / tot) : 0.0f;
                if (shalf == 0) s_stat[srow] = make_float2(mu, tot > 0.0f ? 1.0f / sqrtf(ss / (float) p.K + p.ln_eps) : 0.0f);
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            for (int i0 = 0; i0 < nkb; i0 += 3) {
                if (i0 > 0) load_x(i0);
                float4 gv[3], bv[3];
#pragma unroll
                for (int bb = 0; bb < 3; ++bb) {
                    if (i0 + bb >= nkb) break;
                    const int k = (kb0 + i0 + bb) * TB + 4 * c4;
                    gv[bb] = __ldg(reinterpret_cast<const float4 *>(p.ln_gamma + k));
                    bv[bb] = __ldg(reinterpret_cast<const float4 *>(p.ln_beta + k));
                }
#pragma unroll
                for (int bb = 0; bb < 3; ++bb) {
                    if (i0 + bb >= nkb) break;
                    const int i = i0 + bb, s = i % T_STAGES;
                    if (i >= T_STAGES) ts_wait(&b_empty[s], ((i / T_STAGES) - 1) & 1);
                    uint8_t * xs = smem + s * stage_bytes;
                    const float4 g = gv[bb], be = bv[bb];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int r = r8 + 8 * j;
                        const float2 st = s_stat[r];
                        const float4 x = xv[bb][j];
                        const T16 h[4] = {Half16<T16>::from_f((x.x - st.x) * st.y * g.x + be.x), Half16<T16>::from_f((x.y - st.x) * st.y * g.y + be.y),
                                          Half16<T16>::from_f((x.z - st.x) * st.y * g.z + be.z), Half16<T16>::from_f((x.w - st.x) * st.y * g.w + be.w)};
                        const uint2 pk = r < p.M ? *reinterpret_cast<const uint2 *>(h) : make_uint2(0u, 0u);
                        *reinterpret_cast<uint2 *>(xs + r * 128 + (((c4 >> 1) ^ (r & 7)) << 4) + ((c4 & 1) << 3)) = pk;
                    }
                    ptx::fence_proxy_async_smem();          // generic-proxy stores -> visible to the tensor core's async proxy
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    if (tid == 0) ptx::mbar_arrive(&b_full[s]);
                }
            }
        }
        const int row = warp * 32 + lane;
        if (nkb > 0) {
            ts_wait(&b_acc, 0);
            ptx::tc_fence_after();
        }
        TS_STAMP(2);      // accumulator complete
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
    TS_STAMP(3);          // tile in shared memory

    // ---- cluster-wide reduction and epilogue: every CTA of the cluster (one per K split) finishes a slice of the tile ----
    // All loads of a batch of units -- the KS partial tiles over DSMEM, bias, residual -- are issued before the first use, so a
    // thread pays one round trip per batch, not per element.  A CTA tells its peers that it is done reading their tiles as
    // soon as its last batch is in registers; the wait for the peers' same signal overlaps its arithmetic and stores.
    cg::cluster_group cluster = cg::this_cluster();
    if (p.KS > 1) {
        cluster.sync();
        TS_STAMP(4);      // all K splits of the tile are in shared memory
    }
    T16 * out16 = reinterpret_cast<T16 *>(p.out16);
    auto finish = [&](auto ks_tag) {
        constexpr int KSC = decltype(ks_tag)::value;
        constexpr int U = KSC == 1 ? 4 : (KSC == 2 ? 4 : (KSC <= 4 ? 2 : 1));       // units in flight per thread
        const float * peer[KSC];
#pragma unroll
        for (int r = 0; r < KSC; ++r) peer[r] = KSC > 1 ? cluster.map_shared_rank(tile_sum, r) : tile_sum;
        for (int base = v_lo; base < v_hi; base += U * T_THREADS) {
            float4 part[U][KSC], rs[U], bs[U];
            bool live[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int v = base + u * T_THREADS + tid;
                const int m = v >> 4, n = n0 + ((v & 15) << 2);
                live[u] = v < v_hi && m < p.M && n < p.N;
                if (base == v_lo) {
                    bs[u] = pre_b[u];
                    rs[u] = pre_r[u];
                } else {
                    load_br(v, bs[u], rs[u]);
                }
                if (live[u]) {
                    const int ea = m * T_PITCH + ((v & 15) << 2);
#pragma unroll
                    for (int r = 0; r < KSC; ++r) part[u][r] = *reinterpret_cast<const float4 *>(peer[r] + ea);
                }
            }
            if (KSC > 1 && base + U * T_THREADS >= v_hi)        // last batch is in flight: release the peers' tiles
                asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
            if (KSC > 1 && base + U * T_THREADS >= v_hi) TS_STAMP(7);       // partial tiles of the last batch are in registers
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (!live[u]) continue;
                const int v = base + u * T_THREADS + tid;
                const int m = v >> 4, n = n0 + ((v & 15) << 2);
                float x[4] = {part[u][0].x, part[u][0].y, part[u][0].z, part[u][0].w};
#pragma unroll
                for (int r = 1; r < KSC; ++r) {                                   // fixed rank order
                    x[0] += part[u][r].x; x[1] += part[u][r].y; x[2] += part[u][r].z; x[3] += part[u][r].w;
                }
                const float b4[4] = {bs[u].x, bs[u].y, bs[u].z, bs[u].w}, r4[4] = {rs[u].x, rs[u].y, rs[u].z, rs[u].w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float y = x[i] + b4[i];
                    if (n + i < p.scale_cols) y *= p.scale;
                    if (p.gelu) y = gelu_ts<T16>(y, p.ref_f16_gelu);
                    x[i] = y + r4[i];
                }
                if (p.ln_part_out) {
                    // statistics of this row's 64 output columns for the LayerNorm folded into the next GEMM: the 16 units of a
                    // row sit in 16 consecutive lanes (the host only asks for this when every K split owns whole rows)
                    const unsigned hm = 0xffffu << (lane & 16);
                    float sm = (x[0] + x[1]) + (x[2] + x[3]);
#pragma unroll
                    for (int o = 8; o > 0; o >>= 1) sm += __shfl_xor_sync(hm, sm, o);
                    const float mean = sm * (1.0f / (float) TB);
                    const float d0 = x[0] - mean, d1 = x[1] - mean, d2 = x[2] - mean, d3 = x[3] - mean;
                    float sq = (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
#pragma unroll
                    for (int o = 8; o > 0; o >>= 1) sq += __shfl_xor_sync(hm, sq, o);
                    if ((lane & 15) == 0) p.ln_part_out[(size_t) nt * p.M + m] = make_float2(mean, sq);
                }
                if (p.vec_io && n + 3 < p.N) {
                    if (p.out32) *reinterpret_cast<float4 *>(p.out32 + (size_t) m * p.ldo32 + n) = make_float4(x[0], x[1], x[2], x[3]);
                    if (out16) {
                        const T16 h[4] = {Half16<T16>::from_f(x[0]), Half16<T16>::from_f(x[1]), Half16<T16>::from_f(x[2]), Half16<T16>::from_f(x[3])};
                        *reinterpret_cast<uint2 *>(out16 + (size_t) m * p.ldo16 + n) = *reinterpret_cast<const uint2 *>(h);
                    }
                } else {
                    for (int i = 0; i < 4; ++i)
                        if (n + i < p.N) {
                            if (p.out32) p.out32[(size_t) m * p.ldo32 + n + i] = x[i];
                            if (out16) out16[(size_t) m * p.ldo16 + n + i] = Half16<T16>::from_f(x[i]);
                        }
                }
            }
        }
    };
    switch (p.KS) {
        case 1: finish(std::integral_constant<int, 1>{}); break;
        case 2: finish(std::integral_constant<int, 2>{}); break;
        case 3: finish(std::integral_constant<int, 3>{}); break;
        case 4: finish(std::integral_constant<int, 4>{}); break;
        case 5: finish(std::integral_constant<int, 5>{}); break;
        case 6: finish(std::integral_constant<int, 6>{}); break;
        case 7: finish(std::integral_constant<int, 7>{}); break;
        default: finish(std::integral_constant<int, 8>{}); break;
    }
    TS_STAMP(5);          // outputs written
    if (p.KS > 1) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");     // peers may still be reading this CTA's tile
    TS_STAMP(6);
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

// ---- development trace: WHISPER_B200_TCS_TRACE=<first launch>:<launches> records per-CTA time stamps, printed at exit ----
struct TraceRec { int n_tiles, KS, M, N, K; };
struct TcsTrace {
    bool on = false;
    long first = 0, count = 0, seen = 0;
    unsigned long long * dev = nullptr;
    std::vector<TraceRec> recs;
    static constexpr size_t kPerLaunch = 1024 * 8;       // up to 1024 CTAs x 8 stamps
    TcsTrace() {
        const char * e = getenv("WHISPER_B200_TCS_TRACE");
        if (!e) return;
        first = atol(e);
        const char * c = strchr(e, ':');
        count = c ? atol(c + 1) : 200;
        if (count <= 0) return;
        if (cudaMalloc(&dev, (size_t) count * kPerLaunch * 8) != cudaSuccess) return;
        cudaMemset(dev, 0, (size_t) count * kPerLaunch * 8);
        on = true;
    }
    unsigned long long * slot(const TraceRec & r) {
        if (!on) return nullptr;
        const long i = seen++;
        if (i < first || i >= first + count || (size_t) r.n_tiles * r.KS > 1024) return nullptr;
        recs.push_back(r);
        return dev + (size_t) (recs.size() - 1) * kPerLaunch;
    }
    ~TcsTrace() {
        if (!on || recs.empty()) return;
        cudaDeviceSynchronize();
        std::vector<unsigned long long> h(recs.size() * kPerLaunch);
        if (cudaMemcpy(h.data(), dev, h.size() * 8, cudaMemcpyDeviceToHost) != cudaSuccess) return;
        unsigned long long prev_end = 0;
        fprintf(stderr, "tcs_trace: per launch, ns relative to the first CTA start: [min median max] of each stamp over the CTAs\n");
        fprintf(stderr, "tcs_trace: stamps = s0 start, s1 dep_done, s2 acc_done, s3 tile_smem, s4 cluster_in, s7 partials_loaded, s5 out_written, s6 end\n");
        for (size_t l = 0; l < recs.size(); ++l) {
            const TraceRec & r = recs[l];
            const int n = r.n_tiles * r.KS;
            const unsigned long long * t = h.data() + l * kPerLaunch;
            unsigned long long t0 = ~0ull, t_end = 0;
            for (int c = 0; c < n; ++c) { if (t[c * 8] && t[c * 8] < t0) t0 = t[c * 8]; if (t[c * 8 + 6] > t_end) t_end = t[c * 8 + 6]; }
            fprintf(stderr, "tcs_trace %3zu M=%d N=%d K=%d grid=%dx%d gap_from_prev_end=%lld |", l, r.M, r.N, r.K, r.n_tiles, r.KS,
                    prev_end ? (long long) (t0 - prev_end) : 0ll);
            for (int s = 0; s < 8; ++s) {
                if (r.KS == 1 && (s == 4 || s == 7)) continue;
                std::vector<long long> v;
                for (int c = 0; c < n; ++c) if (t[c * 8 + s]) v.push_back((long long) (t[c * 8 + s] - t0));
                if (v.empty()) continue;
                std::sort(v.begin(), v.end());
                fprintf(stderr, " s%d[%lld %lld %lld]", s, v.front(), v[v.size() / 2], v.back());
            }
            fprintf(stderr, "\n");
            prev_end = t_end;
        }
    }
};

}  // namespace

bool tc_skinny_usable(const GemmArgs & g) {
    const bool base = g.M > 0 && g.M <= 128 && g.K % TB == 0 && g.ldw % 8 == 0 && !g.pos && !(reinterpret_cast<uintptr_t>(g.w) & 15);
    if (!base) return false;
    if (g.ln_x) {       // LayerNorm-fused A operand: 64 rows at most, 16-byte addressable f32 rows and affine parameters
        if (g.M > 64 || g.K > 20 * TB || !g.ln_part_in || !g.ln_gamma || !g.ln_beta || g.ld_lnx % 4 != 0) return false;
        if ((reinterpret_cast<uintptr_t>(g.ln_x) | reinterpret_cast<uintptr_t>(g.ln_gamma) | reinterpret_cast<uintptr_t>(g.ln_beta)) & 15) return false;
    } else if (g.lda % 8 != 0 || (reinterpret_cast<uintptr_t>(g.a) & 15)) {
        return false;
    }
    // statistics out: whole 64-column tiles, and the units of a row must stay inside one half-warp of one K split
    if (g.ln_part_out && (g.N % TB != 0 || !g.out32)) return false;
    return true;
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
    // K splits = cluster size (<= 8 portable).  Enough CTAs to give every SM work, and enough that a CTA's whole K range is
    // in flight at once (T_STAGES k-blocks) while the grid still fits one wave of two CTAs per SM; >= 2 k-blocks per CTA.
    const int rows_pad = g.M <= 64 ? 64 : 128;
    const int cta_per_sm = rows_pad == 64 ? 2 : 1;          // by shared memory: 81 KB / 121 KB per CTA
    int KS = std::max(ceil_div(n_sm, n_tiles), ceil_div(kblocks, T_STAGES));
    KS = std::min(KS, 8);
    while (KS > 1 && (n_tiles * KS > cta_per_sm * n_sm || KS > kblocks / 2)) --KS;
    if (g.ln_part_out)                    // every K split must own whole rows of the tile: a power of two (rows_pad is 64 or 128)
        while (KS & (KS - 1)) --KS;

    static std::mutex mu;
    static std::unordered_map<WKey, TMap, WKeyHash> wmaps;
    TMap tm_x, tm_w;
    if (g.ln_x) memset(&tm_x, 0, sizeof(tm_x));           // never dereferenced: the kernel builds the A tiles itself
    else if (!tc_make_tmap(&tm_x, g.a, g.M, g.K, g.lda, rows_pad, g.dtype)) return false;
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
    p.ln_part_out = g.ln_part_out; p.ln_x = g.ln_x; p.ld_lnx = g.ld_lnx; p.ln_part_in = g.ln_part_in;
    p.ln_gamma = g.ln_gamma; p.ln_beta = g.ln_beta; p.ln_eps = g.ln_eps;
    static TcsTrace trace;
    p.trace = trace.slot({n_tiles, KS, g.M, g.N, g.K});
    auto al = [](const void * q, uintptr_t a) { return (reinterpret_cast<uintptr_t>(q) & (a - 1)) == 0; };
    p.vec_io = al(g.bias, 16) && al(g.resid, 16) && g.ldr % 4 == 0 && al(g.out32, 16) && g.ldo32 % 4 == 0 && al(g.out16, 8) &&
               g.ldo16 % 4 == 0;
    const int smem = T_STAGES * (rows_pad * 128 + TB * 128) + T_CTRL_BYTES;

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
    constexpr int kMaxSmem = T_STAGES * (128 * 128 + TB * 128) + T_CTRL_BYTES;
    if (g.dtype == DType::F16) {
        static DeviceOnce set;      // function attributes are per device
        once_per_device(set, [&] {
            WB_CUDA(cudaFuncSetAttribute(tc_skinny_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        });
        WB_CUDA(cudaLaunchKernelEx(&cfg, tc_skinny_kernel<__half>, tm_x, tm_w, p));
    } else {
        static DeviceOnce set;      // function attributes are per device
        once_per_device(set, [&] {
            WB_CUDA(cudaFuncSetAttribute(tc_skinny_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        });
        WB_CUDA(cudaLaunchKernelEx(&cfg, tc_skinny_kernel<__nv_bfloat16>, tm_x, tm_w, p));
    }
    return !cuda_failed();
}

}  // namespace wb
