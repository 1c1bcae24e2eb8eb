// Persistent chain kernel of the single-token decoder step (see dec_chain.h).
//
// Reference operators covered by the phases (src/whisper.cpp, whisper_build_graph_decoder):
//   CP_ROW     token+position embedding 2515-2519; residual adds 2640, 2757, 2797; ggml_norm*w+b 2525-2535, 2646-2656,
//              2762-2772
//   CP_GEMM    every ggml_mul_mat of the layer: Q/K/V 2539-2557, out 2634, cross q 2660, cross out 2751, mlp 2776, 2790
//              (the 32-wide "direct" form also applies bias / dh^-0.25 scale / ggml_gelu 2780-2786 and stores 16-bit)
//   CP_SELF    K/V append 2559-2590 and KQ / soft_max_ext(mask) / KQV 2594-2632
//
// Code size is a first-order concern here: every phase runs once per launch, so its instructions are fetched cold, and
// straight-line unrolled code was measured to cost 3-4x its issue time in instruction-cache misses.  Hence one GEMM routine
// for both tile widths, accumulators small enough to finish from registers (no cross-warp reduction), rolled loops.
#include "dec_chain.h"

#include <stdlib.h>

#include <algorithm>

#include "ptx.cuh"

namespace wb {

namespace {

constexpr int CB = 64;                                 // K per stage; weight tile is 64 or 32 rows
// warps 0-3: accumulator read-back (tcgen05.ld lets warp w read TMEM lanes 32w..32w+31 = token rows), warp 4: TMA
// producer, warp 5: MMA issuer; all six run the row / self-attention phases
constexpr int C_THREADS = 192;
constexpr int C_WARPS = C_THREADS / 32;
// Ring geometry depends on the row count: R <= 64 -> 3 stages of (8 KB activations + 16 KB weights), else 2 stages of
// (16 KB + 16 KB).  The A descriptor always spans 128 rows = 16 KB from the start of the activation tile; with 64 live rows
// the upper half of that span is the start of the stage's own weight tile, i.e. harmless stale rows.
constexpr int C_RING_BYTES = 72 * 1024;
constexpr int C_MAX_STAGES = 6;
// 73 KB incl. alignment slack: two CTAs per SM inside the 164 KB shared-memory carve-out.  A larger ring would push the SM
// to the 228 KB carve-out, and the kernels that run between chain launches keep whatever carve-out they find: the
// cross-attention kernel was measured at 115 us with the 28 KB of L1 that leaves, against 82 us normally.
constexpr int C_SMEM = C_RING_BYTES + 1024;
// Weight tile = 128 rows: issuing one tcgen05.mma costs ~150 cycles whatever its N (measured), so a unit should carry as
// many weight rows as the tile count of the smallest GEMM allows.
constexpr int NT = 128;
constexpr int C_TMEM_COLS = 128;

// ---- grid-wide barrier ---------------------------------------------------------------------------------------------
// All CTAs of the launch are co-resident (cooperative launch), so a monotonic arrival counter is enough: barrier k of
// this launch is passed once the counter reaches base + (k+1)*gridDim.  Wrap-around safe through the signed difference.
// One release (the arrival) and one acquire fence (after the spin) per CTA; the CTA's other threads are ordered through
// the two bar.sync (cumulativity), which is the cheapest pattern the PTX memory model allows (1.4 us at 296 CTAs).
__device__ __noinline__ void grid_barrier(unsigned * bar, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
        unsigned v;
        unsigned long long t0 = 0;
        int spins = 0;
        for (;;) {
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
            if ((int) (v - target) >= 0) break;
            if (++spins == 4096) {
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (t0 == 0) t0 = now;
                else if (now - t0 > 4000000000ull) __trap();     // 4 s: a lost CTA must fail the launch, not hang the GPU
                spins = 0;
            }
        }
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    __syncthreads();
}

// mbarrier wait that cannot hang the GPU: a protocol error must fail the launch (trap), not spin forever
__device__ __forceinline__ void mbar_wait_bounded(uint64_t * bar, uint32_t parity) {
    for (unsigned spins = 0; !ptx::mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 24)) __trap();
}

template <typename T16> __device__ __forceinline__ float round16(float v) { return Half16<T16>::to_f(Half16<T16>::from_f(v)); }

template <typename T16> __device__ __forceinline__ float gelu_ref(float v, int ref_f16) {
    if (ref_f16) {      // the reference evaluates GELU through an F16 table (ggml/src/ggml-cpu/vec.h:996-1009)
        const float x = __half2float(__float2half_rn(v));
        const float y = __half2float(__float2half_rn(gelu_tanh(x)));
        return v <= -10.0f ? 0.0f : (v >= 10.0f ? v : y);
    }
    return gelu_tanh(v);
}

template <typename T16> __device__ __forceinline__ void store4_16(T16 * dst, float a, float b, float c, float d) {
    union { T16 h[4]; uint2 u; } pk;
    pk.h[0] = Half16<T16>::from_f(a); pk.h[1] = Half16<T16>::from_f(b);
    pk.h[2] = Half16<T16>::from_f(c); pk.h[3] = Half16<T16>::from_f(d);
    *reinterpret_cast<uint2 *>(dst) = pk.u;
}

// ---- GEMM phase --------------------------------------------------------------------------------------------------------
// out[R <= 128][N] = act[R][K] * W[N][K]^T on the 5th-generation tensor cores: legacy mma.sync tops out near 180 TFLOP/s
// on this part (measured: the step's 102 GFLOP of 64-row GEMMs cost 0.5 ms that way, twice their HBM time).  A (virtual)
// CTA walks a contiguous range of units; a unit is one 64-deep k-block of one weight tile:
//   A operand = activation tile, 128 token rows x 64 k (M = 128; rows >= R are zero-filled by TMA or stale -- they only
//               produce accumulator rows nobody reads), B operand = weight tile, nt x 64 k, both K-major, 128-byte swizzle;
//   warp 4 (one lane) feeds a ring of stages with two TMA loads per unit, completion on the stage's `full` mbarrier;
//   warp 5 (one lane) issues 4 tcgen05.mma (K = 16 each) per unit into a 128 x nt f32 accumulator in TMEM; tcgen05.commit
//   hands the stage back (`empty`) and, at the end of an output tile, publishes the accumulator (`acc_full`);
//   warps 0-3 read their 32 accumulator lanes back, finish the tile and release TMEM (`acc_empty`).
// nt = 64: stream-K, every CTA emits one f32 partial tile per output tile its range touches.  nt = 32: "direct" -- the
// range is exactly one output tile over the full K, so bias / scale / GELU and the 16-bit store happen here.
struct GemmRegs {
    int direct, kpt, U, G;
};
__device__ __forceinline__ void range_of(const GemmRegs & g, int vc, int & u0, int & nu) {
    u0 = 0; nu = 0;
    if (vc < g.G) {
        u0 = (int) ((unsigned) g.U * (unsigned) vc / (unsigned) g.G);
        nu = (int) ((unsigned) g.U * (unsigned) (vc + 1) / (unsigned) g.G) - u0;
    }
}

__device__ __forceinline__ float block_sum(float v, float * s_red) {       // C_WARPS warps; safe to call back to back
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.0f;
#pragma unroll
    for (int w = 0; w < C_WARPS; ++w) r += s_red[w];
    return r;
}

// ---- residual / LayerNorm phase: one CTA per token row ----------------------------------------------------------------
template <typename T16>
__device__ __forceinline__ void row_phase(const ChainCommon & p, const ChainPhase & ph, float * s_red, float4 * rowbuf) {
    const int r = blockIdx.x;
    if (r >= p.R) return;
    const int tid = threadIdx.x;
    const int d = p.d, nq = d >> 2;
    constexpr int NB = 8;                     // partial tiles in flight per quad
    float s = 0.0f;
    DecRow row = {};
    if (ph.embed) row = p.rows[r];
#pragma unroll 1
    for (int q = tid; q < nq; q += C_THREADS) {
        const int c = q * 4;
        float4 v;
        if (ph.embed) {
            const T16 * t = reinterpret_cast<const T16 *>(p.te) + (size_t) row.token * d + c;
            const float4 pe = __ldg(reinterpret_cast<const float4 *>(p.pe + (size_t) row.pos * d + c));
            v = make_float4(Half16<T16>::to_f(t[0]) + pe.x, Half16<T16>::to_f(t[1]) + pe.y, Half16<T16>::to_f(t[2]) + pe.z,
                            Half16<T16>::to_f(t[3]) + pe.w);
        } else {
            // x + (partial tiles in contributor order + bias); NB independent loads in flight
            v = __ldcg(reinterpret_cast<const float4 *>(p.x + (size_t) r * d + c));
            const int ot = c / SG_TILE_COLS;
            const int cnt = sg_cta_of(ph.g, ot * ph.g.kpt + ph.g.kpt - 1) - sg_cta_of(ph.g, ot * ph.g.kpt) + 1;
            const float * src = ph.part + ((size_t) ot * ph.g.maxc) * SG_TILE_FLOATS + r * SG_TILE_COLS + (c % SG_TILE_COLS);
            float4 a = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll 1
            for (int j0 = 0; j0 < cnt; j0 += NB) {
                float4 t[NB];
#pragma unroll
                for (int u = 0; u < NB; ++u)
                    t[u] = j0 + u < cnt ? __ldcg(reinterpret_cast<const float4 *>(src + (size_t) (j0 + u) * SG_TILE_FLOATS)) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
                for (int u = 0; u < NB; ++u) { a.x += t[u].x; a.y += t[u].y; a.z += t[u].z; a.w += t[u].w; }
            }
            if (ph.bias) {
                const float4 b = __ldg(reinterpret_cast<const float4 *>(ph.bias + c));
                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            }
            v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
        }
        __stcg(reinterpret_cast<float4 *>(p.x + (size_t) r * d + c), v);
        rowbuf[q] = v;
        s += (v.x + v.y) + (v.z + v.w);
    }
    if (!ph.ln_w) return;
    const float mean = block_sum(s, s_red) / (float) d;
    float qq = 0.0f;
#pragma unroll 1
    for (int q = tid; q < nq; q += C_THREADS) {           // every thread re-reads only what it wrote
        const float4 v = rowbuf[q];
        const float a = v.x - mean, b = v.y - mean, c = v.z - mean, e = v.w - mean;
        qq += (a * a + b * b) + (c * c + e * e);
    }
    const float var = block_sum(qq, s_red) / (float) d;
    const float rstd = 1.0f / sqrtf(var + p.eps);
    T16 * out = reinterpret_cast<T16 *>(ph.out16) + (size_t) r * ph.ldo16;
#pragma unroll 1
    for (int q = tid; q < nq; q += C_THREADS) {
        const float4 v = rowbuf[q];
        const float4 gw = __ldg(reinterpret_cast<const float4 *>(ph.ln_w + q * 4));
        const float4 gb = __ldg(reinterpret_cast<const float4 *>(ph.ln_b + q * 4));
        store4_16<T16>(out + q * 4, (v.x - mean) * rstd * gw.x + gb.x, (v.y - mean) * rstd * gw.y + gb.y,
                       (v.z - mean) * rstd * gw.z + gb.z, (v.w - mean) * rstd * gw.w + gb.w);
    }
}

// ---- masked self-attention phase: one warp per (row, head) --------------------------------------------------------------
// 8 lanes x 8 values cover one 64-wide head row, so a warp-wide 16-byte load reads four whole 128-byte key rows; U such
// loads are issued before any is consumed.  q | k | v of the new token come from the QKV phase (16-bit, bias and the
// dh^-0.25 scale of Q and K already applied, src/whisper.cpp:2506, 2550-2557); k, v are appended to the cache by this warp
// (every sequence owns exactly one row) and enter the attention straight from registers.
template <typename T16>
__device__ __forceinline__ void self_phase(const ChainCommon & p, const ChainPhase & ph, float * s_sc_all) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = lane & 7, grp = lane >> 3;
    float * sc = s_sc_all + warp * (p.n_ctx + 4);
    const int d = p.d, ld = 2 * d;
    constexpr int U = 16;          // 64 keys per batch: the phase is a chain of L2/HBM round trips, so make each one count
    const int items = p.R * p.H;
    const T16 * qkv = reinterpret_cast<const T16 *>(ph.a);
#pragma unroll 1
    for (int item = blockIdx.x * C_WARPS + warp; item < items; item += gridDim.x * C_WARPS) {
        const int r = item / p.H, h = item - r * p.H;
        const DecRow row = p.rows[r];
        const int T = row.pos;                          // keys already in the cache
        const T16 * src = qkv + (size_t) r * ph.lda + h * 64 + sub * 8;
        const uint4 uq = __ldcg(reinterpret_cast<const uint4 *>(src));
        const uint4 uk = __ldcg(reinterpret_cast<const uint4 *>(src + d));
        const uint4 uv = __ldcg(reinterpret_cast<const uint4 *>(src + 2 * d));
        T16 * cache = reinterpret_cast<T16 *>(row.self_kv) + ph.layer_off + h * 64 + sub * 8;
        float qv[8];
        {
            const T16 * eq = reinterpret_cast<const T16 *>(&uq);
#pragma unroll
            for (int j = 0; j < 8; ++j) qv[j] = Half16<T16>::to_f(eq[j]);
        }
        if (grp < 2)                                     // grp 0 appends K, grp 1 appends V: 8 lanes x 16 bytes each
            *reinterpret_cast<uint4 *>(cache + (size_t) T * ld + grp * d) = grp == 0 ? uk : uv;
        float mx = -INFINITY;
#pragma unroll 1
        for (int base = 0; base <= T; base += 4 * U) {   // key T (the token's own) comes from registers
            uint4 kb[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int t = base + grp + 4 * u;
                kb[u] = t < T ? __ldcg(reinterpret_cast<const uint4 *>(cache + (size_t) t * ld)) : (t == T ? uk : make_uint4(0, 0, 0, 0));
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int t = base + grp + 4 * u;
                const T16 * e = reinterpret_cast<const T16 *>(&kb[u]);
                float acc = 0.0f;
#pragma unroll
                for (int j = 0; j < 8; ++j) acc = fmaf(qv[j], Half16<T16>::to_f(e[j]), acc);
                acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                acc += __shfl_xor_sync(0xffffffffu, acc, 4);
                if (t <= T) {
                    if (sub == 0) sc[t] = acc;
                    mx = fmaxf(mx, acc);
                }
            }
        }
        mx = warp_max(mx);
        __syncwarp();
        float sum = 0.0f;
        for (int t = lane; t <= T; t += 32) {
            const float e = expf(sc[t] - mx);
            sc[t] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        __syncwarp();
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = 0.0f;
        const T16 * vcache = cache + d;
#pragma unroll 1
        for (int base = 0; base <= T; base += 4 * U) {
            uint4 vb[U];
            float pr[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int t = base + grp + 4 * u;
                vb[u] = t < T ? __ldcg(reinterpret_cast<const uint4 *>(vcache + (size_t) t * ld)) : (t == T ? uv : make_uint4(0, 0, 0, 0));
                // softmax weights are rounded to 16 bits before the V product, as the reference's F16 KQV matmul does
                pr[u] = t <= T ? round16<T16>(sc[t] * inv) : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const T16 * e = reinterpret_cast<const T16 *>(&vb[u]);
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = fmaf(pr[u], Half16<T16>::to_f(e[j]), o[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            o[j] += __shfl_xor_sync(0xffffffffu, o[j], 8);
            o[j] += __shfl_xor_sync(0xffffffffu, o[j], 16);
        }
        if (grp == 0) {
            union { T16 hh[8]; uint4 u; } pk;
#pragma unroll
            for (int j = 0; j < 8; ++j) pk.hh[j] = Half16<T16>::from_f(o[j]);
            *reinterpret_cast<uint4 *>(reinterpret_cast<T16 *>(ph.out16) + (size_t) r * ph.ldo16 + h * 64 + sub * 8) = pk.u;
        }
        __syncwarp();       // sc is reused by the warp's next item
    }
}

template <typename T16>
__global__ void __launch_bounds__(C_THREADS, 2)
dec_chain_kernel(const __grid_constant__ ChainParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ float s_red[C_WARPS];
    __shared__ __align__(8) uint64_t s_full[C_MAX_STAGES];   // TMA -> MMA: both tiles of a stage have landed
    __shared__ __align__(8) uint64_t s_empty[C_MAX_STAGES];  // MMA -> TMA: the MMAs reading a stage have retired
    __shared__ __align__(8) uint64_t s_acc_full, s_acc_empty;   // MMA -> read-back: tile complete; read-back -> MMA: TMEM drained
    __shared__ uint32_t s_tmem;
    // 128-byte swizzle atoms are 1024 bytes: align the ring
    uint8_t * smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int R = p.c.R;
    auto stamp = [&](int slot) {
        if (p.trace && blockIdx.x == 0 && threadIdx.x == 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            p.trace[slot] = now;
        }
    };
    stamp(15);
    pdl_trigger();
    if (threadIdx.x == 0) {
        for (int s = 0; s < C_MAX_STAGES; ++s) {
            ptx::mbar_init(&s_full[s], 1);
            ptx::mbar_init(&s_empty[s], 1);
        }
        ptx::mbar_init(&s_acc_full, 1);
        ptx::mbar_init(&s_acc_empty, 4);
        ptx::fence_mbar_init();
    }
    if (warp == 0) {
        ptx::tmem_alloc(&s_tmem, C_TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = s_tmem;

    const unsigned S = R <= 64 ? 3u : 2u;
    const uint32_t x_bytes = R <= 64 ? 8192u : 16384u, stage_bytes = x_bytes + NT * 128u;
    auto stage_ptr = [&](unsigned gc) { return smem + (gc % S) * stage_bytes; };
    const bool is_producer = warp == 4 && lane == 0, is_mma = warp == 5 && lane == 0;

    // Every thread tracks the same counters: units (gcount) and output tiles (tcount) this CTA has been through; unit gc
    // lives in ring stage gc % S, and the mbarrier parities follow from the counters alone.
    unsigned gcount = 0, tcount = 0;
    int pf = -1, u0 = 0, nu = 0, npre = 0;      // GEMM phase whose first weight tiles are already in flight, its range
    GemmRegs g = {};
    unsigned target = p.bar_base;
#pragma unroll 1
    for (int i = 0; i < p.n_phase; ++i) {
        // Weights are never written on the device: the weight tiles of the next GEMM phase's first stages are requested
        // before the barrier (or grid dependency) that guards its activations.  Every MMA of earlier phases has retired
        // (their tiles were read back), so the stages are free; one stage is left alone as scratch for the other phases.
        if (pf < i) {
            pf = p.n_phase;
            for (int j = i; j < p.n_phase; ++j)
                if (p.ph[j].type == CP_GEMM) { pf = j; break; }
            if (pf < p.n_phase) {
                const ChainPhase & q = p.ph[pf];
                g.direct = q.direct; g.kpt = q.g.kpt; g.U = q.g.U; g.G = q.g.G;
                range_of(g, blockIdx.x, u0, nu);
                npre = min(nu, (int) S - 1);
                if (is_producer) {
                    const TMap * tw = &p.tm[2 * q.tm + 1];
                    const uint32_t tx = x_bytes + NT * 128u;
                    for (int s = 0; s < npre; ++s) {
                        const unsigned gc = gcount + s;
                        const int u = u0 + s, tile = u / g.kpt, kb = u - tile * g.kpt;
                        if (gc >= S) mbar_wait_bounded(&s_empty[gc % S], ((gc / S) - 1u) & 1u);
                        ptx::mbar_arrive_expect_tx(&s_full[gc % S], tx);
                        ptx::tma_load_2d(stage_ptr(gc) + x_bytes, tw, &s_full[gc % S], kb * CB, tile * NT);
                    }
                }
            }
        }
        if (i == 0) {
            pdl_wait();
        } else {
            target += gridDim.x;
            grid_barrier(p.bar, target);
        }
        stamp(i);
        const ChainPhase & ph = p.ph[i];
        if (ph.type == CP_ROW || ph.type == CP_SELF) {
            // a stage the prefetch above never targets is scratch for these phases
            uint8_t * scratch = stage_ptr(gcount + S - 1);
            if (ph.type == CP_ROW) row_phase<T16>(p.c, ph, s_red, reinterpret_cast<float4 *>(scratch));
            else self_phase<T16>(p.c, ph, reinterpret_cast<float *>(scratch));
            ptx::fence_proxy_async_smem();      // generic-proxy writes to the scratch stage vs. the TMA writes that follow
        } else if (ph.type == CP_GEMM) {
            const TMap * ta = &p.tm[2 * ph.tm], * tw = &p.tm[2 * ph.tm + 1];
            const uint32_t tx = x_bytes + NT * 128u;
#pragma unroll 1
            for (int vc = blockIdx.x; vc < g.G; vc += gridDim.x) {
                if (vc != (int) blockIdx.x) {            // further virtual CTAs of this phase: nothing was prefetched
                    range_of(g, vc, u0, nu);
                    npre = 0;
                }
                const int tile0 = u0 / g.kpt, kb0 = u0 - tile0 * g.kpt;
                if (is_producer) {
                    // ===== TMA producer =====
                    int tile = tile0, kb = kb0;
                    long long t_wait = 0, t_begin = clock64();
                    for (int it = 0; it < nu; ++it) {
                        const unsigned gc = gcount + it;
                        uint8_t * st = stage_ptr(gc);
                        if (it >= npre) {
                            const long long t0 = clock64();
                            if (gc >= S) mbar_wait_bounded(&s_empty[gc % S], ((gc / S) - 1u) & 1u);
                            t_wait += clock64() - t0;
                            ptx::mbar_arrive_expect_tx(&s_full[gc % S], tx);
                            ptx::tma_load_2d(st + x_bytes, tw, &s_full[gc % S], kb * CB, tile * NT);
                        }
                        ptx::tma_load_2d(st, ta, &s_full[gc % S], kb * CB, 0);
                        if (++kb == g.kpt) { kb = 0; ++tile; }
                    }
                    if (p.trace && blockIdx.x == 0 && g.direct && ph.gelu == 0) { p.trace[20] = (unsigned long long) t_wait; p.trace[21] = (unsigned long long) (clock64() - t_begin); }
                } else if (is_mma) {
                    // ===== MMA issuer =====
                    const uint32_t idesc = ptx::make_idesc_f16(Half16<T16>::kind, 128, NT);
                    int kb = kb0;
                    unsigned tc = tcount;
                    bool fresh = true;                   // the next MMA starts a new output tile
                    long long t_wait = 0, t_begin = clock64();
                    for (int it = 0; it < nu; ++it) {
                        const unsigned gc = gcount + it;
                        const long long t0 = clock64();
                        mbar_wait_bounded(&s_full[gc % S], (gc / S) & 1u);
                        t_wait += clock64() - t0;
                        if (fresh && tc >= 1) mbar_wait_bounded(&s_acc_empty, (tc - 1u) & 1u);    // TMEM drained
                        ptx::tc_fence_after();
                        const uint32_t sx = ptx::smem_u32(stage_ptr(gc));
                        const uint64_t da = ptx::make_sw128_kmajor_desc(sx), db = ptx::make_sw128_kmajor_desc(sx + x_bytes);
#pragma unroll
                        for (int k = 0; k < 4; ++k)      // +32 bytes along K inside the swizzle atom = +2 in 16-byte units
                            ptx::umma_f16(tmem, da + (uint64_t) (2 * k), db + (uint64_t) (2 * k), idesc, (uint32_t) (!fresh || k != 0));
                        ptx::umma_commit(&s_empty[gc % S]);
                        fresh = false;
                        if (++kb == g.kpt || it == nu - 1) {
                            ptx::umma_commit(&s_acc_full);
                            kb = 0;
                            ++tc;
                            fresh = true;
                        }
                    }
                    if (p.trace && blockIdx.x == 0 && g.direct && ph.gelu == 0) { p.trace[22] = (unsigned long long) t_wait; p.trace[23] = (unsigned long long) (clock64() - t_begin); }
                } else if (warp < 4) {
                    // ===== accumulator read-back =====
                    int tile = tile0, kb = kb0;
                    unsigned tc = tcount;
                    const int row = warp * 32 + lane;
#pragma unroll 1
                    for (int it = 0; it < nu; ++it) {
                        if (++kb != g.kpt && it != nu - 1) continue;
                        mbar_wait_bounded(&s_acc_full, tc & 1u);
                        ptx::tc_fence_after();
                        if (warp * 32 < R) {             // warp-uniform: this warp's 32 accumulator lanes hold live rows
                            const float * bias = ph.bias;
                            const int gelu = ph.gelu, ref16 = p.c.ref_f16_gelu, scale_cols = ph.scale_cols;
                            const float scale_v = ph.scale;
                            const int first = g.direct ? 0 : sg_cta_of(ph.g, tile * g.kpt);
                            float * dst = g.direct ? nullptr : ph.part + ((size_t) tile * ph.g.maxc + (vc - first)) * SG_TILE_FLOATS + row * NT;
                            T16 * out = g.direct ? reinterpret_cast<T16 *>(ph.out16) + (size_t) row * ph.ldo16 + tile * NT : nullptr;
#pragma unroll 1
                            for (int c32 = 0; c32 < NT / 32; ++c32) {
                                uint32_t r[32];
                                ptx::tmem_ld_32x32(tmem + ((uint32_t) (warp * 32) << 16) + (uint32_t) (c32 * 32), r);
                                ptx::tmem_ld_wait();
                                if (row >= R) continue;
                                if (!g.direct) {         // this CTA's partial tile of output tile `tile`
#pragma unroll
                                    for (int j = 0; j < 32; j += 4)
                                        __stcg(reinterpret_cast<float4 *>(dst + c32 * 32 + j),
                                               make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])));
                                } else {                 // whole tile: bias, scale, GELU, 16-bit store
                                    const int n0 = tile * NT + c32 * 32;
                                    const float scale = n0 < scale_cols ? scale_v : 1.0f;
#pragma unroll 1
                                    for (int j = 0; j < 32; j += 8) {
                                        union { T16 h[8]; uint4 u; } pk;
#pragma unroll
                                        for (int q = 0; q < 8; ++q) {
                                            float v = __uint_as_float(j == 0 ? r[q] : j == 8 ? r[8 + q] : j == 16 ? r[16 + q] : r[24 + q]);
                                            if (bias) v += __ldg(bias + n0 + j + q);
                                            v *= scale;
                                            if (gelu) v = gelu_ref<T16>(v, ref16);
                                            pk.h[q] = Half16<T16>::from_f(v);
                                        }
                                        *reinterpret_cast<uint4 *>(out + c32 * 32 + j) = pk.u;
                                    }
                                }
                            }
                        }
                        ptx::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) ptx::mbar_arrive(&s_acc_empty);
                        kb = 0;
                        ++tile;
                        ++tc;
                    }
                }
                // every thread advances the counters identically
                if (nu > 0) tcount += (unsigned) ((u0 + nu - 1) / g.kpt - tile0 + 1);
                gcount += (unsigned) nu;
            }
        }
    }
    stamp(p.n_phase);
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem, C_TMEM_COLS);
    }
}

}  // namespace

ChainLauncher::~ChainLauncher() {
    if (bar) cudaFree(bar);
}

int chain_init(ChainLauncher & cl, DType dt) {
    if (cl.grid > 0) return cl.grid;
    int dev = 0, n_sm = 0, coop = 0, occ = 0;
    WB_CUDA(cudaGetDevice(&dev));
    WB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    WB_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    if (cuda_failed() || !coop || n_sm <= 0) return 0;
    if (dt == DType::F16) {
        WB_CUDA(cudaFuncSetAttribute(dec_chain_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, C_SMEM));
        WB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dec_chain_kernel<__half>, C_THREADS, C_SMEM));
    } else {
        WB_CUDA(cudaFuncSetAttribute(dec_chain_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, C_SMEM));
        WB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dec_chain_kernel<__nv_bfloat16>, C_THREADS, C_SMEM));
    }
    if (cuda_failed() || occ <= 0) return 0;
    WB_CUDA(cudaMalloc(&cl.bar, 256));
    if (!cl.bar) return 0;
    WB_CUDA(cudaMemset(cl.bar, 0, 256));
    cl.bar_count = 0;
    cl.grid = std::min(occ, 2) * n_sm;
    return cl.grid;
}

SplitGeom chain_geom_direct(int R, int N, int K) {
    SplitGeom g;
    g.tiles = N / NT;
    g.kpt = K / CB;
    (void) R;
    g.G = g.tiles;                            // virtual CTAs: one whole R x 128 tile each
    g.U = g.G * g.kpt;
    g.maxc = 0;
    return g;
}

SplitGeom chain_geom(int grid, int R, int N, int K, int min_units) {
    SplitGeom g;
    g.tiles = N / NT;
    g.kpt = K / CB;
    (void) R;
    g.U = g.tiles * g.kpt;
    g.G = std::max(1, std::min(grid, g.U / std::max(1, min_units)));
    g.maxc = 1;
    for (int ot = 0; ot < g.tiles; ++ot) {
        const int first = (int) ((((long long) ot * g.kpt + 1) * g.G - 1) / g.U);
        const int last = (int) ((((long long) ot * g.kpt + g.kpt) * g.G - 1) / g.U);
        g.maxc = std::max(g.maxc, last - first + 1);
    }
    return g;
}

size_t chain_part_floats(const SplitGeom & g, int R) {
    (void) R;
    return (size_t) g.tiles * g.maxc * SG_TILE_FLOATS;
}

bool chain_launch(ChainLauncher & cl, DType dt, ChainParams & p, cudaStream_t stream) {
    if (cl.grid <= 0 || p.n_phase <= 0 || p.n_phase > CHAIN_MAX_PHASES) return false;
    p.bar = cl.bar;
    p.bar_base = cl.bar_count;
    cl.bar_count += (unsigned) (p.n_phase - 1) * (unsigned) cl.grid;
    const void * fn = dt == DType::F16 ? (const void *) dec_chain_kernel<__half> : (const void *) dec_chain_kernel<__nv_bfloat16>;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cl.grid);
    cfg.blockDim = dim3(C_THREADS);
    cfg.dynamicSmemBytes = C_SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    void * args[] = {(void *) &p};
    static const bool no_coop = getenv("WHISPER_B200_CHAIN_COOP") && atoi(getenv("WHISPER_B200_CHAIN_COOP")) == 0;
    if (no_coop) {      // development aid: plain launch (co-residency then rests on nothing else running on the device)
        cfg.attrs = attr + 1;
        cfg.numAttrs = 1;
        WB_CUDA(cudaLaunchKernelExC(&cfg, fn, args));
        return !cuda_failed();
    }
    if (cl.pdl_ok) {
        cfg.numAttrs = 2;
        const cudaError_t e = cudaLaunchKernelExC(&cfg, fn, args);
        if (e == cudaSuccess) return true;
        (void) cudaGetLastError();          // the combination is refused: fall back to a plain cooperative launch
        cl.pdl_ok = false;
    }
    cfg.numAttrs = 1;
    WB_CUDA(cudaLaunchKernelExC(&cfg, fn, args));
    return !cuda_failed();
}

}  // namespace wb
