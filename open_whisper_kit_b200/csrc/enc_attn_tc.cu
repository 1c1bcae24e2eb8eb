// Encoder self-attention on the 5th-generation tensor cores (tcgen05 / TMEM), fed by TMA.
//
// Replaces ggml_flash_attn_ext (or KQ / soft_max_ext / KQV) of the reference encoder layer, src/whisper.cpp:2131-2190:
// non-causal attention over T = 1500 positions per window, dh = 64, with the optional 36 zero "phantom" keys of the
// reference's padded flash-attention scratch (src/whisper.cpp:2055, 2141-2159).
//
// One CTA = 128 queries of one (window, head); two CTAs per SM so that one CTA's softmax overlaps the other's MMAs.
//   warp 4 (one lane): TMA producer -- Q tile once, then K tiles [64 keys][64] and V^T tiles [80][64 keys] through two-stage
//                      rings; 3-D tensor maps, so rows past T inside a window are out of bounds = zero-filled;
//   warp 5 (one lane): MMA issuer   -- S = Q K^T (4 x tcgen05.mma 128x64x16) into one of two TMEM buffers, and, once the
//                      softmax warps have published P in one of two shared-memory buffers, PV += P V (4 x 128x80x16);
//   warps 0-3: softmax, one query row per thread (tcgen05.ld 32x32b gives a thread its own row): ONE pass over S per key
//              tile -- exp2 against the row maximum of the earlier tiles ("stale" maximum, two exponentials per MUFU op
//              straight into the 128-byte-swizzled 16-bit A-operand layout), O and the row sum accumulate in TMEM across all
//              key tiles; a row whose scores outgrow its maximum by 2^12 takes the new one and rescales its TMEM lane.
//   Eight-softmax-warp variant (default): two threads per query row, and P_j is written back into TENSOR memory over the scores it
//   came from (tcgen05.st) and consumed by the PV MMA as its A operand -- no shared-memory round trip, no proxy fence.
// V is consumed as V^T (K-major B operand), written by the QKV GEMM's epilogue (or a small transpose kernel per layer).
// Measured (B200, large-v3, 64 windows, in the bench step): ~335 TFLOP/s against 245 for the mma.sync kernel it replaces.
// Per 64-key tile the MMA issuer spends ~430 cycles issuing S, ~430 issuing PV and ~900 waiting for P; a softmax thread
// spends ~1200 of its ~1950 cycles in tcgen05.ld + FFMA / cvt / MUFU for its 64 scores: with one softmax warp per SM
// sub-partition and CTA the phase is bound by instruction latency, not by the tensor or MUFU pipes (21 % / 38 % busy).
#include "enc_kernels.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "ptx.cuh"
#include "tc_gemm.h"

namespace wb {

namespace {

// softmax warps per CTA: 4 (thread = query row) or 8 (two threads per row, 32 of the tile's 64 keys each) + TMA producer + MMA issuer
constexpr int fa_threads(int sw) { return (sw + 2) * 32; }
constexpr int FA_BQ = 128, FA_BK = 64, FA_DH = 64;
constexpr int FA_Q_BYTES = 128 * 128;                   // 128 queries x 64 x 16-bit
constexpr int FA_K_BYTES = FA_BK * 128;                 // 64 keys x 64 x 16-bit
// V^T carries 16 extra rows per head: row 64 is all ones (rows 65..79 zero), so column 64 of P V is the row sum of P exactly
// as the tensor core saw it (16-bit P): the softmax warps never add up their exponentials.
constexpr int FA_VROWS = FA_DH + 16;
constexpr int FA_V_BYTES = FA_VROWS * 128;              // [80 rows][64 keys]
constexpr int FA_P_BYTES = 128 * 128;                   // [128 queries][64 keys]
// K / V^T ring depth: two stages next to the two P buffers; FOUR when P lives in tensor memory (the 32 KB of P buffers become ring
// stages: a K or V tile comes from DRAM, ~1 400 cycles = one whole tile period away, so with two stages the TMA producer -- which
// can only ask for K_{j+2} once PV_{j-1} has released its V stage -- delivered just in time at best)
constexpr int fa_ns(bool pt) { return pt ? 4 : 2; }
constexpr int FA_OFF_Q = 0, FA_OFF_K = FA_Q_BYTES;
constexpr int fa_off_v(bool pt) { return FA_OFF_K + fa_ns(pt) * FA_K_BYTES; }
constexpr int fa_off_p(bool pt) { return fa_off_v(pt) + fa_ns(pt) * FA_V_BYTES; }       // P buffers exist only when !pt
constexpr int fa_smem(bool pt) { return fa_off_p(pt) + (pt ? 0 : 2 * FA_P_BYTES) + 1024; }
constexpr int FA_SMEM = fa_smem(true) > fa_smem(false) ? fa_smem(true) : fa_smem(false);   // 85 / 89 KB: two CTAs per SM
constexpr int FA_TMEM_COLS = 256;                       // S (two buffers): columns 0..63, 64..127; PV: columns 128..207
static_assert(fa_off_p(false) % 1024 == 0 && fa_off_v(false) % 1024 == 0 && fa_off_v(true) % 1024 == 0 && FA_V_BYTES % 1024 == 0,
              "128-byte swizzle atoms are 1024 bytes");

__device__ __forceinline__ void fa_wait(uint64_t * bar, uint32_t parity) {       // bounded: a protocol error must trap, not hang
    for (unsigned spins = 0; !ptx::mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 26)) __trap();
}
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// three-input maximum and two-wide f32 FMA (sm_100): half the issue slots of the softmax warps' max / scale passes
__device__ __forceinline__ float max3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ void fma2(float & y0, float & y1, float a0, float a1, float s, float c) {      // (y0, y1) = (a0, a1) * s + c
    asm("{\n\t.reg .b64 ra, rs, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rs, {%4, %4};\n\tmov.b64 rc, {%5, %5};\n\t"
        "fma.rn.f32x2 rd, ra, rs, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(y0), "=f"(y1)
        : "f"(a0), "f"(a1), "f"(s), "f"(c));
}
// two exponentials per MUFU instruction, straight into the 16-bit pair the P tile wants (f16 keeps subnormals)
template <typename T16> __device__ __forceinline__ uint32_t ex2_pack(float a, float b);
template <> __device__ __forceinline__ uint32_t ex2_pack<__half>(float a, float b) {
    uint32_t y;
    asm("{\n\t.reg .b32 t;\n\tcvt.rn.f16x2.f32 t, %2, %1;\n\tex2.approx.f16x2 %0, t;\n\t}" : "=r"(y) : "f"(a), "f"(b));
    return y;
}
template <> __device__ __forceinline__ uint32_t ex2_pack<__nv_bfloat16>(float a, float b) {
    uint32_t y;
    asm("{\n\t.reg .b32 t;\n\tcvt.rn.bf16x2.f32 t, %2, %1;\n\tex2.approx.ftz.bf16x2 %0, t;\n\t}" : "=r"(y) : "f"(a), "f"(b));
    return y;
}
template <typename T16> __device__ __forceinline__ uint32_t pack2(float a, float b);

template <> __device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}

// V [W*T][3d] (columns 2d + h*64 ..) -> vt [(w*H + h)][80][TP]: rows 0..63 = V^T, row 64 = 1, rows 65..79 = 0
// (keys contiguous; columns >= T are never read: TMA bounds)
template <typename T16>
__global__ void __launch_bounds__(256)
v_transpose_kernel(const T16 * __restrict__ qkv, T16 * __restrict__ vt, int T, int TP, int d, int H) {
    __shared__ T16 tile[64][66];
    const int k0 = blockIdx.x * 64, h = blockIdx.y, w = blockIdx.z;
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;     // 64 x 4
    const T16 * src = qkv + ((size_t) w * T) * 3 * d + 2 * d + h * 64;
    for (int r = ty; r < 64; r += 4) {
        const int k = k0 + r;
        tile[r][tx] = k < T ? src[(size_t) k * 3 * d + tx] : T16(0.0f);
    }
    __syncthreads();
    T16 * dst = vt + ((size_t) (w * H + h) * FA_VROWS) * TP;
    for (int c = ty; c < FA_VROWS; c += 4) {
        const int k = k0 + tx;
        if (k < TP) dst[(size_t) c * TP + k] = c < 64 ? tile[tx][c] : T16(c == 64 ? 1.0f : 0.0f);
    }
}

// rows 64..79 of every head's V^T block: row 64 = 1 (row sums through the PV MMA), rows 65..79 = 0.  Written once per encoder
// call when the QKV GEMM's epilogue fills rows 0..63 itself.
template <typename T16>
__global__ void vt_tail_rows_kernel(T16 * __restrict__ vt, int TP, int n_blocks) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    const size_t per = (size_t) 16 * TP;
    if (i >= per * n_blocks) return;
    const size_t b = i / per, r = (i % per) / TP, k = i % TP;
    vt[(b * FA_VROWS + 64 + r) * TP + k] = T16(r == 0 ? 1.0f : 0.0f);
}

template <typename T16, int SW, bool PT>
__global__ void __launch_bounds__(fa_threads(SW), 2)
enc_attn_tc_kernel(const __grid_constant__ TMap tm_q, const __grid_constant__ TMap tm_k, const __grid_constant__ TMap tm_vt,
                   T16 * __restrict__ out, int T, int d, int H, float scale_log2e, int n_phantom, long long * __restrict__ trace) {
    extern __shared__ uint8_t smem_raw[];
    // development aid (WHISPER_B200_FA_TRACE): clock64 stamps of CTA (0, 0, 0) -- [who][tile][8], who 0 / 1 = lane 0 of softmax warps
    // 0 / 4, 2 = the MMA issuer
    const bool tr = trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
#define FA_STAMP(who, j, i) do { if (tr && (j) < 32) trace[((who) * 32 + (j)) * 8 + (i)] = clock64(); } while (0)
    // Ring stage of tile j: j % NS.  b_s[st] / b_pv[st]: "S_j / PV_j complete" (one tcgen05.commit each), waited for by the softmax
    // warps (S_j ready, PV_{j-1} landed) AND by the TMA producer (the K / V stage is free again); b_p[j & 1]: P_j published.
    constexpr int NS = fa_ns(PT), FA_OFF_V = fa_off_v(PT), FA_OFF_P = fa_off_p(PT);
    __shared__ __align__(8) uint64_t b_q, b_kfull[NS], b_vfull[NS], b_s[NS], b_pv[NS], b_p[2];
    __shared__ uint32_t s_tmem;
    __shared__ float s_hmax[SW == 8 ? 2 : 1][2][128];       // SW == 8: the two threads of a row exchange their half-tile maxima here
    uint8_t * smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * FA_BQ, head = blockIdx.y, win = blockIdx.z;
    const int n_tiles = (T + FA_BK - 1) / FA_BK;

    if (threadIdx.x == 0) {
        ptx::mbar_init(&b_q, 1);
        for (int i = 0; i < NS; ++i) {
            ptx::mbar_init(&b_kfull[i], 1);
            ptx::mbar_init(&b_vfull[i], 1);
            ptx::mbar_init(&b_s[i], 1);
            ptx::mbar_init(&b_pv[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&b_p[i], SW == 8 ? 8 : SW * 32);      // SW == 8: one arrival per warp
        }
        ptx::fence_mbar_init();
    }
    if (warp == 0) {
        ptx::tmem_alloc(&s_tmem, FA_TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = s_tmem;
    // tile j uses buffer j & 1 of every ring; its k-th use of that buffer (k = j >> 1) completes phase k of the buffer's barriers
    if (warp == SW) {
        // ===== TMA producer =====
        if (lane == 0) {
            ptx::prefetch_tensormap(&tm_q);
            ptx::prefetch_tensormap(&tm_k);
            ptx::prefetch_tensormap(&tm_vt);
            ptx::mbar_arrive_expect_tx(&b_q, FA_Q_BYTES);
            ptx::tma_load_3d(smem + FA_OFF_Q, &tm_q, &b_q, head * FA_DH, q0, win);
            for (int j = 0; j < n_tiles; ++j) {
                const int st = j % NS;
                const uint32_t prev = (uint32_t) ((j / NS) - 1) & 1u;     // phase of the stage's previous tenant, tile j - NS
                if (j >= NS) fa_wait(&b_s[st], prev);            // S_{j-NS} complete = K stage st has been read (one commit serves both)
                ptx::mbar_arrive_expect_tx(&b_kfull[st], FA_K_BYTES);
                ptx::tma_load_3d(smem + FA_OFF_K + st * FA_K_BYTES, &tm_k, &b_kfull[st], d + head * FA_DH, j * FA_BK, win);
                if (j >= NS) fa_wait(&b_pv[st], prev);           // PV_{j-NS} complete = V stage st has been read
                ptx::mbar_arrive_expect_tx(&b_vfull[st], FA_V_BYTES);
                ptx::tma_load_3d(smem + FA_OFF_V + st * FA_V_BYTES, &tm_vt, &b_vfull[st], j * FA_BK, 0, win * H + head);
            }
        }
    } else if (warp == SW + 1) {
        // ===== MMA issuer =====
        // S is double-buffered in TMEM and P in shared memory: S_{j+1} is already computed while the softmax warps work on
        // S_j, and PV_j reads P_j while they write P_{j+1}, so the two hand-offs per tile (commit -> mbarrier -> wake-up, ~1 us
        // together as measured) overlap with work instead of serialising it.  PV accumulates in TMEM over all key tiles.
        if (lane == 0) {
            const uint32_t idesc_s = ptx::make_idesc_f16(Half16<T16>::kind, 128, FA_BK);
            const uint32_t idesc_pv = ptx::make_idesc_f16(Half16<T16>::kind, 128, FA_VROWS);
            const uint64_t dq = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem + FA_OFF_Q));
            auto wait_k = [&](int j) {
                fa_wait(&b_kfull[j % NS], (uint32_t) (j / NS) & 1u);
                FA_STAMP(2, j, 4);
            };
            auto issue_s = [&](int j) {          // K_j has arrived (wait_k)
                const int s = j & 1, st = j % NS;
                ptx::tc_fence_after();
                const uint64_t dk = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem + FA_OFF_K + st * FA_K_BYTES));
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    ptx::umma_f16(tmem + (uint32_t) (s * FA_BK), dq + (uint64_t) (2 * k), dk + (uint64_t) (2 * k), idesc_s, (uint32_t) (k != 0));
                ptx::umma_commit(&b_s[st]);
                FA_STAMP(2, j, 5);
            };
            fa_wait(&b_q, 0);
            wait_k(0);
            issue_s(0);
            if (n_tiles > 1) {
                wait_k(1);
                issue_s(1);
            }
            // Per tile this lane issues eight MMAs and two commits, ~65 cycles each as measured (profiles/r4_fa_trace.txt): with the
            // P hand-off through tensor memory it is this issue sequence, not the softmax warps, that sets the tile period.  So
            // what does not depend on the softmax warps -- V_j having arrived -- is waited for BEFORE P_j.
            for (int j = 0; j < n_tiles; ++j) {
                const int s = j & 1, st = j % NS;
                FA_STAMP(2, j, 0);
                fa_wait(&b_vfull[st], (uint32_t) (j / NS) & 1u);
                FA_STAMP(2, j, 2);
                fa_wait(&b_p[s], (j >> 1) & 1);      // P_j is published, S_j has been read, the accumulator is consistent
                FA_STAMP(2, j, 1);
                ptx::tc_fence_after();
                const uint64_t dp = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem + FA_OFF_P + s * FA_P_BYTES));
                const uint64_t dv = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem + FA_OFF_V + st * FA_V_BYTES));
#pragma unroll
                for (int k = 0; k < 4; ++k) {        // 64 keys = four 16-key steps
                    if constexpr (PT)                // P_j sits in tensor memory, in the first 32 columns of S buffer s (8 columns per step)
                        ptx::umma_f16_ts(tmem + 128u, tmem + (uint32_t) (s * FA_BK + 8 * k), dv + (uint64_t) (2 * k), idesc_pv, (uint32_t) (j != 0 || k != 0));
                    else
                        ptx::umma_f16(tmem + 128u, dp + (uint64_t) (2 * k), dv + (uint64_t) (2 * k), idesc_pv, (uint32_t) (j != 0 || k != 0));
                }
                ptx::umma_commit(&b_pv[st]);
                FA_STAMP(2, j, 3);
                if (j + 2 < n_tiles) {               // S buffer s is free: the softmax warps published P_j after reading it
                    wait_k(j + 2);
                    issue_s(j + 2);
                }
            }
        }
    } else if constexpr (SW == 8) {
        // ===== softmax, two threads per query row =====
        // Same scheme as the four-warp variant below (one pass over S per key tile against the row maximum of the EARLIER tiles,
        // O and the row sum in TMEM), but a row's 64 scores of a tile are split between warp w (keys 0..31) and warp w + 4 (keys
        // 32..63) -- both may read TMEM lanes 32 (w % 4) .. -- so that every SM sub-partition has four softmax warps to hide the
        // tcgen05.ld / MUFU / convert latencies behind (the phase is bound by instruction latency, not by a pipe: 21 % tensor,
        // 48 % XU with two warps per sub-partition).  The two threads of a row exchange their half-tile maxima through shared
        // memory once per tile (double-buffered, one 256-thread named barrier), so both take identical rescaling decisions.
        constexpr float kGrow = 12.0f;
        const int half = warp >> 2, row = (warp & 3) * 32 + lane;
        const uint32_t t_lane = tmem + ((uint32_t) ((warp & 3) * 32) << 16);
        auto sync_softmax = [] { asm volatile("bar.sync 1, 256;" ::: "memory"); };
        float m_run = -INFINITY;
        const bool trl = tr && lane == 0 && (warp & 3) == 0;
#define FA_SSTAMP(j, i) do { if (trl && (j) < 32) trace[((half) * 32 + (j)) * 8 + (i)] = clock64(); } while (0)
#pragma unroll 1
        for (int j = 0; j < n_tiles; ++j) {
            const int s = j & 1;
            FA_SSTAMP(j, 0);
            fa_wait(&b_s[j % NS], (uint32_t) (j / NS) & 1u);
            FA_SSTAMP(j, 1);
            ptx::tc_fence_after();
            const int key0 = j * FA_BK + half * 32;
            const bool edge = key0 + 32 > T;
            uint32_t r[32];
            ptx::tmem_ld_32x32(t_lane + (uint32_t) (s * FA_BK + half * 32), r);
            ptx::tmem_ld_wait();
            FA_SSTAMP(j, 2);
            if (edge) {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (key0 + i >= T) r[i] = __float_as_uint(-INFINITY);
            }
            // this half's maximum: four independent chains
            float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
                m4[0] = max3(m4[0], __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
                m4[1] = max3(m4[1], __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
                m4[2] = max3(m4[2], __uint_as_float(r[i + 4]), __uint_as_float(r[i + 5]));
                m4[3] = max3(m4[3], __uint_as_float(r[i + 6]), __uint_as_float(r[i + 7]));
            }
            float mx = fmaxf(max3(m4[0], m4[1], m4[2]), m4[3]);
            uint32_t pk[16];
            auto exp_half = [&](float mb) {
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float y0, y1;
                    fma2(y0, y1, __uint_as_float(r[i]), __uint_as_float(r[i + 1]), scale_log2e, -mb);
                    pk[i >> 1] = ex2_pack<T16>(y0, y1);
                }
            };
            // the exponentials against the stale maximum do not wait for the exchange (tile 0 has no maximum yet: after it)
            if (j > 0) exp_half(m_run * scale_log2e);
            s_hmax[s][half][row] = mx;
            FA_SSTAMP(j, 3);
            sync_softmax();
            FA_SSTAMP(j, 4);
            mx = fmaxf(mx, s_hmax[s][half ^ 1][row]);
            const bool grow = j == 0 || (mx - m_run) * scale_log2e > kGrow;
            if (__any_sync(0xffffffffu, grow)) {            // identical in the partner warp (same rows, same maxima)
                const float m_new = grow ? mx : m_run;
                const float f = j == 0 ? 1.0f : ex2((m_run - m_new) * scale_log2e);      // 1 for the rows that keep their maximum
                m_run = m_new;
                exp_half(m_run * scale_log2e);
                if (j > 0) {              // rescale this thread's half of the row's accumulator (+ the row sum); PV_{j-1} must have landed
                    fa_wait(&b_pv[(j - 1) % NS], (uint32_t) ((j - 1) / NS) & 1u);
                    ptx::tc_fence_after();
                    uint32_t o[32];
                    ptx::tmem_ld_32x32(t_lane + 128u + (uint32_t) (half * 32), o);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
                    ptx::tmem_st_32x32(t_lane + 128u + (uint32_t) (half * 32), o);
                    if (half == 0) {
                        const uint32_t l = ptx::tmem_ld_32x1(t_lane + 128u + 64u);
                        ptx::tmem_ld_wait();
                        ptx::tmem_st_32x1(t_lane + 128u + 64u, __float_as_uint(__uint_as_float(l) * f));
                    }
                    ptx::tmem_st_wait();
                }
            }
            // P buffer s is free once PV_{j-2} has read it -- which this thread already knows: the issuer queued S_j behind PV_{j-2},
            // tcgen05 operations of one thread complete in order, and S_j's commit (b_s) was observed at the top of this iteration
            FA_SSTAMP(j, 5);
            if constexpr (PT) {
                // P_j goes back into tensor memory, over the scores it was computed from: 16-bit pairs, this thread's 32 keys = 16
                // columns of its lane, keys 0..31 of the tile in columns 0..15 of S buffer s (warp w), keys 32..63 in columns 16..31
                // (warp w + 4) -- columns the partner warp has finished reading before the exchange barrier above.  The PV MMA takes
                // it from there as its A operand: no shared-memory round trip, no generic -> async proxy fence.
                ptx::tmem_st_32x16(t_lane + (uint32_t) (s * FA_BK + half * 16), pk);
                FA_SSTAMP(j, 6);
                ptx::tmem_st_wait();
            } else {
                uint8_t * prow = smem + FA_OFF_P + s * FA_P_BYTES + row * 128;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *reinterpret_cast<uint4 *>(prow + (((half * 4 + q) ^ (row & 7)) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                FA_SSTAMP(j, 6);
                ptx::fence_proxy_async_smem();       // P was written through the generic proxy; the tensor core reads it through the async one
            }
            ptx::tc_fence_before();
            __syncwarp();                            // one arrival per warp (32 arrivals on one word serialise)
            if (lane == 0) ptx::mbar_arrive(&b_p[s]);
            FA_SSTAMP(j, 7);
        }
        // O and the row sum l (column 64: the ones row of V^T) sit in TMEM, both relative to m_run
        fa_wait(&b_pv[(n_tiles - 1) % NS], (uint32_t) ((n_tiles - 1) / NS) & 1u);
        ptx::tc_fence_after();
        float l = __uint_as_float(ptx::tmem_ld_32x1(t_lane + 128u + 64u)), f = 1.0f;
        ptx::tmem_ld_wait();
        if (n_phantom > 0) {              // phantom keys: score 0, value 0
            const float m_new = fmaxf(m_run, 0.0f);
            f = ex2((m_run - m_new) * scale_log2e);
            l = l * f + (float) n_phantom * ex2(-m_new * scale_log2e);
        }
        const float inv = f / l;
        const int q = q0 + row;
        T16 * orow = out + ((size_t) win * T + (q < T ? q : 0)) * (size_t) d + head * FA_DH + half * 32;
        {
            uint32_t r[32];
            ptx::tmem_ld_32x32(t_lane + 128u + (uint32_t) (half * 32), r);
            ptx::tmem_ld_wait();
            if (q < T) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    *reinterpret_cast<uint4 *>(orow + i) =
                        make_uint4(pack2<T16>(__uint_as_float(r[i]) * inv, __uint_as_float(r[i + 1]) * inv),
                                   pack2<T16>(__uint_as_float(r[i + 2]) * inv, __uint_as_float(r[i + 3]) * inv),
                                   pack2<T16>(__uint_as_float(r[i + 4]) * inv, __uint_as_float(r[i + 5]) * inv),
                                   pack2<T16>(__uint_as_float(r[i + 6]) * inv, __uint_as_float(r[i + 7]) * inv));
                }
            }
        }
        ptx::tc_fence_before();
    } else {
        // ===== softmax: thread = query row =====
        // P = exp2((s - m) * scale) against the row maximum m of the tiles seen BEFORE this one: P may then exceed 1, which a
        // 16-bit float represents just as well, and nothing depends on this tile's own maximum -- one pass over S, no
        // per-tile correction of O.  Only when a tile's scores outgrow m by more than 2^kGrow does the row take the new
        // maximum, redo the tile and rescale its accumulator lane in TMEM (tcgen05.ld / st); the first tile always does.
        constexpr float kGrow = 12.0f;
        const int row = warp * 32 + lane;                               // TMEM lane and row inside the tile
        const uint32_t t_lane = tmem + ((uint32_t) (warp * 32) << 16);
        float m_run = -INFINITY;
#pragma unroll 1
        for (int j = 0; j < n_tiles; ++j) {
            const int s = j & 1;
            fa_wait(&b_s[j % NS], (uint32_t) (j / NS) & 1u);
            ptx::tc_fence_after();
            const int key0 = j * FA_BK;
            const bool edge = key0 + FA_BK > T;
            const uint32_t t_s = t_lane + (uint32_t) (s * FA_BK);
            uint32_t pk[32];
            float mx = -INFINITY;
            auto exp_tile = [&](float mb, bool track_max) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t r[32];
                    ptx::tmem_ld_32x32(t_s + (uint32_t) (c * 32), r);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        float s0 = __uint_as_float(r[i]), s1 = __uint_as_float(r[i + 1]);
                        if (edge) {
                            if (key0 + c * 32 + i >= T) s0 = -INFINITY;
                            if (key0 + c * 32 + i + 1 >= T) s1 = -INFINITY;
                        }
                        if (track_max) mx = fmaxf(mx, fmaxf(s0, s1));
                        pk[c * 16 + (i >> 1)] = ex2_pack<T16>(fmaf(s0, scale_log2e, -mb), fmaf(s1, scale_log2e, -mb));
                    }
                }
            };
            if (j == 0) {                 // no maximum yet: find it first (warp-uniform)
#pragma unroll 1
                for (int c = 0; c < 2; ++c) {
                    uint32_t r[32];
                    ptx::tmem_ld_32x32(t_s + (uint32_t) (c * 32), r);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (!edge || key0 + c * 32 + i < T) m_run = fmaxf(m_run, __uint_as_float(r[i]));
                }
            }
            exp_tile(m_run * scale_log2e, true);
            const bool grow = (mx - m_run) * scale_log2e > kGrow;
            if (__any_sync(0xffffffffu, grow)) {
                const float m_new = grow ? mx : m_run;
                const float f = ex2((m_run - m_new) * scale_log2e);          // 1 for the rows that keep their maximum
                m_run = m_new;
                exp_tile(m_run * scale_log2e, false);
                if (j > 0) {              // rescale this row's accumulator (64 values + the row sum); PV_{j-1} must have landed
                    fa_wait(&b_pv[(j - 1) % NS], (uint32_t) ((j - 1) / NS) & 1u);
                    ptx::tc_fence_after();
#pragma unroll 1
                    for (int c = 0; c < 2; ++c) {
                        uint32_t r[32];
                        ptx::tmem_ld_32x32(t_lane + 128u + (uint32_t) (c * 32), r);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * f);
                        ptx::tmem_st_32x32(t_lane + 128u + (uint32_t) (c * 32), r);
                    }
                    const uint32_t l = ptx::tmem_ld_32x1(t_lane + 128u + 64u);
                    ptx::tmem_ld_wait();
                    ptx::tmem_st_32x1(t_lane + 128u + 64u, __float_as_uint(__uint_as_float(l) * f));
                    ptx::tmem_st_wait();
                }
            }
            // P buffer s is free once PV_{j-2} has read it
            if (j >= 2) fa_wait(&b_pv[(j - 2) % NS], (uint32_t) ((j - 2) / NS) & 1u);
            uint8_t * prow = smem + FA_OFF_P + s * FA_P_BYTES + row * 128;
#pragma unroll
            for (int q = 0; q < 8; ++q)
                *reinterpret_cast<uint4 *>(prow + ((q ^ (row & 7)) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
            ptx::fence_proxy_async_smem();           // P was written through the generic proxy; the tensor core reads it through the async one
            ptx::tc_fence_before();
            ptx::mbar_arrive(&b_p[s]);
        }
        // O and the row sum l (column 64: the ones row of V^T) sit in TMEM, both relative to m_run
        fa_wait(&b_pv[(n_tiles - 1) % NS], (uint32_t) ((n_tiles - 1) / NS) & 1u);
        ptx::tc_fence_after();
        float l = __uint_as_float(ptx::tmem_ld_32x1(t_lane + 128u + 64u)), f = 1.0f;
        ptx::tmem_ld_wait();
        if (n_phantom > 0) {              // phantom keys: score 0, value 0
            const float m_new = fmaxf(m_run, 0.0f);
            f = ex2((m_run - m_new) * scale_log2e);
            l = l * f + (float) n_phantom * ex2(-m_new * scale_log2e);
        }
        const float inv = f / l;
        const int q = q0 + row;
        T16 * orow = out + ((size_t) win * T + (q < T ? q : 0)) * (size_t) d + head * FA_DH;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
            uint32_t r[32];
            ptx::tmem_ld_32x32(t_lane + 128u + (uint32_t) (c * 32), r);
            ptx::tmem_ld_wait();
            if (q < T) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    *reinterpret_cast<uint4 *>(orow + c * 32 + i) =
                        make_uint4(pack2<T16>(__uint_as_float(r[i]) * inv, __uint_as_float(r[i + 1]) * inv),
                                   pack2<T16>(__uint_as_float(r[i + 2]) * inv, __uint_as_float(r[i + 3]) * inv),
                                   pack2<T16>(__uint_as_float(r[i + 4]) * inv, __uint_as_float(r[i + 5]) * inv),
                                   pack2<T16>(__uint_as_float(r[i + 6]) * inv, __uint_as_float(r[i + 7]) * inv));
                }
            }
        }
        ptx::tc_fence_before();
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem, FA_TMEM_COLS);
    }
}

// ---- development trace: WHISPER_B200_FA_TRACE=<launch index> records the stamps of that launch, printed at exit ----
struct FaTrace {
    long want = -1, seen = 0;
    long long * dev = nullptr;
    FaTrace() {
        const char * e = getenv("WHISPER_B200_FA_TRACE");
        if (!e) return;
        want = atol(e);
        if (cudaMalloc(&dev, 3 * 32 * 8 * 8) != cudaSuccess) { want = -1; return; }
        cudaMemset(dev, 0, 3 * 32 * 8 * 8);
    }
    ~FaTrace() {
        if (want < 0 || !dev) return;
        cudaDeviceSynchronize();
        long long h[3 * 32 * 8];
        if (cudaMemcpy(h, dev, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) return;
        const long long t0 = h[(0 * 32 + 4) * 8 + 0];
        fprintf(stderr, "fa_trace: cycles relative to softmax warp 0's loop top of tile 4\n");
        fprintf(stderr, "fa_trace: softmax stamps = 0 top, 1 S ready, 2 S in regs, 3 exp done, 4 partner max, 5 P buffer free, 6 P stored, 7 arrived\n");
        fprintf(stderr, "fa_trace: mma stamps     = 0 top, 1 P ready, 2 V ready, 3 PV issued, 4 K(j) ready, 5 S(j) issued\n");
        for (int j = 4; j < 14; ++j)
            for (int who = 0; who < 3; ++who) {
                fprintf(stderr, "fa_trace tile %2d %s |", j, who == 0 ? "softmax w0" : who == 1 ? "softmax w4" : "mma       ");
                for (int i = 0; i < (who == 2 ? 6 : 8); ++i) fprintf(stderr, " %6lld", h[(who * 32 + j) * 8 + i] ? h[(who * 32 + j) * 8 + i] - t0 : -1ll);
                fprintf(stderr, "\n");
            }
    }
    long long * slot() {
        if (want < 0) return nullptr;
        return seen++ == want ? dev : nullptr;
    }
};
long long * fa_trace_slot() {
    static FaTrace t;
    return t.slot();
}

}  // namespace

size_t enc_attention_tc_scratch_bytes(int n_windows, int T, int n_head) {
    const int TP = round_up(T, 8);
    return (size_t) n_windows * n_head * FA_VROWS * TP * 2;
}

void enc_attention_tc_init_vt(DType dt, void * vt_scratch, int n_windows, int T, int n_head, cudaStream_t st) {
    const int TP = round_up(T, 8), n_blocks = n_windows * n_head;
    const size_t n = (size_t) 16 * TP * n_blocks;
    if (dt == DType::F16) vt_tail_rows_kernel<__half><<<(unsigned) ceil_div<size_t>(n, 256), 256, 0, st>>>(reinterpret_cast<__half *>(vt_scratch), TP, n_blocks);
    else vt_tail_rows_kernel<__nv_bfloat16><<<(unsigned) ceil_div<size_t>(n, 256), 256, 0, st>>>(reinterpret_cast<__nv_bfloat16 *>(vt_scratch), TP, n_blocks);
    WB_CUDA(cudaGetLastError());
}

bool enc_attention_tc(DType dt, const void * qkv, void * out, void * vt_scratch, int n_windows, int T, int d, int n_head,
                      int n_phantom, cudaStream_t st, bool vt_ready) {
    if (d != n_head * FA_DH || (d % 8) != 0) return false;
    const int TP = round_up(T, 8);
    TMap tm_q, tm_k, tm_vt;
    // qkv as {3d, T, W}: a box that runs past T inside a window is zero-filled instead of reading the next window
    if (!tc_make_tmap3d(&tm_q, qkv, 3 * d, T, n_windows, (size_t) 3 * d * 2, (size_t) T * 3 * d * 2, 64, FA_BQ, dt)) return false;
    if (!tc_make_tmap3d(&tm_k, qkv, 3 * d, T, n_windows, (size_t) 3 * d * 2, (size_t) T * 3 * d * 2, 64, FA_BK, dt)) return false;
    // V^T (+ ones row) as {T, 80, W*H} with row pitch TP
    if (!tc_make_tmap3d(&tm_vt, vt_scratch, T, FA_VROWS, n_windows * n_head, (size_t) TP * 2, (size_t) FA_VROWS * TP * 2, 64, FA_VROWS, dt)) return false;
    const float scale_log2e = (1.0f / sqrtf((float) FA_DH)) * 1.4426950408889634f;
    long long * trace = fa_trace_slot();
    static const bool sw8 = !(getenv("WHISPER_B200_FA_WARPS") && atoi(getenv("WHISPER_B200_FA_WARPS")) == 4);
    dim3 tgrid(ceil_div(TP, 64), n_head, n_windows), grid(ceil_div(T, FA_BQ), n_head, n_windows);
    // P is handed to the PV MMA through tensor memory (eight-warp kernel); WHISPER_B200_FA_PTMEM=0: through shared memory
    static const bool p_tmem = !(getenv("WHISPER_B200_FA_PTMEM") && atoi(getenv("WHISPER_B200_FA_PTMEM")) == 0);
    auto launch = [&](auto tag) {
        using T16 = decltype(tag);
        T16 * o = reinterpret_cast<T16 *>(out);
        static DeviceOnce set;      // function attributes are per device (one guard per instantiation of this lambda = per T16)
        once_per_device(set, [&] {
            WB_CUDA(cudaFuncSetAttribute(enc_attn_tc_kernel<T16, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
            WB_CUDA(cudaFuncSetAttribute(enc_attn_tc_kernel<T16, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
            WB_CUDA(cudaFuncSetAttribute(enc_attn_tc_kernel<T16, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
        });
        if (!vt_ready) v_transpose_kernel<T16><<<tgrid, 256, 0, st>>>(reinterpret_cast<const T16 *>(qkv), reinterpret_cast<T16 *>(vt_scratch), T, TP, d, n_head);
        if (!sw8) enc_attn_tc_kernel<T16, 4, false><<<grid, fa_threads(4), FA_SMEM, st>>>(tm_q, tm_k, tm_vt, o, T, d, n_head, scale_log2e, n_phantom, trace);
        else if (p_tmem) enc_attn_tc_kernel<T16, 8, true><<<grid, fa_threads(8), FA_SMEM, st>>>(tm_q, tm_k, tm_vt, o, T, d, n_head, scale_log2e, n_phantom, trace);
        else enc_attn_tc_kernel<T16, 8, false><<<grid, fa_threads(8), FA_SMEM, st>>>(tm_q, tm_k, tm_vt, o, T, d, n_head, scale_log2e, n_phantom, trace);
    };
    if (dt == DType::F16) launch(__half{});
    else launch(__nv_bfloat16{});
    WB_CUDA(cudaGetLastError());
    return !cuda_failed();
}

}  // namespace wb
