// Encoder self-attention on the 5th-generation tensor cores (tcgen05 / TMEM), fed by TMA.
//
// Replaces ggml_flash_attn_ext (or KQ / soft_max_ext / KQV) of the reference encoder layer, src/whisper.cpp:2131-2190:
// non-causal attention over T = 1500 positions per window, dh = 64, with the optional 36 zero "phantom" keys of the
// reference's padded flash-attention scratch (src/whisper.cpp:2055, 2141-2159).
//
// One CTA = 128 queries of one (window, head); two CTAs per SM so that one CTA's softmax overlaps the other's MMAs.
//   warp 4 (one lane): TMA producer -- Q tile once, then K tiles [128 keys][64] (two stages) and V^T tiles [64][128 keys]
//                      (one stage: V_j is only needed after softmax_j, long after PV_{j-1} released the buffer); 3-D tensor
//                      maps, so rows past T inside a window are out of bounds = zero-filled;
//   warp 5 (one lane): MMA issuer   -- S = Q K^T (4 x tcgen05.mma 128x128x16) into TMEM, then, once the softmax warps have
//                      published P in shared memory, PV = P V (8 x tcgen05.mma 128x64x16) into a second TMEM buffer;
//   warps 0-7: softmax, two threads per query row (tcgen05.ld 32x32b gives a thread its own row; warps w and w+4 may read
//              the same 32 lanes and take 64 score columns each): row max, exp2, P as 16-bit into the 128-byte-swizzled
//              A-operand layout, running max / sum, O kept in registers and rescaled there (O = O * corr + PV read back
//              from TMEM) -- no read-modify-write of TMEM.
// V is consumed as V^T (K-major B operand), produced by a small transpose kernel per layer.
// Measured (B200, large-v3, 64 windows): 339 TFLOP/s against 245 for the mma.sync kernel it replaces.  Per key tile the MMA
// issuer spends ~520 cycles issuing S, ~770 issuing PV and ~2400 waiting for the softmax warps, whose own chain (S landed ->
// row max -> exchange -> PV read-back -> exponentials -> publish P) is latency- not throughput-bound; the second CTA per SM
// is what fills the gaps.
#include "enc_kernels.h"

#include <math.h>

#include "ptx.cuh"
#include "tc_gemm.h"

namespace wb {

namespace {

constexpr int FA_THREADS = 320;     // 8 softmax warps (two threads per query row) + TMA producer + MMA issuer
constexpr int FA_BQ = 128, FA_BK = 128, FA_DH = 64;
constexpr int FA_TILE_BYTES = 128 * 128;                // 128 rows x 64 x 16-bit
// V^T carries 16 extra rows per head: row 64 is all ones (rows 65..79 zero), so column 64 of P V is the row sum of P exactly
// as the tensor core saw it (16-bit P): the softmax warps never add up their exponentials.
constexpr int FA_VROWS = FA_DH + 16;
constexpr int FA_VBLK_BYTES = FA_VROWS * 128;           // one k-block of V^T: [80 rows][64 keys]
constexpr int FA_OFF_Q = 0, FA_OFF_K = FA_TILE_BYTES, FA_OFF_P = 3 * FA_TILE_BYTES, FA_OFF_V = 5 * FA_TILE_BYTES;
constexpr int FA_SMEM = 5 * FA_TILE_BYTES + 2 * FA_VBLK_BYTES + 1024;   // Q + 2 K + P (two k-blocks) + V^T + slack = 101 KB: two CTAs per SM
constexpr int FA_TMEM_COLS = 256;                       // S: columns 0..127, PV: columns 128..207

__device__ __forceinline__ void fa_wait(uint64_t * bar, uint32_t parity) {       // bounded: a protocol error must trap, not hang
    for (unsigned spins = 0; !ptx::mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 26)) __trap();
}
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
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

template <typename T16>
__global__ void __launch_bounds__(FA_THREADS, 2)
enc_attn_tc_kernel(const __grid_constant__ TMap tm_qk, const __grid_constant__ TMap tm_vt, T16 * __restrict__ out, int T, int d,
                   int H, float scale_log2e, int n_phantom) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t b_q, b_kfull[2], b_kempty[2], b_vfull, b_vempty, b_s, b_p, b_pv;
    __shared__ uint32_t s_tmem;
    __shared__ float s_mx[2][2][FA_BQ];      // [tile parity][column half][row]: row maxima exchanged between the two halves
    uint8_t * smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * FA_BQ, head = blockIdx.y, win = blockIdx.z;
    const int n_tiles = (T + FA_BK - 1) / FA_BK;

    if (threadIdx.x == 0) {
        ptx::mbar_init(&b_q, 1);
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&b_kfull[i], 1);
            ptx::mbar_init(&b_kempty[i], 1);
        }
        ptx::mbar_init(&b_vfull, 1);
        ptx::mbar_init(&b_vempty, 1);
        ptx::mbar_init(&b_s, 1);
        ptx::mbar_init(&b_p, 256);
        ptx::mbar_init(&b_pv, 1);
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

    if (warp == 8) {
        // ===== TMA producer =====
        if (lane == 0) {
            ptx::prefetch_tensormap(&tm_qk);
            ptx::prefetch_tensormap(&tm_vt);
            ptx::mbar_arrive_expect_tx(&b_q, FA_TILE_BYTES);
            ptx::tma_load_3d(smem + FA_OFF_Q, &tm_qk, &b_q, head * FA_DH, q0, win);
            for (int j = 0; j < n_tiles; ++j) {
                const int s = j & 1;
                if (j >= 2) fa_wait(&b_kempty[s], ((j >> 1) - 1) & 1);
                ptx::mbar_arrive_expect_tx(&b_kfull[s], FA_TILE_BYTES);
                ptx::tma_load_3d(smem + FA_OFF_K + s * FA_TILE_BYTES, &tm_qk, &b_kfull[s], d + head * FA_DH, j * FA_BK, win);
                if (j >= 1) fa_wait(&b_vempty, (j - 1) & 1);
                ptx::mbar_arrive_expect_tx(&b_vfull, 2 * FA_VBLK_BYTES);
                uint8_t * vs = smem + FA_OFF_V;                          // two k-blocks: [80 rows][64 keys] each
                ptx::tma_load_3d(vs, &tm_vt, &b_vfull, j * FA_BK, 0, win * H + head);
                ptx::tma_load_3d(vs + FA_VBLK_BYTES, &tm_vt, &b_vfull, j * FA_BK + 64, 0, win * H + head);
            }
        }
    } else if (warp == 9) {
        // ===== MMA issuer =====
        // Order per tile: (P_j published) -> S_{j+1} first, then PV_j.  Issuing a tcgen05.mma costs ~150 cycles whatever its
        // shape, PV is 8 of them: with S_{j+1} ahead of PV_j the softmax warps start on tile j+1 while PV_j is still being
        // issued, and only need PV_j when they come to rescale O.
        if (lane == 0) {
            const uint32_t idesc_s = ptx::make_idesc_f16(Half16<T16>::kind, 128, FA_BK);
            const uint32_t idesc_pv = ptx::make_idesc_f16(Half16<T16>::kind, 128, FA_VROWS);
            const uint64_t dq = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem + FA_OFF_Q));
            const uint64_t dp = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem + FA_OFF_P));
            const uint64_t dv = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem + FA_OFF_V));
            auto issue_s = [&](int j) {
                const int s = j & 1;
                fa_wait(&b_kfull[s], (j >> 1) & 1);
                ptx::tc_fence_after();
                const uint64_t dk = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem + FA_OFF_K + s * FA_TILE_BYTES));
#pragma unroll
                for (int k = 0; k < 4; ++k) ptx::umma_f16(tmem, dq + (uint64_t) (2 * k), dk + (uint64_t) (2 * k), idesc_s, (uint32_t) (k != 0));
                ptx::umma_commit(&b_kempty[s]);
                ptx::umma_commit(&b_s);
            };
            fa_wait(&b_q, 0);
            issue_s(0);
            for (int j = 0; j < n_tiles; ++j) {
                fa_wait(&b_p, j & 1);               // P_j is in shared memory; S_j and PV_{j-1} have been read back
                if (j + 1 < n_tiles) issue_s(j + 1);
                fa_wait(&b_vfull, j & 1);
                ptx::tc_fence_after();
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    // k-block = k / 4 (P: 16 KB apart, V^T: 10 KB apart, in 16-byte units), 16-key step inside it = k % 4
                    const uint64_t ap = dp + (uint64_t) ((k >> 2) * (FA_TILE_BYTES >> 4) + 2 * (k & 3));
                    const uint64_t bv = dv + (uint64_t) ((k >> 2) * (FA_VBLK_BYTES >> 4) + 2 * (k & 3));
                    ptx::umma_f16(tmem + 128u, ap, bv, idesc_pv, (uint32_t) (k != 0));
                }
                ptx::umma_commit(&b_vempty);
                ptx::umma_commit(&b_pv);
            }
        }
    } else {
        // ===== softmax: two threads per query row (warps w and w+4 read the same 32 TMEM lanes), 64 score columns each =====
        const int quarter = warp & 3, half = warp >> 2;
        const int row = quarter * 32 + lane;                            // TMEM lane and row inside the tile
        const uint32_t t_lane = tmem + ((uint32_t) (quarter * 32) << 16);
        uint8_t * prow = smem + FA_OFF_P + half * FA_TILE_BYTES + row * 128;       // this thread's k-block of P
        float o[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = 0.0f;
        float m_run = -INFINITY, l_run = 0.0f, corr_prev = 0.0f;
#pragma unroll 1
        for (int j = 0; j < n_tiles; ++j) {
            fa_wait(&b_s, j & 1);
            ptx::tc_fence_after();
            const int key0 = j * FA_BK + half * 64;
            const bool edge = key0 + 64 > T;
            // pass 1: maximum of this thread's 64 raw scores, then of the whole row
            float mx = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                uint32_t r[32];
                ptx::tmem_ld_32x32(t_lane + (uint32_t) (half * 64 + c * 32), r);
                ptx::tmem_ld_wait();
                if (edge) {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (key0 + c * 32 + i < T) mx = fmaxf(mx, __uint_as_float(r[i]));
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
                }
            }
            s_mx[j & 1][half][row] = mx;
            asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");      // the two warps that share these rows
            mx = fmaxf(mx, s_mx[j & 1][half ^ 1][row]);
            const float m_new = fmaxf(m_run, mx);
            const float corr = ex2((m_run - m_new) * scale_log2e);       // m_run = -inf on the first tile -> 0
            const float mb = m_new * scale_log2e;
            // O = O * corr + P V of the previous tile (its MMAs were issued behind this tile's S)
            if (j > 0) {
                fa_wait(&b_pv, (j - 1) & 1);
                ptx::tc_fence_after();
                uint32_t r[32];
                ptx::tmem_ld_32x32(t_lane + 128u + (uint32_t) (half * 32), r);
                const uint32_t rsum = ptx::tmem_ld_32x1(t_lane + 128u + 64u);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] = fmaf(o[i], corr_prev, __uint_as_float(r[i]));
                l_run = fmaf(l_run, corr_prev, __uint_as_float(rsum));
            }
            corr_prev = corr;
            // pass 2: P = exp2(s * scale - m), 16-bit, into the swizzled A-operand layout (this thread: one 128-byte row)
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                uint32_t r[32];
                ptx::tmem_ld_32x32(t_lane + (uint32_t) (half * 64 + c * 32), r);
                ptx::tmem_ld_wait();
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float x0 = fmaf(__uint_as_float(r[i]), scale_log2e, -mb);
                    float x1 = fmaf(__uint_as_float(r[i + 1]), scale_log2e, -mb);
                    if (edge) {
                        if (key0 + c * 32 + i >= T) x0 = -INFINITY;
                        if (key0 + c * 32 + i + 1 >= T) x1 = -INFINITY;
                    }
                    pk[i >> 1] = ex2_pack<T16>(x0, x1);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int chunk = c * 4 + q;                            // 16-byte chunk inside the 128-byte row
                    *reinterpret_cast<uint4 *>(prow + ((chunk ^ (row & 7)) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                }
            }
            m_run = m_new;
            ptx::fence_proxy_async_smem();           // P was written through the generic proxy; the tensor core reads it through the async one
            ptx::tc_fence_before();
            ptx::mbar_arrive(&b_p);
        }
        {   // last tile's P V
            fa_wait(&b_pv, (n_tiles - 1) & 1);
            ptx::tc_fence_after();
            uint32_t r[32];
            ptx::tmem_ld_32x32(t_lane + 128u + (uint32_t) (half * 32), r);
            const uint32_t rsum = ptx::tmem_ld_32x1(t_lane + 128u + 64u);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = fmaf(o[i], corr_prev, __uint_as_float(r[i]));
            l_run = fmaf(l_run, corr_prev, __uint_as_float(rsum));
            ptx::tc_fence_before();
        }
        // phantom keys (score 0, value 0), normalise, store this thread's 32 of the 64 values
        float l = l_run, f = 1.0f;
        if (n_phantom > 0) {
            const float m_new = fmaxf(m_run, 0.0f);
            f = ex2((m_run - m_new) * scale_log2e);
            l = l * f + (float) n_phantom * ex2(-m_new * scale_log2e);
        }
        const float inv = f / l;
        const int q = q0 + row;
        if (q < T) {
            T16 * orow = out + ((size_t) win * T + q) * (size_t) d + head * FA_DH + half * 32;
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
                *reinterpret_cast<uint4 *>(orow + i) =
                    make_uint4(pack2<T16>(o[i] * inv, o[i + 1] * inv), pack2<T16>(o[i + 2] * inv, o[i + 3] * inv),
                               pack2<T16>(o[i + 4] * inv, o[i + 5] * inv), pack2<T16>(o[i + 6] * inv, o[i + 7] * inv));
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem, FA_TMEM_COLS);
    }
}

}  // namespace

size_t enc_attention_tc_scratch_bytes(int n_windows, int T, int n_head) {
    const int TP = round_up(T, 8);
    return (size_t) n_windows * n_head * FA_VROWS * TP * 2;
}

bool enc_attention_tc(DType dt, const void * qkv, void * out, void * vt_scratch, int n_windows, int T, int d, int n_head,
                      int n_phantom, cudaStream_t st) {
    if (d != n_head * FA_DH || (d % 8) != 0) return false;
    const int TP = round_up(T, 8);
    TMap tm_qk, tm_vt;
    // qkv as {3d, T, W}: a 128-row box that runs past T inside a window is zero-filled instead of reading the next window
    if (!tc_make_tmap3d(&tm_qk, qkv, 3 * d, T, n_windows, (size_t) 3 * d * 2, (size_t) T * 3 * d * 2, 64, 128, dt)) return false;
    // V^T (+ ones row) as {T, 80, W*H} with row pitch TP
    if (!tc_make_tmap3d(&tm_vt, vt_scratch, T, FA_VROWS, n_windows * n_head, (size_t) TP * 2, (size_t) FA_VROWS * TP * 2, 64, FA_VROWS, dt)) return false;
    const float scale_log2e = (1.0f / sqrtf((float) FA_DH)) * 1.4426950408889634f;
    dim3 tgrid(ceil_div(TP, 64), n_head, n_windows), grid(ceil_div(T, FA_BQ), n_head, n_windows);
    if (dt == DType::F16) {
        static bool set = false;
        if (!set) {
            WB_CUDA(cudaFuncSetAttribute(enc_attn_tc_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
            set = true;
        }
        v_transpose_kernel<__half><<<tgrid, 256, 0, st>>>(reinterpret_cast<const __half *>(qkv), reinterpret_cast<__half *>(vt_scratch), T, TP, d, n_head);
        enc_attn_tc_kernel<__half><<<grid, FA_THREADS, FA_SMEM, st>>>(tm_qk, tm_vt, reinterpret_cast<__half *>(out), T, d, n_head, scale_log2e, n_phantom);
    } else {
        static bool set = false;
        if (!set) {
            WB_CUDA(cudaFuncSetAttribute(enc_attn_tc_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
            set = true;
        }
        v_transpose_kernel<__nv_bfloat16><<<tgrid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16 *>(qkv), reinterpret_cast<__nv_bfloat16 *>(vt_scratch), T, TP, d, n_head);
        enc_attn_tc_kernel<__nv_bfloat16><<<grid, FA_THREADS, FA_SMEM, st>>>(tm_qk, tm_vt, reinterpret_cast<__nv_bfloat16 *>(out), T, d, n_head, scale_log2e, n_phantom);
    }
    WB_CUDA(cudaGetLastError());
    return !cuda_failed();
}

}  // namespace wb
