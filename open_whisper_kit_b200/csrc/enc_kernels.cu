// Encoder-side kernels around the GEMMs: conv-stem im2col (with the deferred mel clamp), fused LayerNorm,
// and a flash-style non-causal self-attention (dh = 64) with the reference's phantom-key quirk.
//
// Reference operators replaced:
//   im2col_*   : ggml_conv_1d_ph = im2col(F16) + mul_mat        ggml/src/ggml.c:4409-4436, src/whisper.cpp:2006-2014
//   layernorm  : ggml_norm + ggml_mul + ggml_add (3 launches)    src/whisper.cpp:2102-2110, 2207-2215, 2241-2249
//   enc_attn   : permute/cpy into kv_pad + ggml_flash_attn_ext, or KQ / soft_max_ext / KQV   src/whisper.cpp:2131-2190
#include "enc_kernels.h"

#include "mel.h"

namespace wb {

namespace {

template <typename T16> struct Pack2;
template <> struct Pack2<__half> {
    static __device__ __forceinline__ uint32_t pack(float a, float b) {
        __half2 h = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t *>(&h);
    }
};
template <> struct Pack2<__nv_bfloat16> {
    static __device__ __forceinline__ uint32_t pack(float a, float b) {
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t *>(&h);
    }
};

// ---- im2col for conv1 (k=3, s=1, p=1) -------------------------------------------------------------------
// A1[(w, t)][k*n_mel + c] = mel_w[c][seek_w + t + k - 1]   (0 outside the 3000-frame window, 0 past n_len,
// the constant -10 for frames the FFT never touched), with the reference's global clamp/normalisation applied
// on the fly:  (max(x, mmax-8) + 4) / 4   (src/whisper.cpp:3228-3244).  Window slicing as whisper_encode_internal
// (src/whisper.cpp:2381-2403).
template <typename T16>
__global__ void im2col1_kernel(const EncWindow * __restrict__ wins, int n_mel, int k_pad, int n_frames /* 2 * audio context */,
                               T16 * __restrict__ out) {
    __shared__ float tile[32][33];
    const EncWindow w = wins[blockIdx.z];
    const int t0 = blockIdx.x * 32;     // output time block
    const int c0 = blockIdx.y * 32;     // mel channel block
    const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
    const float mmax = w.finalized ? 0.0f : mel_decode_max(*w.max_enc);
    // load mel[c0 + row][seek + t0 - 1 + col], col in [0, 34) handled as 32 + 2 halo via three passes below
    for (int k = 0; k < 3; ++k) {
        for (int r = ty; r < 32; r += 8) {
            const int c = c0 + r;
            const int tl = t0 + tx + k - 1;                  // frame inside the window
            const long long f = (long long) w.seek + tl;     // frame inside the stream's mel
            float v = 0.0f;
            if (c < n_mel && tl >= 0 && tl < n_frames && f < w.n_len) {
                if (w.finalized) {
                    v = w.mel[(size_t) c * w.stride + f];
                } else {
                    const float x = f < w.n_frames_fft ? w.mel[(size_t) c * w.stride + f] : -10.0f;
                    v = (fmaxf(x, mmax - 8.0f) + 4.0f) * 0.25f;
                }
            }
            tile[r][tx] = v;
        }
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {
            const int t = t0 + r;
            const int c = c0 + tx;
            if (t < n_frames && c < n_mel) {
                out[((size_t) blockIdx.z * n_frames + t) * k_pad + k * n_mel + c] = Half16<T16>::from_f(tile[tx][r]);
            }
        }
        __syncthreads();
    }
    // zero the K padding columns once (3*n_mel .. k_pad)
    if (blockIdx.y == 0) {
        const int pad0 = 3 * n_mel;
        for (int r = ty; r < 32; r += 8) {
            const int t = t0 + r;
            if (t >= n_frames) continue;
            for (int c = pad0 + tx; c < k_pad; c += 32)
                out[((size_t) blockIdx.z * n_frames + t) * k_pad + c] = Half16<T16>::from_f(0.0f);
        }
    }
}

// ---- im2col for conv2 (k=3, s=2, p=1) on time-major activations ----------------------------------------------
// A2[(w, t)][k*d + c] = act1[(w, 2t + k - 1)][c], zero for 2t+k-1 outside [0, 3000).  16-byte vector copies.
__global__ void im2col2_kernel(const uint4 * __restrict__ act1, int d8 /* d/8 */, int T /* audio context */, uint4 * __restrict__ out,
                               long long n_vec /* total uint4 of out */) {
    const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_vec) return;
    const int row_vec = 3 * d8;
    const long long m = i / row_vec;             // (w, t)
    const int r = (int) (i % row_vec);
    const int k = r / d8, c8 = r % d8;
    const long long w = m / T;
    const int t = (int) (m % T);
    const int ts = 2 * t + k - 1;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (ts >= 0 && ts < 2 * T) v = act1[(w * 2 * T + ts) * d8 + c8];
    out[i] = v;
}

// ---- LayerNorm: y = (x - mean) / sqrt(var + eps) * gamma + beta, f32 in, 16-bit (and/or f32) out -----------------
// One warp per row, row kept in registers (d <= 1280 -> <= 40 values per lane).
template <typename T16, int MAXV>
__global__ void layernorm_kernel(const float * __restrict__ x, int ldx, const float * __restrict__ gamma,
                                 const float * __restrict__ beta, float eps, int M, int d, T16 * __restrict__ y16,
                                 int ldy16, float * __restrict__ y32, int ldy32, const int * __restrict__ row_map) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= M) return;
    const int src_row = row_map ? row_map[warp] : warp;
    const float * xr = x + (size_t) src_row * ldx;
    float v[MAXV];
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int c = lane + 32 * i;
        v[i] = c < d ? xr[c] : 0.0f;
        s += v[i];
    }
    const float mean = warp_sum(s) / (float) d;
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int c = lane + 32 * i;
        const float dv = c < d ? v[i] - mean : 0.0f;
        v[i] = dv;
        q += dv * dv;
    }
    const float var = warp_sum(q) / (float) d;
    const float rstd = 1.0f / sqrtf(var + eps);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int c = lane + 32 * i;
        if (c < d) {
            const float o = v[i] * rstd * gamma[c] + beta[c];
            if (y16) y16[(size_t) warp * ldy16 + c] = Half16<T16>::from_f(o);
            if (y32) y32[(size_t) warp * ldy32 + c] = o;
        }
    }
}

// Same operator for the encoder's big row counts (M = windows x 1500) when d is a multiple of 128: one warp per row, the row
// in registers as NV float4 per lane (16-byte loads, 512 contiguous bytes per warp instruction), 8-byte packed 16-bit stores.
// The scalar version above moves 4 bytes per lane and instruction and stops at 42 % of the HBM peak; this one is bound by the
// 6 bytes per element it has to move.  Two-pass statistics (mean, then centred variance) as before.
template <typename T16, int NV>
__global__ void __launch_bounds__(256)
layernorm_vec_kernel(const float * __restrict__ x, int ldx, const float * __restrict__ gamma, const float * __restrict__ beta,
                     float eps, int M, T16 * __restrict__ y16, int ldy16, float * __restrict__ y32, int ldy32) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    constexpr int d = NV * 128;
    const float4 * xr = reinterpret_cast<const float4 *>(x + (size_t) row * ldx) + lane;
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = __ldcs(xr + 32 * i);          // streamed once
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) * (1.0f / (float) d);
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
        q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / (float) d) + eps);
    const float4 * g4 = reinterpret_cast<const float4 *>(gamma) + lane;
    const float4 * b4 = reinterpret_cast<const float4 *>(beta) + lane;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float4 g = __ldg(g4 + 32 * i), b = __ldg(b4 + 32 * i);
        const float4 o = make_float4(v[i].x * rstd * g.x + b.x, v[i].y * rstd * g.y + b.y, v[i].z * rstd * g.z + b.z,
                                     v[i].w * rstd * g.w + b.w);
        const int c = 4 * lane + 128 * i;
        if (y16) {
            uint2 pk;
            pk.x = Pack2<T16>::pack(o.x, o.y);
            pk.y = Pack2<T16>::pack(o.z, o.w);
            *reinterpret_cast<uint2 *>(y16 + (size_t) row * ldy16 + c) = pk;
        }
        if (y32) *reinterpret_cast<float4 *>(y32 + (size_t) row * ldy32 + c) = o;
    }
}

// ---- flash-style encoder self-attention ----------------------------------------------------------------------
// qkv: [B*T][3d] 16-bit (Q | K | V), out: [B*T][d] 16-bit.  One CTA = 64 queries of one (window, head); 4 warps of 16
// query rows; keys in tiles of 64 through a double-buffered cp.async ring; QK^T and PV on mma.sync m16n8k16 with f32
// accumulation; online softmax in f32.  n_phantom extra keys with score 0 and value 0 reproduce the reference's
// flash_attn=true path, which attends over the zero rows 1500..1535 of its padded K/V scratch
// (src/whisper.cpp:2055, 2141-2159).
constexpr int AT_BR = 64, AT_BC = 64, AT_DH = 64, AT_THREADS = 128;

__device__ __forceinline__ void cp_async16(void * smem, const void * gmem, bool valid) {
    const uint32_t s = (uint32_t) __cvta_generic_to_shared(smem);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t & r0, uint32_t & r1, uint32_t & r2, uint32_t & r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t & r0, uint32_t & r1, uint32_t & r2, uint32_t & r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
template <typename T16> __device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <> __device__ __forceinline__ void mma16816<__half>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <> __device__ __forceinline__ void mma16816<__nv_bfloat16>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// byte offset of (row, 16-byte chunk) inside a [rows][64 x 16-bit] tile with an XOR swizzle on the chunk index
__device__ __forceinline__ uint32_t sw_off(int row, int chunk) { return (uint32_t) (row * 128 + ((chunk ^ (row & 7)) << 4)); }

template <typename T16>
__global__ void __launch_bounds__(AT_THREADS)
enc_attn_kernel(const T16 * __restrict__ qkv, T16 * __restrict__ out, int T, int d, float scale_log2e, int n_phantom) {
    __shared__ __align__(128) uint8_t s_q[AT_BR * 128];
    __shared__ __align__(128) uint8_t s_k[2][AT_BC * 128];
    __shared__ __align__(128) uint8_t s_v[2][AT_BC * 128];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * AT_BR;
    const int head = blockIdx.y;
    const size_t row_base = (size_t) blockIdx.z * T;
    const int ld = 3 * d;
    const T16 * gq = qkv + row_base * ld + head * AT_DH;
    const T16 * gk = gq + d;
    const T16 * gv = gq + 2 * d;

    auto load_tile = [&](uint8_t * dst, const T16 * src, int r0) {
#pragma unroll
        for (int i = 0; i < (AT_BC * 8) / AT_THREADS; ++i) {
            const int idx = tid + i * AT_THREADS;
            const int r = idx >> 3, c = idx & 7;
            const bool ok = (r0 + r) < T;
            cp_async16(dst + sw_off(r, c), src + (size_t) (ok ? r0 + r : 0) * ld + c * 8, ok);
        }
    };

    load_tile(s_q, gq, q0);
    load_tile(s_k[0], gk, 0);
    load_tile(s_v[0], gv, 0);
    cp_async_commit();

    const int n_tiles = (T + AT_BC - 1) / AT_BC;
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.0f;
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.0f, 0.0f};
    uint32_t qf[4][4];

    for (int it = 0; it < n_tiles; ++it) {
        const int buf = it & 1;
        if (it + 1 < n_tiles) {
            load_tile(s_k[buf ^ 1], gk, (it + 1) * AT_BC);
            load_tile(s_v[buf ^ 1], gv, (it + 1) * AT_BC);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (it == 0) {
            const uint32_t qb = (uint32_t) __cvta_generic_to_shared(s_q);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const int r = warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
                const int c = ks * 2 + (lane >> 4);
                ldsm_x4(qb + sw_off(r, c), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
            }
        }
        // S = Q K^T  (16 x 64 per warp)
        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.0f;
        const uint32_t kb = (uint32_t) __cvta_generic_to_shared(s_k[buf]);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
            for (int np = 0; np < 4; ++np) {           // pairs of 8-key blocks
                uint32_t b0, b1, b2, b3;
                const int r = np * 16 + (lane & 7) + 8 * (lane >> 4);
                const int c = ks * 2 + ((lane >> 3) & 1);
                ldsm_x4(kb + sw_off(r, c), b0, b1, b2, b3);
                mma16816<T16>(s[2 * np], qf[ks], b0, b1);
                mma16816<T16>(s[2 * np + 1], qf[ks], b2, b3);
            }
        }
        // mask keys beyond T (only the last tile can have them)
        const int key0 = it * AT_BC;
        if (key0 + AT_BC > T) {
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                const int kcol = key0 + nb * 8 + 2 * (lane & 3);
                if (kcol >= T) s[nb][0] = s[nb][2] = -INFINITY;
                if (kcol + 1 >= T) s[nb][1] = s[nb][3] = -INFINITY;
            }
        }
        // online softmax: rows g (regs 0,1) and g+8 (regs 2,3)
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            mx[0] = fmaxf(mx[0], fmaxf(s[nb][0], s[nb][1]));
            mx[1] = fmaxf(mx[1], fmaxf(s[nb][2], s[nb][3]));
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
            mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
        }
        float corr[2], mnew[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            mnew[h] = fmaxf(m_run[h], mx[h]);
            corr[h] = exp2f((m_run[h] - mnew[h]) * scale_log2e);     // m_run = -inf on the first tile -> 0
            m_run[h] = mnew[h];
        }
        float rs[2] = {0.0f, 0.0f};
        uint32_t pf[4][4];
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            const float p0 = exp2f((s[nb][0] - mnew[0]) * scale_log2e);
            const float p1 = exp2f((s[nb][1] - mnew[0]) * scale_log2e);
            const float p2 = exp2f((s[nb][2] - mnew[1]) * scale_log2e);
            const float p3 = exp2f((s[nb][3] - mnew[1]) * scale_log2e);
            rs[0] += p0 + p1;
            rs[1] += p2 + p3;
            const int j = nb >> 1;
            if ((nb & 1) == 0) {
                pf[j][0] = Pack2<T16>::pack(p0, p1);
                pf[j][1] = Pack2<T16>::pack(p2, p3);
            } else {
                pf[j][2] = Pack2<T16>::pack(p0, p1);
                pf[j][3] = Pack2<T16>::pack(p2, p3);
            }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) l_run[h] = l_run[h] * corr[h] + rs[h];
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            o[nb][0] *= corr[0];
            o[nb][1] *= corr[0];
            o[nb][2] *= corr[1];
            o[nb][3] *= corr[1];
        }
        // O += P V
        const uint32_t vb = (uint32_t) __cvta_generic_to_shared(s_v[buf]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {                  // 16-key steps
#pragma unroll
            for (int np = 0; np < 4; ++np) {           // pairs of 8-wide dh blocks
                uint32_t b0, b1, b2, b3;
                const int r = j * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
                const int c = np * 2 + (lane >> 4);
                ldsm_x4_t(vb + sw_off(r, c), b0, b1, b2, b3);
                mma16816<T16>(o[2 * np], pf[j], b0, b1);
                mma16816<T16>(o[2 * np + 1], pf[j], b2, b3);
            }
        }
        __syncthreads();
    }

    // finish the row sums across the quad, add the phantom keys, normalise, store
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 1);
        l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 2);
    }
    float inv[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float l = l_run[h];
        float f = 1.0f;
        if (n_phantom > 0) {
            const float mnew = fmaxf(m_run[h], 0.0f);
            f = exp2f((m_run[h] - mnew) * scale_log2e);
            l = l * f + (float) n_phantom * exp2f(-mnew * scale_log2e);
        }
        inv[h] = f / l;
    }
    const int g = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int q = q0 + warp * 16 + g + 8 * h;
        if (q >= T) continue;
        T16 * orow = out + (row_base + q) * (size_t) d + head * AT_DH;
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            const uint32_t pk = Pack2<T16>::pack(o[nb][2 * h] * inv[h], o[nb][2 * h + 1] * inv[h]);
            *reinterpret_cast<uint32_t *>(orow + nb * 8 + 2 * tq) = pk;
        }
    }
}

// Few rows (decoder step): one 128-thread CTA per row so a row's loads spread over four warps.
template <typename T16>
__global__ void __launch_bounds__(128)
layernorm_row_kernel(const float * __restrict__ x, int ldx, const float * __restrict__ gamma, const float * __restrict__ beta,
                     float eps, int d, T16 * __restrict__ y16, int ldy16, float * __restrict__ y32, int ldy32,
                     const int * __restrict__ row_map) {
    __shared__ float s_red[8];
    const int row = blockIdx.x, tid = threadIdx.x;
    pdl_trigger();
    // gamma / beta are weights (never written on the device): fetch them while the predecessor is still running
    float gw[10], gb[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const int c = tid + 128 * i;
        gw[i] = c < d ? __ldg(gamma + c) : 0.0f;
        gb[i] = c < d ? __ldg(beta + c) : 0.0f;
    }
    pdl_wait();
    const float * xr = x + (size_t) (row_map ? row_map[row] : row) * ldx;
    float v[10];                           // d <= 1280
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const int c = tid + 128 * i;
        v[i] = c < d ? xr[c] : 0.0f;
        s += v[i];
    }
    s = warp_sum(s);
    if ((tid & 31) == 0) s_red[tid >> 5] = s;
    __syncthreads();
    const float mean = (s_red[0] + s_red[1] + s_red[2] + s_red[3]) / (float) d;
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const int c = tid + 128 * i;
        const float dv = c < d ? v[i] - mean : 0.0f;
        v[i] = dv;
        q += dv * dv;
    }
    q = warp_sum(q);
    if ((tid & 31) == 0) s_red[4 + (tid >> 5)] = q;
    __syncthreads();
    const float rstd = 1.0f / sqrtf((s_red[4] + s_red[5] + s_red[6] + s_red[7]) / (float) d + eps);
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const int c = tid + 128 * i;
        if (c < d) {
            const float o = v[i] * rstd * gw[i] + gb[i];
            if (y16) y16[(size_t) row * ldy16 + c] = Half16<T16>::from_f(o);
            if (y32) y32[(size_t) row * ldy32 + c] = o;
        }
    }
}

template <typename T16>
void layernorm_dispatch(const float * x, int ldx, const float * g, const float * b, float eps, int M, int d, void * y16,
                        int ldy16, float * y32, int ldy32, const int * row_map, cudaStream_t st) {
    const int threads = 256;
    const int blocks = ceil_div(M * 32, threads);
    T16 * y = reinterpret_cast<T16 *>(y16);
    if (M <= 1024 && d <= 1280) {
        launch_pdl(layernorm_row_kernel<T16>, dim3(M), dim3(128), 0, st, x, ldx, g, b, eps, d, y, ldy16, y32, ldy32, row_map);
        return;
    }
    auto al16 = [](const void * q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    if (!row_map && d % 128 == 0 && d <= 1280 && al16(x) && ldx % 4 == 0 && al16(g) && al16(b) && (!y || (al16(y) && ldy16 % 4 == 0)) &&
        (!y32 || (al16(y32) && ldy32 % 4 == 0))) {
        switch (d / 128) {
#define WB_LN_VEC(NV) case NV: layernorm_vec_kernel<T16, NV><<<blocks, threads, 0, st>>>(x, ldx, g, b, eps, M, y, ldy16, y32, ldy32); return;
            WB_LN_VEC(1) WB_LN_VEC(2) WB_LN_VEC(3) WB_LN_VEC(4) WB_LN_VEC(5) WB_LN_VEC(6) WB_LN_VEC(7) WB_LN_VEC(8) WB_LN_VEC(9) WB_LN_VEC(10)
#undef WB_LN_VEC
        }
    }
    if (d <= 512)
        layernorm_kernel<T16, 16><<<blocks, threads, 0, st>>>(x, ldx, g, b, eps, M, d, y, ldy16, y32, ldy32, row_map);
    else if (d <= 1024)
        layernorm_kernel<T16, 32><<<blocks, threads, 0, st>>>(x, ldx, g, b, eps, M, d, y, ldy16, y32, ldy32, row_map);
    else
        layernorm_kernel<T16, 40><<<blocks, threads, 0, st>>>(x, ldx, g, b, eps, M, d, y, ldy16, y32, ldy32, row_map);
}

}  // namespace

void im2col1(DType dt, const EncWindow * d_wins, int n_windows, int n_mel, int k_pad, int T, void * out, cudaStream_t st) {
    dim3 grid(ceil_div(2 * T, 32), ceil_div(n_mel, 32), n_windows), block(32, 8);
    if (dt == DType::F16)
        im2col1_kernel<__half><<<grid, block, 0, st>>>(d_wins, n_mel, k_pad, 2 * T, reinterpret_cast<__half *>(out));
    else
        im2col1_kernel<__nv_bfloat16><<<grid, block, 0, st>>>(d_wins, n_mel, k_pad, 2 * T, reinterpret_cast<__nv_bfloat16 *>(out));
    WB_CUDA(cudaGetLastError());
}

void im2col2(const void * act1, int n_windows, int d, int T, void * out, cudaStream_t st) {
    const long long n_vec = (long long) n_windows * T * 3 * (d / 8);
    const int threads = 256;
    im2col2_kernel<<<(unsigned) ceil_div<long long>(n_vec, threads), threads, 0, st>>>(
        reinterpret_cast<const uint4 *>(act1), d / 8, T, reinterpret_cast<uint4 *>(out), n_vec);
    WB_CUDA(cudaGetLastError());
}

void layernorm(DType dt, const float * x, int ldx, const float * gamma, const float * beta, float eps, int M, int d,
               void * y16, int ldy16, float * y32, int ldy32, const int * row_map, cudaStream_t st) {
    if (M <= 0) return;
    if (dt == DType::F16)
        layernorm_dispatch<__half>(x, ldx, gamma, beta, eps, M, d, y16, ldy16, y32, ldy32, row_map, st);
    else
        layernorm_dispatch<__nv_bfloat16>(x, ldx, gamma, beta, eps, M, d, y16, ldy16, y32, ldy32, row_map, st);
    WB_CUDA(cudaGetLastError());
}

void enc_attention(DType dt, const void * qkv, void * out, int n_windows, int T, int d, int n_head, int n_phantom,
                   cudaStream_t st) {
    dim3 grid(ceil_div(T, AT_BR), n_head, n_windows);
    const float scale_log2e = (1.0f / sqrtf((float) AT_DH)) * 1.4426950408889634f;
    if (dt == DType::F16)
        enc_attn_kernel<__half><<<grid, AT_THREADS, 0, st>>>(reinterpret_cast<const __half *>(qkv),
                                                             reinterpret_cast<__half *>(out), T, d, scale_log2e, n_phantom);
    else
        enc_attn_kernel<__nv_bfloat16><<<grid, AT_THREADS, 0, st>>>(reinterpret_cast<const __nv_bfloat16 *>(qkv),
                                                                    reinterpret_cast<__nv_bfloat16 *>(out), T, d,
                                                                    scale_log2e, n_phantom);
    WB_CUDA(cudaGetLastError());
}

}  // namespace wb
