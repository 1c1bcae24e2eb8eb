// Decoder-step kernels: token+position embedding, KV append, masked self-attention over the device-resident
// KV cache, cross-attention over the per-window cross K/V, and the on-device logit rules + greedy selection.
//
// Reference operators replaced (src/whisper.cpp):
//   embed        ggml_get_rows(d_te) + ggml_get_rows(d_pe) + add                      2515-2519
//   kv_append    ggml_cpy(Kcur/Vcur -> kv_self views)                                  2559-2590
//   self_attn    KQ = K*Q, soft_max_ext(KQ, mask), KQV  (mask built on the host 2908-2940)   2594-2632
//   cross_attn   KQ = Kcross*Q, soft_max_ext(scale), KQV                               2680-2742
//   sample       whisper_process_logits + whisper_sample_token(best) on the host      6177-6445, 6460-6517
#include "dec_kernels.h"

#include <cooperative_groups.h>

#include "dec_chain.h"

#include <stdlib.h>

#include <type_traits>

namespace cg = cooperative_groups;

namespace wb {

namespace {

template <typename T16>
__global__ void embed_kernel(const T16 * __restrict__ te, const float * __restrict__ pe, const DecRow * __restrict__ rows,
                             int d, float * __restrict__ x) {
    const int r = blockIdx.x;
    pdl_trigger();
    pdl_wait();
    const DecRow row = rows[r];
    const T16 * t = te + (size_t) row.token * d;
    const float * p = pe + (size_t) row.pos * d;
    for (int c = threadIdx.x; c < d; c += blockDim.x) x[(size_t) r * d + c] = Half16<T16>::to_f(t[c]) + p[c];
}

// qkv [R][3d] -> self_kv[layer][pos][0..2d) = (K | V)
__global__ void kv_append_kernel(const uint4 * __restrict__ qkv, const DecRow * __restrict__ rows, int d8,
                                 size_t layer_off8) {
    const int r = blockIdx.x;
    const DecRow row = rows[r];
    uint4 * dst = reinterpret_cast<uint4 *>(row.self_kv) + layer_off8 + (size_t) row.pos * (2 * d8);
    const uint4 * src = qkv + (size_t) r * (3 * d8) + d8;
    for (int c = threadIdx.x; c < 2 * d8; c += blockDim.x) dst[c] = src[c];
}

// One CTA (128 threads) per (row, head) streaming the window's cross K/V: 8 lanes x 16 bytes cover one 64-value key
// row, so every warp-wide load instruction reads four whole 128-byte rows.  kv: [xslot][layer][T][2d] (K | V), K already
// carries dh^-0.25 (cross graph, src/whisper.cpp:2300-2305); score = (q.k) * dh^-0.25 (src/whisper.cpp:2695, 2719).
//
// SELF = true runs the same streaming scheme over the row's own self-attention cache (keys 0..pos, scale 1, no
// phantom keys; Q and K already carry dh^-0.25 each, src/whisper.cpp:2506-2557).  With fused_append the CTA first stores
// this token's K/V head slice into the cache (single-token steps only: every sequence owns exactly one row, so no
// other CTA needs the slice).
// 16-byte read of data that is streamed exactly once (cross K/V: 491 MB per launch): no L1 allocation.  With the default
// path every in-flight line needs an L1 line, and after a kernel that configured the SM for ~200 KB of shared memory only
// 28 KB of L1 are left -- measured: the same launch takes 115 us instead of 81 us behind the chain kernel.
__device__ __forceinline__ uint4 ld_stream16(const uint4 * p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

template <typename T16, bool SELF, bool QSPLIT>
__global__ void __launch_bounds__(128, 9)
cross_attn_kernel(const T16 * __restrict__ q, int ldq, const DecRow * __restrict__ rows, int d, size_t layer_off, int T_in,
                  float kq_scale, int n_phantom, int fused_append, T16 * __restrict__ out, const SplitIn qs) {
    extern __shared__ float s_sc[];          // [T]
    __shared__ float s_red[8];
    __shared__ float s_o[4][64];
    const int r = blockIdx.x, h = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sub = lane & 7, grp = lane >> 3;     // 8 lanes per key row, 4 key rows per warp instruction
    // The successor is the next chain / GEMM launch, whose CTAs hold ~100 KB of shared memory while they wait: letting
    // them in at once would squeeze this kernel's own occupancy, so the cross pass triggers after its K sweep.
    // The row descriptors (and the cross K/V) were written before this decoder call began (H2D copy, encoder stage), and the
    // self K/V of earlier positions by earlier decoder calls: all of it may be touched before the grid dependency resolves.
    // Pull the first batch of keys -- and, for the short self pass, values -- into L2 meanwhile.
    // The cross pass's successor is the next GEMM launch, whose CTAs hold ~80 KB of shared memory while they wait: letting
    // them in at once would squeeze this kernel's own occupancy, so the cross pass triggers after its K sweep.
    if (SELF) pdl_trigger();
    const DecRow row = rows[r];
    {
        const int T_pre = SELF ? row.pos : T_in;
        const int ld0 = SELF ? 2 * d : 64;
        const T16 * kb0 = (SELF ? reinterpret_cast<const T16 *>(row.self_kv) + layer_off + h * 64
                                : reinterpret_cast<const T16 *>(row.cross_kv) + layer_off + (size_t) h * 2 * T_in * 64) +
                          (size_t) (warp * 4 + grp) * ld0 + sub * 8;
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (warp * 4 + grp + 16 * u < T_pre) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(kb0 + (size_t) (16 * u) * ld0));
                if (SELF) asm volatile("prefetch.global.L2 [%0];" ::"l"(kb0 + (size_t) (16 * u) * ld0 + d));
            }
    }
    pdl_wait();
    if (QSPLIT && qs.trace && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        qs.trace[0] = now;
    }
    const int T = SELF ? row.pos + 1 : T_in;
    // self K/V: [position][K | V] rows of 2d; cross K/V: [head][K | V][T][64], one contiguous block per (window, head)
    const T16 * kbase = SELF ? reinterpret_cast<const T16 *>(row.self_kv) + layer_off + h * 64
                             : reinterpret_cast<const T16 *>(row.cross_kv) + layer_off + (size_t) h * 2 * T_in * 64;
    const int ld = SELF ? 2 * d : 64;
    // The step's own key / value (position row.pos) come straight from the projection output; their copy into the cache is
    // only for later steps, so nothing in this launch waits for it.
    const T16 * knew = (SELF && fused_append) ? q + (size_t) r * ldq + d + h * 64 : nullptr;
    if (SELF && fused_append) {
        if (tid < 16) {
            const int which = tid >> 3, c = tid & 7;       // 0: K slice, 1: V slice; 8 x 16 bytes each
            const uint4 u = *reinterpret_cast<const uint4 *>(q + (size_t) r * ldq + (1 + which) * d + h * 64 + c * 8);
            T16 * dst = reinterpret_cast<T16 *>(row.self_kv) + layer_off + (size_t) row.pos * ld + which * d + h * 64 + c * 8;
            *reinterpret_cast<uint4 *>(dst) = u;
        }
    }

    float qv[8];
    if (!SELF && QSPLIT) {
        // query straight from the partial tiles of the chain kernel's stream-K GEMM (dec_chain.h).  The 64 values of this
        // head sit inside one tile: thread (quad, slot) adds the contributors slot, slot+8, ... of one float4, the eight
        // slot sums are combined in slot order (fixed, reproducible), bias added, rounded to 16 bits like the unfused path.
        float4 * s_q = reinterpret_cast<float4 *>(s_sc);              // [8 slots][16 quads], free until the K sweep
        {
            const int quad = tid & 15, slot = tid >> 4;
            const int ot = (h * 64) / SG_TILE_COLS;
            const int first = sg_cta_of(qs.g, ot * qs.g.kpt), last = sg_cta_of(qs.g, ot * qs.g.kpt + qs.g.kpt - 1);
            const float * src = qs.part + ((size_t) ot * qs.g.maxc) * SG_TILE_FLOATS + r * SG_TILE_COLS + (h * 64) % SG_TILE_COLS + quad * 4;
            float4 a = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            for (int j = slot; j <= last - first; j += 8) {
                const float4 v = __ldcg(reinterpret_cast<const float4 *>(src + (size_t) j * SG_TILE_FLOATS));
                a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
            }
            s_q[slot * 16 + quad] = a;
        }
        __syncthreads();
        {
            float4 a = s_q[sub * 2], b = s_q[sub * 2 + 1];
#pragma unroll
            for (int sl = 1; sl < 8; ++sl) {
                const float4 u = s_q[sl * 16 + sub * 2], v = s_q[sl * 16 + sub * 2 + 1];
                a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w;
                b.x += v.x; b.y += v.y; b.z += v.z; b.w += v.w;
            }
            const float av[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            const int c = h * 64 + sub * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                qv[j] = Half16<T16>::to_f(Half16<T16>::from_f(av[j] + (qs.bias ? __ldg(qs.bias + c + j) : 0.0f)));
        }
        __syncthreads();            // s_sc is about to receive scores
    } else {
        const uint4 u = *reinterpret_cast<const uint4 *>(q + (size_t) r * ldq + h * 64 + sub * 8);
        const T16 * e = reinterpret_cast<const T16 *>(&u);
#pragma unroll
        for (int j = 0; j < 8; ++j) qv[j] = Half16<T16>::to_f(e[j]);
    }
    // keys t = warp*4 + grp + 16*i.  Eight independent 16-byte loads are issued before any is consumed so every
    // lane keeps 128 bytes in flight (memory-level parallelism instead of one dependent load per iteration).
    constexpr int U = 8;
    float mx = -INFINITY;
    // the trip count must be warp-uniform (full-mask shuffles inside): iterate on a common base, predicate per key
    for (int base = 0; base < T; base += 16 * U) {
        const int tb = base + warp * 4 + grp;
        uint4 kb[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int t = tb + 16 * u;
            const uint4 * src = reinterpret_cast<const uint4 *>((SELF && knew && t == row.pos) ? knew + sub * 8
                                                                                                  : kbase + (size_t) t * ld + sub * 8);
            kb[u] = t < T ? (SELF ? *src : (QSPLIT ? ld_stream16(src) : __ldg(src))) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int t = tb + 16 * u;
            const T16 * e = reinterpret_cast<const T16 *>(&kb[u]);
            float acc = 0.0f;
#pragma unroll
            for (int j = 0; j < 8; ++j) acc = fmaf(qv[j], Half16<T16>::to_f(e[j]), acc);
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            acc += __shfl_xor_sync(0xffffffffu, acc, 4);
            acc *= kq_scale;
            if (t < T) {
                if (sub == 0) s_sc[t] = acc;
                mx = fmaxf(mx, acc);
            }
        }
    }
    if (!SELF) pdl_trigger();
    mx = warp_max(mx);
    if (lane == 0) s_red[warp] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
    if (n_phantom > 0) mx = fmaxf(mx, 0.0f);
    float sum = 0.0f;
    for (int t = tid; t < T; t += 128) {
        const float e = expf(s_sc[t] - mx);
        s_sc[t] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    if (lane == 0) s_red[4 + warp] = sum;
    __syncthreads();
    sum = s_red[4] + s_red[5] + s_red[6] + s_red[7];
    if (n_phantom > 0) sum += (float) n_phantom * expf(-mx);
    const float inv = 1.0f / sum;

    const T16 * vbase = SELF ? kbase + d : kbase + (size_t) T_in * 64;
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = 0.0f;
    for (int base = 0; base < T; base += 16 * U) {
        const int tb = base + warp * 4 + grp;
        uint4 vb[U];
        float pr[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int t = tb + 16 * u;
            const bool ok = t < T;
            const uint4 * src = reinterpret_cast<const uint4 *>((SELF && knew && t == row.pos) ? knew + d + sub * 8
                                                                                                  : vbase + (size_t) t * ld + sub * 8);
            vb[u] = ok ? (SELF ? *src : (QSPLIT ? ld_stream16(src) : __ldg(src))) : make_uint4(0, 0, 0, 0);
            pr[u] = ok ? Half16<T16>::to_f(Half16<T16>::from_f(s_sc[t] * inv)) : 0.0f;
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
#pragma unroll
        for (int j = 0; j < 8; ++j) s_o[warp][sub * 8 + j] = o[j];
    }
    __syncthreads();
    if (tid < 64) {
        const float v = s_o[0][tid] + s_o[1][tid] + s_o[2][tid] + s_o[3][tid];
        out[(size_t) r * d + h * 64 + tid] = Half16<T16>::from_f(v);
    }
    if (QSPLIT && qs.trace && tid == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        atomicMax(qs.trace + 1, now);
    }
}

// ---- masked self-attention of the decoder step with the dot products on mma.sync fragments -------------------------------------
// cross_attn_kernel<SELF = true> spends ~46 instructions per (4 keys x 8 lanes) on f16 -> f32 conversions, FFMAs and shuffle
// reductions: 6.2 M warp instructions per launch at 113 positions, issue slots 57 % busy, 14.3 us (profiles/r4_ncu_summary.txt).
// The tensor cores are used here as a convert-and-accumulate engine (the problem stays a matrix-vector product: one query per
// sequence and head, 1/8 of each MMA is useful work) -- what counts is that a 16-byte load feeds an MMA operand register as it is:
//   scores: A = 16 keys x 16 dims straight from the cache rows (a lane's two 16-byte loads per key row ARE its A registers; the
//           assignment of dims to the k index is a permutation applied to K and q alike, so the dot product does not care),
//           B = q in all eight columns -> every lane of a quad receives the scores of key rows lane/4 and lane/4 + 8;
//   P V   : A = V^T (16 dims x 16 keys): the two keys of a register pair come from two cache rows, one PRMT each,
//           B = the 16-bit probabilities in all eight columns -> a lane receives 8 output dims.
// Same rounding points as before (f32 scores, expf, probabilities rounded to 16 bits, f32 accumulation); only the order of the
// f32 additions differs.  One CTA per (row, head), warp w owns the 16-key groups w, w + 4, ...
template <typename T16> __device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1);
template <> __device__ __forceinline__ void mma_16816<__half>(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
template <> __device__ __forceinline__ void mma_16816<__nv_bfloat16>(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
template <typename T16> __device__ __forceinline__ uint32_t pack16x2(float lo, float hi) {
    const T16 a = Half16<T16>::from_f(lo), b = Half16<T16>::from_f(hi);
    return (uint32_t) *reinterpret_cast<const unsigned short *>(&a) | ((uint32_t) *reinterpret_cast<const unsigned short *>(&b) << 16);
}

template <typename T16>
__global__ void __launch_bounds__(128, 9)
self_attn_mma_kernel(const T16 * __restrict__ qkv, int ldq, const DecRow * __restrict__ rows, int d, size_t layer_off,
                     int fused_append, T16 * __restrict__ out) {
    extern __shared__ float s_sc[];          // [n_ctx]
    __shared__ float s_red[8];
    __shared__ float s_o[4][64];
    const int r = blockIdx.x, h = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g8 = lane >> 2, t4 = lane & 3;
    // the successor is a GEMM that only needs its weights until this grid completes: let it in at once
    pdl_trigger();
    // row descriptors and the self K/V of earlier positions were written before this launch's predecessor began: pull the cache
    // rows into L2 while the QKV projection is still running
    const DecRow row = rows[r];
    const int ld = 2 * d;
    const T16 * kbase = reinterpret_cast<const T16 *>(row.self_kv) + layer_off + h * 64;
    for (int t = tid; t < row.pos; t += 128) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(kbase + (size_t) t * ld));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(kbase + (size_t) t * ld + d));
    }
    pdl_wait();
    const int T = row.pos + 1;
    // this step's own key / value (position row.pos) come straight from the projection output; their copy into the cache is for
    // later steps, nothing in this launch reads it back
    const T16 * qrow = qkv + (size_t) r * ldq + h * 64;
    const T16 * knew = fused_append ? qrow + d : nullptr;
    if (fused_append && tid < 16) {
        const int which = tid >> 3, c = tid & 7;       // 0: K slice, 1: V slice; 8 x 16 bytes each
        const uint4 u = *reinterpret_cast<const uint4 *>(qrow + (1 + which) * d + c * 8);
        T16 * dst = reinterpret_cast<T16 *>(row.self_kv) + layer_off + (size_t) row.pos * ld + which * d + h * 64 + c * 8;
        *reinterpret_cast<uint4 *>(dst) = u;
    }
    const uint4 zero4 = make_uint4(0, 0, 0, 0);
    const int n_groups = (T + 15) >> 4;

    // ---- scores ----
    float mx = -INFINITY;
    {
        const uint4 qx = *reinterpret_cast<const uint4 *>(qrow + t4 * 8), qy = *reinterpret_cast<const uint4 *>(qrow + 32 + t4 * 8);
        const uint32_t Q[8] = {qx.x, qx.y, qx.z, qx.w, qy.x, qy.y, qy.z, qy.w};
        auto krow = [&](int t) { return (knew && t == row.pos) ? knew : kbase + (size_t) t * ld; };
        // two groups per trip: eight independent 16-byte loads per lane in flight
        for (int g = warp; g < n_groups; g += 8) {
            const int ta0 = g * 16 + g8, tb0 = ta0 + 8, ta1 = ta0 + 64, tb1 = tb0 + 64;      // group g and group g + 4
            uint4 k[8];
            const int tt[4] = {ta0, tb0, ta1, tb1};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool ok = tt[i] < T;
                const T16 * p = krow(ok ? tt[i] : 0);
                k[2 * i] = ok ? *reinterpret_cast<const uint4 *>(p + t4 * 8) : zero4;
                k[2 * i + 1] = ok ? *reinterpret_cast<const uint4 *>(p + 32 + t4 * 8) : zero4;
            }
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const uint4 xa = k[4 * half], ya = k[4 * half + 1], xb = k[4 * half + 2], yb = k[4 * half + 3];
                const uint32_t Ra[8] = {xa.x, xa.y, xa.z, xa.w, ya.x, ya.y, ya.z, ya.w};
                const uint32_t Rb[8] = {xb.x, xb.y, xb.z, xb.w, yb.x, yb.y, yb.z, yb.w};
                float c[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                for (int s = 0; s < 4; ++s) mma_16816<T16>(c, Ra[2 * s], Rb[2 * s], Ra[2 * s + 1], Rb[2 * s + 1], Q[2 * s], Q[2 * s + 1]);
                const int ta = tt[2 * half], tb = tt[2 * half + 1];
                if (ta < T) { mx = fmaxf(mx, c[0]); if (t4 == 0) s_sc[ta] = c[0]; }
                if (tb < T) { mx = fmaxf(mx, c[2]); if (t4 == 0) s_sc[tb] = c[2]; }
            }
        }
    }
    mx = warp_max(mx);
    if (lane == 0) s_red[warp] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
    float sum = 0.0f;
    for (int t = tid; t < T; t += 128) {
        const float e = expf(s_sc[t] - mx);
        s_sc[t] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    if (lane == 0) s_red[4 + warp] = sum;
    __syncthreads();
    sum = s_red[4] + s_red[5] + s_red[6] + s_red[7];
    const float inv = 1.0f / sum;

    // ---- P V ----
    const T16 * vbase = kbase + d;
    const T16 * vnew = knew ? knew + d : nullptr;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    for (int g = warp; g < n_groups; g += 4) {
        // this lane's four keys of the group: k = 2 t4, 2 t4 + 1, 2 t4 + 8, 2 t4 + 9 (the B fragment's k indices)
        const int k0 = g * 16 + 2 * t4;
        const int kk[4] = {k0, k0 + 1, k0 + 8, k0 + 9};
        uint4 v[4];
        float p[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const bool ok = kk[i] < T;
            const T16 * src = (vnew && kk[i] == row.pos) ? vnew : vbase + (size_t) (ok ? kk[i] : 0) * ld;
            v[i] = ok ? *reinterpret_cast<const uint4 *>(src + g8 * 8) : zero4;
            p[i] = ok ? s_sc[kk[i]] * inv : 0.0f;
        }
        const uint32_t pab = pack16x2<T16>(p[0], p[1]), pcd = pack16x2<T16>(p[2], p[3]);
        const uint32_t va[4] = {v[0].x, v[0].y, v[0].z, v[0].w}, vb[4] = {v[1].x, v[1].y, v[1].z, v[1].w};
        const uint32_t vc[4] = {v[2].x, v[2].y, v[2].z, v[2].w}, vd[4] = {v[3].x, v[3].y, v[3].z, v[3].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            // rows lane/4 and lane/4 + 8 of A = dims 8 g8 + 2 i and 8 g8 + 2 i + 1; columns = this lane's key pairs
            const uint32_t a0 = __byte_perm(va[i], vb[i], 0x5410), a1 = __byte_perm(va[i], vb[i], 0x7632);
            const uint32_t a2 = __byte_perm(vc[i], vd[i], 0x5410), a3 = __byte_perm(vc[i], vd[i], 0x7632);
            mma_16816<T16>(acc[i], a0, a1, a2, a3, pab, pcd);
        }
    }
    if (t4 == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            s_o[warp][g8 * 8 + 2 * i] = acc[i][0];
            s_o[warp][g8 * 8 + 2 * i + 1] = acc[i][2];
        }
    }
    __syncthreads();
    if (tid < 64) {
        const float v = s_o[0][tid] + s_o[1][tid] + s_o[2][tid] + s_o[3][tid];
        out[(size_t) r * d + h * 64 + tid] = Half16<T16>::from_f(v);
    }
}

// ---- cross-attention on a bulk-copy ring --------------------------------------------------------------------------------
// Same arithmetic as cross_attn_kernel<SELF = false> (same thread -> key / dimension mapping, same reduction orders), but the
// contiguous 2 x 187.5 KB K / V blocks of one (window, head) arrive through cp.async.bulk in 16 KB chunks (128 keys) into a
// four-stage shared-memory ring fed by one producer lane: the bare access pattern streams at 7.0 TB/s against 6.6 for
// 16-byte loads (tools/microbench/readbw.cu).  The first four chunks are requested before the programmatic-launch dependency
// resolves (cross K/V and the row descriptors were written before this decoder call began).
constexpr int CB_STAGES = 4, CB_KEYS = 128, CB_CHUNK = CB_KEYS * 128;
constexpr int CB_WARPS = 8;                       // compute warps; + 1 producer warp
constexpr int CBQ_MAX = DEC_CROSS_GROUP_MAX;      // decoder rows of one window served by one CTA
constexpr int CB_THREADS = (CB_WARPS + 1) * 32;

__device__ __forceinline__ bool cb_try_wait(uint64_t * bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"((uint32_t) __cvta_generic_to_shared(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}
__device__ __forceinline__ void cb_wait(uint64_t * bar, uint32_t parity) {          // bounded: trap instead of hanging the GPU
    for (unsigned spins = 0; !cb_try_wait(bar, parity); ++spins)
        if (spins > (1u << 26)) __trap();
}

// NQ > 1: one CTA serves up to NQ consecutive decoder rows that attend to the SAME window (the tokens of a prompt, the beams of
// a beam search): the K / V stream is read once for all of them.  Per row the arithmetic is exactly the NQ = 1 arithmetic.
// KM: the scores of the K sweep on mma.sync fragments (as self_attn_mma_kernel): a warp takes 16 consecutive keys of a chunk, its
// lanes' 16-byte shared-memory loads of two key rows are the A registers, the (up to eight) queries of the CTA are the columns of B
// -- one MMA chain per 16 keys whatever NQ is.  The V sweep stays on CUDA cores (a transposing read of the unswizzled [key][64]
// chunk would be four-way bank-conflicted).  With one CTA per SM (8 sequences per GPU) the kernel is bound by its own instruction
// stream, not by the K/V bytes: ~46 instructions per (4 keys x 8 lanes) in the K sweep become ~6.
template <typename T16, int NQ, bool KM>
__global__ void __launch_bounds__(CB_THREADS, 2)
cross_attn_bulk_kernel(const T16 * __restrict__ q, int ldq, const DecRow * __restrict__ rows, const int2 * __restrict__ groups, int d,
                       size_t layer_off, int T, float kq_scale, int n_phantom, T16 * __restrict__ out, int evict_first) {
    extern __shared__ __align__(128) uint8_t cb_smem[];         // ring [CB_STAGES][CB_CHUNK] | scores [NQ][T_pad] f32
    __shared__ __align__(8) uint64_t b_full[CB_STAGES], b_empty[CB_STAGES];
    __shared__ float s_red[NQ][2 * CB_WARPS];
    __shared__ float s_o[CB_WARPS][64];
    const int T_pad = (T + 31) & ~31;
    float * s_sc = reinterpret_cast<float *>(cb_smem + CB_STAGES * CB_CHUNK);
    const int h = blockIdx.y;
    const int r0 = NQ > 1 ? groups[blockIdx.x].x : (int) blockIdx.x;
    const int nq = NQ > 1 ? groups[blockIdx.x].y : 1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sub = lane & 7, grp = lane >> 3;      // 8 lanes per key row, 4 key rows per warp instruction
    const int nck = (T + CB_KEYS - 1) / CB_KEYS;    // chunks per sweep; chunk c: sweep c / nck (0 K, 1 V), keys (c % nck) * 128 ..
    if (tid == 0) {
        for (int i = 0; i < CB_STAGES; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t) __cvta_generic_to_shared(&b_full[i])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t) __cvta_generic_to_shared(&b_empty[i])), "r"(CB_WARPS));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == CB_WARPS) {
        // ===== producer =====
        if (lane == 0) {
            const DecRow row = rows[r0];
            const uint8_t * blk = reinterpret_cast<const uint8_t *>(reinterpret_cast<const T16 *>(row.cross_kv) + layer_off +
                                                                    (size_t) h * 2 * T * 64);
            // the stream is read once: evict-first, so that it displaces neither itself nor the K prefixes of the CTAs that have
            // not started yet (requested into L2 by the GEMMs before this launch, tc_skinny.cu)
            uint64_t pol;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
            for (int c = 0; c < 2 * nck; ++c) {
                const int s = c % CB_STAGES, sweep = c / nck, j = c - sweep * nck;
                if (c >= CB_STAGES) cb_wait(&b_empty[s], ((c / CB_STAGES) - 1) & 1);
                const uint32_t bytes = (uint32_t) (min(CB_KEYS, T - j * CB_KEYS) * 128);
                const uint32_t bar = (uint32_t) __cvta_generic_to_shared(&b_full[s]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
                if (evict_first)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                                     (uint32_t) __cvta_generic_to_shared(cb_smem + s * CB_CHUNK)),
                                 "l"(blk + (size_t) sweep * T * 128 + (size_t) j * CB_CHUNK), "r"(bytes), "r"(bar), "l"(pol)
                                 : "memory");
                else
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                     (uint32_t) __cvta_generic_to_shared(cb_smem + s * CB_CHUNK)),
                                 "l"(blk + (size_t) sweep * T * 128 + (size_t) j * CB_CHUNK), "r"(bytes), "r"(bar)
                                 : "memory");
                if (c == nck) pdl_trigger();        // with the compute warps' trigger after the K sweep: let the successor in
            }
        }
        return;
    }

    // ===== compute warps =====
    pdl_wait();
    auto sync_compute = [] { asm volatile("bar.sync 1, %0;" ::"n"(CB_WARPS * 32) : "memory"); };
    float mx[NQ];
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi) mx[qi] = -INFINITY;
    if constexpr (KM) {
        static_assert(NQ <= 8, "the queries of a CTA are the eight columns of the B fragment");
        const int g8 = lane >> 2, t4 = lane & 3;
        // B: query g8 of the CTA (zeros beyond nq), dims 8 t4 .. and 32 + 8 t4 .. -- the same dims -> k assignment as the A loads below
        uint32_t Q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (g8 < nq) {
            const T16 * qrow = q + (size_t) (r0 + g8) * ldq + h * 64;
            const uint4 qx = *reinterpret_cast<const uint4 *>(qrow + t4 * 8), qy = *reinterpret_cast<const uint4 *>(qrow + 32 + t4 * 8);
            Q[0] = qx.x; Q[1] = qx.y; Q[2] = qx.z; Q[3] = qx.w; Q[4] = qy.x; Q[5] = qy.y; Q[6] = qy.z; Q[7] = qy.w;
        }
        float m0 = -INFINITY, m1 = -INFINITY;          // running maxima of queries 2 t4 and 2 t4 + 1 over this lane's key rows
        const int qa = 2 * t4, qb = 2 * t4 + 1;
        for (int c = 0; c < nck; ++c) {
            const int s = c % CB_STAGES;
            cb_wait(&b_full[s], (c / CB_STAGES) & 1);
            const uint8_t * stage = cb_smem + s * CB_CHUNK;
            const int kl = warp * 16 + g8, ta = c * CB_KEYS + kl, tb = ta + 8;
            const uint4 z = make_uint4(0, 0, 0, 0);
            const uint4 xa = ta < T ? *reinterpret_cast<const uint4 *>(stage + kl * 128 + t4 * 16) : z;
            const uint4 ya = ta < T ? *reinterpret_cast<const uint4 *>(stage + kl * 128 + 64 + t4 * 16) : z;
            const uint4 xb = tb < T ? *reinterpret_cast<const uint4 *>(stage + (kl + 8) * 128 + t4 * 16) : z;
            const uint4 yb = tb < T ? *reinterpret_cast<const uint4 *>(stage + (kl + 8) * 128 + 64 + t4 * 16) : z;
            const uint32_t Ra[8] = {xa.x, xa.y, xa.z, xa.w, ya.x, ya.y, ya.z, ya.w};
            const uint32_t Rb[8] = {xb.x, xb.y, xb.z, xb.w, yb.x, yb.y, yb.z, yb.w};
            float cc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_16816<T16>(cc, Ra[2 * k], Rb[2 * k], Ra[2 * k + 1], Rb[2 * k + 1], Q[2 * k], Q[2 * k + 1]);
            // cc[0] / cc[1]: key ta x queries qa / qb; cc[2] / cc[3]: key tb
            if (qa < nq) {
                if (ta < T) { const float v = cc[0] * kq_scale; s_sc[qa * T_pad + ta] = v; m0 = fmaxf(m0, v); }
                if (tb < T) { const float v = cc[2] * kq_scale; s_sc[qa * T_pad + tb] = v; m0 = fmaxf(m0, v); }
            }
            if (NQ > 1 && qb < nq) {
                if (ta < T) { const float v = cc[1] * kq_scale; s_sc[qb * T_pad + ta] = v; m1 = fmaxf(m1, v); }
                if (tb < T) { const float v = cc[3] * kq_scale; s_sc[qb * T_pad + tb] = v; m1 = fmaxf(m1, v); }
            }
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t) __cvta_generic_to_shared(&b_empty[s])) : "memory");
        }
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi) mx[qi] = qi == qa ? m0 : (qi == qb ? m1 : -INFINITY);
    } else {
    float qv[NQ][8];
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi) {
        const uint4 u = qi < nq ? *reinterpret_cast<const uint4 *>(q + (size_t) (r0 + qi) * ldq + h * 64 + sub * 8) : make_uint4(0, 0, 0, 0);
        const T16 * e = reinterpret_cast<const T16 *>(&u);
#pragma unroll
        for (int j = 0; j < 8; ++j) qv[qi][j] = Half16<T16>::to_f(e[j]);
    }
    for (int c = 0; c < nck; ++c) {
        const int s = c % CB_STAGES;
        cb_wait(&b_full[s], (c / CB_STAGES) & 1);
        const uint8_t * stage = cb_smem + s * CB_CHUNK;
#pragma unroll
        for (int u = 0; u < CB_KEYS / (4 * CB_WARPS); ++u) {
            const int kl = warp * 4 + grp + 4 * CB_WARPS * u, t = c * CB_KEYS + kl;
            uint4 kb = make_uint4(0, 0, 0, 0);
            if (t < T) kb = *reinterpret_cast<const uint4 *>(stage + kl * 128 + sub * 16);
            const T16 * e = reinterpret_cast<const T16 *>(&kb);
            float kf[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) kf[j] = Half16<T16>::to_f(e[j]);
#pragma unroll
            for (int qi = 0; qi < NQ; ++qi) {
                if (NQ > 1 && qi >= nq) break;
                float acc = 0.0f;
#pragma unroll
                for (int j = 0; j < 8; ++j) acc = fmaf(qv[qi][j], kf[j], acc);
                acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                acc += __shfl_xor_sync(0xffffffffu, acc, 4);
                acc *= kq_scale;
                if (t < T) {
                    if (sub == 0) s_sc[qi * T_pad + t] = acc;
                    mx[qi] = fmaxf(mx[qi], acc);
                }
            }
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t) __cvta_generic_to_shared(&b_empty[s])) : "memory");
    }
    }
    pdl_trigger();
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi) {
        const float m = warp_max(mx[qi]);
        if (lane == 0) s_red[qi][warp] = m;
    }
    sync_compute();
    float inv[NQ];
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi) {
        float m = s_red[qi][0];
#pragma unroll
        for (int i = 1; i < CB_WARPS; ++i) m = fmaxf(m, s_red[qi][i]);
        if (n_phantom > 0) m = fmaxf(m, 0.0f);
        mx[qi] = m;
        float sum = 0.0f;
        if (qi < nq)
            for (int t = tid; t < T; t += CB_WARPS * 32) {
                const float e = expf(s_sc[qi * T_pad + t] - m);
                s_sc[qi * T_pad + t] = e;
                sum += e;
            }
        sum = warp_sum(sum);
        if (lane == 0) s_red[qi][CB_WARPS + warp] = sum;
    }
    sync_compute();
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi) {
        float sum = s_red[qi][CB_WARPS];
#pragma unroll
        for (int i = 1; i < CB_WARPS; ++i) sum += s_red[qi][CB_WARPS + i];
        if (n_phantom > 0) sum += (float) n_phantom * expf(-mx[qi]);
        inv[qi] = 1.0f / sum;
    }

    float o[NQ][8];
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi)
#pragma unroll
        for (int j = 0; j < 8; ++j) o[qi][j] = 0.0f;
    for (int c = nck; c < 2 * nck; ++c) {
        const int s = c % CB_STAGES, c0 = (c - nck) * CB_KEYS;
        cb_wait(&b_full[s], (c / CB_STAGES) & 1);
        const uint8_t * stage = cb_smem + s * CB_CHUNK;
#pragma unroll
        for (int u = 0; u < CB_KEYS / (4 * CB_WARPS); ++u) {
            const int kl = warp * 4 + grp + 4 * CB_WARPS * u, t = c0 + kl;
            const bool ok = t < T;
            uint4 vb = make_uint4(0, 0, 0, 0);
            if (ok) vb = *reinterpret_cast<const uint4 *>(stage + kl * 128 + sub * 16);
            const T16 * e = reinterpret_cast<const T16 *>(&vb);
            float vf[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) vf[j] = Half16<T16>::to_f(e[j]);
#pragma unroll
            for (int qi = 0; qi < NQ; ++qi) {
                if (NQ > 1 && qi >= nq) break;
                const float pr = ok ? Half16<T16>::to_f(Half16<T16>::from_f(s_sc[qi * T_pad + t] * inv[qi])) : 0.0f;
#pragma unroll
                for (int j = 0; j < 8; ++j) o[qi][j] = fmaf(pr, vf[j], o[qi][j]);
            }
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t) __cvta_generic_to_shared(&b_empty[s])) : "memory");
    }
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi) {
        if (NQ > 1 && qi >= nq) break;          // CTA-uniform
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            o[qi][j] += __shfl_xor_sync(0xffffffffu, o[qi][j], 8);
            o[qi][j] += __shfl_xor_sync(0xffffffffu, o[qi][j], 16);
        }
        if (qi > 0) sync_compute();            // s_o is reused per row
        if (grp == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) s_o[warp][sub * 8 + j] = o[qi][j];
        }
        sync_compute();
        if (tid < 64) {
            float v = s_o[0][tid];
#pragma unroll
            for (int i = 1; i < CB_WARPS; ++i) v += s_o[i][tid];
            out[(size_t) (r0 + qi) * d + h * 64 + tid] = Half16<T16>::from_f(v);
        }
    }
}

// ---- cross-attention on an e4m3 K/V pool (opt-in, WHISPER_B200_CROSS_KV=fp8) ------------------------------------------------
// The K/V stream is what bounds the decoder step (at the HBM roof in 16 bits), so this variant halves the bytes: the pool holds
// e4m3 values in self-contained chunks of 128 keys -- [128 keys][64 dims] bytes followed by the chunk's f32 scale (16 bytes with
// padding) -- written by cross_quant_kernel from the 16-bit GEMM output (scale = max |x| of the chunk / 448; a floating-point
// format keeps its relative precision across binades, so a finer scale granularity would buy nothing).  Per (window, layer,
// head): nck K chunks, then nck V chunks.  Same ring / producer / softmax structure and thread -> key / dimension mapping as
// cross_attn_bulk_kernel; the dot products run as HFMA2 on e4m3 -> f16 converted pairs (exact conversion) with short f16
// accumulation chains (4 products per lane for a score, 4 keys per flush for the output) that end in f32 -- far below the e4m3
// quantisation step.  q is pre-multiplied by 2^-4 so that no f16 partial sum can overflow (|k| <= 448).
// NOT the reference's arithmetic (its cross K/V is F16, src/whisper.cpp:942): a reduced-precision storage format, off by default.
constexpr int C8_STAGES = 6, C8_KEYS = 128, C8_DATA = C8_KEYS * 64, C8_CHUNK = C8_DATA + 16;

__device__ __forceinline__ __half2 e4m3x2_to_half2(unsigned short v) {
    unsigned r;
    asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(r) : "h"(v));
    return *reinterpret_cast<__half2 *>(&r);
}

template <int NQ>
__global__ void __launch_bounds__(CB_THREADS, 2)
cross_attn_fp8_kernel(const __half * __restrict__ q, int ldq, const DecRow * __restrict__ rows, const int2 * __restrict__ groups, int d,
                      size_t layer_off, int T, float kq_scale, int n_phantom, __half * __restrict__ out) {
    extern __shared__ __align__(128) uint8_t cb_smem[];         // ring [C8_STAGES][C8_CHUNK] | scores [NQ][T_pad] f32
    __shared__ __align__(8) uint64_t b_full[C8_STAGES], b_empty[C8_STAGES];
    __shared__ float s_red[NQ][2 * CB_WARPS];
    __shared__ float s_o[CB_WARPS][64];
    const int T_pad = (T + 31) & ~31;
    float * s_sc = reinterpret_cast<float *>(cb_smem + ((C8_STAGES * C8_CHUNK + 127) & ~127));
    const int h = blockIdx.y;
    const int r0 = NQ > 1 ? groups[blockIdx.x].x : (int) blockIdx.x;
    const int nq = NQ > 1 ? groups[blockIdx.x].y : 1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sub = lane & 7, grp = lane >> 3;      // 8 lanes per key row (8 bytes each), 4 key rows per warp instruction
    const int nck = (T + C8_KEYS - 1) / C8_KEYS;
    if (tid == 0) {
        for (int i = 0; i < C8_STAGES; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t) __cvta_generic_to_shared(&b_full[i])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t) __cvta_generic_to_shared(&b_empty[i])), "r"(CB_WARPS));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == CB_WARPS) {
        if (lane == 0) {
            const DecRow row = rows[r0];
            const uint8_t * blk = reinterpret_cast<const uint8_t *>(row.cross_kv) + layer_off * 2 + (size_t) h * 2 * nck * C8_CHUNK;
            for (int c = 0; c < 2 * nck; ++c) {
                const int s = c % C8_STAGES;
                if (c >= C8_STAGES) cb_wait(&b_empty[s], ((c / C8_STAGES) - 1) & 1);
                const uint32_t bar = (uint32_t) __cvta_generic_to_shared(&b_full[s]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(C8_CHUNK) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 (uint32_t) __cvta_generic_to_shared(cb_smem + s * C8_CHUNK)),
                             "l"(blk + (size_t) c * C8_CHUNK), "r"(C8_CHUNK), "r"(bar)
                             : "memory");
                if (c == nck) pdl_trigger();
            }
        }
        return;
    }

    pdl_wait();
    __half2 q2[NQ][4];
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi) {
        const uint4 u = qi < nq ? *reinterpret_cast<const uint4 *>(q + (size_t) (r0 + qi) * ldq + h * 64 + sub * 8) : make_uint4(0, 0, 0, 0);
        const __half2 * e = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) q2[qi][j] = __hmul2(e[j], __float2half2_rn(0.0625f));
    }
    auto sync_compute = [] { asm volatile("bar.sync 1, %0;" ::"n"(CB_WARPS * 32) : "memory"); };
    float mx[NQ];
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi) mx[qi] = -INFINITY;
    for (int c = 0; c < nck; ++c) {
        const int s = c % C8_STAGES;
        cb_wait(&b_full[s], (c / C8_STAGES) & 1);
        const uint8_t * stage = cb_smem + s * C8_CHUNK;
        const float sc = *reinterpret_cast<const float *>(stage + C8_DATA) * kq_scale * 16.0f;
#pragma unroll
        for (int u = 0; u < C8_KEYS / (4 * CB_WARPS); ++u) {
            const int kl = warp * 4 + grp + 4 * CB_WARPS * u, t = c * C8_KEYS + kl;
            const uint2 kb = *reinterpret_cast<const uint2 *>(stage + kl * 64 + sub * 8);
            __half2 k2[4];
            k2[0] = e4m3x2_to_half2((unsigned short) (kb.x & 0xffffu));
            k2[1] = e4m3x2_to_half2((unsigned short) (kb.x >> 16));
            k2[2] = e4m3x2_to_half2((unsigned short) (kb.y & 0xffffu));
            k2[3] = e4m3x2_to_half2((unsigned short) (kb.y >> 16));
#pragma unroll
            for (int qi = 0; qi < NQ; ++qi) {
                if (NQ > 1 && qi >= nq) break;
                __half2 a2 = __hmul2(q2[qi][0], k2[0]);
#pragma unroll
                for (int j = 1; j < 4; ++j) a2 = __hfma2(q2[qi][j], k2[j], a2);
                float acc = __low2float(a2) + __high2float(a2);
                acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                acc += __shfl_xor_sync(0xffffffffu, acc, 4);
                acc *= sc;
                if (t < T) {
                    if (sub == 0) s_sc[qi * T_pad + t] = acc;
                    mx[qi] = fmaxf(mx[qi], acc);
                }
            }
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t) __cvta_generic_to_shared(&b_empty[s])) : "memory");
    }
    pdl_trigger();
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi) {
        const float m = warp_max(mx[qi]);
        if (lane == 0) s_red[qi][warp] = m;
    }
    sync_compute();
    float inv[NQ];
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi) {
        float m = s_red[qi][0];
#pragma unroll
        for (int i = 1; i < CB_WARPS; ++i) m = fmaxf(m, s_red[qi][i]);
        if (n_phantom > 0) m = fmaxf(m, 0.0f);
        mx[qi] = m;
        float sum = 0.0f;
        if (qi < nq)
            for (int t = tid; t < T; t += CB_WARPS * 32) {
                const float e = expf(s_sc[qi * T_pad + t] - m);
                s_sc[qi * T_pad + t] = e;
                sum += e;
            }
        sum = warp_sum(sum);
        if (lane == 0) s_red[qi][CB_WARPS + warp] = sum;
    }
    sync_compute();
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi) {
        float sum = s_red[qi][CB_WARPS];
#pragma unroll
        for (int i = 1; i < CB_WARPS; ++i) sum += s_red[qi][CB_WARPS + i];
        if (n_phantom > 0) sum += (float) n_phantom * expf(-mx[qi]);
        inv[qi] = 1.0f / sum;
    }

    float o[NQ][8];
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi)
#pragma unroll
        for (int j = 0; j < 8; ++j) o[qi][j] = 0.0f;
    for (int c = nck; c < 2 * nck; ++c) {
        const int s = c % C8_STAGES, c0 = (c - nck) * C8_KEYS;
        cb_wait(&b_full[s], (c / C8_STAGES) & 1);
        const uint8_t * stage = cb_smem + s * C8_CHUNK;
        const float sv = *reinterpret_cast<const float *>(stage + C8_DATA);
        __half2 a2[NQ][4];
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi)
#pragma unroll
            for (int j = 0; j < 4; ++j) a2[qi][j] = __float2half2_rn(0.0f);
#pragma unroll
        for (int u = 0; u < C8_KEYS / (4 * CB_WARPS); ++u) {
            const int kl = warp * 4 + grp + 4 * CB_WARPS * u, t = c0 + kl;
            const bool ok = t < T;
            const uint2 vb = *reinterpret_cast<const uint2 *>(stage + kl * 64 + sub * 8);
            __half2 v2[4];
            v2[0] = e4m3x2_to_half2((unsigned short) (vb.x & 0xffffu));
            v2[1] = e4m3x2_to_half2((unsigned short) (vb.x >> 16));
            v2[2] = e4m3x2_to_half2((unsigned short) (vb.y & 0xffffu));
            v2[3] = e4m3x2_to_half2((unsigned short) (vb.y >> 16));
#pragma unroll
            for (int qi = 0; qi < NQ; ++qi) {
                if (NQ > 1 && qi >= nq) break;
                const __half2 pr = __float2half2_rn(ok ? s_sc[qi * T_pad + t] * inv[qi] : 0.0f);
#pragma unroll
                for (int j = 0; j < 4; ++j) a2[qi][j] = __hfma2(pr, v2[j], a2[qi][j]);
            }
        }
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi) {
            if (NQ > 1 && qi >= nq) break;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(a2[qi][j]);
                o[qi][2 * j] = fmaf(f.x, sv, o[qi][2 * j]);
                o[qi][2 * j + 1] = fmaf(f.y, sv, o[qi][2 * j + 1]);
            }
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t) __cvta_generic_to_shared(&b_empty[s])) : "memory");
    }
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi) {
        if (NQ > 1 && qi >= nq) break;          // CTA-uniform
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            o[qi][j] += __shfl_xor_sync(0xffffffffu, o[qi][j], 8);
            o[qi][j] += __shfl_xor_sync(0xffffffffu, o[qi][j], 16);
        }
        if (qi > 0) sync_compute();
        if (grp == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) s_o[warp][sub * 8 + j] = o[qi][j];
        }
        sync_compute();
        if (tid < 64) {
            float v = s_o[0][tid];
#pragma unroll
            for (int i = 1; i < CB_WARPS; ++i) v += s_o[i][tid];
            out[(size_t) (r0 + qi) * d + h * 64 + tid] = __float2half(v);
        }
    }
}

// 16-bit head-major K/V of one text layer ([window][head][K | V][T][64], the cross-K/V GEMM's output) -> the e4m3 chunk pool.
// One CTA per chunk of 128 keys: thread = (key, half of the 64 dims); keys past T are written as zeros.
__global__ void __launch_bounds__(256)
cross_quant_kernel(const __half * __restrict__ src, uint8_t * __restrict__ dst, int T, int nck, size_t dst_window_bytes) {
    __shared__ float s_max[8];
    const int c = blockIdx.x, hk = blockIdx.y, w = blockIdx.z, n_hk = gridDim.y;
    const int key = threadIdx.x >> 1, half = threadIdx.x & 1, t = c * C8_KEYS + key;
    float v[32];
    if (t < T) {
        const uint4 * p = reinterpret_cast<const uint4 *>(src + (((size_t) w * n_hk + hk) * T + t) * 64 + half * 32);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint4 u = p[i];
            const __half2 * e = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(e[j]);
                v[i * 8 + 2 * j] = f.x;
                v[i * 8 + 2 * j + 1] = f.y;
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0.0f;
    }
    float m = 0.0f;
#pragma unroll
    for (int i = 0; i < 32; ++i) m = fmaxf(m, fabsf(v[i]));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = m;
    __syncthreads();
    m = s_max[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) m = fmaxf(m, s_max[i]);
    const float scale = m > 0.0f ? m * (1.0f / 448.0f) : 1.0f, inv = 1.0f / scale;
    uint8_t * chunk = dst + (size_t) w * dst_window_bytes + ((size_t) hk * nck + c) * C8_CHUNK;
    unsigned pk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        unsigned r = 0;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            unsigned short h2;
            const float a = v[i * 4 + 2 * j] * inv, b = v[i * 4 + 2 * j + 1] * inv;
            asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(h2) : "f"(b), "f"(a));      // first source -> upper byte
            r |= (unsigned) h2 << (16 * j);
        }
        pk[i] = r;
    }
    uint4 * o = reinterpret_cast<uint4 *>(chunk + key * 64 + half * 32);
    o[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    o[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    if (threadIdx.x == 0) *reinterpret_cast<float4 *>(chunk + C8_DATA) = make_float4(scale, 0.0f, 0.0f, 0.0f);
}

// ---- logit rules + greedy selection ------------------------------------------------------------------------
struct ArgMax {
    float v;
    int i;
};
__device__ __forceinline__ ArgMax amax(ArgMax a, ArgMax b) {        // larger value, then lower index
    return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}
__device__ __forceinline__ ArgMax warp_amax(ArgMax a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ArgMax b;
        b.v = __shfl_xor_sync(0xffffffffu, a.v, o);
        b.i = __shfl_xor_sync(0xffffffffu, a.i, o);
        a = amax(a, b);
    }
    return a;
}

template <int NT> __device__ float block_max(float v, float * sh) {
    v = warp_max(v);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = sh[0];
    for (int i = 1; i < NT / 32; ++i) r = fmaxf(r, sh[i]);
    __syncthreads();
    return r;
}
template <int NT> __device__ float block_sum(float v, float * sh) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.0f;
    for (int i = 0; i < NT / 32; ++i) r += sh[i];
    __syncthreads();
    return r;
}
template <int NT> __device__ double block_sum_d(double v, double * sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    for (int i = 0; i < NT / 32; ++i) r += sh[i];
    __syncthreads();
    return r;
}
template <int NT> __device__ ArgMax block_amax(ArgMax a, ArgMax * sh) {
    a = warp_amax(a);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
    __syncthreads();
    ArgMax r = sh[0];
    for (int i = 1; i < NT / 32; ++i) r = amax(r, sh[i]);
    __syncthreads();
    return r;
}

constexpr int SAMPLE_THREADS = 512;
constexpr int SAMPLE_CTAS = 2;                 // CTAs (one cluster) per decoder row
constexpr int SAMPLE_HALF = 26624;             // elements per CTA: 2 * 26624 = 53248 >= n_vocab of every whisper model
constexpr int SAMPLE_SMEM = SAMPLE_HALF * 4;   // the CTA's half row after the rules, f32

struct SampleXchg {                            // one slot per cluster-wide reduction: slots are never reused inside a launch
    float f[4];
    double d, d2;
    ArgMax a[2];
};

// One cluster of two CTAs per decoder row, each CTA owns half of it.  The row (207 KB, L2-resident right after the logits
// GEMM) is read ONCE, thirteen 16-byte loads in flight per thread; the rules of whisper_process_logits are applied in its
// order (`allowed`) on the way into shared memory (suppressed = -inf), and every later pass is a short rolled loop over
// shared memory (a fully unrolled register-resident version was instruction-fetch-bound: 10 k SASS instructions).
// Block-wide results are exchanged through DSMEM and combined in rank order by both CTAs.  The log-softmax, the
// timestamp-mass rule and the probabilities follow the reference's own formulation step by step
// (src/whisper.cpp:6137-6171, 6336-6361); arg-max / timestamp statistics are whisper_sample_token's (6460-6517).
//
// Rows with n_draws > 0 do not take the arg-max but draw from the categorical distribution of the processed row
// (whisper_sample_token with best = false, whisper_sample_token_topk: src/whisper.cpp:6504-6511, 6519-6592).  The reference
// draws through std::discrete_distribution, i.e. index = lower_bound(cp, u) with cp the running sum (in double, index order)
// of probs / sum(probs) and u = generate_canonical<double, 53>(mt19937).  The host draws the u's from each decoder's own
// mt19937 -- the only stateful part -- and the kernel does the rest: probabilities into shared memory, their sum, a block /
// cluster-wide scan of the normalised per-thread sums in double (each thread owns 52 consecutive tokens), and for every u the
// one thread whose scan interval contains it walks its tokens.  Logits never leave the device.
__global__ void __cluster_dims__(SAMPLE_CTAS, 1, 1) __launch_bounds__(SAMPLE_THREADS)
sample_kernel(const float * __restrict__ logits, int ld, const SampleRow * __restrict__ srows,
              const uint32_t * __restrict__ static_mask, SampleParams prm, SampleOut * __restrict__ outs,
              const double * __restrict__ uniforms, DrawOut * __restrict__ draws) {
    extern __shared__ __align__(16) float sh_v[];          // [SAMPLE_HALF]
    __shared__ float sh_f[SAMPLE_THREADS / 32];
    __shared__ double sh_d[SAMPLE_THREADS / 32];
    __shared__ ArgMax sh_a[SAMPLE_THREADS / 32];
    __shared__ uint32_t sh_mask[SAMPLE_HALF / 32];
    __shared__ SampleXchg sh_x[5];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int) cluster.block_rank();
    const int r = blockIdx.x / SAMPLE_CTAS;
    const int tid = threadIdx.x;
    const int V = prm.n_vocab, beg = prm.token_beg, eot = prm.token_eot;
    const int lo = rank * SAMPLE_HALF;                                   // first element of this CTA's half
    const int n4 = (min(max(V - lo, 0), SAMPLE_HALF) + 3) >> 2;          // float4 units that hold at least one valid element
    pdl_trigger();
    // the static suppression mask is written once per whisper_full call, before any decoder launch
    for (int w = tid; w < SAMPLE_HALF / 32; w += SAMPLE_THREADS) {
        const int gw = lo / 32 + w;
        sh_mask[w] = gw < (V + 31) / 32 ? __ldg(static_mask + gw) : 0u;
    }
    pdl_wait();
    const SampleRow sr = srows[r];
    const float * l = logits + (size_t) sr.logits_row * ld;
    __syncthreads();      // sh_mask

    const bool is_initial = sr.n_tokens == 0;
    const bool last_ts = sr.n_tokens > 0 && sr.last >= beg;
    const bool pen_ts = sr.n_tokens < 2 || sr.penult >= beg;
    const int init_lim = (is_initial && prm.max_initial_ts > 0.0f) ? beg + prm.tid0 + 1 : V;
    const int mono_lim = sr.has_ts ? beg + sr.seek_delta / 2 : beg;
    const bool use_temp = sr.temperature > 0.0f;
    const float temp = use_temp ? sr.temperature : 1.0f;
    auto allowed = [&](int i, uint32_t mask_word) {
        bool kill = (mask_word >> (i & 31)) & 1u;
        if (is_initial && prm.suppress_blank && (i == eot || i == prm.token_space)) kill = true;
        if (prm.no_timestamps && i >= beg) kill = true;
        if (last_ts) {
            if (pen_ts) {
                if (i >= beg) kill = true;
            } else {
                if (i < eot) kill = true;
            }
        }
        if (i >= init_lim) kill = true;
        if (i >= beg && i < mono_lim) kill = true;
        return !kill;
    };

    // pass 0: global -> rules -> shared memory; maxima of all / timestamp / text logits
    float mx = -INFINITY, mx_ts = -INFINITY, mx_text = -INFINITY;
    {
        constexpr int NB = SAMPLE_HALF / 4 / SAMPLE_THREADS;             // 13 float4 per thread
        const bool vec_ok = (reinterpret_cast<uintptr_t>(l) & 15) == 0;
        float4 q[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            const int u = tid + SAMPLE_THREADS * j, i0 = lo + 4 * u;
            if (vec_ok && i0 + 3 < V) {
                q[j] = *reinterpret_cast<const float4 *>(l + i0);
            } else {
                q[j].x = i0 < V ? l[i0] : -INFINITY;
                q[j].y = i0 + 1 < V ? l[i0 + 1] : -INFINITY;
                q[j].z = i0 + 2 < V ? l[i0 + 2] : -INFINITY;
                q[j].w = i0 + 3 < V ? l[i0 + 3] : -INFINITY;
            }
        }
#pragma unroll 1
        for (int j = 0; j < NB; ++j) {
            const int u = tid + SAMPLE_THREADS * j, i0 = lo + 4 * u;
            const uint32_t mw = sh_mask[u >> 3];
            float x[4];
            {   // q[j] with a rolled j: select instead of a dynamically indexed register array
                float4 t = q[0];
#pragma unroll
                for (int k = 1; k < NB; ++k)
                    if (j == k) t = q[k];
                x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w;
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int i = i0 + c;
                float y = x[c];
                if (use_temp && i < V) y = y / temp;
                if (i >= V || !allowed(i, mw)) y = -INFINITY;
                x[c] = y;
                mx = fmaxf(mx, y);
                if (i >= beg) mx_ts = fmaxf(mx_ts, y);
                else mx_text = fmaxf(mx_text, y);
            }
            *reinterpret_cast<float4 *>(sh_v + 4 * u) = make_float4(x[0], x[1], x[2], x[3]);
        }
    }
    mx = block_max<SAMPLE_THREADS>(mx, sh_f);            // (its barriers also publish sh_v)
    mx_ts = block_max<SAMPLE_THREADS>(mx_ts, sh_f);
    mx_text = block_max<SAMPLE_THREADS>(mx_text, sh_f);
    const SampleXchg * peer_x = cluster.map_shared_rank(sh_x, rank ^ 1);
    if (tid == 0) {
        sh_x[0].f[0] = mx; sh_x[0].f[1] = mx_ts; sh_x[0].f[2] = mx_text;
    }
    cluster.sync();
    mx = fmaxf(mx, peer_x[0].f[0]);
    mx_ts = fmaxf(mx_ts, peer_x[0].f[1]);
    mx_text = fmaxf(mx_text, peer_x[0].f[2]);

    // log-softmax denominator (whisper_compute_logprobs)
    float se;
    {
        float s4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        for (int u = tid; u < n4; u += SAMPLE_THREADS) {
            const float4 t = *reinterpret_cast<const float4 *>(sh_v + 4 * u);
            if (t.x > -INFINITY) s4[0] += expf(t.x - mx);
            if (t.y > -INFINITY) s4[1] += expf(t.y - mx);
            if (t.z > -INFINITY) s4[2] += expf(t.z - mx);
            if (t.w > -INFINITY) s4[3] += expf(t.w - mx);
        }
        se = (s4[0] + s4[1]) + (s4[2] + s4[3]);
    }
    se = block_sum<SAMPLE_THREADS>(se, sh_f);
    if (tid == 0) sh_x[1].f[0] = se;
    cluster.sync();
    se = rank == 0 ? se + peer_x[1].f[0] : peer_x[1].f[0] + se;
    const float logZ = logf(se) + mx;

    // local element range of the timestamp tokens / of the tokens that can still win
    const int ts_lo = min(max(beg - lo, 0), 4 * n4), ts_hi = 4 * n4;
    // if the probability mass of all timestamps exceeds that of any single text token, only timestamps survive
    bool mask_text = false;
    {
        const float lp_max_ts = mx_ts - logZ;
        float ts = 0.0f;
        for (int e = ts_lo + tid; e < ts_hi; e += SAMPLE_THREADS) {
            const float x = sh_v[e];
            if (x > -INFINITY) ts += expf((x - logZ) - lp_max_ts);
        }
        ts = block_sum<SAMPLE_THREADS>(ts, sh_f);
        if (tid == 0) sh_x[2].f[0] = ts;
        cluster.sync();
        ts = rank == 0 ? ts + peer_x[2].f[0] : peer_x[2].f[0] + ts;
        const float timestamp_logprob = ts > 0.0f ? logf(ts) + lp_max_ts : -INFINITY;
        mask_text = timestamp_logprob > mx_text - logZ;
    }

    if (sr.n_draws > 0) {
        // ---- categorical draws ----
        constexpr int PER = SAMPLE_HALF / SAMPLE_THREADS;         // 52 consecutive tokens per thread
        static_assert(PER % 4 == 0 && PER * SAMPLE_THREADS == SAMPLE_HALF, "chunking of the half row");
        const int e0 = tid * PER;
        float4 * v4 = reinterpret_cast<float4 *>(sh_v) + e0 / 4;
        // probabilities (0 for suppressed tokens) replace the logits in shared memory; their sum and the timestamp statistics
        double loc = 0.0, sum_ts = 0.0;
        ArgMax best_ts = {0.0f, 0x7fffffff};
#pragma unroll 1
        for (int j = 0; j < PER / 4; ++j) {
            const float4 t = v4[j];
            float x[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int i = lo + e0 + 4 * j + c;
                float pr = 0.0f;
                if (x[c] > -INFINITY && !(mask_text && i < beg)) pr = expf(x[c] - logZ);
                x[c] = pr;
                loc += (double) pr;
                if (i >= beg && pr > 0.0f) {
                    best_ts = amax(best_ts, ArgMax{pr, i});
                    sum_ts += (double) pr;
                }
            }
            v4[j] = make_float4(x[0], x[1], x[2], x[3]);
        }
        const double cta_total = block_sum_d<SAMPLE_THREADS>(loc, sh_d);
        best_ts = block_amax<SAMPLE_THREADS>(best_ts, sh_a);
        sum_ts = block_sum_d<SAMPLE_THREADS>(sum_ts, sh_d);
        if (tid == 0) {
            sh_x[3].d = cta_total; sh_x[3].d2 = sum_ts; sh_x[3].a[1] = best_ts;
        }
        cluster.sync();
        const double total = rank == 0 ? cta_total + peer_x[3].d : peer_x[3].d + cta_total;
        best_ts = amax(best_ts, peer_x[3].a[1]);
        sum_ts = rank == 0 ? sum_ts + peer_x[3].d2 : peer_x[3].d2 + sum_ts;

        // scan of the normalised per-thread sums.  excl of a thread IS the incl of its predecessor (same value, not a
        // recomputation), across lanes, warps and the two CTAs, so every u has exactly one owner.
        double locq = 0.0;
#pragma unroll 1
        for (int j = 0; j < PER / 4; ++j) {
            const float4 t = v4[j];
            locq += (double) t.x / total;
            locq += (double) t.y / total;
            locq += (double) t.z / total;
            locq += (double) t.w / total;
        }
        const int lane = tid & 31, warp = tid >> 5;
        double v = locq;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        const double vprev = __shfl_up_sync(0xffffffffu, v, 1);
        if (lane == 31) sh_d[warp] = v;
        __syncthreads();
        double wp = 0.0;
        for (int w = 0; w < warp; ++w) wp += sh_d[w];
        double cta_q = 0.0;
        for (int w = 0; w < SAMPLE_THREADS / 32; ++w) cta_q += sh_d[w];
        if (tid == 0) sh_x[4].d = cta_q;
        cluster.sync();
        const double base = rank == 0 ? 0.0 : peer_x[4].d;
        const double excl = base + (lane == 0 ? wp : wp + vprev), incl = base + (wp + v);
        const bool first_thread = rank == 0 && tid == 0, last_thread = rank == SAMPLE_CTAS - 1 && tid == SAMPLE_THREADS - 1;
        for (int dr = 0; dr < sr.n_draws; ++dr) {
            const double u = uniforms[sr.draw_off + dr];
            const bool overflow = last_thread && u > incl;          // cp.back() is forced to 1 in the reference: the last token
            if (!((u > excl && u <= incl) || (first_thread && u <= excl) || overflow)) continue;
            int found = -1, last_nz = -1;
            double c = excl;
            for (int e = 0; e < PER; ++e) {
                const float pr = sh_v[e0 + e];
                if (pr > 0.0f) last_nz = e;
                c += (double) pr / total;
                if (found < 0 && c >= u) found = e;
            }
            int id;
            if (found >= 0) id = lo + e0 + found;
            else if (overflow) id = V - 1;
            else id = lo + e0 + (last_nz >= 0 ? last_nz : PER - 1);      // walk and scan associate differently: last ulp
            if (id >= V) id = V - 1;
            DrawOut o;
            o.id = id;
            o.p = (id >= lo && id < lo + SAMPLE_HALF) ? sh_v[id - lo] : 0.0f;
            float lx = l[id];
            if (use_temp) lx = lx / temp;
            o.plog = o.p > 0.0f ? lx - logZ : -INFINITY;
            draws[sr.draw_off + dr] = o;
        }
        if (tid == 0 && rank == 0) {
            SampleOut o;
            o.id = 0; o.p = 0.0f; o.plog = 0.0f;
            o.tid = best_ts.i == 0x7fffffff ? sr.tid_default : best_ts.i;
            o.pt = (float) ((double) (best_ts.i == 0x7fffffff ? 0.0f : best_ts.v) / (sum_ts + 1e-10));
            o.ptsum = (float) sum_ts;
            o.runner_up = -1; o.gap = INFINITY;
            outs[r] = o;
        }
        cluster.sync();       // the peer may still be reading this CTA's exchange slots
        return;
    }

    // Greedy arg-max = first index of the maximal PROBABILITY expf(logit - logZ) (src/whisper.cpp:6460-6517).  expf is
    // monotone, so the winner is within rounding distance of its thread's largest logit: each thread finds its two largest
    // logits (cheap compares), evaluates probabilities only for the elements that close to its maximum, and the block /
    // cluster arg-max runs on (probability, lower index) exactly as the reference's scan would decide.
    const int u_lo = mask_text ? (ts_lo >> 2) : 0;        // masked text: start at the float4 that holds token_beg
    float b1 = -INFINITY, b2 = -INFINITY;
    int i1 = 0x7fffffff, i2 = 0x7fffffff;
    for (int u = u_lo + tid; u < n4; u += SAMPLE_THREADS) {
        const float4 t = *reinterpret_cast<const float4 *>(sh_v + 4 * u);
        const float x[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int i = lo + 4 * u + c;
            const float y = (mask_text && i < beg) ? -INFINITY : x[c];
            if (y > b1) {               // indices increase along the scan: strict compares keep the lower index on ties
                b2 = b1; i2 = i1;
                b1 = y; i1 = i;
            } else if (y > b2) {
                b2 = y; i2 = i;
            }
        }
    }
    ArgMax best = {0.0f, 0x7fffffff}, best_ts = {0.0f, 0x7fffffff};
    double sum_ts = 0.0;
    if (b1 > -INFINITY) {
        const float margin = 2e-6f * fmaxf(1.0f, fabsf(b1 - logZ));       // > 2 ulp of (logit - logZ) plus expf's error
        for (int u = u_lo + tid; u < n4; u += SAMPLE_THREADS) {
            const float4 t = *reinterpret_cast<const float4 *>(sh_v + 4 * u);
            const float x[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int i = lo + 4 * u + c;
                if (x[c] >= b1 - margin && !(mask_text && i < beg)) {
                    const float pr = expf(x[c] - logZ);
                    if (pr > 0.0f) best = amax(best, ArgMax{pr, i});
                }
            }
        }
    }
    for (int e = ts_lo + tid; e < ts_hi; e += SAMPLE_THREADS) {      // timestamp statistics: every probability is needed
        const float x = sh_v[e];
        if (x > -INFINITY) {
            const float pr = expf(x - logZ);
            if (pr > 0.0f) {
                best_ts = amax(best_ts, ArgMax{pr, lo + e});
                sum_ts += (double) pr;
            }
        }
    }
    best = block_amax<SAMPLE_THREADS>(best, sh_a);
    best_ts = block_amax<SAMPLE_THREADS>(best_ts, sh_a);
    sum_ts = block_sum_d<SAMPLE_THREADS>(sum_ts, sh_d);
    if (tid == 0) {
        sh_x[3].a[0] = best; sh_x[3].a[1] = best_ts; sh_x[3].d = sum_ts;
    }
    cluster.sync();
    best = amax(best, peer_x[3].a[0]);
    best_ts = amax(best_ts, peer_x[3].a[1]);
    sum_ts = rank == 0 ? sum_ts + peer_x[3].d : peer_x[3].d + sum_ts;
    // runner-up by logit (diagnostics): the best of everyone's second and every loser's best
    ArgMax second = (i1 == best.i) ? ArgMax{b2, i2} : ArgMax{b1, i1};
    if (second.v == -INFINITY) second.i = 0x7fffffff;
    second = block_amax<SAMPLE_THREADS>(second, sh_a);
    if (tid == 0) sh_x[4].a[0] = second;
    cluster.sync();
    second = amax(second, peer_x[4].a[0]);
    if (tid == 0 && rank == 0) {
        auto value = [&](int i) {
            float x = l[i];
            if (use_temp) x = x / temp;
            return x;
        };
        SampleOut o;
        o.id = best.i == 0x7fffffff ? 0 : best.i;
        o.p = best.i == 0x7fffffff ? 0.0f : best.v;
        o.plog = best.i == 0x7fffffff ? 0.0f : (value(o.id) - logZ);
        o.tid = best_ts.i == 0x7fffffff ? sr.tid_default : best_ts.i;
        o.pt = (float) ((double) (best_ts.i == 0x7fffffff ? 0.0f : best_ts.v) / (sum_ts + 1e-10));
        o.ptsum = (float) sum_ts;
        if (o.id >= beg) {
            o.tid = o.id;
            o.pt = o.p;
        }
        o.runner_up = second.i == 0x7fffffff ? -1 : second.i;
        o.gap = (best.i != 0x7fffffff && second.i != 0x7fffffff) ? value(best.i) - value(second.i) : INFINITY;
        outs[r] = o;
    }
    cluster.sync();       // the peer may still be reading this CTA's exchange slots
}

// softmax probability of one token on the RAW logits row (no_speech_prob, src/whisper.cpp:7188-7196)
__global__ void __launch_bounds__(SAMPLE_THREADS)
token_prob_kernel(const float * __restrict__ logits, int ld, const SampleRow * __restrict__ srows, int V, int token,
                  float * __restrict__ out) {
    __shared__ float sh_f[SAMPLE_THREADS / 32];
    const float * l = logits + (size_t) srows[blockIdx.x].logits_row * ld;
    float m = -INFINITY;
    for (int i = threadIdx.x; i < V; i += SAMPLE_THREADS) m = fmaxf(m, l[i]);
    m = block_max<SAMPLE_THREADS>(m, sh_f);
    float se = 0.0f;
    for (int i = threadIdx.x; i < V; i += SAMPLE_THREADS) {
        const float v = l[i];
        if (v > -INFINITY) se += expf(v - m);
    }
    se = block_sum<SAMPLE_THREADS>(se, sh_f);
    if (threadIdx.x == 0) out[blockIdx.x] = expf(l[token] - (logf(se) + m));
}

}  // namespace

void dec_embed(DType dt, const void * te, const float * pe, const DecRow * d_rows, int R, int d, float * x,
               cudaStream_t st) {
    if (R <= 0) return;
    if (dt == DType::F16)
        launch_pdl(embed_kernel<__half>, dim3(R), dim3(256), 0, st, reinterpret_cast<const __half *>(te), pe, d_rows, d, x);
    else
        launch_pdl(embed_kernel<__nv_bfloat16>, dim3(R), dim3(256), 0, st, reinterpret_cast<const __nv_bfloat16 *>(te), pe,
                   d_rows, d, x);
    WB_CUDA(cudaGetLastError());
}

void dec_kv_append(const void * qkv, const DecRow * d_rows, int R, int d, size_t layer_off_elems, cudaStream_t st) {
    if (R <= 0) return;
    kv_append_kernel<<<R, 128, 0, st>>>(reinterpret_cast<const uint4 *>(qkv), d_rows, d / 8, layer_off_elems / 8);
    WB_CUDA(cudaGetLastError());
}

void dec_self_attn(DType dt, const void * qkv, const DecRow * d_rows, int R, int d, int n_head, size_t layer_off_elems,
                   int n_ctx, bool fused_append, void * out, cudaStream_t st, int variant) {
    if (R <= 0) return;
    dim3 grid(R, n_head);
    const size_t smem = (size_t) n_ctx * sizeof(float);
    // dot products on mma.sync fragments (self_attn_mma_kernel) unless WHISPER_B200_SELF_MMA=0 (the CUDA-core kernel it replaced)
    static const bool mma_env = !(getenv("WHISPER_B200_SELF_MMA") && atoi(getenv("WHISPER_B200_SELF_MMA")) == 0);
    const bool use_mma = variant < 0 ? mma_env : variant == 1;
    if (use_mma && d == n_head * 64) {
        if (dt == DType::F16)
            launch_pdl(self_attn_mma_kernel<__half>, grid, dim3(128), smem, st, reinterpret_cast<const __half *>(qkv), 3 * d, d_rows, d,
                       layer_off_elems, fused_append ? 1 : 0, reinterpret_cast<__half *>(out));
        else
            launch_pdl(self_attn_mma_kernel<__nv_bfloat16>, grid, dim3(128), smem, st, reinterpret_cast<const __nv_bfloat16 *>(qkv), 3 * d,
                       d_rows, d, layer_off_elems, fused_append ? 1 : 0, reinterpret_cast<__nv_bfloat16 *>(out));
        WB_CUDA(cudaGetLastError());
        return;
    }
    if (dt == DType::F16)
        launch_pdl(cross_attn_kernel<__half, true, false>, grid, dim3(128), smem, st, reinterpret_cast<const __half *>(qkv), 3 * d,
                   d_rows, d, layer_off_elems, 0, 1.0f, 0, fused_append ? 1 : 0, reinterpret_cast<__half *>(out), SplitIn{});
    else
        launch_pdl(cross_attn_kernel<__nv_bfloat16, true, false>, grid, dim3(128), smem, st,
                   reinterpret_cast<const __nv_bfloat16 *>(qkv), 3 * d, d_rows, d, layer_off_elems, 0, 1.0f, 0,
                   fused_append ? 1 : 0, reinterpret_cast<__nv_bfloat16 *>(out), SplitIn{});
    WB_CUDA(cudaGetLastError());
}

void dec_cross_attn(DType dt, const void * q, const DecRow * d_rows, int R, int d, int n_head, size_t layer_off_elems,
                    int T, int n_phantom, void * out, cudaStream_t st, const SplitIn * q_split, const int2 * d_groups, int n_groups) {
    const SplitIn qs = q_split ? *q_split : SplitIn{};
    if (R <= 0) return;
    dim3 grid(R, n_head);
    const float kq_scale = powf(64.0f, -0.25f);
    const size_t smem = (size_t) T * sizeof(float);
    static const bool bulk = !(getenv("WHISPER_B200_CROSS_BULK") && atoi(getenv("WHISPER_B200_CROSS_BULK")) == 0);
    if (bulk && !q_split) {
        const int nq = (d_groups && n_groups > 0) ? CBQ_MAX : 1;
        const size_t bsmem = (size_t) CB_STAGES * CB_CHUNK + (size_t) nq * ((T + 31) & ~31) * sizeof(float);
        // the K sweep's scores run on mma.sync fragments (template parameter KM); WHISPER_B200_CROSS_MMA=0: on CUDA cores
        static const bool k_mma = !(getenv("WHISPER_B200_CROSS_MMA") && atoi(getenv("WHISPER_B200_CROSS_MMA")) == 0);
        if (bsmem <= 100 * 1024) {
            static const int evict_first = !(getenv("WHISPER_B200_CROSS_EVICT") && atoi(getenv("WHISPER_B200_CROSS_EVICT")) == 0);
            const dim3 g(nq > 1 ? n_groups : R, n_head);
            auto launch = [&](auto tag, auto nq_tag, auto km_tag) {
                using T16 = decltype(tag);
                constexpr int NQ_ = decltype(nq_tag)::value;
                constexpr bool KM_ = decltype(km_tag)::value;
                static DeviceOnce set;      // function attributes are per device (one guard per instantiation)
                once_per_device(set, [&] {
                    WB_CUDA(cudaFuncSetAttribute(cross_attn_bulk_kernel<T16, NQ_, KM_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
                });
                launch_pdl(cross_attn_bulk_kernel<T16, NQ_, KM_>, g, dim3(CB_THREADS), bsmem, st, reinterpret_cast<const T16 *>(q), d, d_rows, d_groups, d,
                           layer_off_elems, T, kq_scale, n_phantom, reinterpret_cast<T16 *>(out), evict_first);
            };
            auto by_nq = [&](auto tag, auto km_tag) {
                if (nq > 1) launch(tag, std::integral_constant<int, CBQ_MAX>{}, km_tag);
                else launch(tag, std::integral_constant<int, 1>{}, km_tag);
            };
            auto by_km = [&](auto tag) {
                if (k_mma) by_nq(tag, std::true_type{});
                else by_nq(tag, std::false_type{});
            };
            if (dt == DType::F16) by_km(__half{});
            else by_km(__nv_bfloat16{});
            WB_CUDA(cudaGetLastError());
            return;
        }
    }
    if (dt == DType::F16)
        if (q_split)
            launch_pdl(cross_attn_kernel<__half, false, true>, grid, dim3(128), smem, st, reinterpret_cast<const __half *>(q), d,
                       d_rows, d, layer_off_elems, T, kq_scale, n_phantom, 0, reinterpret_cast<__half *>(out), qs);
        else
            launch_pdl(cross_attn_kernel<__half, false, false>, grid, dim3(128), smem, st, reinterpret_cast<const __half *>(q), d,
                       d_rows, d, layer_off_elems, T, kq_scale, n_phantom, 0, reinterpret_cast<__half *>(out), qs);
    else
        if (q_split)
            launch_pdl(cross_attn_kernel<__nv_bfloat16, false, true>, grid, dim3(128), smem, st,
                       reinterpret_cast<const __nv_bfloat16 *>(q), d, d_rows, d, layer_off_elems, T, kq_scale, n_phantom, 0,
                       reinterpret_cast<__nv_bfloat16 *>(out), qs);
        else
            launch_pdl(cross_attn_kernel<__nv_bfloat16, false, false>, grid, dim3(128), smem, st,
                       reinterpret_cast<const __nv_bfloat16 *>(q), d, d_rows, d, layer_off_elems, T, kq_scale, n_phantom, 0,
                       reinterpret_cast<__nv_bfloat16 *>(out), qs);
    WB_CUDA(cudaGetLastError());
}

size_t cross_fp8_window_bytes(int n_head, int T) { return (size_t) n_head * 2 * ((T + C8_KEYS - 1) / C8_KEYS) * C8_CHUNK; }

void cross_fp8_quantize(const void * kv16, void * dst_layer_win0, int W, int n_head, int T, cudaStream_t st) {
    if (W <= 0) return;
    const int nck = (T + C8_KEYS - 1) / C8_KEYS;
    cross_quant_kernel<<<dim3(nck, 2 * n_head, W), 256, 0, st>>>(reinterpret_cast<const __half *>(kv16), reinterpret_cast<uint8_t *>(dst_layer_win0),
                                                                 T, nck, cross_fp8_window_bytes(n_head, T));
    WB_CUDA(cudaGetLastError());
}

void dec_cross_attn_fp8(const void * q, const DecRow * d_rows, int R, int d, int n_head, size_t layer_off_elems, int T, int n_phantom,
                        void * out, cudaStream_t st, const int2 * d_groups, int n_groups) {
    if (R <= 0) return;
    const float kq_scale = powf(64.0f, -0.25f);
    const int nq = (d_groups && n_groups > 0) ? CBQ_MAX : 1;
    const size_t bsmem = (size_t) ((C8_STAGES * C8_CHUNK + 127) & ~127) + (size_t) nq * ((T + 31) & ~31) * sizeof(float);
    static DeviceOnce set;
    once_per_device(set, [&] {
        WB_CUDA(cudaFuncSetAttribute(cross_attn_fp8_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        WB_CUDA(cudaFuncSetAttribute(cross_attn_fp8_kernel<CBQ_MAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    });
    if (bsmem > 100 * 1024) {
        fprintf(stderr, "%s: audio context %d does not fit the e4m3 cross-attention kernel\n", __func__, T);
        cuda_fail(cudaErrorInvalidValue, "cross_attn_fp8 shared memory", __FILE__, __LINE__);
        return;
    }
    const dim3 g(nq > 1 ? n_groups : R, n_head);
    const __half * qh = reinterpret_cast<const __half *>(q);
    if (nq > 1)
        launch_pdl(cross_attn_fp8_kernel<CBQ_MAX>, g, dim3(CB_THREADS), bsmem, st, qh, d, d_rows, d_groups, d, layer_off_elems, T, kq_scale,
                   n_phantom, reinterpret_cast<__half *>(out));
    else
        launch_pdl(cross_attn_fp8_kernel<1>, g, dim3(CB_THREADS), bsmem, st, qh, d, d_rows, d_groups, d, layer_off_elems, T, kq_scale,
                   n_phantom, reinterpret_cast<__half *>(out));
    WB_CUDA(cudaGetLastError());
}

void dec_sample(const float * logits, int ld, const SampleRow * d_srows, int R, const uint32_t * d_static_mask,
                const SampleParams & prm, SampleOut * d_out, const double * d_uniforms, DrawOut * d_draws, cudaStream_t st) {
    if (R <= 0) return;
    if (prm.n_vocab > SAMPLE_CTAS * SAMPLE_HALF) {       // model_load() rejects such files; never a silent truncation
        cuda_fail(cudaErrorInvalidValue, "n_vocab <= 53248 (greedy selection kernel)", __FILE__, __LINE__);
        return;
    }
    static DeviceOnce set;      // function attributes are per device
    once_per_device(set, [&] {
        WB_CUDA(cudaFuncSetAttribute(sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SAMPLE_SMEM));
    });
    launch_pdl(sample_kernel, dim3(R * SAMPLE_CTAS), dim3(SAMPLE_THREADS), SAMPLE_SMEM, st, logits, ld, d_srows, d_static_mask, prm,
               d_out, d_uniforms, d_draws);
    WB_CUDA(cudaGetLastError());
}

void dec_token_prob(const float * logits, int ld, const SampleRow * d_srows, int R, int n_vocab, int token, float * d_out,
                    cudaStream_t st) {
    if (R <= 0) return;
    token_prob_kernel<<<R, SAMPLE_THREADS, 0, st>>>(logits, ld, d_srows, n_vocab, token, d_out);
    WB_CUDA(cudaGetLastError());
}

}  // namespace wb
