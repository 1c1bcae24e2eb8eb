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

#include <algorithm>

namespace wb {

namespace {

constexpr int CB = 64;                                 // tile edge (M, K per stage; N is 64 or 32)
constexpr int C_THREADS = 128;
constexpr int C_STAGES = 6;
constexpr int C_STAGE_BYTES = 2 * CB * CB * 2;         // X tile + W tile, 16 KB
constexpr int C_SMEM = C_STAGES * C_STAGE_BYTES;       // 96 KB -> two CTAs per SM
// The last ring stage is never a prefetch target, so the row / self-attention phases may use it as scratch while the next
// GEMM's weights are already landing in stages 0 .. C_STAGES-2.
constexpr int C_SCRATCH_OFF = (C_STAGES - 1) * C_STAGE_BYTES;

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
// A (virtual) CTA walks a contiguous range of 64 x nt x 64 units.  nt = 64: stream-K, one f32 partial tile out per output
// tile touched.  nt = 32: "direct" -- the range is exactly one output tile over the full K, so bias / scale / GELU and the
// 16-bit store happen here.  The four warps split the 64 rows (16 each): every accumulator lives in one warp, so a tile is
// finished straight from registers.  Everything the loop needs is held in registers (GemmRegs): the phase descriptors are
// kernel parameters, and this code is inlined at its single call site so they are read from the constant bank once.
struct GemmRegs {
    const void * a, * w;
    int lda, ldw, nt, kpt, tiles, U, G;
};
struct Cursor {          // (k-block, n-tile, m-block) of the next unit a loader will fetch
    int kb, tile, mb;
};
__device__ __forceinline__ GemmRegs gemm_regs(const ChainPhase & ph) {
    GemmRegs g;
    g.a = ph.a; g.w = ph.w; g.lda = ph.lda; g.ldw = ph.ldw; g.nt = ph.nt; g.kpt = ph.g.kpt; g.tiles = ph.g.tiles;
    g.U = ph.g.U; g.G = ph.g.G;
    return g;
}
__device__ __forceinline__ void range_of(const GemmRegs & g, int vc, int & u0, int & nu) {
    u0 = 0; nu = 0;
    if (vc < g.G) {
        u0 = (int) ((unsigned) g.U * (unsigned) vc / (unsigned) g.G);
        nu = (int) ((unsigned) g.U * (unsigned) (vc + 1) / (unsigned) g.G) - u0;
    }
}
__device__ __forceinline__ Cursor cursor_at(const GemmRegs & g, int u) {
    Cursor c;
    const int ot = u / g.kpt;
    c.kb = u - ot * g.kpt;
    c.mb = ot / g.tiles;
    c.tile = ot - c.mb * g.tiles;
    return c;
}
__device__ __forceinline__ void cursor_next(const GemmRegs & g, Cursor & c) {
    if (++c.kb == g.kpt) {
        c.kb = 0;
        if (++c.tile == g.tiles) { c.tile = 0; ++c.mb; }
    }
}
template <typename T16> __device__ __forceinline__ void load_w(const GemmRegs & g, Cursor & c, uint8_t * stage) {
    const int r0 = threadIdx.x >> 3, ch = threadIdx.x & 7;
    const T16 * W = reinterpret_cast<const T16 *>(g.w) + (size_t) (c.tile * g.nt + r0) * g.ldw + c.kb * CB + ch * 8;
    uint8_t * st = stage + CB * CB * 2;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (j * 16 < g.nt) cp16(st + sw(r0 + j * 16, ch), W + (size_t) (j * 16) * g.ldw, true);
    cursor_next(g, c);
}
template <typename T16> __device__ __forceinline__ void load_x(const GemmRegs & g, Cursor & c, int R, uint8_t * stage) {
    const int r0 = threadIdx.x >> 3, ch = threadIdx.x & 7;
    const T16 * X = reinterpret_cast<const T16 *>(g.a) + c.kb * CB + ch * 8;
    const int m0 = c.mb * CB + r0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const bool ok = m0 + j * 16 < R;
        cp16(stage + sw(r0 + j * 16, ch), X + (size_t) (ok ? m0 + j * 16 : 0) * g.lda, ok);
    }
    cursor_next(g, c);
}

__device__ __forceinline__ float block_sum4(float v, float * s_red) {       // 4 warps; safe to call back to back
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    return (s_red[0] + s_red[1]) + (s_red[2] + s_red[3]);
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
    const DecRow row = p.rows[r];
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
            const int ot = (r >> 6) * ph.g.tiles + (c >> 6);
            const int cnt = sg_cta_of(ph.g, ot * ph.g.kpt + ph.g.kpt - 1) - sg_cta_of(ph.g, ot * ph.g.kpt) + 1;
            const float * src = ph.part + ((size_t) ot * ph.g.maxc) * 4096 + (r & 63) * 64 + (c & 63);
            float4 a = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll 1
            for (int j0 = 0; j0 < cnt; j0 += NB) {
                float4 t[NB];
#pragma unroll
                for (int u = 0; u < NB; ++u)
                    t[u] = j0 + u < cnt ? __ldcg(reinterpret_cast<const float4 *>(src + (size_t) (j0 + u) * 4096)) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
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
    const float mean = block_sum4(s, s_red) / (float) d;
    float qq = 0.0f;
#pragma unroll 1
    for (int q = tid; q < nq; q += C_THREADS) {           // every thread re-reads only what it wrote
        const float4 v = rowbuf[q];
        const float a = v.x - mean, b = v.y - mean, c = v.z - mean, e = v.w - mean;
        qq += (a * a + b * b) + (c * c + e * e);
    }
    const float var = block_sum4(qq, s_red) / (float) d;
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
    constexpr int U = 4;
    const int items = p.R * p.H;
    const T16 * qkv = reinterpret_cast<const T16 *>(ph.a);
#pragma unroll 1
    for (int item = blockIdx.x * 4 + warp; item < items; item += gridDim.x * 4) {
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
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ float s_red[4];
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

    // state of the GEMM phase whose first weight stages are in flight (phase index pf)
    int pf = -1, u0 = 0, nu = 0;
    GemmRegs g = {};
    Cursor wcur = {}, xcur = {};
    unsigned target = p.bar_base;
#pragma unroll 1
    for (int i = 0; i < p.n_phase; ++i) {
        // Weights are never written on the device: the first stages of the next GEMM phase are requested before the
        // barrier (or grid dependency) that guards its activations -- one commit group, older than any activation group.
        if (pf < i) {
            pf = p.n_phase;
            for (int j = i; j < p.n_phase; ++j)
                if (p.ph[j].type == CP_GEMM) { pf = j; break; }
            if (pf < p.n_phase) {
                g = gemm_regs(p.ph[pf]);
                range_of(g, blockIdx.x, u0, nu);
                wcur = cursor_at(g, u0);
                xcur = wcur;
#pragma unroll 1
                for (int s = 0; s < C_STAGES - 1 && s < nu; ++s) load_w<T16>(g, wcur, smem + s * C_STAGE_BYTES);
                cp_commit();
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
        if (ph.type == CP_ROW) {
            row_phase<T16>(p.c, ph, s_red, reinterpret_cast<float4 *>(smem + C_SCRATCH_OFF));
        } else if (ph.type == CP_SELF) {
            self_phase<T16>(p.c, ph, reinterpret_cast<float *>(smem + C_SCRATCH_OFF));
        } else if (ph.type == CP_GEMM) {
            const int np_n = g.nt >> 4;                  // 16-column groups per tile: 4 or 2
            const int arow = warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1), acol = lane >> 4;
            const int brow = (lane & 7) + 8 * (lane >> 4), bcol = (lane >> 3) & 1;
            const int gq = lane >> 2, tq = lane & 3;
#pragma unroll 1
            for (int vc = blockIdx.x; vc < g.G; vc += gridDim.x) {
                if (vc != (int) blockIdx.x) {            // further virtual CTAs of this phase: nothing was prefetched
                    range_of(g, vc, u0, nu);
                    wcur = cursor_at(g, u0);
                    xcur = wcur;
#pragma unroll 1
                    for (int s = 0; s < C_STAGES - 1 && s < nu; ++s) load_w<T16>(g, wcur, smem + s * C_STAGE_BYTES);
                    cp_commit();
                }
                Cursor ccur = xcur;                      // unit being multiplied
#pragma unroll 1
                for (int s = 0; s < C_STAGES - 1; ++s) {
                    if (s < nu) load_x<T16>(g, xcur, R, smem + s * C_STAGE_BYTES);
                    cp_commit();
                }
                float acc[8][4];
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0f;
#pragma unroll 1
                for (int it = 0; it < nu; ++it) {
                    cp_wait<C_STAGES - 2>();
                    __syncthreads();
                    if (it + C_STAGES - 1 < nu) {
                        uint8_t * st = smem + ((it + C_STAGES - 1) % C_STAGES) * C_STAGE_BYTES;
                        load_x<T16>(g, xcur, R, st);
                        load_w<T16>(g, wcur, st);
                    }
                    cp_commit();
                    const uint32_t sx = (uint32_t) __cvta_generic_to_shared(smem + (it % C_STAGES) * C_STAGE_BYTES);
                    const uint32_t swt = sx + CB * CB * 2;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        uint32_t a[4];
                        ldsm4(sx + sw(arow, ks * 2 + acol), a[0], a[1], a[2], a[3]);
#pragma unroll
                        for (int np = 0; np < 4; ++np) {
                            if (np < np_n) {
                                uint32_t b0, b1, b2, b3;
                                ldsm4(swt + sw(np * 16 + brow, ks * 2 + bcol), b0, b1, b2, b3);
                                mma<T16>(acc[2 * np], a, b0, b1);
                                mma<T16>(acc[2 * np + 1], a, b2, b3);
                            }
                        }
                    }
                    if (ccur.kb == g.kpt - 1 || it == nu - 1) {
                        const int ot = ccur.mb * g.tiles + ccur.tile;
                        if (g.nt == 64) {        // this CTA's partial tile of output tile `ot`
                            const int first = sg_cta_of(ph.g, ot * g.kpt);
                            float * dst = ph.part + ((size_t) ot * ph.g.maxc + (vc - first)) * 4096 + (warp * 16 + gq) * 64 + 2 * tq;
#pragma unroll
                            for (int n8 = 0; n8 < 8; ++n8) {
                                __stcg(reinterpret_cast<float2 *>(dst + n8 * 8), make_float2(acc[n8][0], acc[n8][1]));
                                __stcg(reinterpret_cast<float2 *>(dst + 8 * 64 + n8 * 8), make_float2(acc[n8][2], acc[n8][3]));
                            }
                        } else {                 // whole tile: bias, scale, GELU, 16-bit store
                            const int m = ccur.mb * CB + warp * 16 + gq, n = ccur.tile * 32 + 2 * tq;
                            T16 * out = reinterpret_cast<T16 *>(ph.out16);
                            const float * bias = ph.bias;
                            const float scale = ph.scale;
                            const int scale_cols = ph.scale_cols, gelu = ph.gelu, ldo = ph.ldo16, ref16 = p.c.ref_f16_gelu;
#pragma unroll 1
                            for (int n8 = 0; n8 < 4; ++n8) {
                                const int col = n + n8 * 8;
                                float v[4];
#pragma unroll
                                for (int e = 0; e < 4; ++e) v[e] = n8 == 0 ? acc[0][e] : n8 == 1 ? acc[1][e] : n8 == 2 ? acc[2][e] : acc[3][e];
                                const float b0 = bias ? __ldg(bias + col) : 0.0f, b1 = bias ? __ldg(bias + col + 1) : 0.0f;
                                v[0] += b0; v[1] += b1; v[2] += b0; v[3] += b1;
                                if (col < scale_cols) { v[0] *= scale; v[1] *= scale; v[2] *= scale; v[3] *= scale; }
                                if (gelu) {
#pragma unroll
                                    for (int e = 0; e < 4; ++e) v[e] = gelu_ref<T16>(v[e], ref16);
                                }
                                union { T16 h[2]; uint32_t u32; } lo, hi;
                                lo.h[0] = Half16<T16>::from_f(v[0]); lo.h[1] = Half16<T16>::from_f(v[1]);
                                hi.h[0] = Half16<T16>::from_f(v[2]); hi.h[1] = Half16<T16>::from_f(v[3]);
                                if (m < R) *reinterpret_cast<uint32_t *>(out + (size_t) m * ldo + col) = lo.u32;
                                if (m + 8 < R) *reinterpret_cast<uint32_t *>(out + (size_t) (m + 8) * ldo + col) = hi.u32;
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0f;
                    }
                    cursor_next(g, ccur);
                }
                cp_wait<0>();
                __syncthreads();
            }
            if (blockIdx.x >= (unsigned) g.G) cp_wait<0>();      // no work here: retire the (empty) prefetch group
        }
    }
    stamp(p.n_phase);
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
    g.tiles = N / 32;
    g.kpt = K / CB;
    g.G = ceil_div(R, CB) * g.tiles;          // virtual CTAs: one whole 64x32 tile each
    g.U = g.G * g.kpt;
    g.maxc = 0;
    return g;
}

SplitGeom chain_geom(int grid, int R, int N, int K, int min_units) {
    SplitGeom g;
    g.tiles = N / CB;
    g.kpt = K / CB;
    const int mblocks = ceil_div(R, CB);
    g.U = mblocks * g.tiles * g.kpt;
    g.G = std::max(1, std::min(grid, g.U / std::max(1, min_units)));
    g.maxc = 1;
    for (int ot = 0; ot < mblocks * g.tiles; ++ot) {
        const int first = (int) ((((long long) ot * g.kpt + 1) * g.G - 1) / g.U);
        const int last = (int) ((((long long) ot * g.kpt + g.kpt) * g.G - 1) / g.U);
        g.maxc = std::max(g.maxc, last - first + 1);
    }
    return g;
}

size_t chain_part_floats(const SplitGeom & g, int R) {
    return (size_t) ceil_div(R, CB) * g.tiles * g.maxc * 4096;
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
