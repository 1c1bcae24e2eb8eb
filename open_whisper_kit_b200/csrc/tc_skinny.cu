// Weight-streaming GEMM for the decoder step on the 5th-generation tensor cores: out = epilogue(X[M,K] * W[N,K]^T), M <= 128.
//
// 64-column output tiles, the K range split over the CTAs of a thread-block cluster (<= 8).  Legacy mma.sync tops out near
// 180 TFLOP/s on this part (measured), which makes the three wide GEMMs of a decoder layer (QKV, MLP up, MLP down: 13 MB of
// weights each) instruction-bound at 14-16 us although their HBM time is 2 us; here the inner product is tcgen05:
//   warp 4 (one lane): TMA producer -- weight tile [64 rows][64 k] and activation tile [M rows][64 k] per k-block into a ring
//                      of stages (128-byte swizzle); the weight tiles of the first stages are requested BEFORE the
//                      programmatic-launch dependency resolves (weights are never written on the device);
//   warp 5 (one lane): MMA issuer   -- 4 x tcgen05.mma 128x64x16 per k-block, f32 accumulator in TMEM (rows >= M of the A
//                      tile are zero-filled by TMA or stale shared memory: they only produce accumulator rows nobody reads).
//                      Its last tcgen05.commit is MULTICAST to every CTA of the cluster: "my operand ring is dead";
//   warps 0-3:         tcgen05.ld of their 32 accumulator lanes (= rows), then the split-K reduction as a PUSH: every output
//                      row has one owner CTA in the cluster (row m -> CTA m * KS / M); a thread stores its row's 64 partial
//                      sums straight from registers into the owner's shared memory (slot = source rank)
//                      with st.async, which reports its bytes to the owner's mbarrier (no fence, no acknowledgement).  The owner adds the KS slots in rank order and
//                      runs the bias / scale / GELU / residual epilogue on its rows.  No partial tile is staged locally and
//                      nobody waits for a remote LOAD: the reduction costs one one-way trip through the cluster network
//                      (round 1 pulled the tiles over DSMEM: 1.4-2.9 us of dependent round trips per GEMM, profiles/r1_tcs_trace.txt).
//
// LayerNorm folded algebraically into the two GEMMs around it (the three LayerNorm launches of a decoder layer disappear):
//   LN(x) W^T + b = rstd * (x*gamma) W^T - rstd * mean * c + b',   c[n] = sum_k gamma[k] W[n][k],  b'[n] = b[n] + sum_k beta[k] W[n][k]
//   producer (the GEMM that writes the residual stream x): also writes the 16-bit rows x * gamma_next (out16_gamma) and, per
//     64-column tile and row, {mean, centred sum of squares} of its f32 output (ln_part_out);
//   consumer (QKV, cross-Q, MLP up): streams those 16-bit rows as its A operand like any other GEMM -- nothing on its critical
//     path -- and the owner CTA of a row combines the row's partial statistics (Chan et al., fixed order) while the MMAs run;
//     the epilogue applies rstd and mean with the per-column sums c (ln_colsum) and takes b' as its bias.
// Replaces the same reference operators (ggml_mul_mat, ggml_norm of whisper_build_graph_decoder, src/whisper.cpp:2520-2799).
#include "tc_skinny.h"

#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>
#include <unordered_map>
#include <vector>

#include "dec_kernels.h"
#include "ptx.cuh"

namespace wb {

namespace {

constexpr int TB = 64;                 // output tile columns and k per stage
constexpr int T_THREADS = 192;
// k-blocks in flight per CTA: the loop is bound by load latency (~1.2 us a round trip), so deeper is faster -- but two CTAs
// of 5 x 16 KB still fit the 164 KB shared-memory carve-out; one more stage moves the SM to the 228 KB carve-out, which
// stays in force for the cross-attention kernel that follows and costs it a third of its speed (28 KB of L1 left).
constexpr int T_STAGES = 5;
constexpr int T_BAR_BYTES = 128;          // mbarriers + TMEM base address; followed by rows_pad x {mean, rstd}

struct TcSkinnyParams {
    int M, N, K, KS, rows_pad;         // rows_pad: 64 or 128 (A-tile rows)
    int recv_off, need_ready;          // where the received partial rows live in shared memory; whether that overlays live ring stages
    const float * bias;
    float scale; int scale_cols;
    int gelu, ref_f16_gelu;
    int w_l2_prefetch;                 // ask L2 for the weight tiles beyond the ring depth before the dependency resolves
    int vec_io;                        // bias / residual / outputs are 16-byte (8-byte for 16-bit) addressable per 4 columns
    const float * resid; int ldr;
    void * out16; int ldo16;
    float * out32; int ldo32;
    unsigned long long * trace;        // development aid (WHISPER_B200_TCS_TRACE): 8 time stamps per CTA, or null
    // LayerNorm fold (GemmArgs)
    float2 * ln_part_out;
    const float * out16_gamma;
    const float2 * ln_part_in; int ln_parts;
    const float * ln_colsum;
    float ln_eps;
    // L2 prefetch of the next cross-attention's K prefix (GemmArgs)
    const DecRow * pf_rows;
    int pf_H, pf_chunks, pf_lo, pf_hi;
    unsigned long long pf_layer_off_bytes; int pf_head_bytes;
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
// same, observing arrivals (and the stores before them) made by other CTAs of the cluster
__device__ __forceinline__ void ts_wait_cluster(uint64_t * bar, uint32_t parity) {
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
        if (spins > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {      // shared::cta address -> shared::cluster address in CTA `rank`
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
// 16 bytes from registers into (another CTA's) shared memory, asynchronously: the store itself reports its 16 bytes to the mbarrier
// `bar` of the destination CTA, so the sender neither fences nor waits for an acknowledgement (a plain st.shared::cluster followed
// by a releasing arrive cost 1.7-2.3 us per thread here: the release waits for every store's round trip)
__device__ __forceinline__ void st_async_f4(uint32_t addr, uint32_t bar, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%2, %3, %4, %5}, [%1];" ::"r"(addr), "r"(bar), "r"(a), "r"(b),
                 "r"(c), "r"(d)
                 : "memory");
}
// arrive on the mbarrier at this offset in every CTA of `mask` once all tcgen05.mma issued so far by this thread have completed
__device__ __forceinline__ void umma_commit_multicast(uint64_t * bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(ptx::smem_u32(bar)),
                 "h"(mask)
                 : "memory");
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
    const int nt = blockIdx.x, ks = blockIdx.y, KS = p.KS;      // cluster = (1, KS, 1): ks is also the CTA's rank in its cluster
    const int n0 = nt * TB;
    const int kblocks = p.K / TB;
    const int kb0 = (int) ((long long) kblocks * ks / KS), kb1 = (int) ((long long) kblocks * (ks + 1) / KS);
    const int nkb = kb1 - kb0;                                  // >= 1 (host: KS <= kblocks)
    const uint32_t x_bytes = (uint32_t) p.rows_pad * 128u, stage_bytes = x_bytes + TB * 128u;
    uint64_t * b_full = reinterpret_cast<uint64_t *>(smem + T_STAGES * stage_bytes);
    uint64_t * b_empty = b_full + T_STAGES;
    uint64_t & b_acc = b_empty[T_STAGES];          // this CTA's accumulator is complete
    uint64_t & b_ready = b_empty[T_STAGES + 1];    // every CTA of the cluster has completed its MMAs (multicast commits)
    uint64_t & b_recv = b_empty[T_STAGES + 2];     // every partial row this CTA owns has arrived
    uint32_t & s_tmem = *reinterpret_cast<uint32_t *>(b_empty + T_STAGES + 3);
    static_assert((2 * T_STAGES + 4) * 8 <= T_BAR_BYTES, "control block");
    float2 * s_stat = reinterpret_cast<float2 *>(reinterpret_cast<uint8_t *>(b_full) + T_BAR_BYTES);      // [n_own] {mean, rstd}
    // Partial rows received from the cluster: [KS sources][own_max rows][64] f32 (16-byte chunks XOR-swizzled by row).  Lives in
    // ring stages no CTA of this launch ever fills when the K slices are short enough (the d x d GEMMs: 2-3 k-blocks per CTA);
    // otherwise it overlays the ring from stage 0, which a peer may only write once this CTA's MMAs have completed (b_ready).
    float * recv = reinterpret_cast<float *>(smem + p.recv_off);

    // rows of the tile this CTA finishes: [own_lo, own_hi); row m belongs to CTA m * KS / M
    const int own_lo = (ks * p.M + KS - 1) / KS, own_hi = ((ks + 1) * p.M + KS - 1) / KS;
    const int n_own = own_hi - own_lo, own_max = (p.M + KS - 1) / KS;
    const int n_units = n_own * (TB / 4);           // epilogue unit = four consecutive columns of one row; thread tid: units tid, tid + 192, ...
    const float * extra = p.ln_part_in ? p.ln_colsum : p.out16_gamma;       // per-column vector of the LayerNorm fold (never both)
    auto load_cols = [&](int u, float4 & b, float4 & r, float4 & e) {       // bias, residual, fold vector of unit u (zeros when absent)
        const int m = own_lo + (u >> 4), n = n0 + ((u & 15) << 2);
        b = r = e = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (u >= n_units || n >= p.N) return;
        if (p.vec_io && n + 3 < p.N) {
            if (p.bias) b = __ldg(reinterpret_cast<const float4 *>(p.bias + n));
            if (extra) e = __ldg(reinterpret_cast<const float4 *>(extra + n));
            if (p.resid) r = *reinterpret_cast<const float4 *>(p.resid + (size_t) m * p.ldr + n);
        } else {
            float b4[4] = {0.0f, 0.0f, 0.0f, 0.0f}, r4[4] = {0.0f, 0.0f, 0.0f, 0.0f}, e4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            for (int i = 0; i < 4; ++i)
                if (n + i < p.N) {
                    if (p.bias) b4[i] = __ldg(p.bias + n + i);
                    if (extra) e4[i] = __ldg(extra + n + i);
                    if (p.resid) r4[i] = p.resid[(size_t) m * p.ldr + n + i];
                }
            b = make_float4(b4[0], b4[1], b4[2], b4[3]);
            r = make_float4(r4[0], r4[1], r4[2], r4[3]);
            e = make_float4(e4[0], e4[1], e4[2], e4[3]);
        }
    };
    // Requested as soon as the predecessor grid is complete, under the main loop: bias / residual / fold vector of the first two
    // units of the thread (all of them in the step's shapes).
    constexpr int PRE = 2;
    float4 pre_b[PRE], pre_r[PRE], pre_e[PRE];
    auto preload = [&]() {
#pragma unroll
        for (int j = 0; j < PRE; ++j) load_cols(tid + j * T_THREADS, pre_b[j], pre_r[j], pre_e[j]);
    };
    // Row statistics of the folded LayerNorm for the rows this CTA owns, one L2 round trip under the main loop: four lanes per
    // row, each combines every fourth partial {mean, centred sum of squares} pair the producer left (all loads issued at once),
    // then the four are merged in lane order (Chan et al.: as accurate as a two-pass variance; every CTA that needs a row derives
    // bit-identical numbers).  Runs on the warps that hold no accumulator rows when there are such (64-row tiles).
    auto row_stats = [&](int t, int n_thr) {
        const int q = t & 3;
        for (int r0 = 0; r0 < n_own; r0 += n_thr >> 2) {
            const int rl = r0 + (t >> 2);
            const bool live = rl < n_own;
            float2 v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
                v[i] = live && 4 * i + q < p.ln_parts ? p.ln_part_in[(size_t) (4 * i + q) * p.M + own_lo + rl] : make_float2(0.0f, 0.0f);
            float mean = 0.0f, m2 = 0.0f, cnt = 0.0f;
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (4 * i + q < p.ln_parts) {
                    const float delta = v[i].x - mean, tot = cnt + (float) TB;
                    mean += delta * ((float) TB / tot);
                    m2 += v[i].y + delta * delta * (cnt * (float) TB / tot);
                    cnt = tot;
                }
#pragma unroll
            for (int o = 1; o <= 2; o <<= 1) {          // (0,1),(2,3) then (01,23): the lower lane's group first
                const float mean_o = __shfl_xor_sync(0xffffffffu, mean, o), m2_o = __shfl_xor_sync(0xffffffffu, m2, o),
                            cnt_o = __shfl_xor_sync(0xffffffffu, cnt, o);
                const bool lo = (q & o) == 0;
                const float mA = lo ? mean : mean_o, qA = lo ? m2 : m2_o, nA = lo ? cnt : cnt_o;
                const float mB = lo ? mean_o : mean, qB = lo ? m2_o : m2, nB = lo ? cnt_o : cnt;
                const float tot = nA + nB, delta = mB - mA;
                mean = tot > 0.0f ? mA + delta * (nB / tot) : 0.0f;
                m2 = tot > 0.0f ? qA + qB + delta * delta * (nA * nB / tot) : 0.0f;
                cnt = tot;
            }
            if (live && q == 0) s_stat[rl] = make_float2(mean, cnt > 0.0f ? 1.0f / sqrtf(m2 / cnt + p.ln_eps) : 0.0f);
        }
    };

    TS_STAMP(0);          // CTA start
    if (tid == 0) {
        if (ptx::smem_u32(smem) & 1023u) __trap();
        for (int s = 0; s < T_STAGES; ++s) {
            ptx::mbar_init(&b_full[s], 1);
            ptx::mbar_init(&b_empty[s], 1);
        }
        ptx::mbar_init(&b_acc, 1);
        ptx::mbar_init(&b_ready, (uint32_t) KS);
        ptx::mbar_init(&b_recv, 1);
        ptx::fence_mbar_init();
        ptx::mbar_arrive_expect_tx(&b_recv, (uint32_t) (KS * n_own * TB * 4));      // completes when every owned row of every source is in
    }
    if (warp == 0) {
        ptx::tmem_alloc(&s_tmem, 64);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = s_tmem;
    // The peers' barriers must exist before anything is sent to them: every thread arrives here, and waits (cl_wait) right before
    // its first remote operation -- long after the last CTA of the cluster has passed this point.
    if (KS > 1) asm volatile("barrier.cluster.arrive.release;" ::: "memory");
    bool cl_waited = KS <= 1;
    auto cl_wait = [&]() {
        if (!cl_waited) asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
        cl_waited = true;
    };
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
            // the weight tiles that have to wait for a free stage (MLP down: ten k-blocks per CTA, MLP up: seven): into L2 meanwhile,
            // so that the second round through the ring is an L2 round trip, not a DRAM one
            if (p.w_l2_prefetch)
                for (int i = npre; i < nkb; ++i)
                    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(&tm_w), "r"((kb0 + i) * TB), "r"(n0) : "memory");
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
            if (KS > 1 && p.need_ready) {
                cl_wait();
                umma_commit_multicast(&b_ready, (uint16_t) ((1u << KS) - 1u));
            }
        }
        pdl_wait();
        preload();
    } else {
        pdl_wait();
        TS_STAMP(1);      // predecessor grid complete
        preload();
        if (p.ln_part_in) {
            if (p.rows_pad == 64) {
                if (warp >= 2) row_stats(tid - 64, 64);
            } else {
                row_stats(tid, 128);
            }
        }
        if (p.pf_rows && warp == 3 && p.rows_pad == 64) {
            // an idle warp (64-row tiles keep their accumulator rows in warps 0-1): this CTA's share of the K prefix of the coming
            // cross-attention launch, 16 KB per request, into L2 with evict-last priority
            uint64_t pol;
            asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
            const int n_cta = gridDim.x * gridDim.y, c = blockIdx.y * gridDim.x + blockIdx.x;
            for (int id = p.pf_lo + c + lane * n_cta; id < p.pf_hi; id += 32 * n_cta) {
                const int blk = id / p.pf_chunks, j = id - blk * p.pf_chunks;
                const int r = blk / p.pf_H, h = blk - r * p.pf_H;
                const uint8_t * src = reinterpret_cast<const uint8_t *>(p.pf_rows[r].cross_kv) + p.pf_layer_off_bytes +
                                      (size_t) h * p.pf_head_bytes + (size_t) j * 16384;
                const uint32_t bytes = (uint32_t) min(16384, p.pf_head_bytes - j * 16384);      // the block's last request is short
                asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(src), "r"(bytes), "l"(pol) : "memory");
            }
        }
        if (warp * 32 < p.rows_pad) {
            // ===== accumulator row -> registers -> the owner CTA's shared memory =====
            const int row = warp * 32 + lane;
            ts_wait(&b_acc, 0);
            ptx::tc_fence_after();
            TS_STAMP(2);      // accumulator complete
            uint32_t r0[32], r1[32];
            ptx::tmem_ld_32x32(tmem + ((uint32_t) (warp * 32) << 16), r0);
            ptx::tmem_ld_32x32(tmem + ((uint32_t) (warp * 32) << 16) + 32u, r1);
            ptx::tmem_ld_wait();
            ptx::tc_fence_before();
            TS_STAMP(6);      // accumulator row in registers
            if (KS > 1) {
                cl_wait();
                if (p.need_ready) ts_wait_cluster(&b_ready, 0);       // every operand ring of the cluster is dead
            }
            TS_STAMP(7);      // peers ready
            if (row < p.M) {
                const int owner = row * KS / p.M;
                const int rl = row - (owner * p.M + KS - 1) / KS;
                const uint32_t slot = ptx::smem_u32(recv) + (uint32_t) ((ks * own_max + rl) * TB) * 4u;
                const uint32_t dst = KS > 1 ? map_to_cta(slot, (uint32_t) owner) : slot;
                const uint32_t bar_l = ptx::smem_u32(&b_recv), bar = KS > 1 ? map_to_cta(bar_l, (uint32_t) owner) : bar_l;
#pragma unroll
                for (int j = 0; j < 8; ++j) st_async_f4(dst + (uint32_t) ((j ^ (rl & 7)) << 4), bar, r0[4 * j], r0[4 * j + 1], r0[4 * j + 2], r0[4 * j + 3]);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    st_async_f4(dst + (uint32_t) (((8 + j) ^ (rl & 7)) << 4), bar, r1[4 * j], r1[4 * j + 1], r1[4 * j + 2], r1[4 * j + 3]);
            }
            TS_STAMP(3);      // partial rows sent
        }
    }
    cl_wait();
    ts_wait_cluster(&b_recv, 0);
    __syncthreads();      // s_stat is visible; every tcgen05.ld has completed
    TS_STAMP(4);          // all K splits of this CTA's rows are in its shared memory

    // ---- epilogue on the owned rows: partial sums added in rank order, then bias / LayerNorm fold / scale / GELU / residual ----
    T16 * out16 = reinterpret_cast<T16 *>(p.out16);
    auto finish = [&](int u, const float4 & bs, const float4 & rs, const float4 & es) {
        const int rl = u >> 4, c4 = u & 15;
        const int m = own_lo + rl, n = n0 + (c4 << 2);
        if (n >= p.N) return;            // whole-warp uniform per half-warp only when N % 64 != 0 (no statistics then: host check)
        const float * src = recv + (size_t) rl * TB + ((c4 ^ (rl & 7)) << 2);
        float4 part[8];
#pragma unroll
        for (int s = 0; s < 8; ++s)
            if (s < KS) part[s] = *reinterpret_cast<const float4 *>(src + (size_t) s * own_max * TB);
        float x[4] = {part[0].x, part[0].y, part[0].z, part[0].w};
#pragma unroll
        for (int s = 1; s < 8; ++s)
            if (s < KS) { x[0] += part[s].x; x[1] += part[s].y; x[2] += part[s].z; x[3] += part[s].w; }       // fixed rank order
        const float b4[4] = {bs.x, bs.y, bs.z, bs.w}, r4[4] = {rs.x, rs.y, rs.z, rs.w}, e4[4] = {es.x, es.y, es.z, es.w};
        if (p.ln_part_in) {
            const float2 st = s_stat[rl];
#pragma unroll
            for (int i = 0; i < 4; ++i) x[i] = st.y * (x[i] - st.x * e4[i]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float y = x[i] + b4[i];
            if (n + i < p.scale_cols) y *= p.scale;
            if (p.gelu) y = gelu_ts<T16>(y, p.ref_f16_gelu);
            x[i] = y + r4[i];
        }
        if (p.ln_part_out) {
            // statistics of this row's 64 output columns for the LayerNorm folded into the next GEMM: the 16 units of a row sit in
            // 16 consecutive lanes (N % 64 == 0: host check)
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
        float h[4] = {x[0], x[1], x[2], x[3]};
        if (p.out16_gamma) {
#pragma unroll
            for (int i = 0; i < 4; ++i) h[i] *= e4[i];
        }
        if (p.vec_io && n + 3 < p.N) {
            if (p.out32) *reinterpret_cast<float4 *>(p.out32 + (size_t) m * p.ldo32 + n) = make_float4(x[0], x[1], x[2], x[3]);
            if (out16) {
                const T16 hh[4] = {Half16<T16>::from_f(h[0]), Half16<T16>::from_f(h[1]), Half16<T16>::from_f(h[2]), Half16<T16>::from_f(h[3])};
                *reinterpret_cast<uint2 *>(out16 + (size_t) m * p.ldo16 + n) = *reinterpret_cast<const uint2 *>(hh);
            }
        } else {
            for (int i = 0; i < 4; ++i)
                if (n + i < p.N) {
                    if (p.out32) p.out32[(size_t) m * p.ldo32 + n + i] = x[i];
                    if (out16) out16[(size_t) m * p.ldo16 + n + i] = Half16<T16>::from_f(h[i]);
                }
        }
    };
#pragma unroll
    for (int j = 0; j < PRE; ++j)
        if (tid + j * T_THREADS < n_units) finish(tid + j * T_THREADS, pre_b[j], pre_r[j], pre_e[j]);
    for (int u = tid + PRE * T_THREADS; u < n_units; u += T_THREADS) {
        float4 b, r, e;
        load_cols(u, b, r, e);
        finish(u, b, r, e);
    }
    TS_STAMP(5);          // outputs written
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
        fprintf(stderr, "tcs_trace: stamps = s0 start, s1 dep_done, s2 acc_done, s6 acc_in_regs, s7 peers_ready, s3 rows_sent, s4 rows_received, s5 out_written\n");
        for (size_t l = 0; l < recs.size(); ++l) {
            const TraceRec & r = recs[l];
            const int n = r.n_tiles * r.KS;
            const unsigned long long * t = h.data() + l * kPerLaunch;
            unsigned long long t0 = ~0ull, t_end = 0;
            for (int c = 0; c < n; ++c) { if (t[c * 8] && t[c * 8] < t0) t0 = t[c * 8]; if (t[c * 8 + 5] > t_end) t_end = t[c * 8 + 5]; }
            fprintf(stderr, "tcs_trace %3zu M=%d N=%d K=%d grid=%dx%d gap_from_prev_end=%lld |", l, r.M, r.N, r.K, r.n_tiles, r.KS,
                    prev_end ? (long long) (t0 - prev_end) : 0ll);
            for (int s : {0, 1, 2, 6, 7, 3, 4, 5}) {
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
    const bool base = g.M > 0 && g.M <= 128 && g.K % TB == 0 && g.ldw % 8 == 0 && !g.pos && !(reinterpret_cast<uintptr_t>(g.w) & 15) &&
                      g.lda % 8 == 0 && !(reinterpret_cast<uintptr_t>(g.a) & 15);
    if (!base) return false;
    // LayerNorm fold, consumer side: one partial per 64 features of the normalised rows, column sums present
    if (g.ln_part_in && (g.ln_parts != g.K / TB || g.ln_parts > 32 || !g.ln_colsum)) return false;
    // producer side: statistics are per whole 64-column tile of the f32 output; a GEMM is never both (one fold vector per launch)
    if (g.ln_part_out && (g.N % TB != 0 || !g.out32)) return false;
    if (g.out16_gamma && (!g.out16 || g.ln_part_in)) return false;
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
    {   // experiment: fewer, longer K slices for the small GEMMs (fewer senders per owner in the split-K exchange)
        static const int ks_small = getenv("WHISPER_B200_TCS_KS_SMALL") ? atoi(getenv("WHISPER_B200_TCS_KS_SMALL")) : 0;
        if (ks_small > 0 && kblocks <= 4 * T_STAGES && n_tiles <= 32) KS = std::min(KS, std::max(ks_small, ceil_div(kblocks, T_STAGES)));
    }
    while (KS > 1 && (n_tiles * KS > cta_per_sm * n_sm || KS > kblocks / 2)) --KS;

    static std::mutex mu;
    static std::unordered_map<WKey, TMap, WKeyHash> wmaps;
    TMap tm_x, tm_w;
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
    {
        static const bool force_ready = getenv("WHISPER_B200_TCS_READY") && atoi(getenv("WHISPER_B200_TCS_READY")) != 0;
        const int stage_bytes = rows_pad * 128 + TB * 128, nkb_max = ceil_div(kblocks, KS);
        const int recv_bytes = KS * ceil_div(g.M, KS) * TB * 4;
        const bool spare = !force_ready && nkb_max < T_STAGES && (T_STAGES - nkb_max) * stage_bytes >= recv_bytes;
        p.recv_off = spare ? nkb_max * stage_bytes : 0;
        p.need_ready = spare ? 0 : 1;
    }
    p.bias = g.bias; p.scale = g.scale; p.scale_cols = g.scale_cols;
    p.gelu = g.gelu ? 1 : 0; p.ref_f16_gelu = g.dtype == DType::F16 ? 1 : 0;
    p.resid = g.resid; p.ldr = g.ldr; p.out16 = g.out16; p.ldo16 = g.ldo16; p.out32 = g.out32; p.ldo32 = g.ldo32;
    p.ln_part_out = g.ln_part_out; p.out16_gamma = g.out16_gamma; p.ln_part_in = g.ln_part_in; p.ln_parts = g.ln_parts;
    p.ln_colsum = g.ln_colsum; p.ln_eps = g.ln_eps;
    p.pf_rows = nullptr; p.pf_H = p.pf_chunks = p.pf_lo = p.pf_hi = 0; p.pf_layer_off_bytes = 0; p.pf_head_bytes = 0;
    if (g.pf_rows && g.pf_chunks > 0 && g.pf_slots > 0 && rows_pad == 64) {
        const long long tot = (long long) g.pf_R * g.pf_H * g.pf_chunks;
        p.pf_rows = reinterpret_cast<const DecRow *>(g.pf_rows); p.pf_H = g.pf_H; p.pf_chunks = g.pf_chunks;
        p.pf_lo = (int) (tot * g.pf_slot / g.pf_slots); p.pf_hi = (int) (tot * (g.pf_slot + 1) / g.pf_slots);
        p.pf_layer_off_bytes = g.pf_layer_off_bytes; p.pf_head_bytes = g.pf_head_bytes;
    }
    static const bool w_pf = !(getenv("WHISPER_B200_TCS_WPF") && atoi(getenv("WHISPER_B200_TCS_WPF")) == 0);
    p.w_l2_prefetch = w_pf ? 1 : 0;
    static TcsTrace trace;
    p.trace = trace.slot({n_tiles, KS, g.M, g.N, g.K});
    auto al = [](const void * q, uintptr_t a) { return (reinterpret_cast<uintptr_t>(q) & (a - 1)) == 0; };
    p.vec_io = al(g.bias, 16) && al(g.resid, 16) && g.ldr % 4 == 0 && al(g.out32, 16) && g.ldo32 % 4 == 0 && al(g.out16, 8) &&
               g.ldo16 % 4 == 0 && al(g.ln_colsum, 16) && al(g.out16_gamma, 16);
    const int smem = T_STAGES * (rows_pad * 128 + TB * 128) + T_BAR_BYTES + rows_pad * 8;

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
    constexpr int kMaxSmem = T_STAGES * (128 * 128 + TB * 128) + T_BAR_BYTES + 128 * 8;
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
