// C-ABI entry points that expose single kernels of the path (declared in include/whisper_b200.h).
// They take HOST buffers, run the CUDA kernel and copy the result back, so the parity tests can compare
// each kernel with the oracle on identical inputs; the *_bench variants time the kernel with CUDA events
// on device-resident data.
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "common.cuh"
#include "mel.h"
#include "dec_chain.h"
#include "dec_kernels.h"
#include "dtw.h"
#include "enc_kernels.h"
#include "skinny_gemm.h"
#include "tc_gemm.h"
#include "tc_skinny.h"
#include "whisper_b200.h"

using namespace wb;

namespace {

struct DevBuf {
    void * p = nullptr;
    explicit DevBuf(size_t n) {
        if (n == 0) n = 16;
        WB_CUDA(cudaMalloc(&p, n));
    }
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    template <typename T> T * as() { return reinterpret_cast<T *>(p); }
};

}  // namespace

extern "C" {

WB200_API int whisper_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

WB200_API int whisper_b200_kernel_log_mel(const float * pcm, int n_samples, const float * filters, int n_mel,
                                          float * mel_out, int mel_cap, int * n_len, int * n_len_org) {
    if (!pcm || n_samples <= 0 || !filters || !n_len || !n_len_org) return -1;
    cuda_clear_failure();
    MelPlan plan;
    if (!mel_plan_init(plan, filters, n_mel, 201)) return -2;
    const MelGeometry g = mel_geometry(n_samples);
    *n_len = g.n_len;
    *n_len_org = g.n_len_org;
    if (!mel_out) return 0;
    if ((long long) mel_cap < (long long) g.n_len * n_mel) return -3;
    DevBuf d_pcm((size_t) n_samples * 4), d_raw((size_t) n_mel * g.stride * 4), d_max(4), d_st(sizeof(MelStream)),
        d_fin((size_t) n_mel * g.n_len * 4);
    WB_CUDA(cudaMemcpy(d_pcm.p, pcm, (size_t) n_samples * 4, cudaMemcpyHostToDevice));
    WB_CUDA(cudaMemset(d_max.p, 0, 4));
    MelStream st;
    st.pcm = d_pcm.as<float>();
    st.n_samples = n_samples;
    st.n_frames_fft = g.n_frames_fft;
    st.out = d_raw.as<float>();
    st.out_stride = g.stride;
    st.max_enc = d_max.as<unsigned>();
    st.pcm_i16 = 0;
    WB_CUDA(cudaMemcpy(d_st.p, &st, sizeof(st), cudaMemcpyHostToDevice));
    mel_launch(plan, d_st.as<MelStream>(), 1, g.n_frames_fft, 0);
    mel_finalize_launch(d_raw.as<float>(), g.stride, g.n_frames_fft, d_max.as<unsigned>(), d_fin.as<float>(), g.n_len,
                        n_mel, 0);
    WB_CUDA(cudaMemcpy(mel_out, d_fin.p, (size_t) n_mel * g.n_len * 4, cudaMemcpyDeviceToHost));
    return cuda_failed() ? -4 : 0;
}

// n_streams independent streams of n_samples each (synthetic, device resident); returns average ms per launch.
WB200_API double whisper_b200_kernel_log_mel_bench(int n_streams, int n_samples, const float * filters, int n_mel,
                                                   int iters, int flush_l2) {
    cuda_clear_failure();
    MelPlan plan;
    if (!mel_plan_init(plan, filters, n_mel, 201)) return -1.0;
    const MelGeometry g = mel_geometry(n_samples);
    std::vector<float> h((size_t) n_samples);
    unsigned s = 12345u;
    for (auto & v : h) {
        s = s * 1664525u + 1013904223u;
        v = ((int) (s >> 8) % 20001 - 10000) / 30000.0f;
    }
    DevBuf d_pcm((size_t) n_streams * n_samples * 4), d_raw((size_t) n_streams * n_mel * g.stride * 4),
        d_max((size_t) n_streams * 4), d_st(sizeof(MelStream) * n_streams), d_flush(flush_l2 ? (256u << 20) : 16);
    std::vector<MelStream> sts(n_streams);
    for (int i = 0; i < n_streams; ++i) {
        WB_CUDA(cudaMemcpy(d_pcm.as<float>() + (size_t) i * n_samples, h.data(), (size_t) n_samples * 4,
                           cudaMemcpyHostToDevice));
        sts[i] = {d_pcm.as<float>() + (size_t) i * n_samples, n_samples, g.n_frames_fft,
                  d_raw.as<float>() + (size_t) i * n_mel * g.stride, g.stride, d_max.as<unsigned>() + i, 0};
    }
    WB_CUDA(cudaMemcpy(d_st.p, sts.data(), sizeof(MelStream) * n_streams, cudaMemcpyHostToDevice));
    WB_CUDA(cudaMemset(d_max.p, 0, (size_t) n_streams * 4));
    cudaEvent_t e0, e1;
    WB_CUDA(cudaEventCreate(&e0));
    WB_CUDA(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) mel_launch(plan, d_st.as<MelStream>(), n_streams, g.n_frames_fft, 0);
    double total = 0.0;
    for (int i = 0; i < iters; ++i) {
        if (flush_l2) WB_CUDA(cudaMemsetAsync(d_flush.p, i, 256u << 20, 0));
        WB_CUDA(cudaEventRecord(e0, 0));
        mel_launch(plan, d_st.as<MelStream>(), n_streams, g.n_frames_fft, 0);
        WB_CUDA(cudaEventRecord(e1, 0));
        WB_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        WB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        total += ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return cuda_failed() ? -1.0 : total / iters;
}

static int gemm_hook(int skinny /* 0 tc_gemm, 1 skinny_gemm, 2 tc_skinny_gemm */, int dtype, int M, int N, int K, const uint16_t * a, const uint16_t * w, const float * bias,
                     float scale, int scale_cols, int gelu, const float * pos, int pos_rows, const float * resid,
                     uint16_t * out16, float * out32) {
    cuda_clear_failure();
    const int ldo = round_up(N, 8);
    DevBuf d_a((size_t) M * K * 2), d_w((size_t) N * K * 2), d_bias((size_t) N * 4), d_pos((size_t) pos_rows * N * 4),
        d_res((size_t) M * ldo * 4), d_o16((size_t) M * ldo * 2), d_o32((size_t) M * ldo * 4);
    WB_CUDA(cudaMemcpy(d_a.p, a, (size_t) M * K * 2, cudaMemcpyHostToDevice));
    WB_CUDA(cudaMemcpy(d_w.p, w, (size_t) N * K * 2, cudaMemcpyHostToDevice));
    if (bias) WB_CUDA(cudaMemcpy(d_bias.p, bias, (size_t) N * 4, cudaMemcpyHostToDevice));
    if (pos) WB_CUDA(cudaMemcpy(d_pos.p, pos, (size_t) pos_rows * N * 4, cudaMemcpyHostToDevice));
    if (resid)
        WB_CUDA(cudaMemcpy2D(d_res.p, (size_t) ldo * 4, resid, (size_t) N * 4, (size_t) N * 4, M, cudaMemcpyHostToDevice));
    WB_CUDA(cudaMemset(d_o16.p, 0xff, (size_t) M * ldo * 2));
    WB_CUDA(cudaMemset(d_o32.p, 0xff, (size_t) M * ldo * 4));
    GemmArgs g;
    g.dtype = dtype == 1 ? DType::BF16 : DType::F16;
    g.M = M; g.N = N; g.K = K;
    g.a = d_a.p; g.lda = K;
    g.w = d_w.p; g.ldw = K;
    g.bias = bias ? d_bias.as<float>() : nullptr;
    g.scale = scale; g.scale_cols = scale_cols;
    g.gelu = gelu != 0;
    g.pos = pos ? d_pos.as<float>() : nullptr; g.pos_rows = pos_rows;
    g.resid = resid ? d_res.as<float>() : nullptr; g.ldr = ldo;
    g.out16 = out16 ? d_o16.p : nullptr; g.ldo16 = ldo;
    g.out32 = out32 ? d_o32.as<float>() : nullptr; g.ldo32 = ldo;
    SkinnyWorkspace sws;
    if (skinny == 2) {
        if (!tc_skinny_gemm(g, 0)) return -2;
    } else if (skinny) {
        // twice: the second launch checks that the arrival counters were left at zero
        if (!skinny_gemm(g, sws, 0) || !skinny_gemm(g, sws, 0)) return -2;
    } else if (!tc_gemm(g, 0)) {
        return -2;
    }
    WB_CUDA(cudaDeviceSynchronize());
    if (out16)
        WB_CUDA(cudaMemcpy2D(out16, (size_t) N * 2, d_o16.p, (size_t) ldo * 2, (size_t) N * 2, M, cudaMemcpyDeviceToHost));
    if (out32)
        WB_CUDA(cudaMemcpy2D(out32, (size_t) N * 4, d_o32.p, (size_t) ldo * 4, (size_t) N * 4, M, cudaMemcpyDeviceToHost));
    return cuda_failed() ? -3 : 0;
}

WB200_API int whisper_b200_kernel_gemm(int dtype, int M, int N, int K, const uint16_t * a, const uint16_t * w,
                                       const float * bias, float scale, int scale_cols, int gelu, const float * pos,
                                       int pos_rows, const float * resid, uint16_t * out16, float * out32) {
    return gemm_hook(0, dtype, M, N, K, a, w, bias, scale, scale_cols, gelu, pos, pos_rows, resid, out16, out32);
}

WB200_API int whisper_b200_kernel_skinny_gemm(int dtype, int M, int N, int K, const uint16_t * a, const uint16_t * w,
                                              const float * bias, float scale, int scale_cols, int gelu,
                                              const float * resid, uint16_t * out16, float * out32) {
    return gemm_hook(1, dtype, M, N, K, a, w, bias, scale, scale_cols, gelu, nullptr, 0, resid, out16, out32);
}

WB200_API int whisper_b200_kernel_tc_skinny_gemm(int dtype, int M, int N, int K, const uint16_t * a, const uint16_t * w,
                                                 const float * bias, float scale, int scale_cols, int gelu,
                                                 const float * resid, uint16_t * out16, float * out32) {
    return gemm_hook(2, dtype, M, N, K, a, w, bias, scale, scale_cols, gelu, nullptr, 0, resid, out16, out32);
}

// Two decoder-step GEMMs with the LayerNorm between them folded in algebraically (tc_skinny.cu): x = a1 * w1^T + bias1 + resid
// (also emits the 16-bit rows x * gamma and the per-tile row statistics), y = LayerNorm(x; gamma, beta) * w2^T + bias2 computed as
// rstd * ((x * gamma) w2^T - mean * c) + b'.  M <= 128, d and K1 multiples of 64.
WB200_API int whisper_b200_kernel_ln_gemm_pair(int dtype, int M, int d, int K1, int N2, const uint16_t * a1, const uint16_t * w1,
                                               const float * bias1, const float * resid, const float * gamma, const float * beta,
                                               float eps, const uint16_t * w2, float * x_out, float * y_out) {
    if (!a1 || !w1 || !resid || !gamma || !beta || !w2 || !x_out || !y_out) return -1;
    cuda_clear_failure();
    const int ld2 = round_up(N2, 8);
    // the fold vectors, as model.cu prepares them at load (double accumulation over the 16-bit weights)
    std::vector<float> c(N2), bf(N2);
    for (int n = 0; n < N2; ++n) {
        double cs = 0.0, bs = 0.0;
        for (int k = 0; k < d; ++k) {
            float wv;
            if (dtype == 1) {
                const unsigned u = (unsigned) w2[(size_t) n * d + k] << 16;
                memcpy(&wv, &u, 4);
            } else {
                __half hh;
                memcpy(&hh, &w2[(size_t) n * d + k], 2);
                wv = __half2float(hh);
            }
            cs += (double) wv * gamma[k];
            bs += (double) wv * beta[k];
        }
        c[n] = (float) cs;
        bf[n] = (float) bs;
    }
    DevBuf d_a((size_t) M * K1 * 2), d_w1((size_t) d * K1 * 2), d_b((size_t) d * 4), d_x((size_t) M * d * 4), d_g((size_t) d * 4),
        d_h((size_t) M * d * 2), d_w2((size_t) N2 * d * 2), d_y((size_t) M * ld2 * 4), d_part((size_t) (d / 64 + 1) * M * sizeof(float2)),
        d_c((size_t) ld2 * 4), d_bf((size_t) ld2 * 4);
    WB_CUDA(cudaMemcpy(d_a.p, a1, (size_t) M * K1 * 2, cudaMemcpyHostToDevice));
    WB_CUDA(cudaMemcpy(d_w1.p, w1, (size_t) d * K1 * 2, cudaMemcpyHostToDevice));
    if (bias1) WB_CUDA(cudaMemcpy(d_b.p, bias1, (size_t) d * 4, cudaMemcpyHostToDevice));
    WB_CUDA(cudaMemcpy(d_x.p, resid, (size_t) M * d * 4, cudaMemcpyHostToDevice));
    WB_CUDA(cudaMemcpy(d_g.p, gamma, (size_t) d * 4, cudaMemcpyHostToDevice));
    WB_CUDA(cudaMemcpy(d_w2.p, w2, (size_t) N2 * d * 2, cudaMemcpyHostToDevice));
    WB_CUDA(cudaMemcpy(d_c.p, c.data(), (size_t) N2 * 4, cudaMemcpyHostToDevice));
    WB_CUDA(cudaMemcpy(d_bf.p, bf.data(), (size_t) N2 * 4, cudaMemcpyHostToDevice));
    GemmArgs g1;
    g1.dtype = dtype == 1 ? DType::BF16 : DType::F16; g1.M = M; g1.N = d; g1.K = K1; g1.a = d_a.p; g1.lda = K1; g1.w = d_w1.p; g1.ldw = K1;
    g1.bias = bias1 ? d_b.as<float>() : nullptr; g1.resid = d_x.as<float>(); g1.ldr = d; g1.out32 = d_x.as<float>(); g1.ldo32 = d;
    g1.ln_part_out = d_part.as<float2>(); g1.out16 = d_h.p; g1.ldo16 = d; g1.out16_gamma = d_g.as<float>();
    GemmArgs g2;
    g2.dtype = g1.dtype; g2.M = M; g2.N = N2; g2.K = d; g2.a = d_h.p; g2.lda = d; g2.w = d_w2.p; g2.ldw = d; g2.out32 = d_y.as<float>(); g2.ldo32 = ld2;
    g2.ln_part_in = d_part.as<float2>(); g2.ln_parts = d / 64; g2.ln_colsum = d_c.as<float>(); g2.bias = d_bf.as<float>(); g2.ln_eps = eps;
    if (!tc_skinny_usable(g1) || !tc_skinny_usable(g2)) return -2;
    if (!tc_skinny_gemm(g1, 0) || !tc_skinny_gemm(g2, 0)) return -3;
    WB_CUDA(cudaDeviceSynchronize());
    WB_CUDA(cudaMemcpy(x_out, d_x.p, (size_t) M * d * 4, cudaMemcpyDeviceToHost));
    WB_CUDA(cudaMemcpy2D(y_out, (size_t) N2 * 4, d_y.p, (size_t) ld2 * 4, (size_t) N2 * 4, M, cudaMemcpyDeviceToHost));
    return cuda_failed() ? -4 : 0;
}

// Device-resident GEMM timing (random-ish operands); returns average ms.  M <= 128 times the weight-streaming kernel.
WB200_API double whisper_b200_kernel_gemm_bench(int dtype, int M, int N, int K, int gelu, int iters) {
    cuda_clear_failure();
    DevBuf d_a((size_t) M * K * 2), d_w((size_t) N * K * 2), d_bias((size_t) N * 4), d_o16((size_t) M * N * 2);
    {
        // pseudo-random operands in [-1, 1): constant data would under-state power draw and over-state clocks
        const size_t blk = 8u << 20;
        std::vector<uint16_t> h(blk);
        unsigned s = 777u;
        for (auto & v : h) {
            s = s * 1664525u + 1013904223u;
            const float f = ((int) (s >> 9) % 4096 - 2048) / 2048.0f;
            if (dtype == 1) {
                unsigned u;
                memcpy(&u, &f, 4);
                v = (uint16_t) (u >> 16);
            } else {
                __half hh = __float2half(f);
                memcpy(&v, &hh, 2);
            }
        }
        auto fill = [&](void * dst, size_t n_elems) {
            for (size_t off = 0; off < n_elems; off += blk) {
                const size_t n = n_elems - off < blk ? n_elems - off : blk;
                WB_CUDA(cudaMemcpy((uint16_t *) dst + off, h.data(), n * 2, cudaMemcpyHostToDevice));
            }
        };
        fill(d_a.p, (size_t) M * K);
        fill(d_w.p, (size_t) N * K);
    }
    WB_CUDA(cudaMemset(d_bias.p, 0, (size_t) N * 4));
    GemmArgs g;
    g.dtype = dtype == 1 ? DType::BF16 : DType::F16;
    g.M = M; g.N = N; g.K = K;
    g.a = d_a.p; g.lda = K;
    g.w = d_w.p; g.ldw = K;
    g.bias = d_bias.as<float>();
    g.gelu = gelu != 0;
    g.out16 = d_o16.p; g.ldo16 = N;
    cudaEvent_t e0, e1;
    WB_CUDA(cudaEventCreate(&e0));
    WB_CUDA(cudaEventCreate(&e1));
    SkinnyWorkspace sws;
    static const bool tcs = !(getenv("WHISPER_B200_TC_SKINNY") && atoi(getenv("WHISPER_B200_TC_SKINNY")) == 0);
    auto run = [&]() { return M <= 128 ? (tcs ? tc_skinny_gemm(g, 0) : skinny_gemm(g, sws, 0)) : tc_gemm(g, 0); };
    for (int i = 0; i < 3; ++i)
        if (!run()) return -1.0;
    WB_CUDA(cudaEventRecord(e0, 0));
    for (int i = 0; i < iters; ++i) run();
    WB_CUDA(cudaEventRecord(e1, 0));
    WB_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    WB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return cuda_failed() ? -1.0 : ms / iters;
}

// Back-to-back launches of one decoder-step kernel at full clocks (events around the whole loop): average microseconds.
// which: 0 layernorm [R][d], 1 cross-attention (R rows, d, 1500 keys), 2 self-attention at position `aux`, 3 kv append
WB200_API double whisper_b200_kernel_step_bench(int which, int dtype, int R, int d, int aux, int iters) {
    cuda_clear_failure();
    const DType dt = dtype == 1 ? DType::BF16 : DType::F16;
    const int H = d / 64, n_ctx = 448;
    DevBuf x((size_t) R * d * 4), g((size_t) d * 4), y((size_t) R * 3 * d * 2), out((size_t) R * d * 2),
        cross((size_t) R * 1500 * 2 * d * 2), selfkv((size_t) R * n_ctx * 2 * d * 2), rows((size_t) R * sizeof(DecRow));
    WB_CUDA(cudaMemset(x.p, 0, (size_t) R * d * 4));
    WB_CUDA(cudaMemset(g.p, 0, (size_t) d * 4));
    WB_CUDA(cudaMemset(y.p, 0, (size_t) R * 3 * d * 2));
    WB_CUDA(cudaMemset(cross.p, 0, (size_t) R * 1500 * 2 * d * 2));
    WB_CUDA(cudaMemset(selfkv.p, 0, (size_t) R * n_ctx * 2 * d * 2));
    std::vector<DecRow> hr(R);
    for (int r = 0; r < R; ++r)
        hr[r] = {0, aux, (char *) selfkv.p + (size_t) r * n_ctx * 2 * d * 2, (const char *) cross.p + (size_t) r * 1500 * 2 * d * 2};
    WB_CUDA(cudaMemcpy(rows.p, hr.data(), R * sizeof(DecRow), cudaMemcpyHostToDevice));
    auto run = [&]() {
        switch (which) {
            case 0: layernorm(dt, x.as<float>(), d, g.as<float>(), g.as<float>(), 1e-5f, R, d, out.p, d, nullptr, 0, nullptr, 0); break;
            case 1: dec_cross_attn(dt, out.p, rows.as<DecRow>(), R, d, H, 0, 1500, 0, y.p, 0); break;
            case 2: dec_self_attn(dt, y.p, rows.as<DecRow>(), R, d, H, 0, n_ctx, true, out.p, 0); break;
            default: dec_kv_append(y.p, rows.as<DecRow>(), R, d, 0, 0); break;
        }
    };
    cudaEvent_t e0, e1;
    WB_CUDA(cudaEventCreate(&e0));
    WB_CUDA(cudaEventCreate(&e1));
    for (int i = 0; i < 5; ++i) run();
    WB_CUDA(cudaEventRecord(e0, 0));
    for (int i = 0; i < iters; ++i) run();
    WB_CUDA(cudaEventRecord(e1, 0));
    WB_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    WB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return cuda_failed() ? -1.0 : 1e3 * ms / iters;
}

// Masked self-attention of R decoder rows over explicit caches: row r sits at position pos[r] of its own cache [n_ctx][2d] (K | V);
// qkv [R][3d] carries the step's projections (K / V of position pos[r] are taken from it and appended when fused_append is set,
// otherwise the cache must already hold them).  variant: 0 CUDA-core kernel, 1 mma.sync kernel, -1 the default.
WB200_API int whisper_b200_kernel_self_attn(int dtype, int R, int d, int n_ctx, const int * pos, const uint16_t * qkv, const uint16_t * cache,
                                            int fused_append, int variant, uint16_t * out, uint16_t * cache_out) {
    if (R <= 0 || d <= 0 || d % 64 != 0 || n_ctx <= 0 || !pos || !qkv || !cache || !out) return -1;
    for (int r = 0; r < R; ++r)
        if (pos[r] < 0 || pos[r] >= n_ctx) return -2;
    cuda_clear_failure();
    const DType dt = dtype == 1 ? DType::BF16 : DType::F16;
    const size_t cache_row = (size_t) n_ctx * 2 * d * 2;
    DevBuf d_qkv((size_t) R * 3 * d * 2), d_cache((size_t) R * cache_row), d_out((size_t) R * d * 2), d_rows((size_t) R * sizeof(DecRow));
    WB_CUDA(cudaMemcpy(d_qkv.p, qkv, (size_t) R * 3 * d * 2, cudaMemcpyHostToDevice));
    WB_CUDA(cudaMemcpy(d_cache.p, cache, (size_t) R * cache_row, cudaMemcpyHostToDevice));
    std::vector<DecRow> hr(R);
    for (int r = 0; r < R; ++r) hr[r] = {0, pos[r], (char *) d_cache.p + (size_t) r * cache_row, nullptr};
    WB_CUDA(cudaMemcpy(d_rows.p, hr.data(), R * sizeof(DecRow), cudaMemcpyHostToDevice));
    dec_self_attn(dt, d_qkv.p, d_rows.as<DecRow>(), R, d, d / 64, 0, n_ctx, fused_append != 0, d_out.p, 0, variant);
    WB_CUDA(cudaDeviceSynchronize());
    WB_CUDA(cudaMemcpy(out, d_out.p, (size_t) R * d * 2, cudaMemcpyDeviceToHost));
    if (cache_out) WB_CUDA(cudaMemcpy(cache_out, d_cache.p, (size_t) R * cache_row, cudaMemcpyDeviceToHost));
    return cuda_failed() ? -4 : 0;
}

static_assert(sizeof(whisper_b200_sample_row) == sizeof(SampleRow) && sizeof(whisper_b200_sample_params) == sizeof(SampleParams) &&
              sizeof(whisper_b200_sample_out) == sizeof(SampleOut) && sizeof(whisper_b200_draw_out) == sizeof(DrawOut),
              "the hook's C structs mirror dec_kernels.h");

WB200_API int whisper_b200_kernel_sample(const float * logits, int n_logit_rows, const struct whisper_b200_sample_row * rows, int n_rows,
                                         const uint32_t * static_mask, struct whisper_b200_sample_params prm, const double * uniforms,
                                         int n_uniforms, struct whisper_b200_sample_out * out, struct whisper_b200_draw_out * draws) {
    if (!logits || !rows || !out || n_rows <= 0 || n_logit_rows <= 0 || prm.n_vocab <= 0 || (n_uniforms > 0 && (!uniforms || !draws))) return -1;
    for (int r = 0; r < n_rows; ++r)
        if (rows[r].logits_row < 0 || rows[r].logits_row >= n_logit_rows || rows[r].n_draws < 0 ||
            (rows[r].n_draws > 0 && (rows[r].draw_off < 0 || rows[r].draw_off + rows[r].n_draws > n_uniforms))) return -2;
    cuda_clear_failure();
    const int V = prm.n_vocab, ld = round_up(V, 8), words = (V + 31) / 32;
    DevBuf d_l((size_t) n_logit_rows * ld * 4), d_r((size_t) n_rows * sizeof(SampleRow)), d_m((size_t) words * 4),
        d_o((size_t) n_rows * sizeof(SampleOut)), d_u((size_t) n_uniforms * 8), d_d((size_t) n_uniforms * sizeof(DrawOut));
    WB_CUDA(cudaMemcpy2D(d_l.p, (size_t) ld * 4, logits, (size_t) V * 4, (size_t) V * 4, n_logit_rows, cudaMemcpyHostToDevice));
    WB_CUDA(cudaMemcpy(d_r.p, rows, (size_t) n_rows * sizeof(SampleRow), cudaMemcpyHostToDevice));
    if (static_mask) WB_CUDA(cudaMemcpy(d_m.p, static_mask, (size_t) words * 4, cudaMemcpyHostToDevice));
    else WB_CUDA(cudaMemset(d_m.p, 0, (size_t) words * 4));
    if (n_uniforms > 0) WB_CUDA(cudaMemcpy(d_u.p, uniforms, (size_t) n_uniforms * 8, cudaMemcpyHostToDevice));
    SampleParams sp;
    memcpy(&sp, &prm, sizeof(sp));
    dec_sample(d_l.as<float>(), ld, d_r.as<SampleRow>(), n_rows, d_m.as<uint32_t>(), sp, d_o.as<SampleOut>(), d_u.as<double>(),
               d_d.as<DrawOut>(), 0);
    WB_CUDA(cudaDeviceSynchronize());
    WB_CUDA(cudaMemcpy(out, d_o.p, (size_t) n_rows * sizeof(SampleOut), cudaMemcpyDeviceToHost));
    if (n_uniforms > 0) WB_CUDA(cudaMemcpy(draws, d_d.p, (size_t) n_uniforms * sizeof(DrawOut), cudaMemcpyDeviceToHost));
    return cuda_failed() ? -4 : 0;
}

// Host-only hook: the alignment stage of the DTW token timestamps (csrc/dtw.cu::dtw_align) on explicit head probabilities.
WB200_API int whisper_b200_dtw_align(const float * probs, int n_heads, int n_tokens, int T, int n_audio, int skip_front, int medfilt_width,
                                     int * first_out) {
    if (!probs || !first_out || n_heads <= 0 || n_tokens <= skip_front + 1 || n_audio <= 0 || n_audio > T || medfilt_width % 2 == 0) return -1;
    const std::vector<int> first = dtw_align(probs, n_heads, n_tokens, T, n_audio, skip_front, medfilt_width);
    for (size_t i = 0; i < first.size(); ++i) first_out[i] = first[i];
    return (int) first.size();
}

WB200_API int whisper_b200_chain_geometry(int grid, int rows, int N, int K, int min_units, int direct, int * out) {
    if (!out || grid <= 0 || rows <= 0 || rows > 128 || N <= 0 || K <= 0 || N % SG_TILE_COLS != 0 || K % 64 != 0) return -1;
    const SplitGeom g = direct ? chain_geom_direct(rows, N, K) : chain_geom(grid, rows, N, K, min_units);
    out[0] = g.tiles; out[1] = g.kpt; out[2] = g.U; out[3] = g.G; out[4] = g.maxc;
    return 0;
}

}  // extern "C"
