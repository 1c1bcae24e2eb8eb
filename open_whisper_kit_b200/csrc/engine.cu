// Batched device engine (see engine.h).  Launch sequences for the three stages of the path:
//   run_mel   <- log_mel_spectrogram                                   reference src/whisper.cpp:3170-3260
//   encode    <- whisper_encode_internal (conv + encoder + cross)      reference src/whisper.cpp:2358-2456
//   decode    <- whisper_decode_internal / whisper_build_graph_decoder reference src/whisper.cpp:2458-2978
// Every GEMM is tc_gemm (tcgen05) with its bias / scale / GELU / residual fused; nothing is computed on the host.
#include "engine.h"

#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "tc_gemm.h"
#include "tc_skinny.h"

namespace wb {

bool DeviceBlock::reserve(size_t bytes, bool keep) {
    if (bytes <= cap) return true;
    const size_t ncap = round_up<size_t>(bytes + bytes / 8, 1 << 20);
    void * np = nullptr;
    WB_CUDA(cudaMalloc(&np, ncap));
    if (!np) return false;
    if (p) {
        if (keep && cap) WB_CUDA(cudaMemcpy(np, p, cap, cudaMemcpyDeviceToDevice));
        WB_CUDA(cudaFree(p));
    }
    p = np;
    cap = ncap;
    return true;
}
DeviceBlock::~DeviceBlock() {
    if (p) cudaFree(p);
}

namespace {
__global__ void kv_copy_prefix_kernel(const uint4 * __restrict__ src, uint4 * __restrict__ dst, size_t layer_stride8,
                                      size_t n8) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n8) dst[blockIdx.y * layer_stride8 + i] = src[blockIdx.y * layer_stride8 + i];
}
// blockIdx.z = pair; every pair copies positions [0, n_pos) of every layer (blockIdx.y)
__global__ void kv_copy_batch_kernel(const Engine::KvCopy * __restrict__ list, size_t layer_stride8, int row8) {
    const Engine::KvCopy c = list[blockIdx.z];
    const size_t n8 = (size_t) c.n_pos * row8;
    const uint4 * src = reinterpret_cast<const uint4 *>(c.src) + blockIdx.y * layer_stride8;
    uint4 * dst = reinterpret_cast<uint4 *>(c.dst) + blockIdx.y * layer_stride8;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t) gridDim.x * blockDim.x) dst[i] = src[i];
}
}  // namespace

bool Engine::init(int dev, bool fa) {
    device = dev;
    flash_attn = fa;
    cuda_clear_failure();
    WB_CUDA(cudaSetDevice(device));
    WB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    if (cuda_failed()) return false;
    if (!mel_plan_init(mel_plan, model.filters.data(), model.filt_n_mel, model.filt_n_fft)) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: mel filterbank of the model is not usable (n_mel=%d, n_fft=%d)\n", __func__,
             model.filt_n_mel, model.filt_n_fft);
        return false;
    }
    ld_logits = round_up(model.hp.n_vocab, 8);
    if (const char * e = getenv("WHISPER_B200_CROSS_KV")) {
        if (!strcmp(e, "fp8")) {
            if (model.dtype == DType::F16) {
                cross_fp8 = true;
                wlog(GGML_LOG_LEVEL_WARN, "%s: cross K/V stored as e4m3 (WHISPER_B200_CROSS_KV=fp8): reduced precision, not the reference's F16 cache\n", __func__);
            } else {
                wlog(GGML_LOG_LEVEL_WARN, "%s: WHISPER_B200_CROSS_KV=fp8 needs f16 operands; ignored\n", __func__);
            }
        }
    }
    return true;
}

void Engine::prof_begin(int cls, double work) {
    if (!prof_on) return;
    ProfRec r;
    r.cls = cls;
    r.work = work;
    auto get = [&]() {
        cudaEvent_t e;
        if (!prof_pool.empty()) {
            e = prof_pool.back();
            prof_pool.pop_back();
        } else {
            WB_CUDA(cudaEventCreate(&e));
        }
        return e;
    };
    r.a = get();
    r.b = get();
    WB_CUDA(cudaEventRecord(r.a, stream));
    prof_recs.push_back(r);
}
void Engine::prof_end() {
    if (!prof_on || prof_recs.empty()) return;
    WB_CUDA(cudaEventRecord(prof_recs.back().b, stream));
    if (prof_recs.size() >= 4096) prof_collect();
}
void Engine::prof_collect() {
    if (prof_recs.empty()) return;
    WB_CUDA(cudaStreamSynchronize(stream));
    for (auto & r : prof_recs) {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            prof_ms[r.cls] += ms;
            prof_work[r.cls] += r.work;
            prof_n[r.cls] += 1;
        }
        prof_pool.push_back(r.a);
        prof_pool.push_back(r.b);
    }
    prof_recs.clear();
}
void Engine::prof_reset() {
    prof_collect();
    for (int i = 0; i < PC_COUNT; ++i) {
        prof_ms[i] = prof_work[i] = 0.0;
        prof_n[i] = 0;
    }
}

Engine::~Engine() {
    prof_collect();
    for (auto e : prof_pool) cudaEventDestroy(e);
    for (int i = 0; i < 3; ++i)
        if (h_pinned[i]) cudaFreeHost(h_pinned[i]);
    if (stream) cudaStreamDestroy(stream);
}

void * Engine::pinned(int which, size_t bytes) {
    if (bytes > h_pinned_cap[which]) {
        WB_CUDA(cudaStreamSynchronize(stream));     // nothing may still be reading the old block
        if (h_pinned[which]) cudaFreeHost(h_pinned[which]);
        h_pinned_cap[which] = round_up<size_t>(bytes, 1 << 16);
        WB_CUDA(cudaMallocHost(&h_pinned[which], h_pinned_cap[which]));
    }
    return h_pinned[which];
}

// ---- mel ------------------------------------------------------------------------------------------------
bool Engine::run_mel(const std::vector<MelJob> & jobs) {
    if (jobs.empty()) return true;
    WB_CUDA(cudaSetDevice(device));
    const int n_mel = model.filt_n_mel;
    size_t stage_floats = 0;
    for (const auto & j : jobs)
        if (j.pcm_host) stage_floats += round_up<size_t>(j.i16 ? (j.n_samples + 1) / 2 : j.n_samples, 4);
    if (!pcm_stage.reserve(stage_floats * 4)) return false;
    if (!meta.reserve(jobs.size() * sizeof(MelStream))) return false;
    std::vector<MelStream> sts(jobs.size());
    size_t off = 0;
    int max_frames = 0;
    for (size_t i = 0; i < jobs.size(); ++i) {
        const MelJob & j = jobs[i];
        MelBuf & mb = *j.out;
        const MelGeometry g = mel_geometry(j.n_samples);
        mb.n_mel = n_mel;
        mb.n_len = g.n_len;
        mb.n_len_org = g.n_len_org;
        mb.n_frames_fft = g.n_frames_fft;
        mb.stride = g.stride;
        mb.finalized = false;
        mb.valid = true;
        if (!mb.data.reserve((size_t) n_mel * g.stride * 4) || !mb.max_enc.reserve(4)) return false;
        WB_CUDA(cudaMemsetAsync(mb.max_enc.p, 0, 4, stream));
        const float * src = j.pcm_dev;
        if (j.pcm_host) {
            float * dst = (float *) pcm_stage.p + off;
            WB_CUDA(cudaMemcpyAsync(dst, j.pcm_host, (size_t) j.n_samples * (j.i16 ? 2 : 4), cudaMemcpyHostToDevice, stream));
            src = dst;
            off += round_up<size_t>(j.i16 ? (j.n_samples + 1) / 2 : j.n_samples, 4);
        }
        sts[i] = {src, j.n_samples, g.n_frames_fft, (float *) mb.data.p, g.stride, (unsigned *) mb.max_enc.p, j.i16 ? 1 : 0};
        max_frames = std::max(max_frames, g.n_frames_fft);
    }
    WB_CUDA(cudaMemcpyAsync(meta.p, sts.data(), sts.size() * sizeof(MelStream), cudaMemcpyHostToDevice, stream));
    double mel_bytes = 0.0;
    for (const auto & j : jobs) mel_bytes += (double) j.n_samples * (j.i16 ? 2.0 : 4.0) + (double) (j.n_samples / 160) * n_mel * 4.0;
    prof_begin(PC_MEL, mel_bytes);
    mel_launch(mel_plan, (const MelStream *) meta.p, (int) jobs.size(), max_frames, stream);
    prof_end();
    n_kernel_launches += 1;
    WB_CUDA(cudaStreamSynchronize(stream));   // sts / host PCM go out of scope
    return !cuda_failed();
}

bool Engine::set_mel(MelBuf & out, const float * data, int n_len, int n_mel) {
    WB_CUDA(cudaSetDevice(device));
    out.n_mel = n_mel;
    out.n_len = n_len;
    out.n_len_org = n_len;
    out.n_frames_fft = n_len;
    out.stride = n_len > 0 ? n_len : 1;
    out.finalized = true;
    out.valid = true;
    if (!out.data.reserve((size_t) std::max(1, n_len) * n_mel * 4) || !out.max_enc.reserve(4)) return false;
    if (n_len > 0) WB_CUDA(cudaMemcpy(out.data.p, data, (size_t) n_len * n_mel * 4, cudaMemcpyHostToDevice));
    return !cuda_failed();
}

bool Engine::get_mel(const MelBuf & mel, float * out) {
    if (!mel.valid) return false;
    WB_CUDA(cudaSetDevice(device));
    const size_t n = (size_t) mel.n_mel * mel.n_len;
    if (mel.finalized) {
        WB_CUDA(cudaMemcpy(out, mel.data.p, n * 4, cudaMemcpyDeviceToHost));
    } else {
        DeviceBlock tmp;
        if (!tmp.reserve(n * 4)) return false;
        mel_finalize_launch((const float *) mel.data.p, mel.stride, mel.n_frames_fft, (const unsigned *) mel.max_enc.p,
                            (float *) tmp.p, mel.n_len, mel.n_mel, stream);
        WB_CUDA(cudaMemcpyAsync(out, tmp.p, n * 4, cudaMemcpyDeviceToHost, stream));
        WB_CUDA(cudaStreamSynchronize(stream));
    }
    return !cuda_failed();
}

// ---- encoder ----------------------------------------------------------------------------------------------
bool Engine::size_cross(CrossKV & kv, int n_windows, int T) {
    const int d = model.hp.n_audio_state;
    kv.n_windows = n_windows;
    kv.T = T;
    kv.fp8 = cross_fp8;
    kv.window_bytes = cross_fp8 ? cross_fp8_window_bytes(model.hp.n_audio_head, T) : (size_t) T * 2 * d * 2;
    kv.layer_stride = (size_t) n_windows * kv.window_bytes / 2;
    return kv.data.reserve(kv.layer_stride * model.hp.n_text_layer * 2);
}

bool Engine::encode(const std::vector<EncJob> & jobs, CrossKV & kv, int win0, bool keep_embd32) {
    const int W = (int) jobs.size();
    if (W == 0) return true;
    WB_CUDA(cudaSetDevice(device));
    const auto & hp = model.hp;
    const int d = hp.n_audio_state, H = hp.n_audio_head, T = kv.T;      // audio context: 1500 or params.audio_ctx
    const DType dt = model.dtype;
    const int n_mel = hp.n_mels;
    const int k1 = model.conv1_kpad;
    const size_t M1 = (size_t) W * 2 * T, M = (size_t) W * T;

    size_t need = 0;
    auto sz = [&](size_t b) { need += round_up<size_t>(b, 256); return b; };
    sz(M1 * k1 * 2); sz(M1 * d * 2); sz(M * 3 * d * 2); sz(M * d * 4); sz(M * d * 2); sz(M * 3 * d * 2); sz(M * d * 2);
    sz(M * 4 * d * 2); sz(M * d * 2); sz(W * sizeof(EncWindow));
    static const bool legacy_attn = getenv("WHISPER_B200_ENC_ATTN") && !strcmp(getenv("WHISPER_B200_ENC_ATTN"), "legacy");
    const size_t vt_bytes = legacy_attn ? 0 : enc_attention_tc_scratch_bytes(W, T, H);
    sz(vt_bytes);
    if (!ws.begin(need)) return false;
    void * A1 = ws.take(M1 * k1 * 2);
    void * act1 = ws.take(M1 * d * 2);
    void * A2 = ws.take(M * 3 * d * 2);
    float * x = (float *) ws.take(M * d * 4);
    void * h16 = ws.take(M * d * 2);
    void * qkv = ws.take(M * 3 * d * 2);
    void * att = ws.take(M * d * 2);
    void * mlp = ws.take(M * 4 * d * 2);
    void * enc16 = ws.take(M * d * 2);
    EncWindow * d_wins = (EncWindow *) ws.take(W * sizeof(EncWindow));
    void * vt = vt_bytes ? ws.take(vt_bytes) : nullptr;

    std::vector<EncWindow> wins(W);
    for (int i = 0; i < W; ++i) {
        const MelBuf & mb = *jobs[i].mel;
        if (!mb.valid || mb.n_mel != n_mel) {
            wlog(GGML_LOG_LEVEL_ERROR, "%s: no mel for window %d (call whisper_pcm_to_mel / whisper_set_mel first)\n", __func__, i);
            return false;
        }
        wins[i] = {(const float *) mb.data.p, mb.stride, mb.n_len, mb.n_frames_fft, jobs[i].seek,
                   (const unsigned *) mb.max_enc.p, mb.finalized ? 1 : 0};
    }
    WB_CUDA(cudaMemcpyAsync(d_wins, wins.data(), W * sizeof(EncWindow), cudaMemcpyHostToDevice, stream));

    bool ok = true;
    int gemm_cls = PC_GEMM_CONV;
    auto gemm = [&](const GemmArgs & g) {
        prof_begin(gemm_cls, 2.0 * g.M * (double) g.N * g.K);
        ok = ok && tc_gemm(g, stream);
        prof_end();
        n_kernel_launches += 1;
    };
    auto ln = [&](const float * xin, const float * gw, const float * gb, void * y16, float * y32) {
        prof_begin(PC_LAYERNORM, (double) M * d * 6.0);
        layernorm(dt, xin, d, gw, gb, hp.eps, (int) M, d, y16, d, y32, d, nullptr, stream);
        prof_end();
        n_kernel_launches += 1;
    };

    // conv stem: two GEMMs over im2col'd time-major activations, GELU fused; positional add fused into the second
    prof_begin(PC_IM2COL, (double) M1 * k1 * 2.0);
    im2col1(dt, d_wins, W, n_mel, k1, T, A1, stream);
    prof_end();
    {
        GemmArgs g;
        g.dtype = dt; g.M = (int) M1; g.N = d; g.K = k1; g.a = A1; g.lda = k1; g.w = model.conv1_w; g.ldw = k1;
        g.bias = model.conv1_b; g.gelu = true; g.out16 = act1; g.ldo16 = d;
        gemm(g);
    }
    prof_begin(PC_IM2COL, (double) M * 3 * d * 4.0);
    im2col2(act1, W, d, T, A2, stream);
    prof_end();
    {
        GemmArgs g;
        g.dtype = dt; g.M = (int) M; g.N = d; g.K = 3 * d; g.a = A2; g.lda = 3 * d; g.w = model.conv2_w; g.ldw = 3 * d;
        g.bias = model.conv2_b; g.gelu = true; g.pos = model.e_pe; g.pos_rows = T; g.out32 = x; g.ldo32 = d;
        gemm(g);
    }
    n_kernel_launches += 2;

    // V^T for the tcgen05 attention comes straight out of the QKV GEMM's epilogue (WHISPER_B200_VT_EPILOGUE=0: transpose kernel)
    static const bool vt_epi_env = !(getenv("WHISPER_B200_VT_EPILOGUE") && atoi(getenv("WHISPER_B200_VT_EPILOGUE")) == 0);
    const bool vt_epi = vt_epi_env && vt != nullptr && d == H * 64;
    const int TP = round_up(T, 8);
    if (vt_epi) {
        enc_attention_tc_init_vt(dt, vt, W, T, H, stream);
        n_kernel_launches += 1;
    }
    gemm_cls = PC_GEMM_ENC;
    for (int il = 0; il < hp.n_audio_layer; ++il) {
        const EncLayer & L = model.enc[il];
        ln(x, L.ln1_w, L.ln1_b, h16, nullptr);
        {
            GemmArgs g;
            g.dtype = dt; g.M = (int) M; g.N = 3 * d; g.K = d; g.a = h16; g.lda = d; g.w = L.wqkv; g.ldw = d;
            g.bias = L.bqkv; g.out16 = qkv; g.ldo16 = 3 * d;
            if (vt_epi) {
                g.vt = vt; g.vt_col0 = 2 * d; g.vt_T = T; g.vt_TP = TP; g.vt_H = H;
            }
            gemm(g);
        }
        prof_begin(PC_ENC_ATTN, 4.0 * (double) W * T * (double) T * d);
        if (legacy_attn) {
            enc_attention(dt, qkv, att, W, T, d, H, n_phantom(T), stream);
            n_kernel_launches += 1;
        } else {
            ok = ok && enc_attention_tc(dt, qkv, att, vt, W, T, d, H, n_phantom(T), stream, vt_epi);
            n_kernel_launches += vt_epi ? 1 : 2;
        }
        prof_end();
        {
            GemmArgs g;
            g.dtype = dt; g.M = (int) M; g.N = d; g.K = d; g.a = att; g.lda = d; g.w = L.wo; g.ldw = d;
            g.bias = L.bo; g.resid = x; g.ldr = d; g.out32 = x; g.ldo32 = d;
            gemm(g);
        }
        ln(x, L.ln2_w, L.ln2_b, h16, nullptr);
        {
            GemmArgs g;
            g.dtype = dt; g.M = (int) M; g.N = 4 * d; g.K = d; g.a = h16; g.lda = d; g.w = L.w1; g.ldw = d;
            g.bias = L.b1; g.gelu = true; g.out16 = mlp; g.ldo16 = 4 * d;
            gemm(g);
        }
        {
            GemmArgs g;
            g.dtype = dt; g.M = (int) M; g.N = d; g.K = 4 * d; g.a = mlp; g.lda = 4 * d; g.w = L.w2; g.ldw = 4 * d;
            g.bias = L.b2; g.resid = x; g.ldr = d; g.out32 = x; g.ldo32 = d;
            gemm(g);
        }
    }
    float * e32 = nullptr;
    if (keep_embd32) {
        if (!embd_enc32.reserve(M * d * 4)) return false;
        e32 = (float *) embd_enc32.p;
    }
    ln(x, model.e_ln_w, model.e_ln_b, enc16, e32);
    gemm_cls = PC_GEMM_CROSS;

    // cross K/V for every text layer straight into the pool: K scaled by dh^-0.25 (no bias), V + bias
    if (kv.fp8 != cross_fp8 || (kv.fp8 && !kv16_tmp.reserve(M * 2 * d * 2))) return false;
    const float kscale = powf(64.0f, -0.25f);
    for (int il = 0; il < hp.n_text_layer; ++il) {
        const DecLayer & L = model.dec[il];
        GemmArgs g;
        g.dtype = dt; g.M = (int) M; g.N = 2 * d; g.K = d; g.a = enc16; g.lda = d; g.w = L.wxkv; g.ldw = d;
        g.bias = L.bxkv; g.scale = kscale; g.scale_cols = d;
        char * pool = (char *) kv.data.p + il * kv.layer_stride * 2 + (size_t) win0 * kv.window_bytes;
        g.out16 = kv.fp8 ? kv16_tmp.p : pool;
        g.ldo16 = 2 * d;
        g.head_major_T = T;       // [window][head][K|V][T][64]: the decoder streams one (window, head) block per CTA
        gemm(g);
        if (kv.fp8) cross_fp8_quantize(kv16_tmp.p, pool, W, H, T, stream);
    }
    WB_CUDA(cudaStreamSynchronize(stream));
    if (!ok) wlog(GGML_LOG_LEVEL_ERROR, "%s: GEMM launch rejected its arguments\n", __func__);
    return ok && !cuda_failed();
}

// ---- decoder ----------------------------------------------------------------------------------------------
bool Engine::decode(const std::vector<DecRow> & rows, const std::vector<int> & logit_rows, size_t cross_layer_stride) {
    const int R = (int) rows.size(), RL = (int) logit_rows.size();
    if (R == 0) return true;
    WB_CUDA(cudaSetDevice(device));
    const auto & hp = model.hp;
    const int d = hp.n_text_state, H = hp.n_text_head, n_ctx = hp.n_text_ctx, V = hp.n_vocab;
    const DType dt = model.dtype;

    size_t need = 0;
    auto sz = [&](size_t b) { need += round_up<size_t>(b, 256); };
    sz((size_t) R * d * 4); sz((size_t) R * d * 2); sz((size_t) R * 3 * d * 2); sz((size_t) R * d * 2); sz((size_t) R * d * 2);
    sz((size_t) R * 4 * d * 2); sz((size_t) std::max(1, RL) * d * 2); sz(R * sizeof(DecRow)); sz(std::max(1, RL) * sizeof(int));
    sz(R * sizeof(int2)); sz((size_t) (d / 64 + 1) * R * sizeof(float2));
    if (!ws.begin(need)) return false;
    float * x = (float *) ws.take((size_t) R * d * 4);
    void * h16 = ws.take((size_t) R * d * 2);
    void * qkv = ws.take((size_t) R * 3 * d * 2);
    void * att = ws.take((size_t) R * d * 2);
    void * q16 = ws.take((size_t) R * d * 2);
    void * mlp = ws.take((size_t) R * 4 * d * 2);
    void * hl16 = ws.take((size_t) std::max(1, RL) * d * 2);
    DecRow * d_rows = (DecRow *) ws.take(R * sizeof(DecRow));
    int * d_lrows = (int *) ws.take(std::max(1, RL) * sizeof(int));
    int2 * d_groups = (int2 *) ws.take(R * sizeof(int2));
    float2 * ln_part = (float2 *) ws.take((size_t) (d / 64 + 1) * R * sizeof(float2));
    if (!logits.reserve((size_t) std::max(1, RL) * ld_logits * 4)) return false;

    // runs of consecutive rows that attend to the same window (prompt tokens of a window, beams of a stream): the cross-
    // attention streams a window's K/V once per run instead of once per row
    std::vector<int2> groups;
    for (int i = 0; i < R; ++i) {
        if (!groups.empty() && rows[i].cross_kv == rows[groups.back().x].cross_kv && groups.back().y < DEC_CROSS_GROUP_MAX) groups.back().y++;
        else groups.push_back(make_int2(i, 1));
    }
    const int n_groups = (int) groups.size() < R ? (int) groups.size() : 0;      // all runs of length 1: plain one-CTA-per-row launch

    // stage the row descriptors through pinned memory so the copy is asynchronous
    char * hp_buf = (char *) pinned(0, R * sizeof(DecRow) + RL * sizeof(int) + R * sizeof(int2));
    memcpy(hp_buf, rows.data(), R * sizeof(DecRow));
    if (RL) memcpy(hp_buf + R * sizeof(DecRow), logit_rows.data(), RL * sizeof(int));
    WB_CUDA(cudaMemcpyAsync(d_rows, hp_buf, R * sizeof(DecRow), cudaMemcpyHostToDevice, stream));
    if (RL) WB_CUDA(cudaMemcpyAsync(d_lrows, hp_buf + R * sizeof(DecRow), RL * sizeof(int), cudaMemcpyHostToDevice, stream));
    if (n_groups) {
        char * hg = hp_buf + R * sizeof(DecRow) + RL * sizeof(int);
        memcpy(hg, groups.data(), n_groups * sizeof(int2));
        WB_CUDA(cudaMemcpyAsync(d_groups, hg, n_groups * sizeof(int2), cudaMemcpyHostToDevice, stream));
    }

    bool ok = true;
    int gemm_cls = PC_GEMM_DEC;
    auto gemm = [&](const GemmArgs & g) {
        prof_begin(gemm_cls, ((double) g.N * g.K + (double) g.M * (g.N + g.K)) * 2.0);   // bytes: weights + activations
        // few rows: stream the weights with every SM -- tcgen05 version (tc_skinny.cu) unless WHISPER_B200_TC_SKINNY=0, else
        // the mma.sync one (skinny_gemm.cu); many rows (long prompts): tensor-core tiles (tc_gemm.cu)
        static const bool tcs = !(getenv("WHISPER_B200_TC_SKINNY") && atoi(getenv("WHISPER_B200_TC_SKINNY")) == 0);
        if (g.ln_part_in || g.ln_part_out) {
            ok = ok && tc_skinny_usable(g) && tc_skinny_gemm(g, stream);       // LayerNorm-folded forms exist on this kernel only
        } else if (g.M <= 128) {
            ok = ok && (tcs && tc_skinny_usable(g) ? tc_skinny_gemm(g, stream) : skinny_gemm(g, skinny_ws, stream));
        } else if (g.M <= 512 && tcs && !g.pos) {
            // short prompts (a handful of tokens per window): 128 x 256 tensor-core tiles would leave most SMs without a tile
            // (N = 1280 -> 10-20 tiles), so stream the weights once per block of 128 rows instead
            for (int r0 = 0; r0 < g.M && ok; r0 += 128) {
                GemmArgs c = g;
                c.M = std::min(128, g.M - r0);
                c.a = (const char *) g.a + (size_t) r0 * g.lda * 2;
                if (g.resid) c.resid = g.resid + (size_t) r0 * g.ldr;
                if (g.out16) c.out16 = (char *) g.out16 + (size_t) r0 * g.ldo16 * 2;
                if (g.out32) c.out32 = g.out32 + (size_t) r0 * g.ldo32;
                ok = ok && (tc_skinny_usable(c) ? tc_skinny_gemm(c, stream) : tc_gemm(c, stream));
                if (r0) n_kernel_launches += 1;
            }
        } else {
            ok = ok && tc_gemm(g, stream);
        }
        prof_end();
        n_kernel_launches += 1;
    };
    auto ln = [&](const float * gw, const float * gb) {
        prof_begin(PC_LAYERNORM_DEC, (double) R * d * 6.0);
        layernorm(dt, x, d, gw, gb, hp.eps, R, d, h16, d, nullptr, 0, nullptr, stream);
        prof_end();
        n_kernel_launches += 1;
    };
    const float qk_scale = powf(64.0f, -0.25f);
    const size_t self_layer = (size_t) n_ctx * 2 * d;
    // single-token step (every sequence contributes exactly one row): the self-attention kernel appends K/V itself
    // (rows of one sequence are always contiguous in a batch, so comparing neighbours is enough)
    bool fuse_append = true;
    for (int i = 1; i < R && fuse_append; ++i)
        if (rows[i].self_kv == rows[i - 1].self_kv) fuse_append = false;
    if (fuse_append && !align.on && !cross_fp8 && chain_usable(R)) return decode_chain(rows, logit_rows, cross_layer_stride);
    if (align.on) {
        if (align.n_heads_total <= 0 || !align.probs.reserve((size_t) align.n_heads_total * R * cross_T * sizeof(float))) return false;
    }
    // At most 128 rows (every GEMM of the step runs on the weight-streaming kernel): the three LayerNorms of a layer are folded
    // algebraically into their neighbours (tc_skinny.cu) -- the GEMM that produces the residual stream (O, cross-O, MLP-down) also
    // writes the 16-bit rows x * gamma of the NEXT LayerNorm and per-tile row statistics; the GEMM that consumes the normalised
    // rows (QKV, cross-Q, MLP-up) streams those rows as an ordinary A operand and applies mean / rstd in its epilogue with the
    // per-column sums prepared at model load (model.cu: ln_fold).  Removes three dependent launches per layer with nothing added
    // to a critical path.  WHISPER_B200_LN_FOLD=0 keeps the separate LayerNorm kernels (the rounding points of the reference).
    static const bool ln_fold_env = !(getenv("WHISPER_B200_LN_FOLD") && atoi(getenv("WHISPER_B200_LN_FOLD")) == 0);
    static const bool tcs_env = !(getenv("WHISPER_B200_TC_SKINNY") && atoi(getenv("WHISPER_B200_TC_SKINNY")) == 0);
    const bool fold = ln_fold_env && tcs_env && !exact_ln && R <= 128 && d % 64 == 0;
    auto ln_consumer = [&](GemmArgs & g, const float * colsum, const float * bias_folded) {      // A operand = x * gamma rows in h16
        g.ln_part_in = ln_part; g.ln_parts = d / 64; g.ln_colsum = colsum; g.bias = bias_folded; g.ln_eps = hp.eps;
    };
    auto ln_producer = [&](GemmArgs & g, const float * gamma_next) {
        g.ln_part_out = ln_part; g.out16 = h16; g.ldo16 = d; g.out16_gamma = gamma_next;
    };
    // L2 prefetch of the coming cross-attention's K prefix by the six GEMMs that run between two cross-attention launches
    // (tc_skinny.cu): slots 0-2 = cross-O, MLP up, MLP down of the previous layer, 3-5 = QKV, O, cross-Q of the layer itself
    // How much: the first four 16 KB chunks of every (row, head) K block, ~84 MB per launch at 64 rows (measured at 64 windows:
    // 2 / 4 / 6 / 8 chunks give 242 / 243 / 241 / 242 ms per 60 steps against 248 without -- L2 keeps ~35-40 MB of it next to the
    // GEMMs' own traffic; an L2 persisting set-aside (cudaLimitPersistingL2CacheSize) does not change that and costs the encoder
    // 18 % of its speed, so none is configured).  Only for large batches: the requests compete with the latency-bound GEMM chain that carries them, and
    // with few rows the cross-attention is too short to pay that back (8 windows: 516 ms per 220-token step with, 500 without).
    static const int pf_chunks_env = getenv("WHISPER_B200_CROSS_PF_CHUNKS") ? atoi(getenv("WHISPER_B200_CROSS_PF_CHUNKS")) : -1;
    const int pf_block_chunks = (2 * cross_T * 128 + 16383) / 16384;         // 16 KB requests that cover one K | V block
    const int pf_auto = R >= 32 ? 4 : 0;
    const int pf_chunks = (tcs_env && !cross_fp8 && R <= 64 && !align.on) ? std::min(pf_chunks_env >= 0 ? pf_chunks_env : pf_auto, pf_block_chunks) : 0;
    auto prefetch = [&](GemmArgs & g, int layer, int slot) {
        if (pf_chunks <= 0 || layer >= hp.n_text_layer) return;
        g.pf_rows = d_rows; g.pf_R = R; g.pf_H = H; g.pf_chunks = pf_chunks; g.pf_slot = slot; g.pf_slots = 6;
        g.pf_layer_off_bytes = (size_t) layer * cross_layer_stride * 2; g.pf_head_bytes = 2 * cross_T * 64 * 2;
    };

    prof_begin(PC_DEC_MISC, (double) R * d * 10.0);
    dec_embed(dt, model.d_te, model.d_pe, d_rows, R, d, x, stream);
    prof_end();
    n_kernel_launches += 1;
    for (int il = 0; il < hp.n_text_layer; ++il) {
        const DecLayer & L = model.dec[il];
        if (!fold || il == 0) ln(L.ln1_w, L.ln1_b);
        {
            GemmArgs g;   // Q and K carry dh^-0.25 each (src/whisper.cpp:2506, 2550, 2557); V is biased only
            g.dtype = dt; g.M = R; g.N = 3 * d; g.K = d; g.a = h16; g.lda = d; g.w = L.wqkv; g.ldw = d;
            g.bias = L.bqkv; g.scale = qk_scale; g.scale_cols = 2 * d; g.out16 = qkv; g.ldo16 = 3 * d;
            if (fold && il > 0) ln_consumer(g, L.qkv_c, L.qkv_b);
            prefetch(g, il, 3);
            gemm(g);
        }
        if (!fuse_append) {
            prof_begin(PC_DEC_MISC, (double) R * 2 * d * 4.0);
            dec_kv_append(qkv, d_rows, R, d, il * self_layer, stream);
            prof_end();
            n_kernel_launches += 1;
        }
        double self_bytes = 0.0;
        for (const auto & rw : rows) self_bytes += (double) (rw.pos + 1) * 2 * d * 2.0;
        prof_begin(PC_SELF_ATTN, self_bytes);
        dec_self_attn(dt, qkv, d_rows, R, d, H, il * self_layer, n_ctx, fuse_append, att, stream);
        prof_end();
        n_kernel_launches += 1;
        {
            GemmArgs g;
            g.dtype = dt; g.M = R; g.N = d; g.K = d; g.a = att; g.lda = d; g.w = L.wo; g.ldw = d;
            g.bias = L.bo; g.resid = x; g.ldr = d; g.out32 = x; g.ldo32 = d;
            if (fold) ln_producer(g, L.lnx_w);
            prefetch(g, il, 4);
            gemm(g);
        }
        if (!fold) ln(L.lnx_w, L.lnx_b);
        {
            GemmArgs g;
            g.dtype = dt; g.M = R; g.N = d; g.K = d; g.a = h16; g.lda = d; g.w = L.wxq; g.ldw = d;
            g.bias = L.bxq; g.out16 = q16; g.ldo16 = d;
            if (fold) ln_consumer(g, L.xq_c, L.xq_b);
            prefetch(g, il, 5);
            gemm(g);
        }
        if (align.on && !align.heads_by_layer[il].empty()) {
            int a0 = 0;
            for (int k = 0; k < il; ++k) a0 += (int) align.heads_by_layer[k].size();
            dtw_capture_layer(dt, q16, d_rows, R, d, (const int *) align.d_heads.p + a0, (int) align.heads_by_layer[il].size(),
                              il * cross_layer_stride, cross_T, a0, (float *) align.probs.p, stream);
            n_kernel_launches += 1;
        }
        prof_begin(PC_CROSS_ATTN, cross_fp8 ? (double) R * cross_fp8_window_bytes(H, cross_T) : (double) R * cross_T * 2.0 * d * 2.0);
        if (cross_fp8)
            dec_cross_attn_fp8(q16, d_rows, R, d, H, il * cross_layer_stride, cross_T, n_phantom(cross_T), att, stream, d_groups, n_groups);
        else
            dec_cross_attn(dt, q16, d_rows, R, d, H, il * cross_layer_stride, cross_T, n_phantom(cross_T), att, stream, nullptr, d_groups, n_groups);
        prof_end();
        n_kernel_launches += 1;
        {
            GemmArgs g;
            g.dtype = dt; g.M = R; g.N = d; g.K = d; g.a = att; g.lda = d; g.w = L.wxo; g.ldw = d;
            g.bias = L.bxo; g.resid = x; g.ldr = d; g.out32 = x; g.ldo32 = d;
            if (fold) ln_producer(g, L.ln2_w);
            prefetch(g, il + 1, 0);
            gemm(g);
        }
        if (!fold) ln(L.ln2_w, L.ln2_b);
        {
            GemmArgs g;
            g.dtype = dt; g.M = R; g.N = 4 * d; g.K = d; g.a = h16; g.lda = d; g.w = L.w1; g.ldw = d;
            g.bias = L.b1; g.gelu = true; g.out16 = mlp; g.ldo16 = 4 * d;
            if (fold) ln_consumer(g, L.m1_c, L.m1_b);
            prefetch(g, il + 1, 1);
            gemm(g);
        }
        {
            GemmArgs g;
            g.dtype = dt; g.M = R; g.N = d; g.K = 4 * d; g.a = mlp; g.lda = 4 * d; g.w = L.w2; g.ldw = 4 * d;
            g.bias = L.b2; g.resid = x; g.ldr = d; g.out32 = x; g.ldo32 = d;
            if (fold && il + 1 < hp.n_text_layer) ln_producer(g, model.dec[il + 1].ln1_w);
            prefetch(g, il + 1, 2);
            gemm(g);
        }
    }
    if (RL > 0) {
        // final LayerNorm only on the rows whose logits are wanted, then the tied-embedding logits GEMM
        prof_begin(PC_LAYERNORM_DEC, (double) RL * d * 6.0);
        layernorm(dt, x, d, model.d_ln_w, model.d_ln_b, hp.eps, RL, d, hl16, d, nullptr, 0, d_lrows, stream);
        prof_end();
        gemm_cls = PC_GEMM_LOGITS;
        GemmArgs g;
        g.dtype = dt; g.M = RL; g.N = V; g.K = d; g.a = hl16; g.lda = d; g.w = model.d_te; g.ldw = d;
        g.out32 = (float *) logits.p; g.ldo32 = ld_logits;
        gemm(g);
        n_kernel_launches += 1;
    }
    if (!ok) wlog(GGML_LOG_LEVEL_ERROR, "%s: GEMM launch rejected its arguments\n", __func__);
    return ok && !cuda_failed();
}

// ---- single-token decoder step through the persistent chain kernel ---------------------------------------------------
bool Engine::chain_usable(int R) {
    if (chain_mode < 0) {
        const char * e = getenv("WHISPER_B200_CHAIN");
        const auto & hp = model.hp;
        const bool geom_ok = hp.n_text_state % 128 == 0 && hp.n_text_state <= 1536 && hp.n_text_ctx <= 2048 &&
                             hp.n_text_state == hp.n_text_head * 64;
        // 1: the whole layer between two cross-attention launches is one chain launch; 2: hybrid -- only the stream-K GEMM +
        // residual/LayerNorm groups are chain launches, the wide GEMMs and self-attention stay separate kernels
        chain_mode = e && geom_ok ? atoi(e) : 0;                  // opt-in until it beats the unfused sequence
        if (chain_mode < 0 || chain_mode > 2) chain_mode = 0;
        if (const char * u = getenv("WHISPER_B200_CHAIN_UNITS")) chain_min_units = std::max(1, atoi(u));
        if (chain_mode && chain_init(chain, model.dtype) <= 0) chain_mode = 0;
    }
    return chain_mode != 0 && R >= 1 && R <= 128;
}

bool Engine::decode_chain(const std::vector<DecRow> & rows, const std::vector<int> & logit_rows, size_t cross_layer_stride) {
    const int R = (int) rows.size(), RL = (int) logit_rows.size();
    const auto & hp = model.hp;
    const int d = hp.n_text_state, H = hp.n_text_head, n_ctx = hp.n_text_ctx, V = hp.n_vocab, L = hp.n_text_layer;
    const DType dt = model.dtype;
    const int G = chain.grid;

    // stream-K geometries (one per GEMM shape) and the partial-tile scratch they share
    // the wide GEMMs (QKV, MLP up) own whole 64x32 tiles and finish in place; the d-wide ones are stream-K
    const SplitGeom g_qkv = chain_geom_direct(R, 3 * d, d), g_m1 = chain_geom_direct(R, 4 * d, d),
                    g_dd = chain_geom(G, R, d, d, chain_min_units), g_m2 = chain_geom(G, R, d, 4 * d, chain_min_units);
    size_t part_floats = 0;
    for (const SplitGeom * g : {&g_dd, &g_m2}) part_floats = std::max(part_floats, chain_part_floats(*g, R));
    if (!chain_part.reserve(part_floats * 4)) return false;
    float * part = (float *) chain_part.p;

    bool identity_logits = RL == R;
    for (int i = 0; i < RL && identity_logits; ++i) identity_logits = logit_rows[i] == i;

    size_t need = 0;
    auto sz = [&](size_t b) { need += round_up<size_t>(b, 256); };
    sz((size_t) R * d * 4); sz((size_t) R * d * 2); sz((size_t) R * 3 * d * 2); sz((size_t) R * d * 2); sz((size_t) R * 4 * d * 2);
    sz((size_t) std::max(1, RL) * d * 2); sz(R * sizeof(DecRow)); sz(std::max(1, RL) * sizeof(int));
    if (!ws.begin(need)) return false;
    float * x = (float *) ws.take((size_t) R * d * 4);
    void * h16 = ws.take((size_t) R * d * 2);
    void * qkv = ws.take((size_t) R * 3 * d * 2);
    void * att = ws.take((size_t) R * d * 2);
    void * mlp = ws.take((size_t) R * 4 * d * 2);
    void * hl16 = ws.take((size_t) std::max(1, RL) * d * 2);
    DecRow * d_rows = (DecRow *) ws.take(R * sizeof(DecRow));
    int * d_lrows = (int *) ws.take(std::max(1, RL) * sizeof(int));
    if (!logits.reserve((size_t) std::max(1, RL) * ld_logits * 4)) return false;

    char * hp_buf = (char *) pinned(0, R * sizeof(DecRow) + RL * sizeof(int));
    memcpy(hp_buf, rows.data(), R * sizeof(DecRow));
    if (RL) memcpy(hp_buf + R * sizeof(DecRow), logit_rows.data(), RL * sizeof(int));
    WB_CUDA(cudaMemcpyAsync(d_rows, hp_buf, R * sizeof(DecRow), cudaMemcpyHostToDevice, stream));
    if (RL) WB_CUDA(cudaMemcpyAsync(d_lrows, hp_buf + R * sizeof(DecRow), RL * sizeof(int), cudaMemcpyHostToDevice, stream));

    const float qk_scale = powf(64.0f, -0.25f);
    const size_t self_layer = (size_t) n_ctx * 2 * d;
    double self_bytes = 0.0;
    for (const auto & rw : rows) self_bytes += (double) (rw.pos + 1) * 2 * d * 2.0;

    // TMA descriptors: the three GEMM inputs of this call, and (cached) one per weight matrix
    const int box_x = R <= 64 ? 64 : 128;
    TMap tm_h16, tm_att, tm_mlp;
    if (!tc_make_tmap(&tm_h16, h16, R, d, d, box_x, dt) || !tc_make_tmap(&tm_att, att, R, d, d, box_x, dt) ||
        !tc_make_tmap(&tm_mlp, mlp, R, 4 * d, 4 * d, box_x, dt)) return false;
    if (chain_wmaps.empty()) {
        chain_wmaps.resize((size_t) L * 6);
        for (int il = 0; il < L; ++il) {
            const DecLayer & Lr = model.dec[il];
            TMap * m = &chain_wmaps[(size_t) il * 6];
            if (!tc_make_tmap(&m[0], Lr.wqkv, 3 * d, d, d, 128, dt) || !tc_make_tmap(&m[1], Lr.wo, d, d, d, 128, dt) ||
                !tc_make_tmap(&m[2], Lr.wxq, d, d, d, 128, dt) || !tc_make_tmap(&m[3], Lr.wxo, d, d, d, 128, dt) ||
                !tc_make_tmap(&m[4], Lr.w1, 4 * d, d, d, 128, dt) || !tc_make_tmap(&m[5], Lr.w2, d, 4 * d, 4 * d, 128, dt)) {
                chain_wmaps.clear();
                return false;
            }
        }
    }
    enum { W_QKV = 0, W_O, W_XQ, W_XO, W_1, W_2 };

    ChainParams cp;
    int n_gemm = 0;
    auto reset = [&]() {
        n_gemm = 0;
        cp = ChainParams();
        cp.c.R = R; cp.c.d = d; cp.c.H = H; cp.c.n_ctx = n_ctx; cp.c.eps = hp.eps; cp.c.ref_f16_gelu = dt == DType::F16 ? 1 : 0;
        cp.c.x = x; cp.c.rows = d_rows; cp.c.te = model.d_te; cp.c.pe = model.d_pe;
    };
    double chain_bytes = 0.0;
    auto add_row = [&](const SplitGeom * g, const float * bias, const float * lw, const float * lb, void * out16, bool embed) {
        ChainPhase & ph = cp.ph[cp.n_phase++];
        ph.type = CP_ROW; ph.embed = embed ? 1 : 0;
        if (g) { ph.g = *g; ph.part = part; }
        ph.bias = bias; ph.ln_w = lw; ph.ln_b = lb; ph.out16 = out16; ph.ldo16 = d;
        chain_bytes += (double) R * d * 10.0;
    };
    auto add_gemm = [&](const SplitGeom & g, const TMap & ta, int K, const TMap & tw, int N) {      // stream-K -> partial tiles
        ChainPhase & ph = cp.ph[cp.n_phase++];
        ph.type = CP_GEMM; ph.direct = 0; ph.N = N; ph.K = K; ph.g = g; ph.part = part;
        ph.tm = n_gemm; cp.tm[2 * n_gemm] = ta; cp.tm[2 * n_gemm + 1] = tw; ++n_gemm;
        chain_bytes += ((double) N * K + (double) R * K) * 2.0;
    };
    auto add_gemm_direct = [&](const SplitGeom & g, const TMap & ta, int K, const TMap & tw, int N, const float * bias, float scale,
                               int scale_cols, bool gelu, void * out16) {
        ChainPhase & ph = cp.ph[cp.n_phase++];
        ph.type = CP_GEMM; ph.direct = 1; ph.N = N; ph.K = K; ph.g = g;
        ph.tm = n_gemm; cp.tm[2 * n_gemm] = ta; cp.tm[2 * n_gemm + 1] = tw; ++n_gemm;
        ph.bias = bias; ph.scale = scale; ph.scale_cols = scale_cols; ph.gelu = gelu ? 1 : 0; ph.out16 = out16; ph.ldo16 = N;
        chain_bytes += ((double) N * K + (double) R * (K + N)) * 2.0;
    };
    auto add_self = [&](int il) {
        ChainPhase & ph = cp.ph[cp.n_phase++];
        ph.type = CP_SELF; ph.a = qkv; ph.lda = 3 * d; ph.out16 = att; ph.ldo16 = d; ph.layer_off = il * self_layer;
        chain_bytes += self_bytes;
    };
    // QKV -> self-attention -> out-projection -> residual + LayerNorm -> cross query of layer il (h16 = LayerNorm(x) on entry)
    auto add_attn_half = [&](int il) {
        const DecLayer & Lr = model.dec[il];
        const TMap * wm = &chain_wmaps[(size_t) il * 6];
        add_gemm_direct(g_qkv, tm_h16, d, wm[W_QKV], 3 * d, Lr.bqkv, qk_scale, 2 * d, false, qkv);
        add_self(il);
        add_gemm(g_dd, tm_att, d, wm[W_O], d);
        add_row(&g_dd, Lr.bo, Lr.lnx_w, Lr.lnx_b, h16, false);
        add_gemm(g_dd, tm_h16, d, wm[W_XQ], d);
    };
    // cross out-projection -> residual + LayerNorm -> MLP of layer il; the closing residual/LayerNorm row phase is added by the caller
    auto add_mlp_half = [&](int il) {
        const DecLayer & Lr = model.dec[il];
        const TMap * wm = &chain_wmaps[(size_t) il * 6];
        add_gemm(g_dd, tm_att, d, wm[W_XO], d);
        add_row(&g_dd, Lr.bxo, Lr.ln2_w, Lr.ln2_b, h16, false);
        add_gemm_direct(g_m1, tm_h16, d, wm[W_1], 4 * d, Lr.b1, 1.0f, 0, true, mlp);
        add_gemm(g_m2, tm_mlp, 4 * d, wm[W_2], d);
    };
    static const bool trace_on = getenv("WHISPER_B200_CHAIN_TRACE") != nullptr;
    if (trace_on && !chain_trace.reserve((size_t) (L + 1) * 32 * 8)) return false;
    if (trace_on) WB_CUDA(cudaMemsetAsync(chain_trace.p, 0, (size_t) (L + 1) * 32 * 8, stream));
    int launch_idx = 0;
    bool ok = true;
    auto launch = [&]() {
        if (trace_on) cp.trace = (unsigned long long *) chain_trace.p + 32 * launch_idx;
        ++launch_idx;
        prof_begin(PC_DEC_CHAIN, chain_bytes);
        ok = ok && chain_launch(chain, dt, cp, stream);
        prof_end();
        n_kernel_launches += 1;
        chain_bytes = 0.0;
    };
    auto cross = [&](int il) {
        SplitIn qs;
        qs.part = part; qs.bias = model.dec[il].bxq; qs.g = g_dd;
        if (trace_on) qs.trace = (unsigned long long *) chain_trace.p + 32 * il + 27;
        prof_begin(PC_CROSS_ATTN, (double) R * cross_T * 2.0 * d * 2.0);
        dec_cross_attn(dt, nullptr, d_rows, R, d, H, il * cross_layer_stride, cross_T, n_phantom(cross_T), att, stream, &qs);
        prof_end();
        n_kernel_launches += 1;
    };
    if (chain_mode == 1) {
        for (int il = 0; il <= L; ++il) {
            reset();
            if (il == 0) {
                add_row(nullptr, nullptr, model.dec[0].ln1_w, model.dec[0].ln1_b, h16, true);
            } else {
                add_mlp_half(il - 1);
                if (il < L) add_row(&g_m2, model.dec[il - 1].b2, model.dec[il].ln1_w, model.dec[il].ln1_b, h16, false);
                else        add_row(&g_m2, model.dec[il - 1].b2, identity_logits ? model.d_ln_w : nullptr,
                                    identity_logits ? model.d_ln_b : nullptr, hl16, false);
            }
            if (il < L) add_attn_half(il);
            launch();
            if (il < L) cross(il);
        }
    } else {
        // hybrid: chain launches only where a stream-K GEMM feeds a residual + LayerNorm (and the cross query GEMM)
        auto skinny = [&](const void * a, int K, const void * w, int N, const float * bias, float scale, int scale_cols, bool gelu,
                          void * out16) {
            GemmArgs g;
            g.dtype = dt; g.M = R; g.N = N; g.K = K; g.a = a; g.lda = K; g.w = w; g.ldw = K; g.bias = bias;
            g.scale = scale; g.scale_cols = scale_cols; g.gelu = gelu; g.out16 = out16; g.ldo16 = N;
            prof_begin(PC_GEMM_DEC, ((double) N * K + (double) R * (N + K)) * 2.0);
            ok = ok && skinny_gemm(g, skinny_ws, stream);
            prof_end();
            n_kernel_launches += 1;
        };
        reset();
        add_row(nullptr, nullptr, model.dec[0].ln1_w, model.dec[0].ln1_b, h16, true);
        launch();
        for (int il = 0; il < L; ++il) {
            const DecLayer & Lr = model.dec[il];
            const TMap * wm = &chain_wmaps[(size_t) il * 6];
            skinny(h16, d, Lr.wqkv, 3 * d, Lr.bqkv, qk_scale, 2 * d, false, qkv);
            prof_begin(PC_SELF_ATTN, self_bytes);
            dec_self_attn(dt, qkv, d_rows, R, d, H, il * self_layer, n_ctx, true, att, stream);
            prof_end();
            n_kernel_launches += 1;
            reset();
            add_gemm(g_dd, tm_att, d, wm[W_O], d);
            add_row(&g_dd, Lr.bo, Lr.lnx_w, Lr.lnx_b, h16, false);
            add_gemm(g_dd, tm_h16, d, wm[W_XQ], d);
            launch();
            cross(il);
            reset();
            add_gemm(g_dd, tm_att, d, wm[W_XO], d);
            add_row(&g_dd, Lr.bxo, Lr.ln2_w, Lr.ln2_b, h16, false);
            launch();
            skinny(h16, d, Lr.w1, 4 * d, Lr.b1, 1.0f, 0, true, mlp);
            reset();
            add_gemm(g_m2, tm_mlp, 4 * d, wm[W_2], d);
            if (il + 1 < L) add_row(&g_m2, Lr.b2, model.dec[il + 1].ln1_w, model.dec[il + 1].ln1_b, h16, false);
            else            add_row(&g_m2, Lr.b2, identity_logits ? model.d_ln_w : nullptr, identity_logits ? model.d_ln_b : nullptr, hl16, false);
            launch();
        }
    }
    if (RL > 0) {
        if (!identity_logits) {
            prof_begin(PC_LAYERNORM_DEC, (double) RL * d * 6.0);
            layernorm(dt, x, d, model.d_ln_w, model.d_ln_b, hp.eps, RL, d, hl16, d, nullptr, 0, d_lrows, stream);
            prof_end();
            n_kernel_launches += 1;
        }
        prof_begin(PC_GEMM_LOGITS, ((double) V * d + (double) RL * (V + d)) * 2.0);
        GemmArgs g;
        g.dtype = dt; g.M = RL; g.N = V; g.K = d; g.a = hl16; g.lda = d; g.w = model.d_te; g.ldw = d;
        g.out32 = (float *) logits.p; g.ldo32 = ld_logits;
        ok = ok && skinny_gemm(g, skinny_ws, stream);
        prof_end();
        n_kernel_launches += 1;
    }
    if (trace_on && L >= 3 && chain_mode == 1) {
        // average over the middle launches (all have the same 11 phases): phase durations, launch-to-launch gap
        std::vector<unsigned long long> h((size_t) (L + 1) * 32);
        WB_CUDA(cudaStreamSynchronize(stream));
        WB_CUDA(cudaMemcpy(h.data(), chain_trace.p, h.size() * 8, cudaMemcpyDeviceToHost));
        if (chain_trace_acc.empty()) chain_trace_acc.assign(32, 0.0);
        for (int l = 1; l < L; ++l) {
            const unsigned long long * t = h.data() + 32 * l;
            for (int i = 0; i < 10; ++i) chain_trace_acc[i] += (double) (t[i + 1] - t[i]) * 1e-3;
            chain_trace_acc[10] += (double) (t[0] - t[15]) * 1e-3;                   // entry -> first phase (PDL wait)
            chain_trace_acc[11] += (double) (t[15] - (t - 32)[l == 1 ? 6 : 10]) * 1e-3;   // previous chain end -> entry
            chain_trace_acc[12] += (double) ((t - 32)[27] - (t - 32)[l == 1 ? 6 : 10]) * 1e-3;   // previous chain end -> cross start
            chain_trace_acc[13] += (double) ((t - 32)[28] - (t - 32)[27]) * 1e-3;                 // cross kernel, first start -> last end
            chain_trace_acc[14] += (double) ((long long) t[15] - (long long) (t - 32)[28]) * 1e-3;   // cross end -> chain entry (negative: overlapped)
        }
        chain_trace_steps += L - 1;
        if (chain_trace_steps % ((L - 1) * 20) == 0) {
            static const char * names[15] = {"xO", "row", "mlp1", "mlp2", "row", "qkv", "self", "O", "row", "xQ", "wait", "gap(cross)",
                                             "end->cross", "cross", "cross->entry"};
            fprintf(stderr, "chain trace (us, R=%d):", R);
            for (int i = 0; i < 15; ++i) fprintf(stderr, " %s %.2f", names[i], chain_trace_acc[i] / chain_trace_steps);
            fprintf(stderr, "  pdl=%d\n", chain.pdl_ok ? 1 : 0);
            {
                const unsigned long long * t = h.data() + 32 * (L / 2);
                fprintf(stderr, "   qkv phase, CTA 0 (cycles): producer waited for free stages %llu of %llu; MMA issuer waited for data %llu of %llu\n",
                        t[20], t[21], t[22], t[23]);
            }
        }
    }
    if (!ok) wlog(GGML_LOG_LEVEL_ERROR, "%s: chain launch failed\n", __func__);
    return ok && !cuda_failed();
}

bool Engine::fetch_logits(int row, float * out) {
    WB_CUDA(cudaMemcpyAsync(out, (const float *) logits.p + (size_t) row * ld_logits, (size_t) model.hp.n_vocab * 4,
                            cudaMemcpyDeviceToHost, stream));
    WB_CUDA(cudaStreamSynchronize(stream));
    return !cuda_failed();
}

bool Engine::fetch_logits_rows(int row0, int n_rows, float * out) {
    if (n_rows <= 0) return true;
    const size_t w = (size_t) model.hp.n_vocab * 4;
    WB_CUDA(cudaMemcpy2DAsync(out, w, (const float *) logits.p + (size_t) row0 * ld_logits, (size_t) ld_logits * 4, w, n_rows,
                              cudaMemcpyDeviceToHost, stream));
    WB_CUDA(cudaStreamSynchronize(stream));
    return !cuda_failed();
}

bool Engine::sample(const std::vector<SampleRow> & srows, const std::vector<double> & uniforms, const uint32_t * d_mask,
                    const SampleParams & prm, std::vector<SampleOut> & out, std::vector<DrawOut> & draws) {
    const int R = (int) srows.size();
    const size_t n_u = uniforms.size();
    out.resize(R);
    draws.resize(n_u);
    if (R == 0) return true;
    // device / pinned layout: [rows | uniforms | outs | draws], each part 256-byte aligned
    const size_t in_b = round_up<size_t>(R * sizeof(SampleRow), 256), u_b = round_up<size_t>(n_u * sizeof(double), 256);
    const size_t out_b = round_up<size_t>(R * sizeof(SampleOut), 256), dr_b = round_up<size_t>(n_u * sizeof(DrawOut), 256);
    if (!meta.reserve(in_b + u_b + out_b + dr_b)) return false;
    char * dev = (char *) meta.p;
    char * hb = (char *) pinned(1, in_b + u_b + out_b + dr_b);
    memcpy(hb, srows.data(), R * sizeof(SampleRow));
    if (n_u) memcpy(hb + in_b, uniforms.data(), n_u * sizeof(double));
    WB_CUDA(cudaMemcpyAsync(dev, hb, in_b + (n_u ? u_b : 0), cudaMemcpyHostToDevice, stream));
    prof_begin(PC_SAMPLE, (double) R * model.hp.n_vocab * 4.0 * 5.0);
    dec_sample((const float *) logits.p, ld_logits, (const SampleRow *) dev, R, d_mask, prm, (SampleOut *) (dev + in_b + u_b),
               (const double *) (dev + in_b), (DrawOut *) (dev + in_b + u_b + out_b), stream);
    prof_end();
    n_kernel_launches += 1;
    WB_CUDA(cudaMemcpyAsync(hb + in_b + u_b, dev + in_b + u_b, out_b + (n_u ? dr_b : 0), cudaMemcpyDeviceToHost, stream));
    WB_CUDA(cudaStreamSynchronize(stream));
    memcpy(out.data(), hb + in_b + u_b, R * sizeof(SampleOut));
    if (n_u) memcpy(draws.data(), hb + in_b + u_b + out_b, n_u * sizeof(DrawOut));
    return !cuda_failed();
}

bool Engine::token_prob(const std::vector<SampleRow> & srows, int token, std::vector<float> & out) {
    const int R = (int) srows.size();
    out.resize(R);
    if (R == 0) return true;
    const size_t in_b = R * sizeof(SampleRow), out_b = R * sizeof(float);
    if (!meta.reserve(round_up<size_t>(in_b, 256) + out_b)) return false;
    SampleRow * d_in = (SampleRow *) meta.p;
    float * d_out = (float *) ((char *) meta.p + round_up<size_t>(in_b, 256));
    char * hb = (char *) pinned(1, in_b + out_b);
    memcpy(hb, srows.data(), in_b);
    WB_CUDA(cudaMemcpyAsync(d_in, hb, in_b, cudaMemcpyHostToDevice, stream));
    dec_token_prob((const float *) logits.p, ld_logits, d_in, R, model.hp.n_vocab, token, d_out, stream);
    n_kernel_launches += 1;
    WB_CUDA(cudaMemcpyAsync(hb + in_b, d_out, out_b, cudaMemcpyDeviceToHost, stream));
    WB_CUDA(cudaStreamSynchronize(stream));
    memcpy(out.data(), hb + in_b, out_b);
    return !cuda_failed();
}

bool Engine::kv_copy_prefix(const void * src, void * dst, int n_pos) {
    if (n_pos <= 0 || src == dst) return true;
    const int d = model.hp.n_text_state;
    const size_t layer8 = (size_t) model.hp.n_text_ctx * 2 * d / 8;
    const size_t n8 = (size_t) n_pos * 2 * d / 8;
    dim3 grid((unsigned) ceil_div<size_t>(n8, 256), model.hp.n_text_layer);
    kv_copy_prefix_kernel<<<grid, 256, 0, stream>>>((const uint4 *) src, (uint4 *) dst, layer8, n8);
    n_kernel_launches += 1;
    WB_CUDA(cudaGetLastError());
    return !cuda_failed();
}

bool Engine::kv_copy_prefix_batch(const std::vector<KvCopy> & copies) {
    const int n = (int) copies.size();
    if (n == 0) return true;
    const int d = model.hp.n_text_state;
    const size_t bytes = n * sizeof(KvCopy);
    if (!kv_copy_list.reserve(bytes)) return false;
    // (the previous list's H2D copy has completed: every decode step ends with a synchronising read-back of the selection)
    void * hb = pinned(2, bytes);
    memcpy(hb, copies.data(), bytes);
    WB_CUDA(cudaMemcpyAsync(kv_copy_list.p, hb, bytes, cudaMemcpyHostToDevice, stream));
    int max_pos = 1;
    for (const auto & c : copies) max_pos = std::max(max_pos, c.n_pos);
    const int row8 = 2 * d / 8;
    dim3 grid((unsigned) std::min<size_t>(ceil_div<size_t>((size_t) max_pos * row8, 256), 64), model.hp.n_text_layer, n);
    kv_copy_batch_kernel<<<grid, 256, 0, stream>>>((const KvCopy *) kv_copy_list.p, (size_t) model.hp.n_text_ctx * row8, row8);
    n_kernel_launches += 1;
    WB_CUDA(cudaGetLastError());
    return !cuda_failed();
}

}  // namespace wb
