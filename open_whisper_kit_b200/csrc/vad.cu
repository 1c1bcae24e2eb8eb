// Voice activity detection (Silero VAD, "silero-16k") on the GPU, behind the reference's whisper_vad_* API and the
// whisper_full_params::vad pre-filter.
//
// Reference: src/whisper.cpp:4341-5496 -- model file, graph (STFT as a strided convolution with a fixed basis, four
// Conv1d + ReLU layers, one LSTM cell, a 128 -> 1 projection, sigmoid) evaluated chunk by chunk on the CPU (the reference
// forces the CPU backend for it, 4658-4660), the probability -> speech segment state machine (5196-5440), and the audio
// filter + time map of whisper_full (6643-6825, 7947-8025).
//
// Here the network is split where its only sequential dependency sits:
//   vad_frontend_kernel  every 512-sample chunk is independent up to and including the LSTM's input projection:
//                        reflect pad -> STFT magnitudes -> 4 x conv + ReLU -> W_ih x + b_ih.  One CTA per few chunks,
//                        all chunks of all streams of a call in one launch.
//   vad_lstm_kernel      the recurrence h, c over the chunks of ONE stream: 512 threads = 512 gate rows of W_hh, one step per
//                        chunk, the 128 -> 1 projection and sigmoid fused.  Streams run side by side (grid.x = streams).
// Rounding points follow the reference's ggml graph: convolution inputs are rounded to f16 (ggml_conv_1d = im2col to F16 +
// mul_mat), convolution weights are F16 in the file, the LSTM matrices and all accumulations are f32.
#include "vad.h"

#include <float.h>
#include <limits.h>
#include <math.h>
#include <string.h>

#include <algorithm>

namespace wb {

namespace {

constexpr int V_WIN = 512;          // samples per chunk (file header n_window)
constexpr int V_PAD = 64;           // reflect padding on both sides
constexpr int V_FRAME = V_WIN + 2 * V_PAD;
constexpr int V_STFT_K = 256, V_STFT_HOP = 128, V_STFT_T = 4, V_BINS = 129;
constexpr int V_H = 128;            // LSTM width
constexpr int V_CH = 4;             // chunks per CTA of the front-end kernel
constexpr int V_THREADS = 256;

__device__ __forceinline__ float r16(float x) { return __half2float(__float2half_rn(x)); }

struct VadWeights {
    const __half * stft;                 // [258][256]
    const __half * w0; const float * b0; // [128][129][3]
    const __half * w1; const float * b1; // [64][128][3]
    const __half * w2; const float * b2; // [64][64][3]
    const __half * w3; const float * b3; // [128][64][3]
    const float * w_ih; const float * b_ih;   // [512][128]
    const float * w_hh; const float * b_hh;   // [512][128]
    const __half * w_f; const float * b_f;    // [128], [1]
};

struct VadStream {
    const float * pcm;       // device
    int n_samples;
    int n_chunks;
    int chunk0;              // first row of this stream in the gate / probability arrays
    float * h;               // [128] LSTM state of this stream (device)
    float * c;
};

// out[o][t] = relu(b[o] + sum_{c,k} w[o][c][k] * in16[c][t * stride + k - 1]), zero padded; in16 already rounded to f16
template <int CIN, int COUT, int TIN, int TOUT, int STRIDE>
__device__ __forceinline__ void conv3_relu(const __half * __restrict__ w, const float * __restrict__ b, const float * s_in /*[V_CH][CIN][TIN]*/,
                                           float * s_out /*[V_CH][COUT][TOUT]*/, bool round_out) {
    for (int idx = threadIdx.x; idx < COUT * TOUT; idx += V_THREADS) {
        const int o = idx / TOUT, t = idx % TOUT;
        float acc[V_CH];
#pragma unroll
        for (int q = 0; q < V_CH; ++q) acc[q] = 0.0f;
        const __half * wr = w + (size_t) o * CIN * 3;
        for (int c = 0; c < CIN; ++c) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int ti = t * STRIDE + k - 1;
                if (ti < 0 || ti >= TIN) continue;
                const float wv = __half2float(__ldg(wr + c * 3 + k));
#pragma unroll
                for (int q = 0; q < V_CH; ++q) acc[q] = fmaf(wv, s_in[(q * CIN + c) * TIN + ti], acc[q]);
            }
        }
        const float bias = __ldg(b + o);
#pragma unroll
        for (int q = 0; q < V_CH; ++q) {
            const float y = fmaxf(acc[q] + bias, 0.0f);
            s_out[(q * COUT + o) * TOUT + t] = round_out ? r16(y) : y;
        }
    }
}

__global__ void __launch_bounds__(V_THREADS)
vad_frontend_kernel(const VadStream * __restrict__ streams, VadWeights W, float * __restrict__ gates /*[rows][512]*/) {
    __shared__ float s_frame[V_CH][V_FRAME];            // padded chunk, rounded to f16
    __shared__ float s_a[V_CH * V_BINS * V_STFT_T];     // ping
    __shared__ float s_b[V_CH * 128 * V_STFT_T];        // pong
    const VadStream st = streams[blockIdx.y];
    const int c0 = blockIdx.x * V_CH;
    if (c0 >= st.n_chunks) return;
    const int tid = threadIdx.x;
    // chunk i = samples [512 i, 512 i + 512), zero filled past the end; reflect 64 samples on both sides
    for (int idx = tid; idx < V_CH * V_FRAME; idx += V_THREADS) {
        const int q = idx / V_FRAME, p = idx % V_FRAME;
        const int j = p < V_PAD ? V_PAD - p : (p < V_PAD + V_WIN ? p - V_PAD : 2 * V_WIN + V_PAD - 2 - p);
        const long long s = (long long) (c0 + q) * V_WIN + j;
        s_frame[q][p] = (c0 + q < st.n_chunks && s < st.n_samples) ? r16(st.pcm[s]) : 0.0f;
    }
    __syncthreads();
    // STFT: 258 basis rows x 4 hops; magnitude of (row j, row 129 + j)
    for (int j = tid; j < V_BINS; j += V_THREADS) {
        float re[V_CH][V_STFT_T], im[V_CH][V_STFT_T];
#pragma unroll
        for (int q = 0; q < V_CH; ++q)
#pragma unroll
            for (int t = 0; t < V_STFT_T; ++t) re[q][t] = im[q][t] = 0.0f;
        const __half * br = W.stft + (size_t) j * V_STFT_K, * bi = W.stft + (size_t) (V_BINS + j) * V_STFT_K;
        for (int k = 0; k < V_STFT_K; ++k) {
            const float wr = __half2float(__ldg(br + k)), wi = __half2float(__ldg(bi + k));
#pragma unroll
            for (int q = 0; q < V_CH; ++q)
#pragma unroll
                for (int t = 0; t < V_STFT_T; ++t) {
                    const float x = s_frame[q][t * V_STFT_HOP + k];
                    re[q][t] = fmaf(wr, x, re[q][t]);
                    im[q][t] = fmaf(wi, x, im[q][t]);
                }
        }
#pragma unroll
        for (int q = 0; q < V_CH; ++q)
#pragma unroll
            for (int t = 0; t < V_STFT_T; ++t)
                s_a[(q * V_BINS + j) * V_STFT_T + t] = r16(sqrtf(re[q][t] * re[q][t] + im[q][t] * im[q][t]));
    }
    __syncthreads();
    conv3_relu<129, 128, 4, 4, 1>(W.w0, W.b0, s_a, s_b, true);
    __syncthreads();
    conv3_relu<128, 64, 4, 2, 2>(W.w1, W.b1, s_b, s_a, true);
    __syncthreads();
    conv3_relu<64, 64, 2, 1, 2>(W.w2, W.b2, s_a, s_b, true);
    __syncthreads();
    conv3_relu<64, 128, 1, 1, 1>(W.w3, W.b3, s_b, s_a, false);      // x_t [128] per chunk, f32 (feeds an f32 matrix product)
    __syncthreads();
    // LSTM input projection: gates_in = W_ih x + b_ih
    for (int g = tid; g < 4 * V_H; g += V_THREADS) {
        float acc[V_CH];
#pragma unroll
        for (int q = 0; q < V_CH; ++q) acc[q] = 0.0f;
        const float4 * wr = reinterpret_cast<const float4 *>(W.w_ih + (size_t) g * V_H);
        for (int k4 = 0; k4 < V_H / 4; ++k4) {
            const float4 w = __ldg(wr + k4);
#pragma unroll
            for (int q = 0; q < V_CH; ++q) {
                const float * x = s_a + q * V_H + 4 * k4;
                acc[q] = fmaf(w.x, x[0], acc[q]);
                acc[q] = fmaf(w.y, x[1], acc[q]);
                acc[q] = fmaf(w.z, x[2], acc[q]);
                acc[q] = fmaf(w.w, x[3], acc[q]);
            }
        }
        const float bias = __ldg(W.b_ih + g);
#pragma unroll
        for (int q = 0; q < V_CH; ++q)
            if (c0 + q < st.n_chunks) gates[(size_t) (st.chunk0 + c0 + q) * (4 * V_H) + g] = acc[q] + bias;
    }
}

// One CTA per stream, thread g = gate row g of W_hh.  The matrix never leaves the SM: columns 0..63 of a row sit in the
// thread's registers (16 float4), columns 64..127 in shared memory, transposed so that the 512 threads read consecutive
// words (128 KB).  One step per chunk: 512 dot products with h, the gate non-linearities on 128 threads, the 128 -> 1
// projection and the sigmoid.
constexpr int V_LSTM_SMEM = (V_H / 2) * 4 * V_H * (int) sizeof(float);
__global__ void __launch_bounds__(4 * V_H, 1)
vad_lstm_kernel(const VadStream * __restrict__ streams, VadWeights W, const float * __restrict__ gates, float * __restrict__ probs) {
    extern __shared__ __align__(16) float s_w[];           // [64][512]: W_hh[g][64 + k] at s_w[k * 512 + g]
    __shared__ __align__(16) float s_h[V_H];
    __shared__ float s_g[4 * V_H];
    __shared__ float s_red[4];
    const VadStream st = streams[blockIdx.x];
    const int g = threadIdx.x;
    float4 w[V_H / 8];
    {
        const float4 * wr = reinterpret_cast<const float4 *>(W.w_hh + (size_t) g * V_H);
#pragma unroll
        for (int k = 0; k < V_H / 8; ++k) w[k] = __ldg(wr + k);
        for (int k = 0; k < V_H / 8; ++k) {
            const float4 v = __ldg(wr + V_H / 8 + k);
            s_w[(4 * k + 0) * (4 * V_H) + g] = v.x;
            s_w[(4 * k + 1) * (4 * V_H) + g] = v.y;
            s_w[(4 * k + 2) * (4 * V_H) + g] = v.z;
            s_w[(4 * k + 3) * (4 * V_H) + g] = v.w;
        }
    }
    const float b_hh = __ldg(W.b_hh + g);
    float c_state = 0.0f, wf = 0.0f;
    if (g < V_H) {
        s_h[g] = st.h[g];
        c_state = st.c[g];
        wf = __half2float(__ldg(W.w_f + g));
    }
    const float b_f = __ldg(W.b_f);
    __syncthreads();
    float gin = st.n_chunks > 0 ? gates[(size_t) st.chunk0 * (4 * V_H) + g] : 0.0f;
    for (int i = 0; i < st.n_chunks; ++i) {
        const float gnext = i + 1 < st.n_chunks ? gates[(size_t) (st.chunk0 + i + 1) * (4 * V_H) + g] : 0.0f;    // off the critical path
        float acc = 0.0f;
#pragma unroll
        for (int k = 0; k < V_H / 8; ++k) {
            const float4 h = *reinterpret_cast<const float4 *>(s_h + 4 * k);
            acc = fmaf(w[k].x, h.x, acc);
            acc = fmaf(w[k].y, h.y, acc);
            acc = fmaf(w[k].z, h.z, acc);
            acc = fmaf(w[k].w, h.w, acc);
        }
#pragma unroll 16
        for (int k = 0; k < V_H / 2; ++k) acc = fmaf(s_w[k * (4 * V_H) + g], s_h[V_H / 2 + k], acc);
        s_g[g] = gin + (acc + b_hh);
        __syncthreads();
        float part = 0.0f;
        if (g < V_H) {
            const float it = 1.0f / (1.0f + expf(-s_g[g])), ft = 1.0f / (1.0f + expf(-s_g[V_H + g]));
            const float gt = tanhf(s_g[2 * V_H + g]), ot = 1.0f / (1.0f + expf(-s_g[3 * V_H + g]));
            c_state = ft * c_state + it * gt;
            const float h = ot * tanhf(c_state);
            s_h[g] = h;
            part = wf * r16(fmaxf(h, 0.0f));                 // the 128 -> 1 projection is a convolution: its input goes through f16
            part = warp_sum(part);
            if ((g & 31) == 0) s_red[g >> 5] = part;
        }
        __syncthreads();
        if (g == 0) {
            const float z = ((s_red[0] + s_red[1]) + (s_red[2] + s_red[3])) + b_f;
            probs[st.chunk0 + i] = 1.0f / (1.0f + expf(-z));
        }
        gin = gnext;
    }
    if (g < V_H) {
        st.h[g] = s_h[g];
        st.c[g] = c_state;
    }
}

template <typename T> bool read_pod(whisper_model_loader * l, T & v) { return l->read(l->context, &v, sizeof(T)) == sizeof(T); }

}  // namespace

VadModel::~VadModel() {
    for (void * p : allocs) cudaFree(p);
}

// File layout (reference src/whisper.cpp:4761-5075; writer models/convert-silero-vad-to-ggml.py): magic, model-type string,
// 3 x i32 version, n_window, n_context, n_encoder_layers x {in, out, kernel}, lstm_input, lstm_hidden, final_in, final_out,
// then tensor records {n_dims, name length, type (0 f32 / 1 f16), dims, name, data}.
bool vad_model_load(whisper_model_loader * loader, VadModel & m, int device) {
    uint32_t magic = 0;
    if (!read_pod(loader, magic) || magic != 0x67676d6c) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: invalid model data (bad magic)\n", __func__);
        return false;
    }
    int32_t len = 0;
    if (!read_pod(loader, len) || len < 0 || len > 256) return false;
    m.type.resize(len);
    if (len && loader->read(loader->context, &m.type[0], len) != (size_t) len) return false;
    int32_t ver[3], n_layers = 0;
    if (!read_pod(loader, ver[0]) || !read_pod(loader, ver[1]) || !read_pod(loader, ver[2]) || !read_pod(loader, m.n_window) ||
        !read_pod(loader, m.n_context) || !read_pod(loader, n_layers))
        return false;
    m.version = std::to_string(ver[0]) + "." + std::to_string(ver[1]) + "." + std::to_string(ver[2]);
    wlog(GGML_LOG_LEVEL_INFO, "%s: model type: %s, version: %s\n", __func__, m.type.c_str(), m.version.c_str());
    if (n_layers != 4) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: %d encoder layers: only the 4-layer silero-16k geometry is implemented\n", __func__, n_layers);
        return false;
    }
    int32_t enc[4][3], tail[4];
    for (auto & e : enc)
        for (int32_t & v : e)
            if (!read_pod(loader, v)) return false;
    for (int32_t & v : tail)
        if (!read_pod(loader, v)) return false;
    static const int32_t want[4][3] = {{129, 128, 3}, {128, 64, 3}, {64, 64, 3}, {64, 128, 3}};
    if (memcmp(enc, want, sizeof(want)) != 0 || tail[0] != 128 || tail[1] != 128 || tail[2] != 128 || tail[3] != 1 || m.n_window != V_WIN) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: unsupported VAD geometry (expected silero-16k: window 512, convs 129-128-64-64-128, LSTM 128)\n", __func__);
        return false;
    }
    if (cudaSetDevice(device) != cudaSuccess) return false;
    struct Want { const char * name; size_t n; bool f16; void ** dst; };
    VadModel::Ptrs & P = m.p;
    const Want tensors[] = {
        {"_model.stft.forward_basis_buffer", 258 * 256, true, (void **) &P.stft},
        {"_model.encoder.0.reparam_conv.weight", 128 * 129 * 3, true, (void **) &P.w0}, {"_model.encoder.0.reparam_conv.bias", 128, false, (void **) &P.b0},
        {"_model.encoder.1.reparam_conv.weight", 64 * 128 * 3, true, (void **) &P.w1},  {"_model.encoder.1.reparam_conv.bias", 64, false, (void **) &P.b1},
        {"_model.encoder.2.reparam_conv.weight", 64 * 64 * 3, true, (void **) &P.w2},   {"_model.encoder.2.reparam_conv.bias", 64, false, (void **) &P.b2},
        {"_model.encoder.3.reparam_conv.weight", 128 * 64 * 3, true, (void **) &P.w3},  {"_model.encoder.3.reparam_conv.bias", 128, false, (void **) &P.b3},
        {"_model.decoder.rnn.weight_ih", 512 * 128, false, (void **) &P.w_ih}, {"_model.decoder.rnn.bias_ih", 512, false, (void **) &P.b_ih},
        {"_model.decoder.rnn.weight_hh", 512 * 128, false, (void **) &P.w_hh}, {"_model.decoder.rnn.bias_hh", 512, false, (void **) &P.b_hh},
        {"_model.decoder.decoder.2.weight", 128, true, (void **) &P.w_f}, {"_model.decoder.decoder.2.bias", 1, false, (void **) &P.b_f},
    };
    int n_loaded = 0;
    std::vector<char> buf;
    while (true) {
        int32_t n_dims = 0, name_len = 0, ttype = 0;
        if (!read_pod(loader, n_dims) || !read_pod(loader, name_len) || !read_pod(loader, ttype)) break;
        if (loader->eof(loader->context)) break;
        if (n_dims < 0 || n_dims > 4 || name_len <= 0 || name_len > 256 || (ttype != 0 && ttype != 1)) return false;
        size_t n = 1;
        for (int i = 0; i < n_dims; ++i) {
            int32_t d = 0;
            if (!read_pod(loader, d) || d <= 0) return false;
            n *= (size_t) d;
        }
        std::string name(name_len, '\0');
        if (loader->read(loader->context, &name[0], name_len) != (size_t) name_len) return false;
        const Want * w = nullptr;
        for (const Want & t : tensors)
            if (name == t.name) w = &t;
        if (!w) {
            wlog(GGML_LOG_LEVEL_ERROR, "%s: unknown tensor '%s' in model file\n", __func__, name.c_str());
            return false;
        }
        if (n != w->n || (ttype == 1) != w->f16) {
            wlog(GGML_LOG_LEVEL_ERROR, "%s: tensor '%s' has wrong size / type in model file\n", __func__, name.c_str());
            return false;
        }
        const size_t bytes = n * (w->f16 ? 2 : 4);
        buf.resize(bytes);
        if (loader->read(loader->context, buf.data(), bytes) != bytes) return false;
        void * d = nullptr;
        if (cudaMalloc(&d, bytes) != cudaSuccess) return false;
        m.allocs.push_back(d);
        if (cudaMemcpy(d, buf.data(), bytes, cudaMemcpyHostToDevice) != cudaSuccess) return false;
        *w->dst = d;
        ++n_loaded;
    }
    if (n_loaded != (int) (sizeof(tensors) / sizeof(tensors[0]))) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: ERROR not all tensors loaded from model file - expected %d, got %d\n", __func__,
             (int) (sizeof(tensors) / sizeof(tensors[0])), n_loaded);
        return false;
    }
    return true;
}

bool vad_run(const VadModel & m, const std::vector<VadJob> & jobs, cudaStream_t stream, DeviceBlock & scratch) {
    const int n = (int) jobs.size();
    if (n == 0) return true;
    std::vector<VadStream> sts(n);
    int rows = 0, max_chunks = 0;
    for (int i = 0; i < n; ++i) {
        const int nc = (jobs[i].n_samples + V_WIN - 1) / V_WIN;
        sts[i] = {jobs[i].pcm_dev, jobs[i].n_samples, nc, rows, jobs[i].h, jobs[i].c};
        rows += nc;
        max_chunks = std::max(max_chunks, nc);
    }
    if (rows == 0) return true;
    // scratch: stream descriptors | gates [rows][512] | probs [rows]
    const size_t off_g = round_up<size_t>(n * sizeof(VadStream), 256), off_p = off_g + round_up<size_t>((size_t) rows * 512 * 4, 256);
    if (!scratch.reserve(off_p + (size_t) rows * 4)) return false;
    WB_CUDA(cudaMemcpyAsync(scratch.p, sts.data(), n * sizeof(VadStream), cudaMemcpyHostToDevice, stream));
    VadWeights W;
    W.stft = (const __half *) m.p.stft;
    W.w0 = (const __half *) m.p.w0; W.b0 = m.p.b0; W.w1 = (const __half *) m.p.w1; W.b1 = m.p.b1;
    W.w2 = (const __half *) m.p.w2; W.b2 = m.p.b2; W.w3 = (const __half *) m.p.w3; W.b3 = m.p.b3;
    W.w_ih = m.p.w_ih; W.b_ih = m.p.b_ih; W.w_hh = m.p.w_hh; W.b_hh = m.p.b_hh; W.w_f = (const __half *) m.p.w_f; W.b_f = m.p.b_f;
    float * gates = (float *) ((char *) scratch.p + off_g), * probs = (float *) ((char *) scratch.p + off_p);
    vad_frontend_kernel<<<dim3(ceil_div(max_chunks, V_CH), n), V_THREADS, 0, stream>>>((const VadStream *) scratch.p, W, gates);
    static DeviceOnce once;
    once_per_device(once, [&] { WB_CUDA(cudaFuncSetAttribute(vad_lstm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, V_LSTM_SMEM)); });
    vad_lstm_kernel<<<n, 4 * V_H, V_LSTM_SMEM, stream>>>((const VadStream *) scratch.p, W, gates, probs);
    WB_CUDA(cudaGetLastError());
    std::vector<float> host(rows);
    WB_CUDA(cudaMemcpyAsync(host.data(), probs, (size_t) rows * 4, cudaMemcpyDeviceToHost, stream));
    WB_CUDA(cudaStreamSynchronize(stream));
    if (cuda_failed()) return false;
    for (int i = 0; i < n; ++i) jobs[i].probs->assign(host.begin() + sts[i].chunk0, host.begin() + sts[i].chunk0 + sts[i].n_chunks);
    return true;
}

// ---- probabilities -> speech segments (reference whisper_vad_segments_from_probs, src/whisper.cpp:5196-5440) ------------------
// A hysteresis detector over one probability per 512-sample chunk: speech starts at p >= threshold, ends after
// min_silence of p < threshold - 0.15, is split when it outgrows max_speech (at the last silence of >= 98 ms if there was one),
// short speeches are dropped, near neighbours merged, and the survivors padded.  Sample arithmetic is the reference's.
std::vector<VadSegment> vad_segments_from_probs(const float * probs, int n_probs, int n_window, const whisper_vad_params & prm) {
    const int rate = WHISPER_SAMPLE_RATE;
    const int total = n_probs * n_window;
    const int min_silence = rate * prm.min_silence_duration_ms / 1000, min_speech = rate * prm.min_speech_duration_ms / 1000;
    const int pad = rate * prm.speech_pad_ms / 1000, min_silence_at_max = rate * 98 / 1000;
    int max_speech;
    if (prm.max_speech_duration_s > 100000.0f) {
        max_speech = INT_MAX / 2;
    } else {
        const int64_t t = (int64_t) rate * (int64_t) (prm.max_speech_duration_s) - n_window - 2 * pad;
        max_speech = t > INT_MAX ? INT_MAX / 2 : (int) t;
        if (max_speech < 0) max_speech = INT_MAX / 2;
    }
    const float hi = prm.threshold, lo = std::max(prm.threshold - 0.15f, 0.01f);

    struct Span { int start, end; };
    std::vector<Span> spans;
    struct {
        bool in_speech = false, open = false;   // inside a speech / a speech has been opened and not yet closed or dropped
        int start = 0;                          // first sample of the current speech
        int quiet_since = 0;                    // first sample of the current sub-threshold stretch (0: none)
        int last_quiet = 0;                     // start of the last stretch that lasted >= 98 ms: where a too-long speech is cut
        int resume = 0;                         // where speech came back after that stretch
    } s;
    auto forget_quiet = [&]() { s.last_quiet = s.resume = s.quiet_since = 0; };
    for (int i = 0; i < n_probs; ++i) {
        const float p = probs[i];
        const int at = n_window * i;
        if (p >= hi && s.quiet_since) {                       // speech again: the quiet stretch is over
            s.quiet_since = 0;
            if (s.resume < s.last_quiet) s.resume = at;
        }
        if (p >= hi && !s.in_speech) {
            s.in_speech = s.open = true;
            s.start = at;
            continue;
        }
        if (s.in_speech && at - s.start > max_speech) {        // too long: cut at the last long-enough silence, or right here
            if (s.last_quiet) {
                spans.push_back({s.start, s.last_quiet});
                s.open = true;
                if (s.resume < s.last_quiet) s.in_speech = s.open = false;
                else s.start = s.resume;
                forget_quiet();
            } else {
                spans.push_back({s.start, at});
                forget_quiet();
                s.in_speech = s.open = false;
                continue;
            }
        }
        if (p < lo && s.in_speech) {
            if (!s.quiet_since) s.quiet_since = at;
            if (at - s.quiet_since > min_silence_at_max) s.last_quiet = s.quiet_since;
            if (at - s.quiet_since < min_silence) continue;
            if (s.quiet_since - s.start > min_speech) spans.push_back({s.start, s.quiet_since});
            forget_quiet();
            s.in_speech = s.open = false;
            continue;
        }
    }
    if (s.open && total - s.start > min_speech) spans.push_back({s.start, total});

    // neighbours closer than 200 ms become one; anything still shorter than min_speech goes
    const int merge_gap = rate * 200 / 1000;
    for (size_t i = 0; i + 1 < spans.size();) {
        if (spans[i + 1].start - spans[i].end < merge_gap) {
            spans[i].end = spans[i + 1].end;
            spans.erase(spans.begin() + i + 1);
        } else {
            ++i;
        }
    }
    spans.erase(std::remove_if(spans.begin(), spans.end(), [&](const Span & x) { return x.end - x.start < min_speech; }), spans.end());

    // padding: full pad at the outer ends and where the gap allows, otherwise the gap is split
    std::vector<VadSegment> out(spans.size());
    for (size_t i = 0; i < spans.size(); ++i) {
        if (i == 0) spans[i].start = spans[i].start > pad ? spans[i].start - pad : 0;
        if (i + 1 < spans.size()) {
            const int gap = spans[i + 1].start - spans[i].end;
            if (gap < 2 * pad) {
                spans[i].end += gap / 2;
                spans[i + 1].start = spans[i + 1].start > gap / 2 ? spans[i + 1].start - gap / 2 : 0;
            } else {
                spans[i].end = spans[i].end + pad < total ? spans[i].end + pad : total;
                spans[i + 1].start = spans[i + 1].start > pad ? spans[i + 1].start - pad : 0;
            }
        } else {
            spans[i].end = spans[i].end + pad < total ? spans[i].end + pad : total;
        }
        out[i].start = vad_samples_to_cs(spans[i].start);
        out[i].end = vad_samples_to_cs(spans[i].end);
    }
    return out;
}

}  // namespace wb
