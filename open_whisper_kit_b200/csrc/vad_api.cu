// whisper_vad_* of include/whisper.h on the GPU detector (vad.cu), and the VAD pre-filter of whisper_full / whisper_full_parallel:
// speech segments are cut out of the audio, joined with 0.1 s of silence, transcribed, and segment times are mapped back through
// a piecewise-linear table.  Reference: src/whisper.cpp:4429-5496 (API), 6643-6825 (filter), 7947-8025 (time map).
#include <float.h>
#include <string.h>

#include <algorithm>
#include <fstream>

#include "full.h"
#include "vad.h"
#include "vad_api.h"
#include "whisper_b200.h"

using namespace wb;

struct whisper_vad_context {
    VadModel model;
    int device = 0;
    cudaStream_t stream = nullptr;
    DeviceBlock lstm_state;      // h | c, 128 floats each; zero = reset
    DeviceBlock pcm, scratch;
    std::vector<float> probs;
    int64_t t_vad_us = 0;
    std::mutex mu;
    ~whisper_vad_context() {
        if (stream) cudaStreamDestroy(stream);
    }
};

struct whisper_vad_segments {
    std::vector<VadSegment> data;
};

namespace {

bool detect(whisper_vad_context * v, const float * samples, int n_samples, bool reset) {
    if (!v || !samples || n_samples < 0) return false;
    std::lock_guard<std::mutex> lock(v->mu);
    cuda_clear_failure();
    WB_CUDA(cudaSetDevice(v->device));
    const int64_t t0 = time_us();
    const int n_chunks = (n_samples + v->model.n_window - 1) / v->model.n_window;
    wlog(GGML_LOG_LEVEL_INFO, "%s: detecting speech in %d samples (reset=%d), n_chunks: %d\n", __func__, n_samples, reset ? 1 : 0, n_chunks);
    if (reset) WB_CUDA(cudaMemsetAsync(v->lstm_state.p, 0, 256 * sizeof(float), v->stream));
    v->probs.clear();
    if (n_samples == 0) return !cuda_failed();
    if (!v->pcm.reserve((size_t) n_samples * sizeof(float))) return false;
    WB_CUDA(cudaMemcpyAsync(v->pcm.p, samples, (size_t) n_samples * sizeof(float), cudaMemcpyHostToDevice, v->stream));
    std::vector<VadJob> jobs(1);
    jobs[0] = {(const float *) v->pcm.p, n_samples, (float *) v->lstm_state.p, (float *) v->lstm_state.p + 128, &v->probs};
    const bool ok = vad_run(v->model, jobs, v->stream, v->scratch);
    v->t_vad_us += time_us() - t0;
    wlog(GGML_LOG_LEVEL_INFO, "%s: vad time = %.2f ms processing %d samples\n", __func__, 1e-3f * v->t_vad_us, n_samples);
    return ok;
}

struct VadFile {
    std::ifstream fin;
};

}  // namespace

extern "C" {

struct whisper_vad_context_params whisper_vad_default_context_params(void) {
    // (use_gpu is false in the reference, which forces its CPU backend for the detector; this library has no CPU path and
    // always runs it on gpu_device)
    whisper_vad_context_params r = {4, false, 0};
    return r;
}

struct whisper_vad_context * whisper_vad_init_with_params(struct whisper_model_loader * loader, struct whisper_vad_context_params params) {
    if (!loader || !loader->read) return nullptr;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0 || params.gpu_device < 0 || params.gpu_device >= n_dev) {
        cudaGetLastError();
        wlog(GGML_LOG_LEVEL_ERROR, "%s: CUDA device %d is not available; the VAD has no CPU fallback\n", __func__, params.gpu_device);
        if (loader->close) loader->close(loader->context);
        return nullptr;
    }
    whisper_vad_context * v = nullptr;
    try {
        v = new whisper_vad_context();
        v->device = params.gpu_device;
        cuda_clear_failure();
        bool ok = vad_model_load(loader, v->model, v->device);
        if (loader->close) loader->close(loader->context);
        ok = ok && cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking) == cudaSuccess && v->lstm_state.reserve(256 * sizeof(float));
        if (ok) ok = cudaMemset(v->lstm_state.p, 0, 256 * sizeof(float)) == cudaSuccess;
        if (!ok) {
            wlog(GGML_LOG_LEVEL_ERROR, "%s: failed to load the VAD model\n", __func__);
            delete v;
            return nullptr;
        }
    } catch (const std::exception & ex) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: %s\n", __func__, ex.what());
        delete v;
        return nullptr;
    }
    return v;
}

struct whisper_vad_context * whisper_vad_init_from_file_with_params(const char * path_model, struct whisper_vad_context_params params) {
    if (!path_model) return nullptr;
    wlog(GGML_LOG_LEVEL_INFO, "%s: loading VAD model from '%s'\n", __func__, path_model);
    VadFile * f = new VadFile();
    f->fin.open(path_model, std::ios::binary);
    if (!f->fin) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: failed to open VAD model '%s'\n", __func__, path_model);
        delete f;
        return nullptr;
    }
    whisper_model_loader loader = {};
    loader.context = f;
    loader.read = [](void * c, void * out, size_t n) -> size_t {
        auto * vf = (VadFile *) c;
        vf->fin.read((char *) out, n);
        return (size_t) vf->fin.gcount();
    };
    loader.eof = [](void * c) -> bool { return ((VadFile *) c)->fin.eof(); };
    loader.close = [](void * c) { delete (VadFile *) c; };
    return whisper_vad_init_with_params(&loader, params);
}

bool whisper_vad_detect_speech(struct whisper_vad_context * vctx, const float * samples, int n_samples) {
    return detect(vctx, samples, n_samples, true);
}
bool whisper_vad_detect_speech_stateful(struct whisper_vad_context * vctx, const float * samples, int n_samples) {
    return detect(vctx, samples, n_samples, false);
}
void whisper_vad_reset_state(struct whisper_vad_context * vctx) {
    if (!vctx) return;
    std::lock_guard<std::mutex> lock(vctx->mu);
    cudaSetDevice(vctx->device);
    cudaMemset(vctx->lstm_state.p, 0, 256 * sizeof(float));
}
int whisper_vad_n_probs(struct whisper_vad_context * vctx) { return (int) vctx->probs.size(); }
float * whisper_vad_probs(struct whisper_vad_context * vctx) { return vctx->probs.data(); }

struct whisper_vad_segments * whisper_vad_segments_from_probs(struct whisper_vad_context * vctx, struct whisper_vad_params params) {
    if (!vctx) return nullptr;
    try {
        whisper_vad_segments * s = new whisper_vad_segments;
        s->data = vad_segments_from_probs(vctx->probs.data(), (int) vctx->probs.size(), vctx->model.n_window, params);
        for (size_t i = 0; i < s->data.size(); ++i)
            wlog(GGML_LOG_LEVEL_INFO, "%s: VAD segment %d: start = %.2f, end = %.2f (duration: %.2f)\n", __func__, (int) i, s->data[i].start / 100.0,
                 s->data[i].end / 100.0, (s->data[i].end - s->data[i].start) / 100.0);
        return s;
    } catch (const std::exception &) {
        return nullptr;
    }
}

struct whisper_vad_segments * whisper_vad_segments_from_samples(struct whisper_vad_context * vctx, struct whisper_vad_params params,
                                                                const float * samples, int n_samples) {
    if (!whisper_vad_detect_speech(vctx, samples, n_samples)) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: failed to detect speech\n", __func__);
        return nullptr;
    }
    return whisper_vad_segments_from_probs(vctx, params);
}

int whisper_vad_segments_n_segments(struct whisper_vad_segments * segments) { return (int) segments->data.size(); }
float whisper_vad_segments_get_segment_t0(struct whisper_vad_segments * segments, int i) { return (float) segments->data[i].start; }
float whisper_vad_segments_get_segment_t1(struct whisper_vad_segments * segments, int i) { return (float) segments->data[i].end; }
void whisper_vad_free_segments(struct whisper_vad_segments * segments) { delete segments; }
void whisper_vad_free(struct whisper_vad_context * ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    delete ctx;
}

}  // extern "C"

// Host-only hook (needs no device): the probability -> segment logic on explicit probabilities; seg_out[2i], seg_out[2i+1] =
// start, end in centiseconds.  Returns the number of segments (may exceed cap; only cap are written).
extern "C" WB200_API int whisper_b200_vad_segments_from_probs(const float * probs, int n_probs, int n_window, struct whisper_vad_params params,
                                                              long long * seg_out, int cap) {
    if (!probs || n_probs < 0 || n_window <= 0) return -1;
    const std::vector<VadSegment> segs = vad_segments_from_probs(probs, n_probs, n_window, params);
    for (size_t i = 0; i < segs.size() && (int) i < cap; ++i) {
        seg_out[2 * i] = segs[i].start;
        seg_out[2 * i + 1] = segs[i].end;
    }
    return (int) segs.size();
}

// Test hooks mirroring the two internal stages: the filter (needs the device) and the time map on an explicit table (host only).
extern "C" WB200_API int whisper_b200_vad_filter(struct whisper_context * ctx, struct whisper_full_params params, const float * samples,
                                                 int n_samples, float * out, int cap, long long * table, int cap_pairs, int * n_pairs) {
    if (!ctx || !ctx->state || !samples || !n_pairs) return -1;
    std::vector<float> filtered;
    if (!wb::vad_filter(*ctx, *ctx->state, params, samples, n_samples, filtered)) return -2;
    const auto & tab = ctx->state->vad_mapping_table;
    *n_pairs = (int) tab.size();
    for (int i = 0; i < (int) tab.size() && i < cap_pairs; ++i) {
        table[2 * i] = tab[i].processed_time;
        table[2 * i + 1] = tab[i].original_time;
    }
    for (int i = 0; i < (int) filtered.size() && i < cap; ++i) out[i] = filtered[i];
    return (int) filtered.size();
}

extern "C" WB200_API long long whisper_b200_vad_map_time(const long long * table, int n_pairs, long long t) {
    std::vector<whisper_state::vad_time_mapping> tab((size_t) std::max(n_pairs, 0));
    for (int i = 0; i < n_pairs; ++i) tab[i] = {table[2 * i], table[2 * i + 1]};
    return wb::vad_map_time(tab, t);
}

namespace wb {

void vad_free_state_context(whisper_state * st) {
    if (st && st->vad_context) {
        whisper_vad_free(st->vad_context);
        st->vad_context = nullptr;
    }
}

// The pre-filter of whisper_full (reference whisper_vad, src/whisper.cpp:6643-6825): detect, cut the speech out (every segment
// but the last extended by samples_overlap), join the pieces with 0.1 s of zeros, and record (processed time -> original time)
// pairs: both ends of every piece, a point every 200 ms inside pieces longer than 1 s, both ends of every inserted silence.
bool vad_filter(whisper_context & ctx, whisper_state & state, const whisper_full_params & params, const float * samples, int n_samples,
                std::vector<float> & filtered) {
    wlog(GGML_LOG_LEVEL_INFO, "%s: VAD is enabled, processing speech segments only\n", __func__);
    state.vad_mapping_table.clear();
    state.has_vad_segments = false;
    filtered.clear();
    if (!state.vad_context) {
        whisper_vad_context_params vp = whisper_vad_default_context_params();
        vp.gpu_device = ctx.eng.device;
        state.vad_context = whisper_vad_init_from_file_with_params(params.vad_model_path, vp);
        if (!state.vad_context) {
            wlog(GGML_LOG_LEVEL_ERROR, "%s: failed to initialize VAD context\n", __func__);
            return false;
        }
    }
    whisper_vad_segments * segs = whisper_vad_segments_from_samples(state.vad_context, params.vad_params, samples, n_samples);
    if (!segs) return false;
    const std::vector<VadSegment> & sg = segs->data;
    if (!sg.empty()) {
        state.has_vad_segments = true;
        const int overlap = (int) (params.vad_params.samples_overlap * WHISPER_SAMPLE_RATE);
        const int gap = (int) (0.1 * WHISPER_SAMPLE_RATE);
        struct Piece { int first, len; int64_t orig_start, orig_end; };
        std::vector<Piece> pieces;
        long long total = 0;
        for (size_t i = 0; i < sg.size(); ++i) {
            int a = vad_cs_to_samples(sg[i].start), b = vad_cs_to_samples(sg[i].end);
            if (i + 1 < sg.size()) b += overlap;
            b = std::min(b, n_samples - 1);
            total += b - a;                                         // the reference sizes its buffer before clamping the start
            a = std::min(a, n_samples - 1);
            pieces.push_back({a, b - a, sg[i].start, sg[i].end});
        }
        total += (long long) (pieces.size() - 1) * gap;
        long long need = 0;
        for (size_t i = 0; i < pieces.size(); ++i) need += pieces[i].len > 0 ? pieces[i].len + (i + 1 < pieces.size() ? gap : 0) : 0;
        filtered.assign((size_t) std::max(total, need), 0.0f);
        auto & table = state.vad_mapping_table;
        int off = 0;
        for (size_t i = 0; i < pieces.size(); ++i) {
            const Piece & pc = pieces[i];
            if (pc.len <= 0) continue;
            const int64_t v0 = vad_samples_to_cs(off), v1 = vad_samples_to_cs(off + pc.len);
            table.push_back({v0, pc.orig_start});
            table.push_back({v1, pc.orig_end});
            if (v1 - v0 > 100) {                                   // longer than 1 s: interpolation points every 200 ms
                const int n_points = (int) ((v1 - v0) / 20) - 1;
                for (int j = 1; j <= n_points; ++j) {
                    const int64_t vt = v0 + (int64_t) j * 20;
                    if (vt >= v1) continue;
                    table.push_back({vt, pc.orig_start + ((vt - v0) * (pc.orig_end - pc.orig_start)) / (v1 - v0)});
                }
            }
            memcpy(filtered.data() + off, samples + pc.first, (size_t) pc.len * sizeof(float));
            off += pc.len;
            if (i + 1 < pieces.size()) {
                table.push_back({vad_samples_to_cs(off), pc.orig_end});
                table.push_back({vad_samples_to_cs(off + gap), sg[i + 1].start});
                off += gap;                                         // already zero
            }
        }
        // (the buffer keeps the reference's size: when a piece was dropped it is longer than `off` and ends in zeros)
        std::sort(table.begin(), table.end(), [](const whisper_state::vad_time_mapping & a, const whisper_state::vad_time_mapping & b) {
            return a.processed_time < b.processed_time;
        });
        table.erase(std::unique(table.begin(), table.end(), [](const whisper_state::vad_time_mapping & a, const whisper_state::vad_time_mapping & b) {
                        return a.processed_time == b.processed_time;
                    }), table.end());
        wlog(GGML_LOG_LEVEL_INFO, "%s: Reduced audio from %d to %d samples (%.1f%% reduction)\n", __func__, n_samples, off,
             100.0f * (1.0f - (float) off / n_samples));
    }
    whisper_vad_free_segments(segs);
    return true;
}

// processed (filtered-audio) time -> original time: clamp outside the table, exact hit, else linear interpolation in integers
// (reference map_processed_to_original_time, src/whisper.cpp:7947-7989)
int64_t vad_map_time(const std::vector<whisper_state::vad_time_mapping> & tab, int64_t t) {
    if (tab.empty()) return t;
    if (t <= tab.front().processed_time) return tab.front().original_time;
    if (t >= tab.back().processed_time) return tab.back().original_time;
    auto hi = std::lower_bound(tab.begin(), tab.end(), t, [](const whisper_state::vad_time_mapping & e, int64_t v) { return e.processed_time < v; });
    if (hi->processed_time == t) return hi->original_time;
    auto lo = hi - 1;
    const int64_t dp = hi->processed_time - lo->processed_time;
    if (dp == 0) return lo->original_time;
    return lo->original_time + ((t - lo->processed_time) * (hi->original_time - lo->original_time)) / dp;
}

}  // namespace wb
