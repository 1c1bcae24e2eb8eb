// Voice activity detection on the GPU (vad.cu): Silero VAD model, batched front end + per-stream LSTM recurrence, and the
// probability -> speech segment logic.  Reference: src/whisper.cpp:4341-5496.
#pragma once

#include <string>
#include <vector>

#include "engine.h"

namespace wb {

struct VadModel {
    std::string type, version;
    int32_t n_window = 512, n_context = 64;
    struct Ptrs {          // device
        void *stft = nullptr, *w0 = nullptr, *w1 = nullptr, *w2 = nullptr, *w3 = nullptr, *w_f = nullptr;      // f16
        float *b0 = nullptr, *b1 = nullptr, *b2 = nullptr, *b3 = nullptr, *w_ih = nullptr, *b_ih = nullptr, *w_hh = nullptr, *b_hh = nullptr,
              *b_f = nullptr;
    } p;
    std::vector<void *> allocs;
    VadModel() = default;
    VadModel(const VadModel &) = delete;
    VadModel & operator=(const VadModel &) = delete;
    ~VadModel();
};

bool vad_model_load(whisper_model_loader * loader, VadModel & m, int device);

struct VadJob {
    const float * pcm_dev;          // device PCM of one stream
    int n_samples;
    float * h, * c;                 // device LSTM state [128] each, read and updated
    std::vector<float> * probs;     // out: one probability per 512-sample chunk
};
// all streams in two launches (front end over every chunk, recurrence with one CTA per stream)
bool vad_run(const VadModel & m, const std::vector<VadJob> & jobs, cudaStream_t stream, DeviceBlock & scratch);

struct VadSegment { int64_t start, end; };      // centiseconds
inline int vad_cs_to_samples(int64_t cs) { return (int) ((cs / 100.0) * WHISPER_SAMPLE_RATE + 0.5); }
inline int64_t vad_samples_to_cs(int samples) { return (int64_t) ((samples / (double) WHISPER_SAMPLE_RATE) * 100.0 + 0.5); }
std::vector<VadSegment> vad_segments_from_probs(const float * probs, int n_probs, int n_window, const whisper_vad_params & prm);

}  // namespace wb
