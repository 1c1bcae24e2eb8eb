// Host interface of the fused log-mel kernel (mel.cu).
#pragma once

#include "common.cuh"

namespace wb {

// One audio stream of a batched launch (lives in device memory; blockIdx.y selects it).
struct MelStream {
    const float * pcm;        // device, mono 16 kHz: f32 in [-1, 1], or (pcm_i16 != 0) int16 behind the same pointer
    int           n_samples;
    int           n_frames_fft;  // frames that see audio (the rest of the reference's n_len frames are the constant -10)
    float *       out;        // device, raw log10 mel [n_mel][out_stride]
    int           out_stride;
    unsigned *    max_enc;    // device, ordered-uint encoding of the running max (init 0)
    int           pcm_i16;    // 16-bit PCM ingest fused into the load: x = s / 32768 (what the reference's decoder front-end,
                              // examples/common-whisper.cpp:42-134 via miniaudio's s16 -> f32, hands to whisper_pcm_to_mel)
};

struct MelGeometry {
    int n_len;         // frames the reference's container would hold   (src/whisper.cpp:3206)
    int n_len_org;     // frames reported by whisper_n_len()            (src/whisper.cpp:3208)
    int n_frames_fft;  // frames actually run through the FFT           (src/whisper.cpp:3117)
    int stride;        // row stride of our raw buffer
};

struct MelPlan {
    int     n_mel = 0;
    void *  d_tables = nullptr;
    void *  d_w = nullptr;      // packed sparse filterbank: group meta | per-bin k_start | taps [group][tap][lane]
    int     n_groups = 0;       // groups of 8 mel bins
    int     filt_floats = 0;
    size_t  smem_bytes = 0;
    MelPlan() = default;
    MelPlan(const MelPlan &) = delete;
    MelPlan & operator=(const MelPlan &) = delete;
    ~MelPlan();
};

#if defined(__CUDACC__)
__host__ __device__
#endif
inline float mel_decode_max(unsigned enc) {
    const unsigned b = (enc & 0x80000000u) ? (enc & 0x7fffffffu) : ~enc;
    float f;
#if defined(__CUDA_ARCH__)
    f = __uint_as_float(b);
#else
    __builtin_memcpy(&f, &b, 4);
#endif
    return f;
}

MelGeometry mel_geometry(int n_samples);
bool mel_plan_init(MelPlan & plan, const float * filters /*[n_mel][201] host*/, int n_mel, int n_fft_bins);
void mel_launch(const MelPlan & plan, const MelStream * d_streams, int n_streams, int max_frames_fft,
                cudaStream_t stream);
void mel_finalize_launch(const float * raw, int raw_stride, int n_frames_fft, const unsigned * max_enc, float * out,
                         int n_len, int n_mel, cudaStream_t stream);

}  // namespace wb
