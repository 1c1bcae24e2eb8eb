// whisper_context / whisper_state of the B200 library (opaque behind include/whisper.h).
#pragma once

#include <random>
#include <string>
#include <vector>

#include "engine.h"

// reference: whisper_segment, src/whisper.cpp:460-470
struct whisper_segment {
    int64_t t0;
    int64_t t1;
    std::string text;
    float no_speech_prob;
    std::vector<whisper_token_data> tokens;
    bool speaker_turn_next;
};

// reference: whisper_sequence / whisper_decoder, src/whisper.cpp:783-820
struct whisper_sequence {
    std::vector<whisper_token_data> tokens;
    int result_len = 0;
    double sum_logprobs_all = 0.0;
    double sum_logprobs = 0.0;
    double avg_logprobs = 0.0;
    double entropy = 0.0;
    double score = 0.0;
};

struct whisper_decoder {
    whisper_sequence sequence;
    int i_batch = 0;
    int seek_delta = 0;
    bool failed = false, completed = false, has_ts = false;

    // host sampling path (temperature > 0, beam search, logits_filter_callback): per-decoder copies as the reference
    std::vector<float> probs, logits, logprobs;
    std::vector<std::pair<double, int>> logits_id;
    std::mt19937 rng;
    std::vector<whisper_token_data> sampled;   // this iteration's host draws (made in parallel across decoders, consumed in order)

    // device path: the token selected on the device right after the decode step, consumed by the next iteration
    bool has_pending = false;
    whisper_token_data pending = {};

    wb::DeviceBlock kv, kv_alt;   // self-attention cache of this sequence [n_text_layer][n_text_ctx][2d]
};

#define WHISPER_MAX_DECODERS 8

struct whisper_state {
    int64_t t_sample_us = 0, t_encode_us = 0, t_decode_us = 0, t_batchd_us = 0, t_prompt_us = 0, t_mel_us = 0;
    int32_t n_sample = 0, n_encode = 0, n_decode = 0, n_batchd = 0, n_prompt = 0, n_fail_p = 0, n_fail_h = 0;

    wb::MelBuf mel;
    wb::CrossKV cross;                 // own single-window pool (low-level API); batched runs use a shared pool
    const void * cross_base = nullptr; // layer-0 K/V of the current window
    size_t cross_layer_stride = 0;
    int cross_T = 1500;                // audio context of that K/V

    whisper_decoder decoders[WHISPER_MAX_DECODERS];

    std::vector<float> logits;         // host copy, [n_tokens][n_vocab] with the last row valid
    std::vector<whisper_segment> result_all;
    std::vector<whisper_token> prompt_past0, prompt_past1;
    int lang_id = 0;
    float no_speech_prob = 0.0f;
    int32_t exp_n_audio_ctx = 0;

    // [EXPERIMENTAL] token-level timestamps (reference whisper_state, src/whisper.cpp:919-927)
    std::vector<float> energy;         // PCM signal energy
    int64_t t_beg = 0, t_last = 0;
    whisper_token tid_last = 0;

    // VAD pre-filter of whisper_full (reference whisper_state, src/whisper.cpp:923-934): the detector is created on first use;
    // the table maps times of the filtered ("processed") audio back to the original recording
    struct whisper_vad_context * vad_context = nullptr;
    bool has_vad_segments = false;
    struct vad_time_mapping { int64_t processed_time, original_time; };
    std::vector<vad_time_mapping> vad_mapping_table;

    struct whisper_context * ctx = nullptr;
    int device = 0;                    // CUDA device of the buffers above (kept here so the state can outlive its context)
};

struct whisper_vad_context;

struct whisper_context {
    int64_t t_load_us = 0;
    int64_t t_start_us = 0;
    whisper_context_params params;
    wb::Engine eng;
    whisper_state * state = nullptr;
    std::string path_model;
    wb::CrossKV batch_cross;           // shared pool of the batched whisper_full path
    std::vector<whisper_state *> spare_states;   // worker states of whisper_full_parallel, kept with their device buffers
    wb::DeviceBlock static_mask;       // device bitmask of always-suppressed tokens for the current run
};
