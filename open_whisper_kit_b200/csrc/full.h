// Batched whisper_full driver and the single-window helpers behind the low-level API (full.cu).
#pragma once

#include <string>
#include <vector>

#include "state.h"

namespace wb {

struct StreamSpec {
    whisper_state * state = nullptr;
    whisper_full_params params;
    const float * samples = nullptr;
    int n_samples = 0;
    bool samples_on_device = false;
    bool samples_i16 = false;          // `samples` points at int16 PCM (s / 32768 is applied by the mel kernel's load)
    int rc = 0;
};

int64_t time_us();
std::string to_timestamp(int64_t t, bool comma);

// Runs every stream to completion as one device batch; returns the first non-zero stream status (reference codes).
int run_streams(whisper_context & ctx, std::vector<StreamSpec> & specs);

bool encode_single(whisper_context & ctx, whisper_state & st, int seek, bool keep_embd32);
bool decode_single(whisper_context & ctx, whisper_state & st, const whisper_token * tokens, int n_tokens, int n_past);
int lang_auto_detect(whisper_context & ctx, whisper_state & st, int offset_ms, float * lang_probs);

}  // namespace wb
