// VAD pre-filter of whisper_full and the time map of its results (vad_api.cu).
#pragma once

#include <vector>

#include "state.h"

namespace wb {

bool vad_filter(whisper_context & ctx, whisper_state & state, const whisper_full_params & params, const float * samples, int n_samples,
                std::vector<float> & filtered);
int64_t vad_map_time(const std::vector<whisper_state::vad_time_mapping> & table, int64_t processed_time);
void vad_free_state_context(whisper_state * st);

}  // namespace wb
