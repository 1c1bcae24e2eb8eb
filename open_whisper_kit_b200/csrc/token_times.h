// Token-level timestamps and max_len segment splitting of the finished segments (token_times.cu), host side.
#pragma once

#include <vector>

#include "state.h"

namespace wb {

// |PCM| envelope the timestamps snap to (reference get_signal_energy, src/whisper.cpp:8425-8442)
void envelope_abs_mean(const float * pcm, int n_samples, int half_width, std::vector<float> & env);
// fills t0 / t1 / vlen of the tokens of state.result_all[i_segment] (reference src/whisper.cpp:8455-8660)
void assign_token_times(const Vocab & vocab, whisper_state & state, int i_segment, float thold_pt, float thold_ptsum);
// replaces the last segment by pieces of at most max_len characters; returns their number (reference src/whisper.cpp:6077-6130)
int split_last_segment(const Vocab & vocab, whisper_state & state, int max_len, bool split_on_word);

}  // namespace wb
