// DTW token timestamps (dtw.cu): alignment-head selection, the device capture of their cross-attention probabilities, and the
// host alignment.  Reference: src/whisper.cpp:384-410, 2721-2737, 8683-8998.
#pragma once

#include <vector>

#include "dec_kernels.h"
#include "model.h"

namespace wb {

// heads per text layer in capture order; false (with a log line) when the context parameters do not name valid heads
bool dtw_alignment_heads(const whisper_context_params & cp, int n_text_layer, int n_head, std::vector<std::vector<int>> & by_layer);

// softmax(q k^T) of `n_heads` heads of one layer for all R rows -> out[(a0 + i) * R + row][T]  (d_heads: device array)
void dtw_capture_layer(DType dt, const void * q, const DecRow * d_rows, int R, int d, const int * d_heads, int n_heads, size_t layer_off,
                       int T, int a0, float * out, cudaStream_t st);

// probs [n_heads][n_tokens][T]; returns per token skip_front .. n_tokens - 2 the first audio position of its run on the path
std::vector<int> dtw_align(const float * probs, int n_heads, int n_tokens, int T, int n_audio, int skip_front, int medfilt_width);

}  // namespace wb
