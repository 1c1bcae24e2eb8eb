// Encoder-side kernels around the GEMMs (enc_kernels.cu).
#pragma once

#include "common.cuh"

namespace wb {

// One 30 s encoder window of a batched launch (device memory).
struct EncWindow {
    const float *    mel;           // stream's mel, [n_mel][stride]: raw log10 (finalized == 0) or final values
    int              stride;
    int              n_len;         // frames the reference container holds; beyond it the window is zero
    int              n_frames_fft;  // frames stored in the raw buffer; [n_frames_fft, n_len) is the constant -10
    int              seek;          // first mel frame of the window (mel_offset, src/whisper.cpp:2381-2403)
    const unsigned * max_enc;       // ordered-uint max of the raw mel (mel.cu); unused when finalized
    int              finalized;
};

void im2col1(DType dt, const EncWindow * d_wins, int n_windows, int n_mel, int k_pad, int T /* audio context */,
             void * out /*[W*2T][k_pad]*/, cudaStream_t st);
void im2col2(const void * act1 /*[W*2T][d]*/, int n_windows, int d, int T, void * out /*[W*T][3d]*/, cudaStream_t st);

// y = LayerNorm(x[row_map ? row_map[m] : m]) for m < M; 16-bit and/or f32 outputs.
void layernorm(DType dt, const float * x, int ldx, const float * gamma, const float * beta, float eps, int M, int d,
               void * y16, int ldy16, float * y32, int ldy32, const int * row_map, cudaStream_t st);

// Non-causal self-attention over T positions per window, dh = 64.  qkv [W*T][3d] -> out [W*T][d].
void enc_attention(DType dt, const void * qkv, void * out, int n_windows, int T, int d, int n_head, int n_phantom,
                   cudaStream_t st);

// The same attention on tcgen05 / TMEM (enc_attn_tc.cu); vt_scratch holds V^T, enc_attention_tc_scratch_bytes() bytes.
size_t enc_attention_tc_scratch_bytes(int n_windows, int T, int n_head);
// vt_ready: rows 0..63 of the V^T blocks were written by the QKV GEMM's epilogue (GemmArgs::vt) and rows 64..79 by
// enc_attention_tc_init_vt(); otherwise a transpose kernel fills the scratch from qkv first.
void enc_attention_tc_init_vt(DType dt, void * vt_scratch, int n_windows, int T, int n_head, cudaStream_t st);
bool enc_attention_tc(DType dt, const void * qkv, void * out, void * vt_scratch, int n_windows, int T, int d, int n_head,
                      int n_phantom, cudaStream_t st, bool vt_ready = false);

}  // namespace wb
