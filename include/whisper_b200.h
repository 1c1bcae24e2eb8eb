/* whisper_b200.h -- extension entry points of the B200-native library, next to the reference C API.
 *
 * include/whisper.h is the drop-in boundary (ABI-identical to the reference's include/whisper.h); nothing in
 * this header is needed by a caller of the reference.  It adds:
 *   (1) device-resident / batched variants of the hot path, so a caller that already holds PCM in HBM
 *       (or wants many 30 s windows decoded in one call) does not pay the host round trip;
 *   (2) single-kernel hooks used by the parity tests and by bench.py's roofline measurements.
 * Plain pointers and sizes only; no C++ or torch types.
 */
#ifndef WHISPER_B200_H
#define WHISPER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define WB200_API __attribute__((visibility("default")))
#else
#define WB200_API
#endif

struct whisper_context;
struct whisper_state;

/* Number of CUDA devices visible to the library (0 when there is none: every compute entry point then fails). */
WB200_API int whisper_b200_device_count(void);

/* ---- (2) kernel hooks ------------------------------------------------------------------------------------- */

/* Fused log-mel kernel on one PCM buffer (host pointers).  Output is the reference's final container
 * [n_mel][n_len] f32 (clamped and normalised) -- replaces log_mel_spectrogram, reference src/whisper.cpp:3170-3260.
 * With mel_out == NULL only the geometry is returned.  Returns 0 on success. */
WB200_API int whisper_b200_kernel_log_mel(const float * pcm, int n_samples, const float * filters /*[n_mel][201]*/,
                                          int n_mel, float * mel_out, int mel_cap, int * n_len, int * n_len_org);

/* Average milliseconds of one launch of the log-mel kernel over n_streams device-resident streams. */
WB200_API double whisper_b200_kernel_log_mel_bench(int n_streams, int n_samples, const float * filters, int n_mel,
                                                   int iters, int flush_l2);

/* tcgen05 GEMM with the fused epilogue on host buffers: out = epi(A[M,K] * W[N,K]^T).
 * dtype 0 = f16, 1 = bf16 (bit patterns in uint16_t).  Replaces ggml_mul_mat + bias/scale/GELU/residual adds,
 * reference src/whisper.cpp:2112-2237.  Any of bias/pos/resid/out16/out32 may be NULL. */
WB200_API int whisper_b200_kernel_gemm(int dtype, int M, int N, int K, const uint16_t * a, const uint16_t * w,
                                       const float * bias, float scale, int scale_cols, int gelu, const float * pos,
                                       int pos_rows, const float * resid, uint16_t * out16, float * out32);

WB200_API double whisper_b200_kernel_gemm_bench(int dtype, int M, int N, int K, int gelu, int iters);

#ifdef __cplusplus
}
#endif

#endif /* WHISPER_B200_H */
