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

#include "whisper.h"

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define WB200_API __attribute__((visibility("default")))
#else
#define WB200_API
#endif

/* Number of CUDA devices visible to the library (0 when there is none: every compute entry point then fails). */
WB200_API int whisper_b200_device_count(void);

/* ---- (1) device-resident / batched path and read-back of intermediates ------------------------------------ */

/* whisper_full_parallel for PCM that already lives in device memory (d_samples is a CUDA device pointer on the
 * context's GPU): the audio is split in n_processors equal chunks (one 30 s window each when
 * n_samples == n_processors * 480000) which are decoded as ONE device batch.  Results are read with the usual
 * whisper_full_get_* accessors.  Same return codes as whisper_full (reference include/whisper.h:603-620). */
WB200_API int whisper_b200_full_device(struct whisper_context * ctx, struct whisper_full_params params,
                                       const float * d_samples, int n_samples, int n_processors);

/* The reference's mel container [n_mel][n_len] (f32, clamped + normalised) of a state (NULL = default state),
 * which the reference API never exposes (whisper_state::mel, reference src/whisper.cpp:414-420).  out == NULL
 * returns only the geometry. */
WB200_API int whisper_b200_get_mel(struct whisper_context * ctx, struct whisper_state * state, float * out, int cap,
                                   int * n_len, int * n_mel);

/* Encoder output [1500][n_audio_state] f32 of the last whisper_encode on this context
 * (whisper_state::embd_enc, reference src/whisper.cpp:2241-2251). */
WB200_API int whisper_b200_get_encoder_output(struct whisper_context * ctx, float * out, int n_floats);

/* Cross-attention K|V of one text layer after whisper_encode: [1500][2*n_text_state] 16-bit patterns
 * (kv_cross, reference src/whisper.cpp:2300-2339; K carries the dh^-0.25 scale). */
WB200_API int whisper_b200_get_cross_kv(struct whisper_context * ctx, int layer, uint16_t * out, int n_elems);

/* 16-bit operand type of the tensor path: 0 = f16 (default; bit-exact weights of F16 model files), 1 = bf16
 * (environment WHISPER_B200_DTYPE=bf16 at context creation). */
WB200_API int whisper_b200_dtype(struct whisper_context * ctx);

/* Number of kernels of this library launched on the context so far. */
WB200_API long long whisper_b200_kernel_launches(struct whisper_context * ctx);

/* Per-kernel-class timing with CUDA events on the library's stream.  enable(1) resets and starts, enable(0) stops.
 * read() fills out[3*c + {0,1,2}] = {milliseconds, launches, algorithmic work} for class c and returns the number of
 * classes, in this order: mel(bytes), im2col(bytes), gemm_conv(flop), layernorm(bytes), gemm_encoder(flop),
 * encoder_attention(flop), gemm_cross_kv(flop), decoder_misc(bytes), gemm_decoder(bytes), self_attention(bytes),
 * cross_attention(bytes), gemm_logits(bytes), sample(bytes), layernorm_decoder(bytes), decoder_chain(bytes). */
WB200_API void whisper_b200_profile_enable(struct whisper_context * ctx, int on);
WB200_API int whisper_b200_profile_read(struct whisper_context * ctx, double * out, int cap);

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

/* Weight-streaming GEMM of the decoder step (M <= 128 rows), same epilogue; replaces mul_mat_vec_f / mul_mat_f,
 * reference ggml/src/ggml-cuda/mmvf.cu:8, mmf.cuh:50. */
WB200_API int whisper_b200_kernel_skinny_gemm(int dtype, int M, int N, int K, const uint16_t * a, const uint16_t * w,
                                              const float * bias, float scale, int scale_cols, int gelu,
                                              const float * resid, uint16_t * out16, float * out32);

/* The same GEMM on tcgen05 / TMEM fed by TMA (csrc/tc_skinny.cu), the default of the decoder step. */
WB200_API int whisper_b200_kernel_tc_skinny_gemm(int dtype, int M, int N, int K, const uint16_t * a, const uint16_t * w,
                                                 const float * bias, float scale, int scale_cols, int gelu,
                                                 const float * resid, uint16_t * out16, float * out32);

/* Two decoder-step GEMMs with the LayerNorm between them folded in (csrc/tc_skinny.cu): x = a1 * w1^T + bias1 + resid, whose
 * epilogue also emits per-tile row statistics, then y = LayerNorm(x; gamma, beta, eps) * w2^T with the normalised A operand built
 * inside the second kernel.  M <= 64; d, K1 multiples of 64.  x_out [M][d], y_out [M][N2] f32.  Returns 0 or a negative error. */
WB200_API int whisper_b200_kernel_ln_gemm_pair(int dtype, int M, int d, int K1, int N2, const uint16_t * a1, const uint16_t * w1,
                                               const float * bias1, const float * resid, const float * gamma, const float * beta,
                                               float eps, const uint16_t * w2, float * x_out, float * y_out);
WB200_API double whisper_b200_kernel_gemm_bench(int dtype, int M, int N, int K, int gelu, int iters);

/* Average microseconds of one decoder-step kernel launched back to back: which = 0 LayerNorm, 1 cross-attention,
 * 2 self-attention at position aux, 3 KV append; R rows of width d. */
WB200_API double whisper_b200_kernel_step_bench(int which, int dtype, int R, int d, int aux, int iters);

/* Masked self-attention of the decoder step (reference: KQ / soft_max_ext / KQV of whisper_build_graph_decoder,
 * src/whisper.cpp:2594-2632) on explicit inputs: row r attends to positions 0..pos[r] of its own cache [n_ctx][2d] (K | V);
 * qkv [R][3d]; with fused_append the K / V of position pos[r] come from qkv and are stored into the cache.  variant: -1 the
 * default kernel, 0 the CUDA-core kernel, 1 the mma.sync-fragment kernel.  out [R][d]; cache_out (optional) receives the caches. */
WB200_API int whisper_b200_kernel_self_attn(int dtype, int R, int d, int n_ctx, const int * pos, const uint16_t * qkv, const uint16_t * cache,
                                            int fused_append, int variant, uint16_t * out, uint16_t * cache_out);

/* Beam bookkeeping hook: copy positions [0, n_pos) of the self-attention history (decoder 0) of state `src` into state `dst`
 * with the batched copy kernel the beam search uses when a beam changes parent (reference: whisper_kv_cache_seq_cp,
 * src/whisper.cpp:1100-1137); `dst` then also points at `src`'s cross K/V, so whisper_decode_with_state(dst, ..., n_past = n_pos)
 * continues `src`'s sequence. */
WB200_API int whisper_b200_kv_copy(struct whisper_context * ctx, struct whisper_state * src, struct whisper_state * dst, int n_pos);

/* ---- 16-bit PCM ingest -------------------------------------------------------------------------------------------------------
 * whisper_pcm_to_mel / whisper_full_parallel for int16 mono 16 kHz samples: x = s / 32768 -- the conversion the reference's
 * callers run on the host before the API (examples/common-whisper.cpp:42-134, miniaudio s16 -> f32) -- is applied inside the
 * mel kernel's load stage, so half the bytes cross PCIe and HBM.  Results are bit-identical to passing the converted floats. */
WB200_API int whisper_b200_pcm16_to_mel(struct whisper_context * ctx, const int16_t * samples, int n_samples);
WB200_API int whisper_b200_full_parallel_i16(struct whisper_context * ctx, struct whisper_full_params params, const int16_t * samples,
                                             int n_samples, int n_processors);

/* ---- multi-GPU: a group of contexts, one model replica per GPU ------------------------------------------------------------
 * whisper_full_parallel (reference src/whisper.cpp:7801-7929) splits the audio into n_processors chunks and decodes them
 * independently; whisper_b200_group_full_parallel does the same split, deals the chunks out to the GPUs of the group in
 * contiguous blocks (chunk i -> GPU floor(i * n_gpus / n_processors)), runs one batch per GPU on one worker thread each -- no
 * collective on the path -- and gathers the segments on the host in chunk order with the reference's timestamp fix-up
 * (7879-7889).  The result is read from context 0 with the whisper_full_get_* accessors and equals what whisper_full_parallel
 * produces on one GPU.  devices == NULL: the first n_devices visible GPUs (n_devices <= 0: all of them). */
struct whisper_b200_group;
WB200_API struct whisper_b200_group * whisper_b200_group_init_from_file(const char * path_model, struct whisper_context_params params,
                                                                        const int * devices, int n_devices);
WB200_API void whisper_b200_group_free(struct whisper_b200_group * g);
WB200_API int whisper_b200_group_size(struct whisper_b200_group * g);
WB200_API struct whisper_context * whisper_b200_group_context(struct whisper_b200_group * g, int i);
WB200_API int whisper_b200_group_full_parallel(struct whisper_b200_group * g, struct whisper_full_params params, const float * samples,
                                               int n_samples, int n_processors);
/* GPU index (0 .. n_gpus-1) that owns chunk `chunk` of n_chunks; host logic only -- needs no device. */
WB200_API int whisper_b200_partition_owner(int chunk, int n_chunks, int n_gpus);

/* The on-device logit rules + token selection kernel (csrc/dec_kernels.cu: whisper_process_logits + whisper_sample_token /
 * whisper_sample_token_topk, src/whisper.cpp:6177-6592) on explicit logits rows and decoder states.
 * logits: host [n_logit_rows][n_vocab] f32.  rows[r].logits_row picks the row; n_tokens / last / penult / has_ts / seek_delta are
 * the decoder state the rules read; temperature > 0 divides the logits; n_draws == 0 -> arg-max into out[r], else n_draws
 * categorical draws into draws[draw_off ..] from uniforms[draw_off ..] (u in [0,1), what std::discrete_distribution would take
 * from the decoder's generator), out[r] then carries tid / pt / ptsum only.  static_mask: (n_vocab + 31) / 32 words, bit = token
 * suppressed at every step (may be NULL).  Returns 0, or a negative error. */
struct whisper_b200_sample_row {
    int logits_row, n_tokens, last, penult, has_ts, seek_delta;
    float temperature;
    int n_draws, draw_off, tid_default;
};
struct whisper_b200_sample_params {
    int n_vocab, token_eot, token_beg, token_space, suppress_blank, no_timestamps;
    float max_initial_ts;
    int tid0;
};
struct whisper_b200_sample_out {
    int id, tid;
    float p, plog, pt, ptsum;
    int runner_up;
    float gap;
};
struct whisper_b200_draw_out {
    int id;
    float p, plog;
};
WB200_API int whisper_b200_kernel_sample(const float * logits, int n_logit_rows, const struct whisper_b200_sample_row * rows, int n_rows,
                                         const uint32_t * static_mask, struct whisper_b200_sample_params prm, const double * uniforms,
                                         int n_uniforms, struct whisper_b200_sample_out * out, struct whisper_b200_draw_out * draws);

/* The alignment stage of the DTW token timestamps (csrc/dtw.cu, host logic only -- needs no device): probs [n_heads][n_tokens][T]
 * cross-attention probabilities of the alignment heads -> standardise over the tokens per (head, audio position), median filter
 * of medfilt_width along the first n_audio positions, mean over heads, negate, dynamic time warping over the tokens
 * skip_front .. n_tokens - 2 (reference src/whisper.cpp:8712-8998).  first_out[k] = first audio position of token
 * skip_front + k on the path.  Returns the number of aligned tokens, or -1. */
WB200_API int whisper_b200_dtw_align(const float * probs, int n_heads, int n_tokens, int T, int n_audio, int skip_front,
                                     int medfilt_width, int * first_out);

/* Voice activity detection (reference src/whisper.cpp:4341-5496, 6643-6825, 7947-8033).  The whisper_vad_* functions of
 * whisper.h run the Silero model on the GPU (csrc/vad.cu); these hooks expose the two host stages around it for parity tests.
 *   _vad_segments_from_probs: the probability -> speech-segment state machine on explicit probabilities (host only; replaces
 *     whisper_vad_segments_from_probs at 5209-5420).  seg_out[2i], [2i+1] = start, end in centiseconds; returns the count.
 *   _vad_filter: the audio filter whisper_full applies when params.vad is set (whisper_vad, 6643-6825): filtered samples into
 *     out[0..cap), the (processed, original) time pairs into table; returns the filtered length, -1/-2 on failure.
 *   _vad_map_time: processed -> original time through such a table (map_processed_to_original_time, 7947-7989; host only). */
WB200_API int whisper_b200_vad_segments_from_probs(const float * probs, int n_probs, int n_window, struct whisper_vad_params params,
                                                   long long * seg_out, int cap);
WB200_API int whisper_b200_vad_filter(struct whisper_context * ctx, struct whisper_full_params params, const float * samples,
                                      int n_samples, float * out, int cap, long long * table, int cap_pairs, int * n_pairs);
WB200_API long long whisper_b200_vad_map_time(const long long * table, int n_pairs, long long t);

/* Stream-K geometry of one GEMM phase of the persistent decoder-step kernel (csrc/dec_chain.h, host logic only -- needs
 * no device): out[0..4] = {tiles, k-blocks per tile, units, CTAs taking part, partial-tile slots per output tile}.
 * direct != 0: every CTA owns one whole 128-column tile.  Returns 0, or -1 for shapes the kernel does not take. */
WB200_API int whisper_b200_chain_geometry(int grid, int rows, int N, int K, int min_units, int direct, int * out);

/* The model loader's expansion of quantised GGML blocks to f16 (csrc/model.cu, host logic only -- needs no device).
 * ggml_type: 2 q4_0, 3 q4_1, 6 q5_0, 7 q5_1, 8 q8_0 (32 elements per block), 10..14 q2_K..q6_K (256 per block); restates
 * dequantize_row_* of the reference (ggml/src/ggml-quants.c:307-415, 784-1791).  out16 receives n_blocks * 32 (or 256) IEEE
 * half bit patterns.  Returns the number of elements written, or -1 for a type the loader does not take. */
WB200_API long long whisper_b200_dequantize_blocks(int ggml_type, const void * raw, long long n_blocks, uint16_t * out16);

/* The token-level timestamp heuristic and the max_len re-wrapping (csrc/full.cu, host logic only -- needs no device) applied
 * to ONE segment whose inputs are all given explicitly; restates whisper_exp_compute_token_level_timestamps and
 * whisper_wrap_segment (src/whisper.cpp:8455-8660, 6077-6130).  token_texts: the n_vocab token strings; tok_state =
 * {t_beg, t_last, tid_last} in / out; tokens are updated in place; segment k after wrapping is seg_t[2k] .. seg_t[2k+1] with
 * seg_ntok[k] tokens.  Returns the number of segments, or -1. */
WB200_API int whisper_b200_token_timestamps(const char * const * token_texts, int n_vocab, int token_eot, int token_beg,
                                            const float * pcm, int n_samples, long long seg_t0, long long seg_t1,
                                            struct whisper_token_data * tokens, int n_tokens, float thold_pt, float thold_ptsum,
                                            long long * tok_state, int max_len, int split_on_word, long long * seg_t, int * seg_ntok,
                                            int seg_cap);

/* The host sampling path's restatement of whisper_process_logits followed by the greedy whisper_sample_token
 * (csrc/full.cu <- src/whisper.cpp:6177-6517; host logic only -- needs no device) on ONE explicit logits row and decoder
 * state.  token_texts: the n_vocab token strings; special = {eot, sot, translate, transcribe, solm, prev, nosp, not, beg};
 * hist: the tokens sampled so far.  topk_out (optional): topk_k draws of the restated whisper_sample_token_topk (6519-6592)
 * with the decoder's mt19937 seeded with topk_seed.  Any of the outputs may be NULL.  Returns 0, or -1. */
WB200_API int whisper_b200_process_logits(const char * const * token_texts, int n_vocab, const int * special, int n_audio_ctx,
                                          struct whisper_full_params params, float temperature, const float * logits_row,
                                          const whisper_token * hist, int n_hist, int has_ts, int seek_delta, float * logits_out,
                                          float * logprobs_out, float * probs_out, struct whisper_token_data * tok_out, int topk_k,
                                          unsigned topk_seed, struct whisper_token_data * topk_out);

/* whisper_sequence_score restated (src/whisper.cpp:6595-6641; host logic only): out = {sum_logprobs, avg_logprobs, entropy,
 * score} of the first result_len of n tokens given by ids and log-probabilities. */
WB200_API int whisper_b200_sequence_score(struct whisper_full_params params, const float * plog, const whisper_token * ids, int n,
                                          int result_len, double * out);

/* The tokenizer of whisper_tokenize (csrc/whisper_api.cu <- src/whisper.cpp:3272-3320; host logic only -- needs no device) on
 * a vocabulary given as its n_vocab token strings.  Returns the token count, or -needed if n_max_tokens is too small. */
WB200_API int whisper_b200_tokenize(const char * const * token_texts, int n_vocab, const char * text, whisper_token * tokens,
                                    int n_max_tokens);

#ifdef __cplusplus
}
#endif

#endif /* WHISPER_B200_H */
