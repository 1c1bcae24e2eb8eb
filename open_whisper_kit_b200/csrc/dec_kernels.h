// Decoder-step kernels (dec_kernels.cu).
#pragma once

#include "common.cuh"

namespace wb {

// One token row of a decoder batch (device memory).  Replaces whisper_batch + the host-built KQ mask
// (reference src/whisper.cpp:472-523, 2908-2940): causality is "positions 0..pos of the row's own slot".
struct DecRow {
    int          token;
    int          pos;       // position in the text context (n_past + i)
    void *       self_kv;   // this sequence's self-attention cache   [n_text_layer][n_text_ctx][2d]  (K | V)
    const void * cross_kv;  // this window's cross K/V, layer 0       [n_head][K | V][1500][64]; layers are layer_stride apart
};

// Per-row decoder state the logit rules depend on (reference whisper_decoder, src/whisper.cpp:797-820).
struct SampleRow {
    int logits_row;   // row of the logits buffer (several decoders may share one: the fan-out after the prompt pass)
    int n_tokens;     // tokens sampled so far in this window (0 -> "initial")
    int last;         // last / penultimate sampled token ids (valid if n_tokens >= 1 / 2)
    int penult;
    int has_ts;
    int seek_delta;
    float temperature;   // > 0: logits are divided by it first (whisper_process_logits, src/whisper.cpp:6199-6203)
    int n_draws;         // 0: arg-max (whisper_sample_token, best); k > 0: k draws from the categorical distribution of the
                         // processed row, one per uniform (whisper_sample_token best = false / whisper_sample_token_topk)
    int draw_off;        // first of this row's n_draws entries in the uniforms / draws arrays
    int tid_default;     // tid reported when no timestamp token has probability mass (0 for sample_token, token_beg for topk)
};

struct SampleParams {
    int n_vocab;
    int token_eot, token_beg, token_space;
    int suppress_blank, no_timestamps;
    float max_initial_ts;
    int tid0;            // round(max_initial_ts / 0.02)
};

struct DrawOut {         // one categorical draw: token, its probability and log-probability
    int id;
    float p, plog;
};

struct SampleOut {       // the float fields of whisper_token_data (include/whisper.h)
    int id, tid;
    float p, plog, pt, ptsum;
    int runner_up;       // second-best token after all rules and its logit distance to the winner (diagnostics:
    float gap;           // lets a parity test tell a genuine divergence from a near-tie flip)
};

void dec_embed(DType dt, const void * te, const float * pe, const DecRow * d_rows, int R, int d, float * x,
               cudaStream_t st);
void dec_kv_append(const void * qkv, const DecRow * d_rows, int R, int d, size_t layer_off_elems, cudaStream_t st);
// fused_append: also store this token's K/V into the cache (only valid when every sequence has exactly one row).
// variant: -1 default (mma.sync fragments unless WHISPER_B200_SELF_MMA=0), 0 CUDA-core kernel, 1 mma.sync kernel (tests compare them).
void dec_self_attn(DType dt, const void * qkv, const DecRow * d_rows, int R, int d, int n_head, size_t layer_off_elems,
                   int n_ctx, bool fused_append, void * out, cudaStream_t st, int variant = -1);
struct SplitIn;   // dec_chain.h: the query as partial tiles of the chain kernel's cross-q GEMM (q is ignored then)
// d_groups (optional): n_groups runs {first row, count <= DEC_CROSS_GROUP_MAX} of consecutive rows that share one window's cross
// K/V (prompt tokens, beams): one CTA per (run, head) then streams the K/V once for all rows of the run.
constexpr int DEC_CROSS_GROUP_MAX = 5;
void dec_cross_attn(DType dt, const void * q, const DecRow * d_rows, int R, int d, int n_head, size_t layer_off_elems,
                    int T, int n_phantom, void * out, cudaStream_t st, const SplitIn * q_split = nullptr,
                    const int2 * d_groups = nullptr, int n_groups = 0);
// Opt-in e4m3 cross-K/V pool (WHISPER_B200_CROSS_KV=fp8, f16 models only): bytes of one window in one text layer; the 16-bit
// head-major K/V of W windows of one layer -> the pool (dst = that layer's first window of the batch); the attention over it
// (layer_off_elems in 2-byte units like everywhere else).
size_t cross_fp8_window_bytes(int n_head, int T);
void cross_fp8_quantize(const void * kv16, void * dst_layer_win0, int W, int n_head, int T, cudaStream_t st);
void dec_cross_attn_fp8(const void * q, const DecRow * d_rows, int R, int d, int n_head, size_t layer_off_elems, int T, int n_phantom,
                        void * out, cudaStream_t st, const int2 * d_groups, int n_groups);
// Logit rules + selection for R decoder rows: arg-max rows fill d_out[r] completely; rows with n_draws > 0 fill the timestamp
// statistics of d_out[r] (tid, pt, ptsum) and d_draws[draw_off .. draw_off + n_draws) from d_uniforms[draw_off ..]
// (uniforms in [0, 1) drawn by the host from each decoder's mt19937, exactly as std::discrete_distribution consumes them).
void dec_sample(const float * logits, int ld, const SampleRow * d_srows, int R, const uint32_t * d_static_mask,
                const SampleParams & prm, SampleOut * d_out, const double * d_uniforms, DrawOut * d_draws, cudaStream_t st);
void dec_token_prob(const float * logits, int ld, const SampleRow * d_srows, int R, int n_vocab, int token, float * d_out,
                    cudaStream_t st);

}  // namespace wb
