// Batched device engine: mel -> encoder (+ cross K/V) -> decoder rows -> logits / on-device selection.
// One Engine per whisper_context (one model replica on one GPU); calls are serialised by a (recursive) mutex.
#pragma once

#include <mutex>
#include <vector>

#include "dec_chain.h"
#include "dec_kernels.h"
#include "dtw.h"
#include "enc_kernels.h"
#include "mel.h"
#include "model.h"
#include "skinny_gemm.h"

namespace wb {

struct DeviceBlock {           // growable device allocation (contents are transient unless stated otherwise)
    void * p = nullptr;
    size_t cap = 0;
    bool reserve(size_t bytes, bool keep = false);
    ~DeviceBlock();
    DeviceBlock() = default;
    DeviceBlock(const DeviceBlock &) = delete;
    DeviceBlock & operator=(const DeviceBlock &) = delete;
};

struct Arena {                 // bump allocator over a DeviceBlock, reset per engine call
    DeviceBlock blk;
    size_t off = 0;
    bool begin(size_t total) { off = 0; return blk.reserve(total); }
    void * take(size_t bytes) {
        void * r = (char *) blk.p + off;
        off += round_up<size_t>(bytes, 256);
        return r;
    }
};

// Mel of one audio stream, device resident (reference: whisper_mel, src/whisper.cpp:414-420).
struct MelBuf {
    DeviceBlock data;          // raw log10 [n_mel][stride] (finalized == false) or final values [n_mel][n_len]
    DeviceBlock max_enc;       // 1 x unsigned
    int n_mel = 0, n_len = 0, n_len_org = 0, n_frames_fft = 0, stride = 0;
    bool finalized = false;
    bool valid = false;
};

// Cross-attention K/V of a set of windows: [n_text_layer][n_windows][n_head][K | V][T][64], 16-bit.
// T = audio context of the encoder run that filled it (1500, or whisper_full_params::audio_ctx).
// With Engine::cross_fp8 the pool holds e4m3 chunks instead (dec_kernels.cu, cross_attn_fp8_kernel): same nesting, window_bytes
// per (layer, window); layer_stride stays in 2-byte units either way.
struct CrossKV {
    DeviceBlock data;
    int n_windows = 0;
    int T = 1500;
    size_t layer_stride = 0;   // elements (2-byte units)
    size_t window_bytes = 0;   // one window in one text layer
    bool fp8 = false;
    const void * window_base(int w, int d) const {
        return (const char *) data.p + (size_t) w * (window_bytes ? window_bytes : (size_t) T * 2 * d * 2);
    }
};

struct MelJob {
    const float * pcm_host = nullptr;   // exactly one of pcm_host / pcm_dev
    const float * pcm_dev = nullptr;
    bool i16 = false;                   // the pointer holds int16 samples (converted inside the mel kernel's load)
    int n_samples = 0;
    MelBuf * out = nullptr;
};

struct EncJob {
    const MelBuf * mel = nullptr;
    int seek = 0;
};

// Optional per-kernel-class timing with CUDA events on the engine's stream (bench.py's roofline numbers).
enum ProfClass {
    PC_MEL = 0, PC_IM2COL, PC_GEMM_CONV, PC_LAYERNORM, PC_GEMM_ENC, PC_ENC_ATTN, PC_GEMM_CROSS, PC_DEC_MISC, PC_GEMM_DEC,
    PC_SELF_ATTN, PC_CROSS_ATTN, PC_GEMM_LOGITS, PC_SAMPLE, PC_LAYERNORM_DEC, PC_DEC_CHAIN, PC_COUNT
};

struct Engine {
    Model model;
    bool flash_attn = true;
    int device = 0;
    cudaStream_t stream = nullptr;
    MelPlan mel_plan;
    // Serialises the API calls on this context (one GPU = one queue).  Recursive: the progress / encoder_begin / new_segment /
    // abort callbacks of whisper_full run on the calling thread under the lock and may call the low-level entry points
    // (whisper_pcm_to_mel, whisper_encode, whisper_decode, ... on ANOTHER state); a nested whisper_full* is refused.
    std::recursive_mutex mu;
    bool in_full = false;

    Arena ws;                  // activations
    DeviceBlock meta;          // small per-call device arrays
    DeviceBlock pcm_stage;     // H2D staging for PCM
    DeviceBlock logits;        // f32 [rows][ld_logits], valid until the next decode
    SkinnyWorkspace skinny_ws; // split-K scratch of the decoder-step GEMM
    DeviceBlock embd_enc32;    // f32 [windows*1500][d] of the last encode when requested
    ChainLauncher chain;       // persistent single-token decoder-step kernel (dec_chain.cu)
    DeviceBlock chain_part;    // its stream-K partial tiles
    std::vector<TMap> chain_wmaps;   // TMA descriptors of the decoder weights, [layer][qkv, o, xq, xo, w1, w2]
    int chain_mode = -1;       // -1 undecided, 0 off (WHISPER_B200_CHAIN=0 or unsupported geometry), 1 on
    int chain_min_units = 2;
    DeviceBlock chain_trace;   // WHISPER_B200_CHAIN_TRACE=1: per-phase device timestamps of the chain launches
    std::vector<double> chain_trace_acc;
    long long chain_trace_steps = 0;
    int ld_logits = 0;
    void * h_pinned[3] = {nullptr, nullptr, nullptr};   // pinned host staging: [0] decoder rows (H2D), [1] sampler I/O, [2] KV copy lists
    size_t h_pinned_cap[3] = {0, 0, 0};
    DeviceBlock kv_copy_list;
    long long n_kernel_launches = 0;   // launches of our own kernels (bench.py reports them)

    // profiler
    struct ProfRec { int cls; cudaEvent_t a, b; double work; };
    bool prof_on = false;
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> prof_pool;
    double prof_ms[PC_COUNT] = {}, prof_work[PC_COUNT] = {};
    long long prof_n[PC_COUNT] = {};
    void prof_begin(int cls, double work);
    void prof_end();
    void prof_collect();     // synchronises, folds the pending records into prof_ms / prof_work / prof_n
    void prof_reset();

    // zero keys the reference's flash-attention path attends to: its K/V scratch is padded to a multiple of 256 positions
    // (src/whisper.cpp:2055, 2481)
    int n_phantom(int T = 1500) const { return flash_attn ? (T + 255) / 256 * 256 - T : 0; }
    int cross_T = 1500;        // audio context of the cross K/V the next decode() call reads (set by the caller)
    // WHISPER_B200_CROSS_KV=fp8 (f16 models, no DTW): cross K/V stored as e4m3 chunks -- half the bytes of the stream that bounds
    // the decoder step, at reduced precision (NOT the reference's arithmetic; off by default)
    bool cross_fp8 = false;
    DeviceBlock kv16_tmp;      // 16-bit K/V of one text layer of an encoder chunk, the quantiser's input

    // DTW token timestamps: when `on`, the next decode() also writes the cross-attention probabilities of the alignment heads
    // (heads_by_layer, capture order = layer, then list order) for every row to `probs` [n_heads_total][rows][cross_T]
    // set for the duration of a whisper_full* call that asks for token-level timestamps: decode() keeps the separate LayerNorm
    // kernels (the reference's rounding points) instead of the algebraic fold
    bool exact_ln = false;
    struct AlignCapture {
        bool on = false;
        std::vector<std::vector<int>> heads_by_layer;
        int n_heads_total = 0;
        DeviceBlock d_heads;       // the head lists, flattened in capture order
        DeviceBlock probs;
    } align;

    bool init(int device, bool flash_attn);
    ~Engine();

    void * pinned(int which, size_t bytes);

    bool run_mel(const std::vector<MelJob> & jobs);
    bool set_mel(MelBuf & out, const float * data, int n_len, int n_mel);
    // copies the reference's final [n_mel][n_len] container to the host (parity hook)
    bool get_mel(const MelBuf & mel, float * out);

    // Encodes jobs.size() windows; K/V rows for job i land at window index win0 + i of kv (kv must be sized already).
    bool encode(const std::vector<EncJob> & jobs, CrossKV & kv, int win0, bool keep_embd32);
    bool size_cross(CrossKV & kv, int n_windows, int T = 1500);

    // Runs the decoder over `rows`; logits are produced for rows[logit_rows[i]] into logits row i.
    // cross_layer_stride: element distance between text layers in the cross-K/V pool the rows point into.
    bool decode(const std::vector<DecRow> & rows, const std::vector<int> & logit_rows, size_t cross_layer_stride);
    // single-token step (one row per sequence, R <= 128) through the chain kernel; same outputs as decode()
    bool decode_chain(const std::vector<DecRow> & rows, const std::vector<int> & logit_rows, size_t cross_layer_stride);
    bool chain_usable(int R);
    bool fetch_logits(int row, float * out);                         // D2H one row (n_vocab floats)
    bool fetch_logits_rows(int row0, int n_rows, float * out);       // D2H rows [row0, row0 + n_rows) packed [n_rows][n_vocab]
    // logit rules + arg-max / categorical draws on the rows of the last decode(); uniforms: one per requested draw, indexed by
    // SampleRow::draw_off (empty when every row is an arg-max row)
    bool sample(const std::vector<SampleRow> & srows, const std::vector<double> & uniforms, const uint32_t * d_mask,
                const SampleParams & prm, std::vector<SampleOut> & out, std::vector<DrawOut> & draws);
    bool token_prob(const std::vector<SampleRow> & srows, int token, std::vector<float> & out);
    bool kv_copy_prefix(const void * src, void * dst, int n_pos);    // self-KV: positions [0, n_pos) of every layer
    struct KvCopy { const void * src; void * dst; int n_pos; };
    bool kv_copy_prefix_batch(const std::vector<KvCopy> & copies);   // the same for many (src, dst) pairs in ONE launch

    size_t self_kv_bytes() const {
        return (size_t) model.hp.n_text_layer * model.hp.n_text_ctx * 2 * model.hp.n_text_state * 2;
    }
};

}  // namespace wb
