// Model container: hyper-parameters, mel filters, vocabulary and device-resident weights.
// Reads the reference's legacy GGML whisper file (reference src/whisper.cpp:1485-1956) without any ggml code.
#pragma once

#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "whisper.h"

namespace wb {

void wlog(ggml_log_level level, const char * fmt, ...) __attribute__((format(printf, 2, 3)));
void wlog_set(ggml_log_callback cb, void * user);

// reference: whisper_hparams, src/whisper.cpp:590-603
struct HParams {
    int32_t n_vocab = 51864, n_audio_ctx = 1500, n_audio_state = 384, n_audio_head = 6, n_audio_layer = 4;
    int32_t n_text_ctx = 448, n_text_state = 384, n_text_head = 6, n_text_layer = 4, n_mels = 80, ftype = 1;
    float eps = 1e-5f;
};

// reference: whisper_vocab, src/whisper.cpp:429-458 and the special-token arithmetic at 1624-1639
struct Vocab {
    int n_vocab = 51864;
    std::map<std::string, int> token_to_id;
    std::vector<std::string> id_to_token;
    int token_eot = 50256, token_sot = 50257, token_translate = 50357, token_transcribe = 50358, token_solm = 50359,
        token_prev = 50360, token_nosp = 50361, token_not = 50362, token_beg = 50363;
    bool is_multilingual() const { return n_vocab >= 51865; }
    int num_languages() const { return n_vocab - 51765 - (is_multilingual() ? 1 : 0); }
};

struct EncLayer {
    float *ln1_w, *ln1_b, *ln2_w, *ln2_b;
    void * wqkv;  float * bqkv;     // [3d][d]   rows: Wq | Wk | Wv ; bias bq | 0 | bv  (key has no bias)
    void * wo;    float * bo;       // [d][d]
    void * w1;    float * b1;       // [4d][d]
    void * w2;    float * b2;       // [d][4d]
};

struct DecLayer {
    float *ln1_w, *ln1_b;
    void * wqkv;  float * bqkv;     // self-attention, fused as in the encoder
    void * wo;    float * bo;
    float *lnx_w, *lnx_b;
    void * wxq;   float * bxq;      // cross-attention query
    void * wxkv;  float * bxkv;     // [2d][d]  rows: Wk | Wv ; bias 0 | bv   (runs in the encoder stage)
    void * wxo;   float * bxo;
    float *ln2_w, *ln2_b;
    void * w1;    float * b1;
    void * w2;    float * b2;
    // LayerNorm folded into the GEMM that consumes it (tc_skinny.cu): c[n] = sum_k gamma[k] W[n][k], b'[n] = b[n] + sum_k beta[k] W[n][k]
    float *qkv_c, *qkv_b;           // ln1 into wqkv
    float *xq_c, *xq_b;             // lnx into wxq
    float *m1_c, *m1_b;             // ln2 into w1
};

struct Model {
    HParams hp;
    int type = 0;                  // e_model: 1 tiny .. 5 large
    int n_loaded = 0;              // tensors read from the file (0 = header-only test model, zero weights)
    DType dtype = DType::F16;
    int device = 0;

    int filt_n_mel = 0, filt_n_fft = 0;
    std::vector<float> filters;    // host, [n_mel][201]

    Vocab vocab;

    float * e_pe = nullptr;        // [1500][d] f32
    void * conv1_w = nullptr;      // [d][conv1_kpad]  K index = k*n_mel + c  (zero padded to a multiple of 64)
    int conv1_kpad = 0;
    float * conv1_b = nullptr;
    void * conv2_w = nullptr;      // [d][3d]          K index = k*d + c
    float * conv2_b = nullptr;
    float *e_ln_w = nullptr, *e_ln_b = nullptr;
    std::vector<EncLayer> enc;

    float * d_pe = nullptr;        // [448][d] f32
    void * d_te = nullptr;         // [n_vocab][d] 16-bit: token embedding and (tied) logits matrix
    float *d_ln_w = nullptr, *d_ln_b = nullptr;
    std::vector<DecLayer> dec;

    std::vector<void *> allocs;    // every device allocation, freed by the destructor
    size_t bytes_device = 0;

    Model() = default;
    Model(const Model &) = delete;
    Model & operator=(const Model &) = delete;
    ~Model();
};

// Returns false (and logs) on any malformed input; never throws.
bool model_load(whisper_model_loader * loader, Model & m, DType dtype, int device);

const char * lang_str(int id);
const char * lang_str_full(int id);
int lang_id(const char * s);
int lang_max_id();

}  // namespace wb
