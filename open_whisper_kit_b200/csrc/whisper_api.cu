// extern "C" implementation of include/whisper.h on the B200 engine: every WHISPER_API symbol of the reference is
// exported with the reference's argument meaning, ownership rules and error values (reference src/whisper.cpp:
// 3547-4339 init/free/accessors, 5912-6034 defaults, 7778-8098 full/result accessors, 9000-9038 logging).
// No exception crosses this boundary.
#include <float.h>
#include <string.h>

#include <algorithm>
#include <fstream>
#include <regex>
#include <thread>

#include "full.h"
#include "vad_api.h"
#include "whisper_b200.h"

using namespace wb;

namespace {

DType env_dtype() {
    const char * e = getenv("WHISPER_B200_DTYPE");
    if (e && (strcmp(e, "bf16") == 0 || strcmp(e, "BF16") == 0)) return DType::BF16;
    return DType::F16;
}

whisper_context * init_with_loader(whisper_model_loader * loader, whisper_context_params params, bool with_state) {
    if (!loader || !loader->read) return nullptr;
    if (!params.use_gpu) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: use_gpu=false requested, but this library is the CUDA (sm_100a) path and has no CPU "
             "fallback\n", __func__);
        if (loader->close) loader->close(loader->context);
        return nullptr;
    }
    if (params.flash_attn && params.dtw_token_timestamps) {       // as the reference (src/whisper.cpp:3708-3711)
        wlog(GGML_LOG_LEVEL_WARN, "%s: dtw_token_timestamps is not supported with flash_attn - disabling\n", __func__);
        params.dtw_token_timestamps = false;
    }
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0 || params.gpu_device < 0 || params.gpu_device >= n_dev) {
        cudaGetLastError();
        wlog(GGML_LOG_LEVEL_ERROR, "%s: CUDA device %d is not available (%d devices visible); no CPU fallback exists\n", __func__,
             params.gpu_device, n_dev);
        if (loader->close) loader->close(loader->context);
        return nullptr;
    }
    whisper_context * ctx = nullptr;
    try {
        ctx = new whisper_context();
        ctx->params = params;
        const int64_t t0 = time_us();
        ctx->t_start_us = t0;
        const bool ok = model_load(loader, ctx->eng.model, env_dtype(), params.gpu_device);
        if (loader->close) loader->close(loader->context);
        if (!ok || !ctx->eng.init(params.gpu_device, params.flash_attn)) {
            wlog(GGML_LOG_LEVEL_ERROR, "%s: failed to load model\n", __func__);
            delete ctx;
            return nullptr;
        }
        ctx->t_load_us = time_us() - t0;
        if (params.dtw_token_timestamps && ctx->eng.cross_fp8) {
            wlog(GGML_LOG_LEVEL_WARN, "%s: dtw_token_timestamps reads the 16-bit cross K/V - WHISPER_B200_CROSS_KV=fp8 ignored for this context\n", __func__);
            ctx->eng.cross_fp8 = false;
        }
        if (params.dtw_token_timestamps) {
            // alignment heads (src/whisper.cpp:1160-1272): an invalid selection fails the initialisation, as aheads_masks_init does
            auto & al = ctx->eng.align;
            const auto & hp = ctx->eng.model.hp;
            if (!dtw_alignment_heads(params, hp.n_text_layer, hp.n_text_head, al.heads_by_layer)) {
                delete ctx;
                return nullptr;
            }
            std::vector<int> flat;
            for (const auto & l : al.heads_by_layer) flat.insert(flat.end(), l.begin(), l.end());
            al.n_heads_total = (int) flat.size();
            if (flat.empty() || !al.d_heads.reserve(flat.size() * sizeof(int)) ||
                cudaMemcpy(al.d_heads.p, flat.data(), flat.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) {
                wlog(GGML_LOG_LEVEL_ERROR, "%s: no usable alignment heads for dtw_token_timestamps\n", __func__);
                delete ctx;
                return nullptr;
            }
        }
        if (with_state) {
            ctx->state = whisper_init_state(ctx);
            if (!ctx->state) {
                delete ctx;
                return nullptr;
            }
        }
    } catch (const std::exception & ex) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: %s\n", __func__, ex.what());
        delete ctx;
        return nullptr;
    }
    return ctx;
}

struct FileCtx {
    std::ifstream fin;
};
struct BufCtx {
    const uint8_t * p;
    size_t size, off;
};

whisper_context * init_from_file(const char * path, whisper_context_params params, bool with_state) {
    if (!path) return nullptr;
    wlog(GGML_LOG_LEVEL_INFO, "%s: loading model from '%s'\n", __func__, path);
    FileCtx * fc = new FileCtx();
    fc->fin.open(path, std::ios::binary);
    if (!fc->fin) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: failed to open '%s'\n", __func__, path);
        delete fc;
        return nullptr;
    }
    whisper_model_loader loader = {};
    loader.context = fc;
    loader.read = [](void * c, void * out, size_t n) -> size_t {
        auto * f = (FileCtx *) c;
        f->fin.read((char *) out, n);
        return (size_t) f->fin.gcount();
    };
    loader.eof = [](void * c) -> bool { return ((FileCtx *) c)->fin.eof(); };
    loader.close = [](void * c) { delete (FileCtx *) c; };
    whisper_context * ctx = init_with_loader(&loader, params, with_state);
    if (ctx) ctx->path_model = path;
    return ctx;
}

whisper_context * init_from_buffer(void * buffer, size_t size, whisper_context_params params, bool with_state) {
    if (!buffer) return nullptr;
    BufCtx * bc = new BufCtx{(const uint8_t *) buffer, size, 0};
    whisper_model_loader loader = {};
    loader.context = bc;
    loader.read = [](void * c, void * out, size_t n) -> size_t {
        auto * b = (BufCtx *) c;
        const size_t k = std::min(n, b->size - b->off);
        memcpy(out, b->p + b->off, k);
        b->off += k;
        return k;
    };
    loader.eof = [](void * c) -> bool { return ((BufCtx *) c)->off >= ((BufCtx *) c)->size; };
    loader.close = [](void * c) { delete (BufCtx *) c; };
    return init_with_loader(&loader, params, with_state);
}

// reference tokenizer: regex word split + greedy longest-match (src/whisper.cpp:3272-3320)
std::vector<int> tokenize(const Vocab & vocab, const std::string & text) {
    std::vector<std::string> words;
    {
        std::string str = text;
        const std::string pat = R"('s|'t|'re|'ve|'m|'ll|'d| ?[[:alpha:]]+| ?[[:digit:]]+| ?[^\s[:alpha:][:digit:]]+|\s+(?!\S)|\s+)";
        std::regex re(pat);
        std::smatch m;
        while (std::regex_search(str, m, re)) {
            for (auto x : m) words.push_back(x);
            str = m.suffix();
        }
    }
    std::vector<int> tokens;
    for (const auto & word : words) {
        if (word.empty()) continue;
        int i = 0;
        const int n = (int) word.size();
        while (i < n) {
            int j = n;
            bool found = false;
            while (j > i) {
                auto it = vocab.token_to_id.find(word.substr(i, j - i));
                if (it != vocab.token_to_id.end()) {
                    tokens.push_back(it->second);
                    i = j;
                    found = true;
                    break;
                }
                --j;
            }
            if (!found) {
                wlog(GGML_LOG_LEVEL_ERROR, "unknown token\n");
                ++i;
            }
        }
    }
    return tokens;
}

// Worker states of the batched calls are recycled: their self-attention caches and mel buffers stay allocated, so a
// steady stream of whisper_full_parallel calls does not pay cudaMalloc / cudaFree per chunk.
whisper_state * borrow_state(whisper_context * ctx) {
    whisper_state * st = nullptr;
    {
        std::lock_guard<std::recursive_mutex> lock(ctx->eng.mu);
        if (!ctx->spare_states.empty()) {
            st = ctx->spare_states.back();
            ctx->spare_states.pop_back();
        }
    }
    if (!st) return whisper_init_state(ctx);
    st->t_sample_us = st->t_encode_us = st->t_decode_us = st->t_batchd_us = st->t_prompt_us = st->t_mel_us = 0;
    st->n_sample = st->n_encode = st->n_decode = st->n_batchd = st->n_prompt = st->n_fail_p = st->n_fail_h = 0;
    st->result_all.clear();
    st->prompt_past0.clear();
    st->prompt_past1.clear();
    st->lang_id = 0;
    st->no_speech_prob = 0.0f;
    st->cross_base = nullptr;
    for (int j = 0; j < WHISPER_MAX_DECODERS; ++j) st->decoders[j].rng = std::mt19937(j);
    return st;
}
void return_state(whisper_context * ctx, whisper_state * st) {
    std::lock_guard<std::recursive_mutex> lock(ctx->eng.mu);
    if (ctx->spare_states.size() < 256) {
        ctx->spare_states.push_back(st);
    } else {
        cudaSetDevice(ctx->eng.device);
        delete st;
    }
}

const char * kModelNames[] = {"unknown", "tiny", "base", "small", "medium", "large"};

}  // namespace

extern "C" {

void ggml_backend_load_all(void) {}

const char * whisper_version(void) { return "1.8.3-b200"; }

// ---- init / free ---------------------------------------------------------------------------------------
struct whisper_context * whisper_init_from_file_with_params(const char * path_model, struct whisper_context_params params) {
    return init_from_file(path_model, params, true);
}
struct whisper_context * whisper_init_from_buffer_with_params(void * buffer, size_t buffer_size, struct whisper_context_params params) {
    return init_from_buffer(buffer, buffer_size, params, true);
}
struct whisper_context * whisper_init_with_params(struct whisper_model_loader * loader, struct whisper_context_params params) {
    return init_with_loader(loader, params, true);
}
struct whisper_context * whisper_init_from_file_with_params_no_state(const char * path_model, struct whisper_context_params params) {
    return init_from_file(path_model, params, false);
}
struct whisper_context * whisper_init_from_buffer_with_params_no_state(void * buffer, size_t buffer_size, struct whisper_context_params params) {
    return init_from_buffer(buffer, buffer_size, params, false);
}
struct whisper_context * whisper_init_with_params_no_state(struct whisper_model_loader * loader, struct whisper_context_params params) {
    return init_with_loader(loader, params, false);
}
struct whisper_context * whisper_init_from_file(const char * path_model) {
    return init_from_file(path_model, whisper_context_default_params(), true);
}
struct whisper_context * whisper_init_from_buffer(void * buffer, size_t buffer_size) {
    return init_from_buffer(buffer, buffer_size, whisper_context_default_params(), true);
}
struct whisper_context * whisper_init(struct whisper_model_loader * loader) {
    return init_with_loader(loader, whisper_context_default_params(), true);
}
struct whisper_context * whisper_init_from_file_no_state(const char * path_model) {
    return init_from_file(path_model, whisper_context_default_params(), false);
}
struct whisper_context * whisper_init_from_buffer_no_state(void * buffer, size_t buffer_size) {
    return init_from_buffer(buffer, buffer_size, whisper_context_default_params(), false);
}
struct whisper_context * whisper_init_no_state(struct whisper_model_loader * loader) {
    return init_with_loader(loader, whisper_context_default_params(), false);
}

struct whisper_state * whisper_init_state(struct whisper_context * ctx) {
    if (!ctx) return nullptr;
    try {
        whisper_state * st = new whisper_state();
        st->ctx = ctx;
        st->device = ctx->eng.device;
        st->decoders[0].rng = std::mt19937(0);
        st->logits.reserve((size_t) ctx->eng.model.hp.n_vocab);
        return st;
    } catch (const std::exception &) {
        return nullptr;
    }
}

int whisper_ctx_init_openvino_encoder_with_state(struct whisper_context *, struct whisper_state *, const char *, const char *, const char *) { return 1; }
int whisper_ctx_init_openvino_encoder(struct whisper_context *, const char *, const char *, const char *) { return 1; }

void whisper_free_state(struct whisper_state * state) {
    if (!state) return;
    // The context may already be gone (the reference lets a caller free its states after the context): everything the state
    // needs to release its device buffers is the device index it keeps itself.  cudaFree synchronises with in-flight work.
    cudaSetDevice(state->device);
    vad_free_state_context(state);
    delete state;
}
void whisper_free(struct whisper_context * ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->eng.device);
    cudaDeviceSynchronize();
    if (ctx->state) {
        whisper_state * st = ctx->state;
        ctx->state = nullptr;
        st->ctx = nullptr;
        vad_free_state_context(st);
        delete st;
    }
    for (whisper_state * st : ctx->spare_states) {
        vad_free_state_context(st);
        delete st;
    }
    ctx->spare_states.clear();
    delete ctx;
}
void whisper_free_params(struct whisper_full_params * params) { delete params; }
void whisper_free_context_params(struct whisper_context_params * params) { delete params; }

// ---- mel / encode / decode -------------------------------------------------------------------------------
int whisper_pcm_to_mel_with_state(struct whisper_context * ctx, struct whisper_state * state, const float * samples, int n_samples, int) {
    if (!ctx || !state || !samples || n_samples <= 0) return -1;
    std::lock_guard<std::recursive_mutex> lock(ctx->eng.mu);
    cuda_clear_failure();
    const int64_t t0 = time_us();
    std::vector<MelJob> jobs(1);
    jobs[0].pcm_host = samples;
    jobs[0].n_samples = n_samples;
    jobs[0].out = &state->mel;
    const bool ok = ctx->eng.run_mel(jobs);
    state->t_mel_us += time_us() - t0;
    if (!ok) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: failed to compute mel spectrogram\n", __func__);
        return -1;
    }
    return 0;
}
int whisper_pcm_to_mel(struct whisper_context * ctx, const float * samples, int n_samples, int n_threads) {
    return ctx ? whisper_pcm_to_mel_with_state(ctx, ctx->state, samples, n_samples, n_threads) : -1;
}
int whisper_set_mel_with_state(struct whisper_context * ctx, struct whisper_state * state, const float * data, int n_len, int n_mel) {
    if (!ctx || !state) return -1;
    if (n_mel != ctx->eng.model.filt_n_mel) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: invalid number of mel bands: %d (expected %d)\n", __func__, n_mel, ctx->eng.model.filt_n_mel);
        return -1;
    }
    if (n_len < 0 || (n_len > 0 && !data)) return -1;
    std::lock_guard<std::recursive_mutex> lock(ctx->eng.mu);
    cuda_clear_failure();
    return ctx->eng.set_mel(state->mel, data, n_len, n_mel) ? 0 : -1;
}
int whisper_set_mel(struct whisper_context * ctx, const float * data, int n_len, int n_mel) {
    return ctx ? whisper_set_mel_with_state(ctx, ctx->state, data, n_len, n_mel) : -1;
}
int whisper_encode_with_state(struct whisper_context * ctx, struct whisper_state * state, int offset, int) {
    if (!ctx || !state) return -1;
    std::lock_guard<std::recursive_mutex> lock(ctx->eng.mu);
    cuda_clear_failure();
    if (!encode_single(*ctx, *state, offset, true)) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: failed to eval\n", __func__);
        return -1;
    }
    return 0;
}
int whisper_encode(struct whisper_context * ctx, int offset, int n_threads) {
    return ctx ? whisper_encode_with_state(ctx, ctx->state, offset, n_threads) : -1;
}
int whisper_decode_with_state(struct whisper_context * ctx, struct whisper_state * state, const whisper_token * tokens, int n_tokens, int n_past, int) {
    if (!ctx || !state || !tokens) return 1;
    std::lock_guard<std::recursive_mutex> lock(ctx->eng.mu);
    cuda_clear_failure();
    if (!decode_single(*ctx, *state, tokens, n_tokens, n_past)) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: failed to eval\n", __func__);
        return 1;
    }
    return 0;
}
int whisper_decode(struct whisper_context * ctx, const whisper_token * tokens, int n_tokens, int n_past, int n_threads) {
    if (!ctx) return -1;
    if (ctx->state == nullptr) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: ERROR state was not loaded.\n", __func__);
        return -1;
    }
    return whisper_decode_with_state(ctx, ctx->state, tokens, n_tokens, n_past, n_threads);
}

// ---- tokenizer / languages ---------------------------------------------------------------------------------
int whisper_tokenize(struct whisper_context * ctx, const char * text, whisper_token * tokens, int n_max_tokens) {
    if (!ctx || !text) return 0;
    std::vector<int> res;
    try {
        res = tokenize(ctx->eng.model.vocab, text);
    } catch (const std::exception &) {
        return 0;
    }
    if (n_max_tokens < (int) res.size()) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: too many resulting tokens: %d (max %d)\n", __func__, (int) res.size(), n_max_tokens);
        return -(int) res.size();
    }
    for (int i = 0; i < (int) res.size(); ++i) tokens[i] = res[i];
    return (int) res.size();
}
__attribute__((visibility("default"))) int whisper_token_count(struct whisper_context * ctx, const char * text) {
    return -whisper_tokenize(ctx, text, nullptr, 0);
}
int whisper_lang_max_id(void) { return lang_max_id(); }
int whisper_lang_id(const char * lang) { return lang_id(lang); }
const char * whisper_lang_str(int id) { return lang_str(id); }
const char * whisper_lang_str_full(int id) { return lang_str_full(id); }

int whisper_lang_auto_detect_with_state(struct whisper_context * ctx, struct whisper_state * state, int offset_ms, int, float * lang_probs) {
    if (!ctx || !state) return -1;
    std::lock_guard<std::recursive_mutex> lock(ctx->eng.mu);
    cuda_clear_failure();
    return lang_auto_detect(*ctx, *state, offset_ms, lang_probs);
}
int whisper_lang_auto_detect(struct whisper_context * ctx, int offset_ms, int n_threads, float * lang_probs) {
    return ctx ? whisper_lang_auto_detect_with_state(ctx, ctx->state, offset_ms, n_threads, lang_probs) : -1;
}

// ---- accessors ----------------------------------------------------------------------------------------------
int whisper_n_len_from_state(struct whisper_state * state) { return state->mel.n_len_org; }
int whisper_n_len(struct whisper_context * ctx) { return ctx->state->mel.n_len_org; }
int whisper_n_vocab(struct whisper_context * ctx) { return ctx->eng.model.vocab.n_vocab; }
int whisper_n_text_ctx(struct whisper_context * ctx) { return ctx->eng.model.hp.n_text_ctx; }
int whisper_n_audio_ctx(struct whisper_context * ctx) { return ctx->eng.model.hp.n_audio_ctx; }
int whisper_is_multilingual(struct whisper_context * ctx) { return ctx->eng.model.vocab.is_multilingual() ? 1 : 0; }
int whisper_model_n_vocab(struct whisper_context * ctx) { return ctx->eng.model.hp.n_vocab; }
int whisper_model_n_audio_ctx(struct whisper_context * ctx) { return ctx->eng.model.hp.n_audio_ctx; }
int whisper_model_n_audio_state(struct whisper_context * ctx) { return ctx->eng.model.hp.n_audio_state; }
int whisper_model_n_audio_head(struct whisper_context * ctx) { return ctx->eng.model.hp.n_audio_head; }
int whisper_model_n_audio_layer(struct whisper_context * ctx) { return ctx->eng.model.hp.n_audio_layer; }
int whisper_model_n_text_ctx(struct whisper_context * ctx) { return ctx->eng.model.hp.n_text_ctx; }
int whisper_model_n_text_state(struct whisper_context * ctx) { return ctx->eng.model.hp.n_text_state; }
int whisper_model_n_text_head(struct whisper_context * ctx) { return ctx->eng.model.hp.n_text_head; }
int whisper_model_n_text_layer(struct whisper_context * ctx) { return ctx->eng.model.hp.n_text_layer; }
int whisper_model_n_mels(struct whisper_context * ctx) { return ctx->eng.model.hp.n_mels; }
int whisper_model_ftype(struct whisper_context * ctx) { return ctx->eng.model.hp.ftype; }
int whisper_model_type(struct whisper_context * ctx) { return ctx->eng.model.type; }
const char * whisper_model_type_readable(struct whisper_context * ctx) {
    const int t = ctx->eng.model.type;
    return kModelNames[t >= 0 && t <= 5 ? t : 0];
}
float * whisper_get_logits(struct whisper_context * ctx) { return ctx->state->logits.data(); }
float * whisper_get_logits_from_state(struct whisper_state * state) { return state->logits.data(); }
const char * whisper_token_to_str(struct whisper_context * ctx, whisper_token token) {
    const auto & v = ctx->eng.model.vocab.id_to_token;
    if (token < 0 || token >= (int) v.size()) return "";
    return v[token].c_str();
}
whisper_token whisper_token_eot(struct whisper_context * ctx) { return ctx->eng.model.vocab.token_eot; }
whisper_token whisper_token_sot(struct whisper_context * ctx) { return ctx->eng.model.vocab.token_sot; }
whisper_token whisper_token_solm(struct whisper_context * ctx) { return ctx->eng.model.vocab.token_solm; }
whisper_token whisper_token_prev(struct whisper_context * ctx) { return ctx->eng.model.vocab.token_prev; }
whisper_token whisper_token_nosp(struct whisper_context * ctx) { return ctx->eng.model.vocab.token_nosp; }
whisper_token whisper_token_not(struct whisper_context * ctx) { return ctx->eng.model.vocab.token_not; }
whisper_token whisper_token_beg(struct whisper_context * ctx) { return ctx->eng.model.vocab.token_beg; }
whisper_token whisper_token_lang(struct whisper_context * ctx, int lang_id) { return ctx->eng.model.vocab.token_sot + 1 + lang_id; }
whisper_token whisper_token_translate(struct whisper_context * ctx) { return ctx->eng.model.vocab.token_translate; }
whisper_token whisper_token_transcribe(struct whisper_context * ctx) { return ctx->eng.model.vocab.token_transcribe; }

// ---- timings --------------------------------------------------------------------------------------------------
struct whisper_timings * whisper_get_timings(struct whisper_context * ctx) {
    if (!ctx || ctx->state == nullptr) return nullptr;
    const whisper_state * s = ctx->state;
    whisper_timings * t = new whisper_timings;
    t->sample_ms = 1e-3f * s->t_sample_us / std::max(1, s->n_sample);
    t->encode_ms = 1e-3f * s->t_encode_us / std::max(1, s->n_encode);
    t->decode_ms = 1e-3f * s->t_decode_us / std::max(1, s->n_decode);
    t->batchd_ms = 1e-3f * s->t_batchd_us / std::max(1, s->n_batchd);
    t->prompt_ms = 1e-3f * s->t_prompt_us / std::max(1, s->n_prompt);
    return t;
}
void whisper_print_timings(struct whisper_context * ctx) {
    if (!ctx) return;
    const int64_t t_end = time_us();
    wlog(GGML_LOG_LEVEL_INFO, "\n");
    wlog(GGML_LOG_LEVEL_INFO, "%s:     load time = %8.2f ms\n", __func__, ctx->t_load_us / 1000.0f);
    if (ctx->state) {
        const whisper_state * s = ctx->state;
        const int n_sample = std::max(1, s->n_sample), n_encode = std::max(1, s->n_encode), n_decode = std::max(1, s->n_decode);
        const int n_batchd = std::max(1, s->n_batchd), n_prompt = std::max(1, s->n_prompt);
        wlog(GGML_LOG_LEVEL_INFO, "%s:     fallbacks = %3d p / %3d h\n", __func__, s->n_fail_p, s->n_fail_h);
        wlog(GGML_LOG_LEVEL_INFO, "%s:      mel time = %8.2f ms\n", __func__, s->t_mel_us / 1000.0f);
        wlog(GGML_LOG_LEVEL_INFO, "%s:   sample time = %8.2f ms / %5d runs ( %8.2f ms per run)\n", __func__, 1e-3f * s->t_sample_us, n_sample, 1e-3f * s->t_sample_us / n_sample);
        wlog(GGML_LOG_LEVEL_INFO, "%s:   encode time = %8.2f ms / %5d runs ( %8.2f ms per run)\n", __func__, 1e-3f * s->t_encode_us, n_encode, 1e-3f * s->t_encode_us / n_encode);
        wlog(GGML_LOG_LEVEL_INFO, "%s:   decode time = %8.2f ms / %5d runs ( %8.2f ms per run)\n", __func__, 1e-3f * s->t_decode_us, n_decode, 1e-3f * s->t_decode_us / n_decode);
        wlog(GGML_LOG_LEVEL_INFO, "%s:   batchd time = %8.2f ms / %5d runs ( %8.2f ms per run)\n", __func__, 1e-3f * s->t_batchd_us, n_batchd, 1e-3f * s->t_batchd_us / n_batchd);
        wlog(GGML_LOG_LEVEL_INFO, "%s:   prompt time = %8.2f ms / %5d runs ( %8.2f ms per run)\n", __func__, 1e-3f * s->t_prompt_us, n_prompt, 1e-3f * s->t_prompt_us / n_prompt);
    }
    wlog(GGML_LOG_LEVEL_INFO, "%s:    total time = %8.2f ms\n", __func__, (t_end - ctx->t_start_us) / 1000.0f);
}
void whisper_reset_timings(struct whisper_context * ctx) {
    if (!ctx) return;
    ctx->t_start_us = time_us();
    if (ctx->state) {
        whisper_state * s = ctx->state;
        s->t_mel_us = s->t_sample_us = s->t_encode_us = s->t_decode_us = s->t_batchd_us = s->t_prompt_us = 0;
        s->n_sample = s->n_encode = s->n_decode = s->n_batchd = s->n_prompt = 0;
    }
}
const char * whisper_print_system_info(void) {
    static std::string s;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    s = "WHISPER : COREML = 0 | OPENVINO = 0 | B200 : ARCH = sm_100a | TCGEN05 = 1 | TMA = 1 | CUDA_DEVICES = " + std::to_string(n) +
        " | DTYPE = " + (env_dtype() == DType::F16 ? "f16" : "bf16") + " | ";
    return s.c_str();
}

// ---- defaults --------------------------------------------------------------------------------------------------
struct whisper_context_params whisper_context_default_params(void) {
    struct whisper_context_params r;
    memset(&r, 0, sizeof(r));
    r.use_gpu = true;
    r.flash_attn = true;
    r.gpu_device = 0;
    r.dtw_token_timestamps = false;
    r.dtw_aheads_preset = WHISPER_AHEADS_NONE;
    r.dtw_n_top = -1;
    r.dtw_aheads.n_heads = 0;
    r.dtw_aheads.heads = nullptr;
    r.dtw_mem_size = 1024 * 1024 * 128;
    return r;
}
struct whisper_context_params * whisper_context_default_params_by_ref(void) {
    return new whisper_context_params(whisper_context_default_params());
}
struct whisper_vad_params whisper_vad_default_params(void) {
    whisper_vad_params r = {0.5f, 250, 100, FLT_MAX, 30, 0.1f};
    return r;
}
struct whisper_full_params whisper_full_default_params(enum whisper_sampling_strategy strategy) {
    struct whisper_full_params r;
    memset(&r, 0, sizeof(r));
    r.strategy = strategy;
    r.n_threads = std::min(4, (int) std::thread::hardware_concurrency());
    r.n_max_text_ctx = 16384;
    r.no_context = true;
    r.print_progress = true;
    r.print_timestamps = true;
    r.thold_pt = 0.01f;
    r.thold_ptsum = 0.01f;
    r.language = "en";
    r.suppress_blank = true;
    r.temperature = 0.0f;
    r.max_initial_ts = 1.0f;
    r.length_penalty = -1.0f;
    r.temperature_inc = 0.2f;
    r.entropy_thold = 2.4f;
    r.logprob_thold = -1.0f;
    r.no_speech_thold = 0.6f;
    r.greedy.best_of = -1;
    r.beam_search.beam_size = -1;
    r.beam_search.patience = -1.0f;
    r.grammar_penalty = 100.0f;
    r.vad_params = whisper_vad_default_params();
    if (strategy == WHISPER_SAMPLING_GREEDY) r.greedy.best_of = 5;
    if (strategy == WHISPER_SAMPLING_BEAM_SEARCH) {
        r.beam_search.beam_size = 5;
        r.beam_search.patience = -1.0f;
    }
    return r;
}
struct whisper_full_params * whisper_full_default_params_by_ref(enum whisper_sampling_strategy strategy) {
    return new whisper_full_params(whisper_full_default_params(strategy));
}

// ---- whisper_full ---------------------------------------------------------------------------------------------
int whisper_full_with_state(struct whisper_context * ctx, struct whisper_state * state, struct whisper_full_params params,
                            const float * samples, int n_samples) {
    if (!ctx || !state) return -1;
    try {
        std::vector<StreamSpec> specs(1);
        specs[0].state = state;
        specs[0].params = params;
        specs[0].samples = samples;
        specs[0].n_samples = n_samples;
        return run_streams(*ctx, specs);
    } catch (const std::exception & ex) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: %s\n", __func__, ex.what());
        return -6;
    }
}
int whisper_full(struct whisper_context * ctx, struct whisper_full_params params, const float * samples, int n_samples) {
    if (!ctx || !ctx->state) return -1;
    std::vector<float> vad_samples;
    if (params.vad) {                       // src/whisper.cpp:7784-7797
        if (!vad_filter(*ctx, *ctx->state, params, samples, n_samples, vad_samples)) {
            wlog(GGML_LOG_LEVEL_ERROR, "%s: failed to compute VAD\n", __func__);
            return -1;
        }
        if (vad_samples.empty()) {
            ctx->state->result_all.clear();
            return 0;
        }
        samples = vad_samples.data();
        n_samples = (int) vad_samples.size();
    }
    return whisper_full_with_state(ctx, ctx->state, params, samples, n_samples);
}

// Chunking, timestamp fix-up and timing merge as the reference (src/whisper.cpp:7801-7929); the chunks are decoded as
// rows of ONE device batch per GPU instead of n_processors host threads.  With a group of contexts (one model replica per
// GPU, include/whisper_b200.h) the chunks are dealt out in contiguous blocks, one worker thread per GPU, no collective: the
// only exchange is this host-side gather of the segments in chunk order.
static int full_parallel_on(const std::vector<whisper_context *> & ctxs, struct whisper_full_params params, const float * samples,
                            int n_samples, int n_processors, bool i16 = false) {
    // i16: `samples` really points at int16 PCM; offsets are in samples either way
    auto at = [&](int start) {
        return i16 ? reinterpret_cast<const float *>(reinterpret_cast<const int16_t *>(samples) + start) : samples + start;
    };
    whisper_context * ctx = ctxs[0];
    const int G = (int) ctxs.size();
    const int offset_samples = (WHISPER_SAMPLE_RATE * params.offset_ms) / 1000;
    const int n_per = (n_samples - offset_samples) / n_processors;
    std::vector<whisper_state *> states(n_processors, nullptr);     // [0] is the caller's state, the rest are borrowed
    std::vector<int> owner(n_processors, 0);
    std::vector<std::vector<StreamSpec>> specs(G);
    auto give_back = [&]() {
        for (int i = 1; i < n_processors; ++i)
            if (states[i]) return_state(ctxs[owner[i]], states[i]);
    };
    for (int i = 0; i < n_processors; ++i) {
        owner[i] = whisper_b200_partition_owner(i, n_processors, G);
        states[i] = i == 0 ? ctx->state : borrow_state(ctxs[owner[i]]);
        if (!states[i]) {
            give_back();
            return -7;
        }
        StreamSpec sp;
        sp.state = states[i];
        sp.params = params;
        sp.params.print_realtime = false;
        sp.samples_i16 = i16;
        if (i == 0) {
            sp.samples = samples;
            sp.n_samples = offset_samples + n_per;
        } else {
            const int start = offset_samples + i * n_per;
            sp.params.offset_ms = 0;
            sp.params.print_progress = false;
            sp.params.new_segment_callback = nullptr;
            sp.params.new_segment_callback_user_data = nullptr;
            sp.params.progress_callback = nullptr;
            sp.params.progress_callback_user_data = nullptr;
            sp.samples = at(start);
            sp.n_samples = (i == n_processors - 1) ? n_samples - start : n_per;
        }
        specs[owner[i]].push_back(sp);
    }
    // GPU 0 runs on the calling thread (its chunk 0 fires the user's callbacks there, as the reference's main thread does)
    std::vector<std::thread> workers;
    for (int g = 1; g < G; ++g) {
        if (specs[g].empty()) continue;
        workers.emplace_back([&, g]() {
            try {
                run_streams(*ctxs[g], specs[g]);
            } catch (const std::exception & ex) {
                wlog(GGML_LOG_LEVEL_ERROR, "%s: GPU %d: %s\n", __func__, ctxs[g]->eng.device, ex.what());
                for (auto & sp : specs[g]) sp.rc = -6;
            }
        });
    }
    try {
        run_streams(*ctx, specs[0]);         // call-wide failures are written to every spec's rc
    } catch (...) {
        for (auto & th : workers) th.join();
        give_back();
        throw;
    }
    for (auto & th : workers) th.join();
    int ret = specs[0][0].rc;                // the reference reports the first chunk's status (src/whisper.cpp:7857) ...
    for (int g = 1; g < G && ret == 0; ++g)  // ... a replica that failed as a whole must not pass silently either
        for (const auto & sp : specs[g])
            if (sp.rc != 0 && ret == 0) ret = sp.rc;

    const int64_t offset_t = (int64_t) (params.offset_ms / 10.0);
    whisper_state * st0 = ctx->state;
    for (int i = 1; i < n_processors; ++i) {
        whisper_state * sti = states[i];
        const int64_t shift = 100 * ((int64_t) i * n_per) / WHISPER_SAMPLE_RATE + offset_t;
        for (auto & result : sti->result_all) {
            result.t0 += shift;
            result.t1 += shift;
            if (!st0->result_all.empty()) result.t0 = std::max(result.t0, st0->result_all.back().t1);
            st0->result_all.push_back(std::move(result));
            if (params.new_segment_callback) params.new_segment_callback(ctx, st0, 1, params.new_segment_callback_user_data);
        }
        st0->t_mel_us += sti->t_mel_us;
        st0->t_sample_us += sti->t_sample_us;
        st0->t_encode_us += sti->t_encode_us;
        st0->t_decode_us += sti->t_decode_us;
        st0->t_batchd_us += sti->t_batchd_us;
        st0->t_prompt_us += sti->t_prompt_us;
        st0->n_sample += sti->n_sample;
        st0->n_encode += sti->n_encode;
        st0->n_decode += sti->n_decode;
        st0->n_batchd += sti->n_batchd;
        st0->n_prompt += sti->n_prompt;
    }
    give_back();
    st0->t_mel_us /= n_processors;
    st0->t_sample_us /= n_processors;
    st0->t_encode_us /= n_processors;
    st0->t_decode_us /= n_processors;
    wlog(GGML_LOG_LEVEL_WARN, "\n");
    wlog(GGML_LOG_LEVEL_WARN, "%s: the audio has been split into %d chunks at the following times:\n", __func__, n_processors);
    for (int i = 1; i < n_processors; ++i)
        wlog(GGML_LOG_LEVEL_WARN, "%s: split %d - %s\n", __func__, i,
             to_timestamp(100 * ((int64_t) i * n_per) / WHISPER_SAMPLE_RATE + offset_t, false).c_str());
    wlog(GGML_LOG_LEVEL_WARN, "%s: the transcription quality may be degraded near these boundaries\n", __func__);
    return ret;
}

int whisper_full_parallel(struct whisper_context * ctx, struct whisper_full_params params, const float * samples, int n_samples,
                          int n_processors) {
    if (!ctx || !ctx->state) return -1;
    if (n_processors <= 1) return whisper_full(ctx, params, samples, n_samples);
    std::vector<float> vad_samples;
    if (params.vad) {                       // src/whisper.cpp:7812-7824
        if (!vad_filter(*ctx, *ctx->state, params, samples, n_samples, vad_samples)) {
            wlog(GGML_LOG_LEVEL_ERROR, "%s: failed to compute VAD\n", __func__);
            return -1;
        }
        if (vad_samples.empty()) return 0;
        samples = vad_samples.data();
        n_samples = (int) vad_samples.size();
    }
    try {
        return full_parallel_on({ctx}, params, samples, n_samples, n_processors);
    } catch (const std::exception & ex) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: %s\n", __func__, ex.what());
        return -6;
    }
}

// ---- context group: one replica per GPU (include/whisper_b200.h) ------------------------------------------------------------
struct whisper_b200_group {
    std::vector<whisper_context *> ctxs;
};

WB200_API int whisper_b200_partition_owner(int chunk, int n_chunks, int n_gpus) {
    if (n_gpus <= 1 || n_chunks <= 0 || chunk < 0) return 0;
    // contiguous blocks whose sizes differ by at most one: chunk i -> floor(i * G / n)
    return (int) std::min<long long>((long long) chunk * n_gpus / n_chunks, n_gpus - 1);
}

WB200_API struct whisper_b200_group * whisper_b200_group_init_from_file(const char * path_model, struct whisper_context_params params,
                                                                        const int * devices, int n_devices) {
    if (!path_model) return nullptr;
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess || visible <= 0) {
        cudaGetLastError();
        wlog(GGML_LOG_LEVEL_ERROR, "%s: no CUDA device is visible; no CPU fallback exists\n", __func__);
        return nullptr;
    }
    std::vector<int> devs;
    if (devices && n_devices > 0) devs.assign(devices, devices + n_devices);
    else for (int i = 0; i < (n_devices > 0 ? std::min(n_devices, visible) : visible); ++i) devs.push_back(i);
    try {
        auto * g = new whisper_b200_group();
        g->ctxs.assign(devs.size(), nullptr);
        // the replicas load concurrently (file -> pinned staging -> HBM per device); only replica 0 owns a default state that
        // callers read results from, the others get theirs for symmetry with whisper_init_from_file_with_params
        std::vector<std::thread> loaders;
        for (size_t i = 0; i < devs.size(); ++i)
            loaders.emplace_back([&, i]() {
                whisper_context_params cp = params;
                cp.gpu_device = devs[i];
                g->ctxs[i] = init_from_file(path_model, cp, true);
            });
        for (auto & th : loaders) th.join();
        for (whisper_context * c : g->ctxs)
            if (!c) {
                whisper_b200_group_free(g);
                return nullptr;
            }
        return g;
    } catch (const std::exception & ex) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: %s\n", __func__, ex.what());
        return nullptr;
    }
}

WB200_API void whisper_b200_group_free(struct whisper_b200_group * g) {
    if (!g) return;
    for (whisper_context * c : g->ctxs) whisper_free(c);
    delete g;
}

WB200_API int whisper_b200_group_size(struct whisper_b200_group * g) { return g ? (int) g->ctxs.size() : 0; }

WB200_API struct whisper_context * whisper_b200_group_context(struct whisper_b200_group * g, int i) {
    return g && i >= 0 && i < (int) g->ctxs.size() ? g->ctxs[i] : nullptr;
}

WB200_API int whisper_b200_group_full_parallel(struct whisper_b200_group * g, struct whisper_full_params params, const float * samples,
                                               int n_samples, int n_processors) {
    if (!g || g->ctxs.empty() || !g->ctxs[0]->state) return -1;
    if (n_processors <= 1) return whisper_full(g->ctxs[0], params, samples, n_samples);
    std::vector<float> vad_samples;
    if (params.vad) {                       // the filter runs once, on the first GPU; results land in its state as always
        if (!vad_filter(*g->ctxs[0], *g->ctxs[0]->state, params, samples, n_samples, vad_samples)) {
            wlog(GGML_LOG_LEVEL_ERROR, "%s: failed to compute VAD\n", __func__);
            return -1;
        }
        if (vad_samples.empty()) return 0;
        samples = vad_samples.data();
        n_samples = (int) vad_samples.size();
    }
    try {
        return full_parallel_on(g->ctxs, params, samples, n_samples, n_processors);
    } catch (const std::exception & ex) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: %s\n", __func__, ex.what());
        return -6;
    }
}

// ---- 16-bit PCM ingest: the s16 -> f32 conversion of the reference's audio front-end is fused into the
// mel kernel's load, so half the bytes cross PCIe and HBM ---------------------------------------------------------------------
WB200_API int whisper_b200_pcm16_to_mel(struct whisper_context * ctx, const int16_t * samples, int n_samples) {
    if (!ctx || !ctx->state || !samples || n_samples <= 0) return -1;
    std::lock_guard<std::recursive_mutex> lock(ctx->eng.mu);
    cuda_clear_failure();
    const int64_t t0 = time_us();
    std::vector<MelJob> jobs(1);
    jobs[0].pcm_host = reinterpret_cast<const float *>(samples);
    jobs[0].i16 = true;
    jobs[0].n_samples = n_samples;
    jobs[0].out = &ctx->state->mel;
    const bool ok = ctx->eng.run_mel(jobs);
    ctx->state->t_mel_us += time_us() - t0;
    return ok ? 0 : -1;
}

WB200_API int whisper_b200_full_parallel_i16(struct whisper_context * ctx, struct whisper_full_params params, const int16_t * samples,
                                             int n_samples, int n_processors) {
    if (!ctx || !ctx->state || !samples) return -1;
    if (params.vad) {                       // the detector takes f32 PCM: widen on the host, then the f32 entry point
        std::vector<float> f((size_t) std::max(n_samples, 0));
        for (int i = 0; i < n_samples; ++i) f[i] = samples[i] * (1.0f / 32768.0f);
        return whisper_full_parallel(ctx, params, f.data(), n_samples, n_processors);
    }
    try {
        if (n_processors <= 1) {
            std::vector<StreamSpec> specs(1);
            specs[0].state = ctx->state;
            specs[0].params = params;
            specs[0].samples = reinterpret_cast<const float *>(samples);
            specs[0].n_samples = n_samples;
            specs[0].samples_i16 = true;
            return run_streams(*ctx, specs);
        }
        return full_parallel_on({ctx}, params, reinterpret_cast<const float *>(samples), n_samples, n_processors, true);
    } catch (const std::exception & ex) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: %s\n", __func__, ex.what());
        return -6;
    }
}

// ---- results ---------------------------------------------------------------------------------------------------
int whisper_full_n_segments_from_state(struct whisper_state * state) { return (int) state->result_all.size(); }
int whisper_full_n_segments(struct whisper_context * ctx) { return (int) ctx->state->result_all.size(); }
int whisper_full_lang_id_from_state(struct whisper_state * state) { return state->lang_id; }
int whisper_full_lang_id(struct whisper_context * ctx) { return ctx->state->lang_id; }
// with the VAD pre-filter the stored times are those of the filtered audio: map them back (src/whisper.cpp:7991-8033)
int64_t whisper_full_get_segment_t0_from_state(struct whisper_state * state, int i) {
    const int64_t t0 = state->result_all[i].t0;
    if (!state->has_vad_segments || state->vad_mapping_table.empty()) return t0;
    return vad_map_time(state->vad_mapping_table, t0);
}
int64_t whisper_full_get_segment_t1_from_state(struct whisper_state * state, int i) {
    const int64_t t1 = state->result_all[i].t1;
    if (!state->has_vad_segments || state->vad_mapping_table.empty()) return t1;
    const int64_t o1 = vad_map_time(state->vad_mapping_table, t1), o0 = whisper_full_get_segment_t0_from_state(state, i);
    return o1 - o0 < 10 ? o0 + 10 : o1;          // never a zero-length segment
}
int64_t whisper_full_get_segment_t0(struct whisper_context * ctx, int i) { return whisper_full_get_segment_t0_from_state(ctx->state, i); }
int64_t whisper_full_get_segment_t1(struct whisper_context * ctx, int i) { return whisper_full_get_segment_t1_from_state(ctx->state, i); }
bool whisper_full_get_segment_speaker_turn_next_from_state(struct whisper_state * state, int i) { return state->result_all[i].speaker_turn_next; }
bool whisper_full_get_segment_speaker_turn_next(struct whisper_context * ctx, int i) { return ctx->state->result_all[i].speaker_turn_next; }
const char * whisper_full_get_segment_text_from_state(struct whisper_state * state, int i) { return state->result_all[i].text.c_str(); }
const char * whisper_full_get_segment_text(struct whisper_context * ctx, int i) { return ctx->state->result_all[i].text.c_str(); }
int whisper_full_n_tokens_from_state(struct whisper_state * state, int i) { return (int) state->result_all[i].tokens.size(); }
int whisper_full_n_tokens(struct whisper_context * ctx, int i) { return (int) ctx->state->result_all[i].tokens.size(); }
const char * whisper_full_get_token_text_from_state(struct whisper_context * ctx, struct whisper_state * state, int i, int j) {
    return whisper_token_to_str(ctx, state->result_all[i].tokens[j].id);
}
const char * whisper_full_get_token_text(struct whisper_context * ctx, int i, int j) {
    return whisper_token_to_str(ctx, ctx->state->result_all[i].tokens[j].id);
}
whisper_token whisper_full_get_token_id_from_state(struct whisper_state * state, int i, int j) { return state->result_all[i].tokens[j].id; }
whisper_token whisper_full_get_token_id(struct whisper_context * ctx, int i, int j) { return ctx->state->result_all[i].tokens[j].id; }
whisper_token_data whisper_full_get_token_data_from_state(struct whisper_state * state, int i, int j) { return state->result_all[i].tokens[j]; }
whisper_token_data whisper_full_get_token_data(struct whisper_context * ctx, int i, int j) { return ctx->state->result_all[i].tokens[j]; }
float whisper_full_get_token_p_from_state(struct whisper_state * state, int i, int j) { return state->result_all[i].tokens[j].p; }
float whisper_full_get_token_p(struct whisper_context * ctx, int i, int j) { return ctx->state->result_all[i].tokens[j].p; }
float whisper_full_get_segment_no_speech_prob_from_state(struct whisper_state * state, int i) { return state->result_all[i].no_speech_prob; }
float whisper_full_get_segment_no_speech_prob(struct whisper_context * ctx, int i) { return ctx->state->result_all[i].no_speech_prob; }

// ---- misc ---------------------------------------------------------------------------------------------------------
int whisper_bench_memcpy(int n_threads) {
    fputs(whisper_bench_memcpy_str(n_threads), stderr);
    return 0;
}
const char * whisper_bench_memcpy_str(int) { return "memcpy: ggml host micro-benchmark is not part of the B200 path\n"; }
int whisper_bench_ggml_mul_mat(int n_threads) {
    fputs(whisper_bench_ggml_mul_mat_str(n_threads), stderr);
    return 0;
}
const char * whisper_bench_ggml_mul_mat_str(int) { return "ggml_mul_mat: ggml host micro-benchmark is not part of the B200 path\n"; }
void whisper_log_set(ggml_log_callback log_callback, void * user_data) { wlog_set(log_callback, user_data); }

// ---- extension entry points (include/whisper_b200.h) -------------------------------------------------------------------
WB200_API int whisper_b200_get_mel(struct whisper_context * ctx, struct whisper_state * state, float * out, int cap, int * n_len,
                                   int * n_mel) {
    if (!ctx) return -1;
    whisper_state * st = state ? state : ctx->state;
    if (!st || !st->mel.valid) return -1;
    if (n_len) *n_len = st->mel.n_len;
    if (n_mel) *n_mel = st->mel.n_mel;
    if (!out) return 0;
    if ((long long) cap < (long long) st->mel.n_len * st->mel.n_mel) return -2;
    std::lock_guard<std::recursive_mutex> lock(ctx->eng.mu);
    return ctx->eng.get_mel(st->mel, out) ? 0 : -3;
}

WB200_API int whisper_b200_get_encoder_output(struct whisper_context * ctx, float * out, int n_floats) {
    if (!ctx || !out) return -1;
    const int want = 1500 * ctx->eng.model.hp.n_audio_state;
    if (n_floats != want || ctx->eng.embd_enc32.cap < (size_t) want * 4) return -2;
    std::lock_guard<std::recursive_mutex> lock(ctx->eng.mu);
    cudaSetDevice(ctx->eng.device);
    return cudaMemcpy(out, ctx->eng.embd_enc32.p, (size_t) want * 4, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -3;
}

WB200_API int whisper_b200_get_cross_kv(struct whisper_context * ctx, int layer, uint16_t * out, int n_elems) {
    if (!ctx || !ctx->state || !out || !ctx->state->cross.data.p) return -1;
    const int d = ctx->eng.model.hp.n_text_state;
    const int T = ctx->state->cross.T;
    const int want = T * 2 * d;
    if (n_elems != want || layer < 0 || layer >= ctx->eng.model.hp.n_text_layer) return -2;
    std::lock_guard<std::recursive_mutex> lock(ctx->eng.mu);
    cudaSetDevice(ctx->eng.device);
    const char * src = (const char *) ctx->state->cross.data.p + (size_t) layer * ctx->state->cross.layer_stride * 2;
    if (ctx->state->cross.fp8) {
        // e4m3 pool (opt-in): expand the chunks [128 keys][64] + f32 scale back to 16 bits in the same [T][K | V] order
        const size_t bytes = ctx->state->cross.window_bytes;
        std::vector<uint8_t> raw(bytes);
        if (cudaMemcpy(raw.data(), src, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) return -3;
        const int nck = (T + 127) / 128;
        const size_t chunk = 128 * 64 + 16;
        auto e4m3 = [](uint8_t b) {
            const int e = (b >> 3) & 15, m = b & 7;
            const float v = e == 0 ? ldexpf((float) m, -9) : ldexpf(1.0f + m / 8.0f, e - 7);
            return (b & 0x80) ? -v : v;
        };
        for (int h = 0; h < d / 64; ++h)
            for (int kv = 0; kv < 2; ++kv)
                for (int t = 0; t < T; ++t) {
                    const uint8_t * ck = raw.data() + ((size_t) (h * 2 + kv) * nck + t / 128) * chunk;
                    float sc;
                    memcpy(&sc, ck + 128 * 64, 4);
                    for (int j = 0; j < 64; ++j) {
                        const __half hv = __float2half(e4m3(ck[(t % 128) * 64 + j]) * sc);
                        memcpy(out + (size_t) t * 2 * d + kv * d + h * 64 + j, &hv, 2);
                    }
                }
        return 0;
    }
    // the pool keeps a window as [head][K | V][T][64]; hand it out in the reference's [T][K(d) | V(d)] order
    std::vector<uint16_t> hm((size_t) want);
    if (cudaMemcpy(hm.data(), src, (size_t) want * 2, cudaMemcpyDeviceToHost) != cudaSuccess) return -3;
    for (int h = 0; h < d / 64; ++h)
        for (int kv = 0; kv < 2; ++kv)
            for (int t = 0; t < T; ++t)
                memcpy(out + (size_t) t * 2 * d + kv * d + h * 64, hm.data() + ((size_t) (h * 2 + kv) * T + t) * 64, 128);
    return 0;
}

// Copies positions [0, n_pos) of the self-attention history of decoder 0 of `src` into decoder 0 of `dst` with the batched
// copy kernel the beam search uses when a beam changes parent (the role of whisper_kv_cache_seq_cp, src/whisper.cpp:1100-1137).
WB200_API int whisper_b200_kv_copy(struct whisper_context * ctx, struct whisper_state * src, struct whisper_state * dst, int n_pos) {
    if (!ctx || !src || !dst || n_pos < 0 || n_pos > ctx->eng.model.hp.n_text_ctx) return -1;
    std::lock_guard<std::recursive_mutex> lock(ctx->eng.mu);
    cuda_clear_failure();
    cudaSetDevice(ctx->eng.device);
    if (!src->decoders[0].kv.p || !dst->decoders[0].kv.reserve(ctx->eng.self_kv_bytes())) return -2;
    std::vector<Engine::KvCopy> copies(1);
    copies[0] = {src->decoders[0].kv.p, dst->decoders[0].kv.p, n_pos};
    if (!ctx->eng.kv_copy_prefix_batch(copies)) return -3;
    WB_CUDA(cudaStreamSynchronize(ctx->eng.stream));
    // the copy shares the source window's cross K/V too (a beam inherits its stream's audio)
    dst->cross_base = src->cross_base;
    dst->cross_layer_stride = src->cross_layer_stride;
    dst->cross_T = src->cross_T;
    return cuda_failed() ? -4 : 0;
}

WB200_API int whisper_b200_dtype(struct whisper_context * ctx) { return ctx ? (int) ctx->eng.model.dtype : -1; }

WB200_API long long whisper_b200_kernel_launches(struct whisper_context * ctx) { return ctx ? ctx->eng.n_kernel_launches : 0; }

WB200_API void whisper_b200_profile_enable(struct whisper_context * ctx, int on) {
    if (!ctx) return;
    std::lock_guard<std::recursive_mutex> lock(ctx->eng.mu);
    cudaSetDevice(ctx->eng.device);
    ctx->eng.prof_reset();
    ctx->eng.prof_on = on != 0;
}

WB200_API int whisper_b200_profile_read(struct whisper_context * ctx, double * out, int cap) {
    if (!ctx || !out || cap < 3 * PC_COUNT) return -1;
    std::lock_guard<std::recursive_mutex> lock(ctx->eng.mu);
    cudaSetDevice(ctx->eng.device);
    ctx->eng.prof_collect();
    for (int i = 0; i < PC_COUNT; ++i) {
        out[3 * i + 0] = ctx->eng.prof_ms[i];
        out[3 * i + 1] = (double) ctx->eng.prof_n[i];
        out[3 * i + 2] = ctx->eng.prof_work[i];
    }
    return PC_COUNT;
}

WB200_API int whisper_b200_full_device(struct whisper_context * ctx, struct whisper_full_params params, const float * d_samples,
                                       int n_samples, int n_processors) {
    if (!ctx || !ctx->state || !d_samples || n_processors < 1) return -1;
    try {
        const int n_per = n_samples / n_processors;
        std::vector<whisper_state *> states;
        std::vector<StreamSpec> specs(n_processors);
        for (int i = 0; i < n_processors; ++i) {
            whisper_state * st = i == 0 ? ctx->state : borrow_state(ctx);
            if (!st) {
                for (whisper_state * b : states) return_state(ctx, b);
                return -7;
            }
            if (i > 0) states.push_back(st);
            auto pc = params;
            pc.offset_ms = 0;
            pc.print_progress = false;
            pc.print_realtime = false;
            const int start = i * n_per;
            specs[i].state = st;
            specs[i].params = pc;
            specs[i].samples = d_samples + start;
            specs[i].n_samples = (i == n_processors - 1) ? n_samples - start : n_per;
            specs[i].samples_on_device = true;
        }
        const int rc = run_streams(*ctx, specs);
        whisper_state * st0 = ctx->state;
        for (int i = 0; i < n_processors - 1; ++i) {
            for (auto & r : states[i]->result_all) {
                r.t0 += 100 * ((int64_t) (i + 1) * n_per) / WHISPER_SAMPLE_RATE;
                r.t1 += 100 * ((int64_t) (i + 1) * n_per) / WHISPER_SAMPLE_RATE;
                if (!st0->result_all.empty()) r.t0 = std::max(r.t0, st0->result_all.back().t1);
                st0->result_all.push_back(std::move(r));
            }
            return_state(ctx, states[i]);
        }
        return rc;
    } catch (const std::exception & ex) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: %s\n", __func__, ex.what());
        return -6;
    }
}

}  // extern "C"

// Host-only hook (needs no device): the tokenizer on a vocabulary given as its n_vocab token strings; same return convention
// as whisper_tokenize (count, or -needed when n_max_tokens is too small).
extern "C" WB200_API int whisper_b200_tokenize(const char * const * token_texts, int n_vocab, const char * text, whisper_token * tokens,
                                               int n_max_tokens) {
    if (!token_texts || !text || n_vocab <= 0) return -1;
    try {
        wb::Vocab vocab;
        vocab.n_vocab = n_vocab;
        for (int i = 0; i < n_vocab; ++i) vocab.token_to_id[token_texts[i]] = i;
        const std::vector<int> res = tokenize(vocab, text);
        if (n_max_tokens < (int) res.size()) return -(int) res.size();
        for (size_t i = 0; i < res.size(); ++i) tokens[i] = res[i];
        return (int) res.size();
    } catch (const std::exception &) {
        return -1;
    }
}
