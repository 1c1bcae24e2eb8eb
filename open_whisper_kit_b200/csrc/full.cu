// whisper_full semantics on the batched engine.
//
// The reference runs ONE audio stream per call and walks it window by window (whisper_full_with_state,
// src/whisper.cpp:6827-7776); whisper_full_parallel spawns one thread + state per chunk (7801-7929).  Here a set of
// streams (one per chunk) advances in lock-step: every round encodes the current 30 s window of every live stream as
// one device batch and decodes all their sequences as rows of one decoder batch.  Per stream the control flow --
// seek loop, temperature ladder, prompt construction, per-token state machine, fallback decision, segment emission --
// is the reference's, so each stream produces exactly what a stand-alone whisper_full call would.
//
// Two selection paths per stream and temperature round:
//   device path : greedy at temperature 0 without a logits callback -> rules + log-softmax + arg-max run in
//                 dec_kernels.cu right after the decode step; 24 bytes per sequence come back to the host.
//   host path   : temperature > 0 (mt19937 + discrete_distribution draws), "beam search" (which in the reference is k
//                 draws per beam, src/whisper.cpp:6519-6592) or a user logits callback -> the logits row is copied to
//                 the host and the restated reference rules/samplers below run there.
#include "full.h"
#include "token_times.h"
#include "whisper_b200.h"

#include <math.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <map>
#include <regex>
#include <thread>

namespace wb {

int64_t time_us() {
    return std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static const char * const kNonSpeech[] = {
    "\"", "#", "(", ")", "*", "+", "/", ":", ";", "<", "=", ">", "@", "[", "\\", "]", "^", "_", "`", "{", "|", "}", "~",
    "「", "」", "『", "』", "<<", ">>", "<<<", ">>>", "--", "---", "-(", "-[", "('", "(\"", "((", "))", "(((",
    ")))", "[[", "]]", "{{", "}}", "♪♪", "♪♪♪", "♩", "♪", "♫", "♬", "♭",
    "♮", "♯"};

std::string to_timestamp(int64_t t, bool comma) {
    int64_t msec = t * 10;
    const int64_t hr = msec / (1000 * 60 * 60);
    msec -= hr * (1000 * 60 * 60);
    const int64_t mn = msec / (1000 * 60);
    msec -= mn * (1000 * 60);
    const int64_t sec = msec / 1000;
    msec -= sec * 1000;
    char buf[32];
    snprintf(buf, sizeof(buf), "%02d:%02d:%02d%s%03d", (int) hr, (int) mn, (int) sec, comma ? "," : ".", (int) msec);
    return buf;
}

// ---- token ids that are suppressed at every step (everything in whisper_process_logits that does not depend on the
// decoder state: src/whisper.cpp:6224-6292) ------------------------------------------------------------------------
static void build_static_suppress(const Vocab & v, const whisper_full_params & p, std::vector<uint32_t> & bits) {
    const int V = v.n_vocab;
    bits.assign((V + 31) / 32, 0u);
    auto kill = [&](int id) {
        if (id >= 0 && id < V) bits[id >> 5] |= 1u << (id & 31);
    };
    kill(v.token_not);
    kill(v.token_sot);
    kill(v.token_nosp);
    if (!p.tdrz_enable) kill(v.token_solm);
    kill(v.token_translate);
    kill(v.token_transcribe);
    kill(v.token_prev);
    for (int i = 0; i <= lang_max_id(); ++i) kill(v.token_sot + 1 + i);
    if (p.suppress_regex) {
        try {
            std::regex re(p.suppress_regex);
            for (const auto & kv : v.token_to_id)
                if (std::regex_match(kv.first, re)) kill(kv.second);
        } catch (const std::regex_error &) {
            wlog(GGML_LOG_LEVEL_ERROR, "%s: invalid suppress_regex ignored\n", __func__);
        }
    }
    if (p.suppress_nst) {
        for (const char * t : kNonSpeech) {
            const std::string a = t, b = std::string(" ") + t;
            auto ia = v.token_to_id.find(a);
            if (ia != v.token_to_id.end()) kill(ia->second);
            auto ib = v.token_to_id.find(b);
            if (ib != v.token_to_id.end()) kill(ib->second);
        }
        auto i1 = v.token_to_id.find(" -");
        if (i1 != v.token_to_id.end()) kill(i1->second);
        auto i2 = v.token_to_id.find(" '");
        if (i2 != v.token_to_id.end()) kill(i2->second);
    }
}

// ---- host restatement of the reference's logit processing (used only by the host path) -------------------------
// whisper_compute_logprobs / whisper_compute_probs, src/whisper.cpp:6137-6171
static void compute_logprobs(const float * logits, int n, std::vector<float> & logprobs) {
    float logit_max = -INFINITY;
    for (int i = 0; i < n; ++i) logit_max = std::max(logit_max, logits[i]);
    float logsumexp = 0.0f;
    for (int i = 0; i < n; ++i)
        if (logits[i] > -INFINITY) logsumexp += expf(logits[i] - logit_max);
    logsumexp = logf(logsumexp) + logit_max;
    logprobs.resize(n);
    for (int i = 0; i < n; ++i) logprobs[i] = logits[i] > -INFINITY ? logits[i] - logsumexp : -INFINITY;
}

// The host sampling path works per decoder on private state (its logits / probs copies, its mt19937), as the reference's
// worker threads do (src/whisper.cpp:7504-7538): run fn(0..n-1) on up to 32 host threads.  Exactly the same arithmetic
// per item as the serial loop, so results do not depend on the thread count.
template <typename F> static void parallel_for(int n, F && fn, int max_threads = 32) {
    const int hw = (int) std::thread::hardware_concurrency();
    const int n_thr = std::max(1, std::min(n, std::min(hw > 0 ? hw : 4, std::max(1, max_threads))));
    if (n_thr <= 1) {
        for (int i = 0; i < n; ++i) fn(i);
        return;
    }
    std::atomic<int> next{0};
    std::vector<std::thread> pool;
    pool.reserve(n_thr - 1);
    auto work = [&]() {
        for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) fn(i);
    };
    for (int t = 1; t < n_thr; ++t) pool.emplace_back(work);
    work();
    for (auto & th : pool) th.join();
}

// whisper_process_logits, src/whisper.cpp:6177-6445
// cb_ctx: only handed to the user's logits_filter_callback; vocab / n_audio_ctx: the model's
static void process_logits_host(whisper_context * cb_ctx, const Vocab & vocab, int n_audio_ctx, whisper_state & state, whisper_decoder & dec,
                                const whisper_full_params & params, const std::vector<uint32_t> & static_bits,
                                const float * logits_row, float temperature) {
    const int n = vocab.n_vocab;
    const auto & toks = dec.sequence.tokens;
    const bool is_initial = toks.empty();
    auto & logits = dec.logits;
    auto & logprobs = dec.logprobs;
    auto & probs = dec.probs;
    logits.assign(logits_row, logits_row + n);
    if (temperature > 0.0f)
        for (int i = 0; i < n; ++i) logits[i] /= temperature;
    probs.resize(n);

    if (params.suppress_blank && is_initial) {
        logits[vocab.token_eot] = -INFINITY;
        auto it = vocab.token_to_id.find(" ");
        if (it != vocab.token_to_id.end()) logits[it->second] = -INFINITY;
    }
    // state-independent suppressions before the user callback: <|notimestamps|>, sot, nosp, solm, task, prev, languages
    logits[vocab.token_not] = -INFINITY;
    if (params.no_timestamps)
        for (int i = vocab.token_beg; i < n; ++i) logits[i] = -INFINITY;
    logits[vocab.token_sot] = -INFINITY;
    logits[vocab.token_nosp] = -INFINITY;
    if (!params.tdrz_enable) logits[vocab.token_solm] = -INFINITY;
    logits[vocab.token_translate] = -INFINITY;
    logits[vocab.token_transcribe] = -INFINITY;
    logits[vocab.token_prev] = -INFINITY;
    for (int i = 0; i <= lang_max_id(); ++i) {
        const int id = vocab.token_sot + 1 + i;
        if (id < n) logits[id] = -INFINITY;
    }
    if (params.logits_filter_callback) {
        params.logits_filter_callback(cb_ctx, &state, toks.data(), (int) toks.size(), logits.data(),
                                      params.logits_filter_callback_user_data);
    }
    // regex / non-speech suppressions (after the callback, as in the reference)
    for (int i = 0; i < n; ++i)
        if ((static_bits[i >> 5] >> (i & 31)) & 1u) logits[i] = -INFINITY;

    {
        const bool last_ts = !toks.empty() && toks.back().id >= vocab.token_beg;
        const bool pen_ts = toks.size() < 2 || toks[toks.size() - 2].id >= vocab.token_beg;
        if (last_ts) {
            if (pen_ts) {
                for (int i = vocab.token_beg; i < n; ++i) logits[i] = -INFINITY;
            } else {
                for (int i = 0; i < vocab.token_eot; ++i) logits[i] = -INFINITY;
            }
        }
    }
    if (is_initial && params.max_initial_ts > 0.0f) {
        const float precision = 30.0f / n_audio_ctx;
        const int tid0 = (int) std::round(params.max_initial_ts / precision);
        for (int i = vocab.token_beg + tid0 + 1; i < n; ++i) logits[i] = -INFINITY;
    }
    if (dec.has_ts) {
        const int tid0 = dec.seek_delta / 2;
        for (int i = vocab.token_beg; i < vocab.token_beg + tid0 && i < n; ++i) logits[i] = -INFINITY;
    }
    compute_logprobs(logits.data(), n, logprobs);
    {
        float timestamp_logprob = -INFINITY;
        {
            float logsumexp = 0.0f;
            float logprob_max = -INFINITY;
            for (int i = vocab.token_beg; i < n; ++i) logprob_max = std::max(logprob_max, logprobs[i]);
            for (int i = vocab.token_beg; i < n; ++i)
                if (logprobs[i] > -INFINITY) logsumexp += expf(logprobs[i] - logprob_max);
            if (logsumexp > 0.0f) timestamp_logprob = logf(logsumexp) + logprob_max;
        }
        float max_text = -INFINITY;
        for (int i = 0; i < vocab.token_beg; ++i) max_text = std::max(max_text, logprobs[i]);
        if (timestamp_logprob > max_text) {
            for (int i = 0; i < vocab.token_beg; ++i) {
                logits[i] = -INFINITY;
                logprobs[i] = -INFINITY;
            }
        }
    }
    for (int i = 0; i < n; ++i) probs[i] = logits[i] == -INFINITY ? 0.0f : expf(logprobs[i]);
}

// whisper_sample_token, src/whisper.cpp:6460-6517
static whisper_token_data sample_token_host(const Vocab & vocab, whisper_decoder & dec, bool best) {
    whisper_token_data r = {0, 0, 0.0f, 0.0f, 0.0f, 0.0f, -1, -1, -1, 0.0f};
    const auto & probs = dec.probs;
    const auto & logprobs = dec.logprobs;
    const int n = vocab.n_vocab;
    {
        double sum_ts = 0.0, max_ts = 0.0;
        for (int i = vocab.token_beg; i < n; ++i) {
            sum_ts += probs[i];
            if (max_ts < probs[i]) {
                max_ts = probs[i];
                r.tid = i;
            }
        }
        r.pt = (float) (max_ts / (sum_ts + 1e-10));
        r.ptsum = (float) sum_ts;
    }
    if (best) {
        for (int i = 0; i < n; ++i)
            if (r.p < probs[i]) {
                r.id = i;
                r.p = probs[i];
                r.plog = logprobs[i];
            }
    } else {
        std::discrete_distribution<> dist(probs.begin(), probs.end());
        r.id = dist(dec.rng);
        r.p = probs[r.id];
        r.plog = logprobs[r.id];
    }
    if (r.id >= vocab.token_beg) {
        r.tid = r.id;
        r.pt = r.p;
    }
    return r;
}

// whisper_sample_token_topk, src/whisper.cpp:6519-6592: k draws from the categorical distribution
static std::vector<whisper_token_data> sample_token_topk_host(const Vocab & vocab, whisper_decoder & dec, int k) {
    const auto & probs = dec.probs;
    const auto & logprobs = dec.logprobs;
    const int n = vocab.n_vocab;
    whisper_token tid = vocab.token_beg;
    float pt = 0.0f, ptsum = 0.0f;
    {
        double sum_ts = 0.0, max_ts = 0.0;
        for (int i = vocab.token_beg; i < n; ++i) {
            sum_ts += probs[i];
            if (max_ts < probs[i]) {
                max_ts = probs[i];
                tid = i;
            }
        }
        pt = (float) (max_ts / (sum_ts + 1e-10));
        ptsum = (float) sum_ts;
    }
    std::discrete_distribution<> dist(probs.begin(), probs.end());
    std::vector<whisper_token_data> out;
    out.reserve(k);
    for (int i = 0; i < k; ++i) {
        const int id = dist(dec.rng);
        whisper_token_data t = {id, tid, probs[id], logprobs[id], pt, ptsum, -1, -1, -1, 0.0f};
        if (t.id >= vocab.token_beg) {
            t.tid = t.id;
            t.pt = t.p;
        }
        out.push_back(t);
    }
    return out;
}

// whisper_sequence_score, src/whisper.cpp:6595-6641
static void sequence_score(const whisper_full_params & params, whisper_sequence & s) {
    if (s.result_len == 0) return;
    double result = 0.0;
    for (int i = 0; i < s.result_len; ++i) result += s.tokens[i].plog;
    s.sum_logprobs = result;
    s.avg_logprobs = result / s.result_len;
    double penalty = s.result_len;
    if (params.length_penalty > 0.0f) penalty = pow((5.0 + penalty) / 6.0, params.length_penalty);
    s.score = result / penalty;
    const int n = 32;
    int cnt = 0;
    double entropy = 0.0;
    std::map<whisper_token, int> counts;
    for (int i = std::max(0, s.result_len - n); i < s.result_len; ++i) {
        counts[s.tokens[i].id]++;
        cnt++;
    }
    for (const auto & kv : counts) {
        const double p = kv.second / (double) cnt;
        entropy -= p * log(p);
    }
    s.entropy = entropy;
}

static bool sequences_equal(const whisper_sequence & a, const whisper_sequence & b) {
    if (a.tokens.size() != b.tokens.size()) return false;
    for (int i = (int) a.tokens.size() - 1; i >= 0; --i)
        if (a.tokens[i].id != b.tokens[i].id) return false;
    return true;
}

// ---- single-window helpers used by the low-level API and by language detection --------------------------------
bool encode_single(whisper_context & ctx, whisper_state & st, int seek, bool keep_embd32) {
    Engine & e = ctx.eng;
    // audio context: the state's experimental override (set by whisper_full from params.audio_ctx, src/whisper.cpp:6986)
    if (!e.size_cross(st.cross, 1, st.exp_n_audio_ctx > 0 ? st.exp_n_audio_ctx : 1500)) return false;
    std::vector<EncJob> jobs(1);
    jobs[0].mel = &st.mel;
    jobs[0].seek = seek;
    const int64_t t0 = time_us();
    const bool ok = e.encode(jobs, st.cross, 0, keep_embd32);
    st.cross_base = st.cross.data.p;
    st.cross_layer_stride = st.cross.layer_stride;
    st.cross_T = st.cross.T;
    st.t_encode_us += time_us() - t0;
    st.n_encode++;
    return ok;
}

bool decode_single(whisper_context & ctx, whisper_state & st, const whisper_token * tokens, int n_tokens, int n_past) {
    Engine & e = ctx.eng;
    const auto & hp = e.model.hp;
    if (n_tokens <= 0 || n_past < 0 || n_past + n_tokens > hp.n_text_ctx) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: %d tokens at n_past=%d do not fit the text context (%d)\n", __func__, n_tokens, n_past,
             hp.n_text_ctx);
        return false;
    }
    if (!st.cross_base) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: whisper_encode() must be called first\n", __func__);
        return false;
    }
    whisper_decoder & dec = st.decoders[0];
    if (!dec.kv.reserve(e.self_kv_bytes())) return false;
    std::vector<DecRow> rows(n_tokens);
    for (int i = 0; i < n_tokens; ++i) {
        if (tokens[i] < 0 || tokens[i] >= hp.n_vocab) {
            wlog(GGML_LOG_LEVEL_ERROR, "%s: token %d out of range\n", __func__, tokens[i]);
            return false;
        }
        rows[i] = {tokens[i], n_past + i, dec.kv.p, st.cross_base};
    }
    const int64_t t0 = time_us();
    std::vector<int> lrows(1, n_tokens - 1);
    const int outer_T = e.cross_T;                   // a callback of a running whisper_full may call in here
    e.cross_T = st.cross_T;
    bool ok = e.decode(rows, lrows, st.cross_layer_stride);
    e.cross_T = outer_T;
    st.logits.resize((size_t) n_tokens * hp.n_vocab);
    ok = ok && e.fetch_logits(0, st.logits.data() + (size_t) (n_tokens - 1) * hp.n_vocab);
    const int64_t dt = time_us() - t0;
    if (n_tokens == 1) {
        st.t_decode_us += dt;
        st.n_decode++;
    } else if (n_tokens < 16) {
        st.t_batchd_us += dt;
        st.n_batchd += n_tokens;
    } else {
        st.t_prompt_us += dt;
        st.n_prompt += n_tokens;
    }
    return ok;
}

static int lang_from_logits(const Vocab & vocab, const float * logits, float * lang_probs);

// whisper_lang_auto_detect_with_state, src/whisper.cpp:4021-4094
int lang_auto_detect(whisper_context & ctx, whisper_state & st, int offset_ms, float * lang_probs) {
    const int seek = offset_ms / 10;
    if (seek < 0) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: offset %dms is before the start of the audio\n", __func__, offset_ms);
        return -1;
    }
    if (seek >= st.mel.n_len_org) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: offset %dms is past the end of the audio (%dms)\n", __func__, offset_ms,
             st.mel.n_len_org * 10);
        return -2;
    }
    if (!encode_single(ctx, st, seek, false)) return -6;
    const Vocab & vocab = ctx.eng.model.vocab;
    const whisper_token sot = vocab.token_sot;
    if (!decode_single(ctx, st, &sot, 1, 0)) return -7;
    return lang_from_logits(vocab, st.logits.data(), lang_probs);
}

// softmax over the language tokens of one logits row -> most probable language id (src/whisper.cpp:4060-4091)
static int lang_from_logits(const Vocab & vocab, const float * logits, float * lang_probs) {
    std::vector<std::pair<double, int>> li;
    const int n_lang = std::min(lang_max_id() + 1, std::max(0, vocab.n_vocab - (vocab.token_sot + 1)));
    for (int i = 0; i < n_lang; ++i) li.emplace_back(logits[vocab.token_sot + 1 + i], i);
    std::sort(li.begin(), li.end(), [](const std::pair<double, int> & a, const std::pair<double, int> & b) { return a.first > b.first; });
    const double mx = li[0].first;
    double sum = 0.0;
    for (auto & kv : li) {
        kv.first = exp(kv.first - mx);
        sum += kv.first;
    }
    for (auto & kv : li) kv.first /= sum;
    if (lang_probs)
        for (const auto & kv : li) lang_probs[kv.second] = (float) kv.first;
    return li[0].second;
}

// ---- the batched whisper_full ---------------------------------------------------------------------------------
namespace {

// One hypothesis of a beam-search step: beam `decoder_idx` continued by `token`.  The reference copies the parent's whole
// whisper_sequence into every candidate (src/whisper.cpp:7280-7291); here a candidate is a (parent, token) pair and sequences are
// materialised only for the beams that actually change parent.
struct beam_candidate {
    int decoder_idx;
    int seek_delta;
    bool has_ts;
    whisper_token_data token;
    double sum_logprobs_all;
};

enum class Phase { WINDOW, PROMPT, STEPPING, RANK, DONE };

struct Stream {
    whisper_state * state = nullptr;
    whisper_full_params params;
    const float * samples = nullptr;
    int n_samples = 0;
    bool samples_on_device = false;
    bool samples_i16 = false;
    int detected_lang = -1;      // >= 0: language found by the batched detection pass of run_streams
    std::vector<float> detected_probs;
    int rc = 0;
    Phase phase = Phase::WINDOW;

    int seek_start = 0, seek_end = 0, seek = 0;
    std::vector<float> temperatures;
    int n_decoders = 1;
    int max_prompt_ctx = 0;
    std::vector<whisper_token> prompt_tokens_buf, prompt_init, prompt;
    std::string language_buf;

    int it = 0;                  // temperature index
    float t_cur = 0.0f;
    int n_decoders_cur = 1;
    bool device_path = false;
    int k_draws = 0;             // device path: categorical draws per live decoder and step (0 = arg-max)
    int best_decoder_id = 0;
    int window = 0;              // index into the shared cross pool
    bool no_timestamps = false;
};

}  // namespace

static int stream_begin(whisper_context & ctx, Stream & s) {
    whisper_state * state = s.state;
    auto & params = s.params;
    const Vocab & vocab = ctx.eng.model.vocab;
    const auto & hp = ctx.eng.model.hp;

    // language (auto-detection runs the single-window path for this stream)
    if (params.language == nullptr || strlen(params.language) == 0 || strcmp(params.language, "auto") == 0 ||
        params.detect_language) {
        std::vector<float> probs(lang_max_id() + 1, 0.0f);
        if (s.detected_lang >= 0) probs = s.detected_probs;
        const int id = s.detected_lang >= 0 ? s.detected_lang : lang_auto_detect(ctx, *state, 0, probs.data());
        if (id < 0) {
            wlog(GGML_LOG_LEVEL_ERROR, "%s: failed to auto-detect language\n", __func__);
            return -3;
        }
        state->lang_id = id;
        s.language_buf = lang_str(id);
        params.language = s.language_buf.c_str();
        wlog(GGML_LOG_LEVEL_INFO, "%s: auto-detected language: %s (p = %f)\n", __func__, params.language, probs[id]);
        if (params.detect_language) {
            s.phase = Phase::DONE;
            return 0;
        }
    }
    if (params.grammar_rules != nullptr && params.n_grammar_rules > 0)
        wlog(GGML_LOG_LEVEL_WARN, "%s: grammar-constrained sampling is outside the B200 path (SURVEY section 2): grammar_rules ignored\n", __func__);
    if (params.token_timestamps) {              // src/whisper.cpp:6863-6871
        state->t_beg = 0;
        state->t_last = 0;
        state->tid_last = 0;
        if (s.n_samples > 0) {
            // the envelope is computed on the host from f32 samples: fetch / convert the stream when it is not that already
            std::vector<float> host;
            const float * pcm = s.samples;
            if (s.samples_on_device || s.samples_i16) {
                host.resize((size_t) s.n_samples);
                std::vector<int16_t> h16;
                const void * src = s.samples;
                if (s.samples_on_device) {
                    void * dst = s.samples_i16 ? (h16.resize((size_t) s.n_samples), (void *) h16.data()) : (void *) host.data();
                    if (cudaMemcpy(dst, s.samples, (size_t) s.n_samples * (s.samples_i16 ? 2 : 4), cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
                    src = dst;
                }
                if (s.samples_i16) {
                    const int16_t * q = (const int16_t *) src;
                    for (int k = 0; k < s.n_samples; ++k) host[k] = (float) q[k] * (1.0f / 32768.0f);
                }
                pcm = host.data();
            }
            envelope_abs_mean(pcm, s.n_samples, 32, state->energy);
        }
    }
    s.seek_start = params.offset_ms / 10;
    s.seek_end = params.duration_ms == 0 ? state->mel.n_len_org : s.seek_start + params.duration_ms / 10;
    if (s.seek_end < s.seek_start + 10) {
        wlog(GGML_LOG_LEVEL_WARN, "%s: input is too short - %d ms < 100 ms. consider padding the input audio with silence\n",
             __func__, (s.seek_end - s.seek_start) * 10);
        s.phase = Phase::DONE;
        return 0;
    }
    if (params.temperature_inc > 0.0f) {
        for (float t = params.temperature; t < 1.0f + 1e-6f; t += params.temperature_inc) s.temperatures.push_back(t);
    } else {
        s.temperatures.push_back(params.temperature);
    }
    int n_decoders = 1;
    switch (params.strategy) {
        case WHISPER_SAMPLING_GREEDY: n_decoders = params.greedy.best_of; break;
        case WHISPER_SAMPLING_BEAM_SEARCH: n_decoders = std::max(params.greedy.best_of, params.beam_search.beam_size); break;
    }
    n_decoders = std::max(1, n_decoders);
    if (n_decoders > WHISPER_MAX_DECODERS) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: too many decoders requested (%d), max = %d\n", __func__, n_decoders, WHISPER_MAX_DECODERS);
        return -4;
    }
    s.n_decoders = n_decoders;
    for (int j = 1; j < n_decoders; ++j) state->decoders[j].rng = std::mt19937(j);

    if (params.no_context) {
        state->prompt_past0.clear();
        state->prompt_past1.clear();
    }
    s.max_prompt_ctx = std::min(params.n_max_text_ctx, hp.n_text_ctx / 2);
    {
        if (!params.prompt_tokens && params.initial_prompt) {
            // greedy longest-match tokenisation, see whisper_tokenize in whisper_api.cu
            s.prompt_tokens_buf.resize(1024);
            int n = whisper_tokenize(&ctx, params.initial_prompt, s.prompt_tokens_buf.data(), (int) s.prompt_tokens_buf.size());
            if (n < 0) {
                s.prompt_tokens_buf.resize(-n);
                n = whisper_tokenize(&ctx, params.initial_prompt, s.prompt_tokens_buf.data(), (int) s.prompt_tokens_buf.size());
            }
            s.prompt_tokens_buf.resize(std::max(0, n));
            params.prompt_tokens = s.prompt_tokens_buf.data();
            params.prompt_n_tokens = (int) s.prompt_tokens_buf.size();
        }
        if (params.prompt_tokens && params.prompt_n_tokens > 0) {
            if (params.carry_initial_prompt) {
                if (state->prompt_past0.empty()) {
                    const int max_tokens = std::max(1, s.max_prompt_ctx - 1);
                    if (params.prompt_n_tokens > max_tokens)
                        wlog(GGML_LOG_LEVEL_WARN, "%s: initial prompt is too long (%d tokens), will use only the last %d tokens\n",
                             __func__, params.prompt_n_tokens, max_tokens);
                    const int n_tokens = std::min(params.prompt_n_tokens, max_tokens);
                    state->prompt_past0.assign(params.prompt_tokens + (params.prompt_n_tokens - n_tokens),
                                               params.prompt_tokens + params.prompt_n_tokens);
                }
            } else {
                for (int i = 0; i < params.prompt_n_tokens; ++i) state->prompt_past1.push_back(params.prompt_tokens[i]);
                std::rotate(state->prompt_past1.begin(), state->prompt_past1.end() - params.prompt_n_tokens,
                            state->prompt_past1.end());
            }
        }
    }
    if (params.audio_ctx > hp.n_audio_ctx) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: audio_ctx is larger than the maximum allowed (%d > %d)\n", __func__, params.audio_ctx,
             hp.n_audio_ctx);
        return -5;
    }
    if (params.audio_ctx < 0) return -5;
    state->exp_n_audio_ctx = params.audio_ctx;          // 0: the model's 1500 positions (src/whisper.cpp:6981-6986)

    s.prompt_init = {vocab.token_sot};
    if (vocab.is_multilingual()) {
        const int id = lang_id(params.language);
        state->lang_id = id;
        s.prompt_init.push_back(vocab.token_sot + 1 + std::max(0, id));
        s.prompt_init.push_back(params.translate ? vocab.token_translate : vocab.token_transcribe);
    }
    {
        const bool is_distil = hp.n_text_layer == 2 && hp.n_vocab != 51866;
        if (is_distil && !params.no_timestamps) {
            wlog(GGML_LOG_LEVEL_WARN, "%s: using first release distilled models - forcing no_timestamps\n", __func__);
            params.no_timestamps = true;
        }
    }
    if (params.no_timestamps) s.prompt_init.push_back(vocab.token_not);
    s.no_timestamps = params.no_timestamps;
    s.seek = s.seek_start;
    s.phase = Phase::WINDOW;
    return 0;
}

// DTW token timestamps of the segments [i_segment, i_segment + n_segments) of the window that starts at `seek`: one more
// prompt-style decoder pass over [sot, (lang), notimestamps, text tokens, eot] with the alignment heads' cross-attention
// probabilities captured on the device, then the alignment on the host (dtw.cu; src/whisper.cpp:8850-8998).
static void dtw_window(whisper_context & ctx, Stream & s, int i_segment, int n_segments, int seek, int n_frames) {
    Engine & e = ctx.eng;
    whisper_state * state = s.state;
    const Vocab & vocab = e.model.vocab;
    std::vector<whisper_token> tokens = {vocab.token_sot};
    if (vocab.is_multilingual()) {
        const int id = lang_id(s.params.language);
        state->lang_id = id;
        tokens.push_back(vocab.token_sot + 1 + std::max(0, id));
    }
    const int sot_len = (int) tokens.size();
    tokens.push_back(vocab.token_not);
    std::vector<whisper_token_data *> text;
    for (int i = i_segment; i < i_segment + n_segments; ++i)
        for (auto & t : state->result_all[i].tokens)
            if (t.id < vocab.token_eot) {
                tokens.push_back(t.id);
                text.push_back(&t);
            }
    tokens.push_back(vocab.token_eot);
    const int n_tokens = (int) tokens.size(), n_audio = n_frames / 2, T = e.cross_T;
    if (text.empty() || n_tokens > e.model.hp.n_text_ctx || n_audio <= 7 || n_audio > T || !state->cross_base) return;
    whisper_decoder & dec = state->decoders[0];
    if (!dec.kv.reserve(e.self_kv_bytes())) return;
    std::vector<DecRow> rows(n_tokens);
    for (int i = 0; i < n_tokens; ++i) rows[i] = {tokens[i], i, dec.kv.p, state->cross_base};
    e.align.on = true;
    const bool ok = e.decode(rows, {}, ctx.batch_cross.layer_stride);
    e.align.on = false;
    if (!ok) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: decoder pass for the alignment heads failed\n", __func__);
        return;
    }
    const int A = e.align.n_heads_total;
    std::vector<float> probs((size_t) A * n_tokens * T);
    WB_CUDA(cudaMemcpyAsync(probs.data(), e.align.probs.p, probs.size() * sizeof(float), cudaMemcpyDeviceToHost, e.stream));
    WB_CUDA(cudaStreamSynchronize(e.stream));
    if (cuda_failed()) return;
    const std::vector<int> first = dtw_align(probs.data(), A, n_tokens, T, n_audio, sot_len, 7);
    // row 0 of the alignment is <|notimestamps|>; text token k owns row k + 1 (one audio position = 20 ms = 2 ticks)
    for (size_t k = 0; k < text.size() && k + 1 < first.size(); ++k)
        if (first[k + 1] >= 0) text[k]->t_dtw = (int64_t) first[k + 1] * 2 + seek;
}

// emit the segments of the finished window and advance the seek position (src/whisper.cpp:7609-7772)
static void stream_finish_window(whisper_context & ctx, Stream & s) {
    whisper_state * state = s.state;
    const auto & params = s.params;
    const Vocab & vocab = ctx.eng.model.vocab;
    auto & result_all = state->result_all;
    const size_t n_segments_before = result_all.size();
    const bool dtw = ctx.eng.align.n_heads_total > 0;
    const whisper_decoder & best = state->decoders[s.best_decoder_id];
    int seek_delta = best.seek_delta;
    const int result_len = best.sequence.result_len;
    const auto & tokens_cur = best.sequence.tokens;
    const int seek = s.seek;

    const bool is_no_speech = (state->no_speech_prob > params.no_speech_thold && best.sequence.avg_logprobs < params.logprob_thold);

    state->prompt_past1.clear();
    if (!params.carry_initial_prompt && !s.prompt.empty() && s.prompt.front() == vocab.token_prev) {
        state->prompt_past1.insert(state->prompt_past1.end(), s.prompt.begin() + 1, s.prompt.end() - s.prompt_init.size());
    }
    if (!is_no_speech)
        for (int i = 0; i < result_len; ++i) state->prompt_past1.push_back(tokens_cur[i].id);

    auto emit = [&](int64_t tt0, int64_t tt1, const std::string & text, int i0, int i1_incl, bool turn) {
        if (params.print_realtime) {
            if (params.print_timestamps) printf("[%s --> %s]  %s\n", to_timestamp(tt0, false).c_str(), to_timestamp(tt1, false).c_str(), text.c_str());
            else {
                printf("%s", text.c_str());
                fflush(stdout);
            }
        }
        result_all.push_back({tt0, tt1, text, state->no_speech_prob, {}, turn});
        for (int j = i0; j <= i1_incl; ++j) result_all.back().tokens.push_back(tokens_cur[j]);
        int n_new = 1;
        if (params.token_timestamps) {
            assign_token_times(vocab, *state, (int) result_all.size() - 1, params.thold_pt, params.thold_ptsum);
            if (params.max_len > 0) n_new = split_last_segment(vocab, *state, params.max_len, params.split_on_word);
        }
        if (params.new_segment_callback && !dtw) params.new_segment_callback(&ctx, state, n_new, params.new_segment_callback_user_data);
    };

    if (!tokens_cur.empty() && ctx.eng.model.n_loaded > 0 && !is_no_speech) {
        int i0 = 0;
        int64_t t0 = seek + 2 * (tokens_cur.front().tid - vocab.token_beg);
        std::string text;
        bool speaker_turn_next = false;
        for (int i = 0; i < (int) tokens_cur.size(); ++i) {
            if (params.print_special || tokens_cur[i].id < vocab.token_eot) text += vocab.id_to_token[tokens_cur[i].id];
            if (params.tdrz_enable && tokens_cur[i].id == vocab.token_solm) speaker_turn_next = true;
            if (tokens_cur[i].id > vocab.token_beg && !params.single_segment) {
                const int64_t t1 = seek + 2 * (tokens_cur[i].tid - vocab.token_beg);
                if (!text.empty()) emit(t0, t1, text, i0, i, speaker_turn_next);
                text = "";
                while (i < (int) tokens_cur.size() && tokens_cur[i].id > vocab.token_beg) i++;
                i--;
                t0 = t1;
                i0 = i + 1;
                speaker_turn_next = false;
            }
        }
        if (!text.empty()) emit(t0, seek + seek_delta, text, i0, (int) tokens_cur.size() - 1, speaker_turn_next);
    }
    {   // [EXPERIMENTAL] token-level timestamps with DTW (src/whisper.cpp:7745-7760)
        const int n_segments = (int) (result_all.size() - n_segments_before);
        if (dtw && n_segments) {
            const int n_frames = std::min(std::min(3000, seek_delta), s.seek_end - seek);
            dtw_window(ctx, s, (int) n_segments_before, n_segments, seek, n_frames);
            if (params.new_segment_callback)
                for (int seg = (int) result_all.size() - n_segments; seg < n_segments; seg++)
                    params.new_segment_callback(&ctx, state, seg, params.new_segment_callback_user_data);
        }
    }
    const bool single_timestamp_ending = tokens_cur.size() > 1 && tokens_cur[tokens_cur.size() - 2].id < vocab.token_beg &&
                                         tokens_cur[tokens_cur.size() - 1].id > vocab.token_beg;
    if (single_timestamp_ending) seek_delta = std::min(s.seek_end - seek, 3000);
    s.seek += seek_delta;
    s.phase = Phase::WINDOW;
}

static int run_streams_locked(whisper_context & ctx, std::vector<StreamSpec> & specs);

int run_streams(whisper_context & ctx, std::vector<StreamSpec> & specs) {
    Engine & e = ctx.eng;
    std::lock_guard<std::recursive_mutex> lock(e.mu);
    if (e.in_full) {
        // only reachable from a callback of the running call (same thread, recursive lock): the shared cross-K/V pool and the
        // suppression mask of the outer run are live
        wlog(GGML_LOG_LEVEL_ERROR, "%s: whisper_full* called from inside a callback of a running whisper_full* on the same context\n", __func__);
        for (auto & sp : specs) sp.rc = -1;
        return -1;
    }
    for (auto & sp : specs) {
        if (!sp.state) return -1;
        sp.rc = 0;
    }
    e.in_full = true;
    // Token-level timestamps threshold the token probabilities (src/whisper.cpp:8455-8660): a call that asks for them keeps the
    // reference's rounding points (separate LayerNorm kernels) instead of the algebraic LayerNorm fold of the decoder step
    e.exact_ln = false;
    for (const auto & sp : specs) e.exact_ln = e.exact_ln || sp.params.token_timestamps;
    int rc = -6;
    try {
        rc = run_streams_locked(ctx, specs);
    } catch (...) {
        e.in_full = false;
        e.exact_ln = false;
        throw;
    }
    e.in_full = false;
    e.exact_ln = false;
    bool any = false;
    for (const auto & sp : specs) any = any || sp.rc != 0;
    if (rc != 0 && !any)                        // a call-wide failure (allocation, mixed audio_ctx): every stream failed with it
        for (auto & sp : specs) sp.rc = rc;
    // the windows' cross K/V live in the context's shared pool, which the next batched call overwrites (or reallocates): a
    // later whisper_decode on one of these states needs a fresh whisper_encode
    for (auto & sp : specs) sp.state->cross_base = nullptr;
    return rc;
}

static int run_streams_locked(whisper_context & ctx, std::vector<StreamSpec> & specs) {
    Engine & e = ctx.eng;
    const bool dbg = getenv("WHISPER_B200_DEBUG_TIMING") != nullptr;
    // diagnostics for parity tests: stash the runner-up token / its logit distance in the (otherwise unused, DTW-only)
    // fields t_dtw / vlen of whisper_token_data
    const bool dbg_gaps = getenv("WHISPER_B200_DEBUG_GAPS") != nullptr;
    const int64_t dbg_t0 = time_us();
    int64_t dbg_mel = 0, dbg_enc = 0, dbg_prompt = 0, dbg_dec_submit = 0, dbg_sel = 0, dbg_host = 0;
    int dbg_steps = 0;
    cuda_clear_failure();
    const Vocab & vocab = e.model.vocab;
    const auto & hp = e.model.hp;
    const int n_streams = (int) specs.size();
    std::vector<Stream> S(n_streams);

    // ---- mel of every stream in one launch ----
    {
        const int64_t t0 = time_us();
        std::vector<MelJob> jobs;
        for (int i = 0; i < n_streams; ++i) {
            S[i].state = specs[i].state;
            S[i].params = specs[i].params;
            S[i].samples = specs[i].samples;
            S[i].n_samples = specs[i].n_samples;
            S[i].samples_on_device = specs[i].samples_on_device;
            S[i].samples_i16 = specs[i].samples_i16;
            S[i].window = i;
            S[i].state->result_all.clear();
            if (specs[i].n_samples > 0) {
                MelJob j;
                if (specs[i].samples_on_device) j.pcm_dev = specs[i].samples;
                else j.pcm_host = specs[i].samples;
                j.n_samples = specs[i].n_samples;
                j.i16 = specs[i].samples_i16;
                j.out = &S[i].state->mel;
                jobs.push_back(j);
            }
        }
        if (!e.run_mel(jobs)) {
            wlog(GGML_LOG_LEVEL_ERROR, "%s: failed to compute log mel spectrogram\n", __func__);
            for (auto & sp : specs) sp.rc = -2;
            return -2;
        }
        const int64_t dt = time_us() - t0;
        dbg_mel += dt;
        for (int i = 0; i < n_streams; ++i) S[i].state->t_mel_us += dt / std::max(1, n_streams);
    }
    // Language detection for all the streams that ask for it ("auto" / detect_language) as ONE encoder batch and ONE decoder
    // step on <|sot|> instead of a single-window pass per stream (what whisper_lang_auto_detect does; the arithmetic per stream
    // is the same).  Streams this pass cannot take (no audio, a left-over audio_ctx override) fall back to the per-stream path.
    {
        std::vector<int> need;
        for (int i = 0; i < n_streams; ++i) {
            const auto & p = S[i].params;
            const bool wants = p.language == nullptr || strlen(p.language) == 0 || strcmp(p.language, "auto") == 0 || p.detect_language;
            if (wants && vocab.is_multilingual() && S[i].state->mel.valid && S[i].state->mel.n_len_org > 0 && S[i].state->exp_n_audio_ctx == 0)
                need.push_back(i);
        }
        if (need.size() > 1 && e.size_cross(ctx.batch_cross, n_streams, 1500)) {
            const int64_t t0 = time_us();
            bool ok = true;
            for (size_t k = 0; k < need.size() && ok; k += 32) {
                std::vector<EncJob> jobs;
                for (size_t q = k; q < std::min(need.size(), k + 32); ++q) jobs.push_back({&S[need[q]].state->mel, 0});
                ok = e.encode(jobs, ctx.batch_cross, (int) k, false);
            }
            std::vector<DecRow> rows;
            std::vector<int> lrows;
            for (size_t q = 0; q < need.size() && ok; ++q) {
                whisper_decoder & dec = S[need[q]].state->decoders[0];
                ok = dec.kv.reserve(e.self_kv_bytes());
                rows.push_back({vocab.token_sot, 0, dec.kv.p, ctx.batch_cross.window_base((int) q, hp.n_text_state)});
                lrows.push_back((int) q);
            }
            e.cross_T = 1500;
            ok = ok && e.decode(rows, lrows, ctx.batch_cross.layer_stride);
            std::vector<float> lg;
            if (ok) {
                lg.resize(need.size() * (size_t) hp.n_vocab);
                ok = e.fetch_logits_rows(0, (int) need.size(), lg.data());
            }
            const int64_t dt = time_us() - t0;
            for (size_t q = 0; q < need.size() && ok; ++q) {
                Stream & s = S[need[q]];
                s.detected_probs.assign(lang_max_id() + 1, 0.0f);
                s.detected_lang = lang_from_logits(vocab, lg.data() + q * (size_t) hp.n_vocab, s.detected_probs.data());
                s.state->t_encode_us += dt / (int64_t) need.size();
                s.state->n_encode++;
                s.state->n_decode++;
            }
            if (!ok) cuda_clear_failure();        // the per-stream path reports the error with the reference's code
        }
    }
    for (int i = 0; i < n_streams; ++i) {
        S[i].rc = stream_begin(ctx, S[i]);
        if (S[i].rc != 0) S[i].phase = Phase::DONE;
    }

    // static suppression mask (state independent rules); identical params across streams of one call
    std::vector<uint32_t> static_bits;
    build_static_suppress(vocab, S[0].params, static_bits);
    if (!ctx.static_mask.reserve(static_bits.size() * 4)) return -7;
    WB_CUDA(cudaMemcpy(ctx.static_mask.p, static_bits.data(), static_bits.size() * 4, cudaMemcpyHostToDevice));
    int space_id = -1;
    {
        auto it = vocab.token_to_id.find(" ");
        if (it != vocab.token_to_id.end()) space_id = it->second;
    }

    // one audio context per batch: the streams of a whisper_full_parallel call share their parameters
    int batch_T = 0;
    for (int i = 0; i < n_streams; ++i)
        if (S[i].rc == 0 && S[i].phase != Phase::DONE) {
            const int t = S[i].state->exp_n_audio_ctx > 0 ? S[i].state->exp_n_audio_ctx : 1500;
            if (batch_T != 0 && t != batch_T) return -5;
            batch_T = t;
        }
    if (batch_T == 0) batch_T = 1500;
    if (!e.size_cross(ctx.batch_cross, n_streams, batch_T)) return -7;
    e.cross_T = batch_T;
    struct CrossTReset { Engine & e; ~CrossTReset() { e.cross_T = 1500; } } cross_t_reset{e};
    const int n_max = hp.n_text_ctx / 2 - 4;
    std::vector<float> logits_host((size_t) hp.n_vocab), logits_rows_host;
    const int host_threads = S[0].params.logits_filter_callback ? std::max(1, S[0].params.n_threads) : 32;

    auto fail_stream = [&](Stream & s, int rc) {
        s.rc = rc;
        s.phase = Phase::DONE;
    };

    // Device-side selection for a set of (stream, decoder, logits row) triples: builds the per-row decoder state the rules need,
    // draws the uniforms of the sampling rows from the decoders' own mt19937 in row order (std::discrete_distribution consumes
    // generate_canonical<double, 53> per draw, i.e. the reference's RNG streams stay in step), runs dec_kernels.cu's sampler and
    // hands the result to the decoders: `pending` for arg-max rows, `sampled` for drawing rows.
    struct SelItem { int si, j, logits_row; };
    auto device_select = [&](const std::vector<SelItem> & items) -> bool {
        if (items.empty()) return true;
        std::vector<SampleRow> sr;
        std::vector<double> uniforms;
        sr.reserve(items.size());
        for (const auto & it : items) {
            Stream & s = S[it.si];
            whisper_decoder & d = s.state->decoders[it.j];
            const auto & tk = d.sequence.tokens;
            const int n = (int) tk.size();
            SampleRow r = {};
            r.logits_row = it.logits_row;
            r.n_tokens = n;
            r.last = n > 0 ? tk[n - 1].id : 0;
            r.penult = n > 1 ? tk[n - 2].id : 0;
            r.has_ts = d.has_ts ? 1 : 0;
            r.seek_delta = d.seek_delta;
            r.temperature = s.t_cur;
            r.n_draws = s.k_draws;
            r.draw_off = (int) uniforms.size();
            r.tid_default = s.params.strategy == WHISPER_SAMPLING_BEAM_SEARCH ? vocab.token_beg : 0;
            for (int q = 0; q < s.k_draws; ++q) uniforms.push_back(std::generate_canonical<double, 53>(d.rng));
            sr.push_back(r);
        }
        const auto & p = S[items[0].si].params;       // the streams of one call share the parameters that enter the kernel
        SampleParams prm;
        prm.n_vocab = hp.n_vocab;
        prm.token_eot = vocab.token_eot;
        prm.token_beg = vocab.token_beg;
        prm.token_space = space_id;
        prm.suppress_blank = p.suppress_blank;
        prm.no_timestamps = p.no_timestamps;
        prm.max_initial_ts = p.max_initial_ts;
        prm.tid0 = (int) std::round(p.max_initial_ts / (30.0f / hp.n_audio_ctx));
        std::vector<SampleOut> so;
        std::vector<DrawOut> dr;
        if (!e.sample(sr, uniforms, (const uint32_t *) ctx.static_mask.p, prm, so, dr)) return false;
        for (size_t q = 0; q < items.size(); ++q) {
            whisper_decoder & d = S[items[q].si].state->decoders[items[q].j];
            if (sr[q].n_draws == 0) {
                d.pending = {so[q].id, so[q].tid, so[q].p, so[q].plog, so[q].pt, so[q].ptsum, -1, -1, -1, 0.0f};
                if (dbg_gaps) {
                    d.pending.t_dtw = so[q].runner_up;
                    d.pending.vlen = so[q].gap;
                }
                d.has_pending = true;
            } else {
                d.sampled.clear();
                for (int k = 0; k < sr[q].n_draws; ++k) {
                    const DrawOut & o = dr[sr[q].draw_off + k];
                    whisper_token_data t = {o.id, so[q].tid, o.p, o.plog, so[q].pt, so[q].ptsum, -1, -1, -1, 0.0f};
                    if (t.id >= vocab.token_beg) {
                        t.tid = t.id;
                        t.pt = t.p;
                    }
                    d.sampled.push_back(t);
                }
            }
        }
        return true;
    };

    while (true) {
        // ---- A. window start: progress / end-of-audio / encoder_begin, then one batched encode ----
        std::vector<int> enc_ids;
        for (int si = 0; si < n_streams; ++si) {
            Stream & s = S[si];
            if (s.phase != Phase::WINDOW) continue;
            const auto & p = s.params;
            if (p.progress_callback) {
                const int cur = (100 * (s.seek - s.seek_start)) / (s.seek_end - s.seek_start);
                p.progress_callback(&ctx, s.state, cur, p.progress_callback_user_data);
            }
            if (s.seek + 10 >= s.seek_end) {
                s.phase = Phase::DONE;
                continue;
            }
            if (p.encoder_begin_callback && !p.encoder_begin_callback(&ctx, s.state, p.encoder_begin_callback_user_data)) {
                wlog(GGML_LOG_LEVEL_ERROR, "%s: encoder_begin_callback returned false - aborting\n", __func__);
                s.phase = Phase::DONE;
                continue;
            }
            enc_ids.push_back(si);
        }
        if (enc_ids.empty()) break;
        {
            const int64_t t0 = time_us();
            // windows are encoded in chunks to bound the activation workspace; K/V rows land at the stream's slot
            const int chunk = 32;
            size_t k = 0;
            bool ok = true;
            while (k < enc_ids.size() && ok) {
                // consecutive slots only (the GEMM writes a contiguous row range of the pool)
                size_t k1 = k + 1;
                while (k1 < enc_ids.size() && (int) (k1 - k) < chunk && S[enc_ids[k1]].window == S[enc_ids[k1 - 1]].window + 1) ++k1;
                std::vector<EncJob> jobs;
                for (size_t q = k; q < k1; ++q) jobs.push_back({&S[enc_ids[q]].state->mel, S[enc_ids[q]].seek});
                ok = e.encode(jobs, ctx.batch_cross, S[enc_ids[k]].window, false);
                k = k1;
            }
            const int64_t dt = time_us() - t0;
            dbg_enc += dt;
            for (int si : enc_ids) {
                Stream & s = S[si];
                s.state->t_encode_us += dt / (int64_t) enc_ids.size();
                s.state->n_encode++;
                if (!ok) {
                    wlog(GGML_LOG_LEVEL_ERROR, "%s: failed to encode\n", __func__);
                    fail_stream(s, -6);
                    continue;
                }
                if (s.params.abort_callback && s.params.abort_callback(s.params.abort_callback_user_data)) {
                    fail_stream(s, -6);
                    continue;
                }
                s.state->cross_base = ctx.batch_cross.window_base(s.window, hp.n_text_state);
                s.state->cross_layer_stride = ctx.batch_cross.layer_stride;
                s.state->cross_T = ctx.batch_cross.T;
                if (s.seek > s.seek_start && s.seek + 500 >= s.seek_end) {
                    s.state->prompt_past0.clear();
                    s.state->prompt_past1.clear();
                }
                s.best_decoder_id = 0;
                s.it = 0;
                s.phase = Phase::PROMPT;
            }
        }

        // ---- B. temperature rounds: streams that need a (re)decode at their current temperature go together ----
        while (true) {
            std::vector<int> act;
            for (int si = 0; si < n_streams; ++si)
                if (S[si].phase == Phase::PROMPT) act.push_back(si);
            if (act.empty()) break;

            // B1. per-stream decoder init + prompt; one batched prompt pass
            std::vector<DecRow> rows;
            std::vector<int> lrows;
            for (int si : act) {
                Stream & s = S[si];
                auto & p = s.params;
                whisper_state * st = s.state;
                s.t_cur = s.temperatures[s.it];
                int ndc = 1;
                switch (p.strategy) {
                    case WHISPER_SAMPLING_GREEDY:
                        if (s.t_cur > 0.0f) ndc = p.greedy.best_of;
                        break;
                    case WHISPER_SAMPLING_BEAM_SEARCH:
                        ndc = s.t_cur > 0.0f ? p.greedy.best_of : p.beam_search.beam_size;
                        break;
                }
                s.n_decoders_cur = std::max(1, ndc);
                // Selection runs on the device unless the user wants to see (and edit) the logits.  Greedy strategy: arg-max below
                // temperature 1e-6, else one categorical draw per decoder; beam-search strategy: beam_size draws per beam at any
                // temperature (src/whisper.cpp:7247-7268).  (Several decoders that would all take the arg-max -- 0 < t < 1e-6 with
                // best_of > 1 -- are left to the host path.)
                const bool argmax = p.strategy == WHISPER_SAMPLING_GREEDY && s.t_cur < 1e-6f;
                s.k_draws = argmax ? 0 : (p.strategy == WHISPER_SAMPLING_BEAM_SEARCH ? p.beam_search.beam_size : 1);
                s.device_path = !p.logits_filter_callback && !(argmax && s.n_decoders_cur > 1) && s.k_draws <= WHISPER_MAX_DECODERS;
                for (int j = 0; j < s.n_decoders_cur; ++j) {
                    whisper_decoder & d = st->decoders[j];
                    d.sequence.tokens.clear();
                    d.sequence.result_len = 0;
                    d.sequence.sum_logprobs_all = 0.0;
                    d.sequence.sum_logprobs = -INFINITY;
                    d.sequence.avg_logprobs = -INFINITY;
                    d.sequence.entropy = 0.0;
                    d.sequence.score = -INFINITY;
                    d.seek_delta = 3000;
                    d.failed = d.completed = d.has_ts = false;
                    d.has_pending = false;
                    if (!d.kv.reserve(e.self_kv_bytes())) return -7;
                }
                s.prompt.clear();
                if (p.n_max_text_ctx > 0 && s.t_cur < 0.5f) {
                    const bool can0 = p.carry_initial_prompt && !st->prompt_past0.empty();
                    const bool can1 = !st->prompt_past1.empty();
                    if (s.max_prompt_ctx > 0 && (can0 || can1)) {
                        s.prompt.push_back(vocab.token_prev);
                        int n_take0 = 0;
                        if (can0) {
                            n_take0 = (int) st->prompt_past0.size();
                            s.prompt.insert(s.prompt.end(), st->prompt_past0.end() - n_take0, st->prompt_past0.end());
                        }
                        const int n_take1 = std::min<int>(s.max_prompt_ctx - n_take0 - 1, (int) st->prompt_past1.size());
                        s.prompt.insert(s.prompt.end(), st->prompt_past1.end() - n_take1, st->prompt_past1.end());
                    }
                }
                s.prompt.insert(s.prompt.end(), s.prompt_init.begin(), s.prompt_init.end());
                for (int i = 0; i < (int) s.prompt.size(); ++i)
                    rows.push_back({s.prompt[i], i, st->decoders[0].kv.p, st->cross_base});
                lrows.push_back((int) rows.size() - 1);
            }
            {
                const int64_t t0 = time_us();
                bool ok = e.decode(rows, lrows, ctx.batch_cross.layer_stride);
                // no_speech_prob on the raw logits of the prompt pass (src/whisper.cpp:7188-7196)
                std::vector<SampleRow> sr(act.size());
                for (size_t a = 0; a < act.size(); ++a) sr[a] = {(int) a, 0, 0, 0, 0, 0};
                std::vector<float> nosp;
                ok = ok && e.token_prob(sr, vocab.token_nosp, nosp);
                const int64_t dt = time_us() - t0;
                dbg_prompt += dt;
                for (size_t a = 0; a < act.size(); ++a) {
                    Stream & s = S[act[a]];
                    s.state->t_prompt_us += dt / (int64_t) act.size();
                    s.state->n_prompt += (int) s.prompt.size();
                    if (!ok) {
                        wlog(GGML_LOG_LEVEL_ERROR, "%s: failed to decode\n", __func__);
                        fail_stream(s, -8);
                        continue;
                    }
                    s.state->no_speech_prob = nosp[a];
                    s.phase = Phase::STEPPING;
                }
                if (!ok) continue;
            }
            // B2. process the prompt logits for decoder 0 and fan out to the other decoders
            {
                const int64_t t0 = time_us();
                std::vector<SelItem> sel;
                for (size_t a = 0; a < act.size(); ++a) {
                    Stream & s = S[act[a]];
                    if (s.phase != Phase::STEPPING) continue;
                    whisper_state * st = s.state;
                    for (int j = 1; j < s.n_decoders_cur; ++j)
                        e.kv_copy_prefix(st->decoders[0].kv.p, st->decoders[j].kv.p, (int) s.prompt.size());
                    if (s.device_path) {
                        // every decoder of the stream selects from the same processed prompt row (with its own RNG)
                        for (int j = 0; j < s.n_decoders_cur; ++j) sel.push_back({act[a], j, (int) a});
                    } else {
                        if (!e.fetch_logits((int) a, logits_host.data())) {
                            fail_stream(s, -8);
                            continue;
                        }
                        process_logits_host(&ctx, vocab, hp.n_audio_ctx, *st, st->decoders[0], s.params, static_bits, logits_host.data(), s.t_cur);
                        for (int j = 1; j < s.n_decoders_cur; ++j) {
                            whisper_decoder & d = st->decoders[j];
                            d.probs = st->decoders[0].probs;
                            d.logits = st->decoders[0].logits;
                            d.logprobs = st->decoders[0].logprobs;
                        }
                    }
                }
                if (!device_select(sel))
                    for (const auto & it : sel) fail_stream(S[it.si], -8);
                const int64_t dt = time_us() - t0;
                for (int si : act) S[si].state->t_sample_us += dt / (int64_t) act.size();
            }

            // B3. token loop (src/whisper.cpp:7219-7544)
            std::vector<beam_candidate> cands;
            std::vector<Engine::KvCopy> kv_copies;
            for (int i = 0; i < n_max; ++i) {
                const int64_t ts0 = time_us();
                bool any_live = false;
                {   // this iteration's host draws: every live decoder draws from its own distribution with its own mt19937, so the
                    // draws of different decoders are independent and are made on the host threads; consumed in order below
                    std::vector<std::pair<int, int>> jobs;
                    for (int si : act) {
                        Stream & s = S[si];
                        if (s.phase != Phase::STEPPING || s.device_path) continue;
                        for (int j = 0; j < s.n_decoders_cur; ++j)
                            if (!s.state->decoders[j].completed && !s.state->decoders[j].failed) jobs.push_back({si, j});
                    }
                    parallel_for((int) jobs.size(), [&](int q) {
                        Stream & s = S[jobs[q].first];
                        whisper_decoder & d = s.state->decoders[jobs[q].second];
                        if (s.params.strategy == WHISPER_SAMPLING_BEAM_SEARCH) d.sampled = sample_token_topk_host(vocab, d, s.params.beam_search.beam_size);
                        else d.sampled.assign(1, sample_token_host(vocab, d, s.t_cur < 1e-6f));
                    });
                }
                for (int si : act) {
                    Stream & s = S[si];
                    if (s.phase != Phase::STEPPING) continue;
                    whisper_state * st = s.state;
                    const auto & p = s.params;
                    const bool beam = p.strategy == WHISPER_SAMPLING_BEAM_SEARCH;
                    if (beam) cands.clear();
                    // sampling
                    for (int j = 0; j < s.n_decoders_cur; ++j) {
                        whisper_decoder & d = st->decoders[j];
                        if (d.completed || d.failed) continue;
                        if (!beam) {
                            whisper_token_data tok;
                            if (s.device_path && s.k_draws == 0) {
                                tok = d.pending;
                                d.has_pending = false;
                            } else {
                                tok = d.sampled[0];
                            }
                            d.sequence.tokens.push_back(tok);
                            d.sequence.sum_logprobs_all += tok.plog;
                        } else {
                            for (const auto & t : d.sampled)
                                cands.push_back({j, d.seek_delta, d.has_ts, t, d.sequence.sum_logprobs_all + t.plog});
                            if (!d.sampled.empty()) st->n_sample += 1;
                        }
                    }
                    if (beam) {
                        // rank the hypotheses (same comparator, same input order as the reference -> same permutation)
                        std::sort(cands.begin(), cands.end(), [](const beam_candidate & a, const beam_candidate & b) {
                            if (a.sum_logprobs_all != b.sum_logprobs_all) return a.sum_logprobs_all > b.sum_logprobs_all;
                            return a.decoder_idx < b.decoder_idx;
                        });
                        // two hypotheses are the same sequence if they append the same token to the same history
                        auto same_sequence = [&](const beam_candidate & a, const beam_candidate & b) {
                            if (a.token.id != b.token.id) return false;
                            if (a.decoder_idx == b.decoder_idx) return true;
                            const auto & ta = st->decoders[a.decoder_idx].sequence.tokens;
                            const auto & tb = st->decoders[b.decoder_idx].sequence.tokens;
                            if (ta.size() != tb.size()) return false;
                            for (int k = (int) ta.size() - 1; k >= 0; --k)
                                if (ta[k].id != tb[k].id) return false;
                            return true;
                        };
                        // pass 1 (reads only): which hypothesis every live beam continues with (src/whisper.cpp:7305-7334)
                        int chosen[WHISPER_MAX_DECODERS];
                        uint32_t cur_c = 0;
                        for (int j = 0; j < s.n_decoders_cur; ++j) {
                            chosen[j] = -1;
                            const whisper_decoder & d = st->decoders[j];
                            if (d.completed || d.failed) continue;
                            if (cur_c >= cands.size()) cur_c = 0;
                            const beam_candidate & cur = cands[cur_c];
                            chosen[j] = (int) cur_c++;
                            while (cands.size() > cur_c && same_sequence(cands[cur_c], cur) && i > 0) ++cur_c;
                        }
                        // pass 2: beams that change parent take a copy of the parent's OLD sequence and self-attention history
                        // (into the alternate cache; all copies of the step go out as one launch after the stream loop) ...
                        const int n_past_kv = (int) s.prompt.size() + i;
                        whisper_sequence moved[WHISPER_MAX_DECODERS];
                        for (int j = 0; j < s.n_decoders_cur; ++j) {
                            if (chosen[j] < 0 || cands[chosen[j]].decoder_idx == j) continue;
                            whisper_decoder & d = st->decoders[j];
                            moved[j] = st->decoders[cands[chosen[j]].decoder_idx].sequence;
                            if (!d.kv_alt.reserve(e.self_kv_bytes())) return -7;
                            kv_copies.push_back({st->decoders[cands[chosen[j]].decoder_idx].kv.p, d.kv_alt.p, n_past_kv});
                        }
                        // ... then every live beam appends its token
                        for (int j = 0; j < s.n_decoders_cur; ++j) {
                            if (chosen[j] < 0) continue;
                            const beam_candidate & cur = cands[chosen[j]];
                            whisper_decoder & d = st->decoders[j];
                            if (cur.decoder_idx != j) {
                                d.sequence = std::move(moved[j]);
                                std::swap(d.kv.p, d.kv_alt.p);          // valid once this step's copy launch has run (before the decode)
                                std::swap(d.kv.cap, d.kv_alt.cap);
                            }
                            d.seek_delta = cur.seek_delta;
                            d.has_ts = cur.has_ts;
                            d.sequence.tokens.push_back(cur.token);
                            d.sequence.sum_logprobs_all = cur.sum_logprobs_all;
                        }
                    }
                    // per-decoder state machine
                    for (int j = 0; j < s.n_decoders_cur; ++j) {
                        whisper_decoder & d = st->decoders[j];
                        if (d.completed || d.failed) continue;
                        const auto & token = d.sequence.tokens.back();
                        if (token.id > vocab.token_beg) {
                            const int seek_delta_new = 2 * (token.id - vocab.token_beg);
                            if (d.has_ts && d.seek_delta > seek_delta_new && d.sequence.result_len < i) {
                                d.failed = true;
                                continue;
                            }
                            d.seek_delta = seek_delta_new;
                            d.sequence.result_len = i + 1;
                            d.has_ts = true;
                        }
                        if (token.id == vocab.token_eot || (p.max_tokens > 0 && i >= p.max_tokens) ||
                            (d.has_ts && s.seek + d.seek_delta + 10 >= s.seek_end)) {
                            if (d.sequence.result_len == 0 && !p.no_timestamps) {
                                if (s.seek + d.seek_delta + 10 >= s.seek_end) {
                                    d.sequence.result_len = i + 1;
                                } else {
                                    d.failed = true;
                                    continue;
                                }
                            }
                            if (p.single_segment || p.no_timestamps) {
                                d.sequence.result_len = i + 1;
                                d.seek_delta = 3000;
                            }
                            d.completed = true;
                            continue;
                        }
                        if (e.model.n_loaded == 0) {
                            d.seek_delta = 3000;
                            d.completed = true;
                            continue;
                        }
                        if (i == n_max - 1 && (d.sequence.result_len == 0 || d.seek_delta < 1500)) {
                            d.failed = true;
                            continue;
                        }
                    }
                    bool all_done = true;
                    for (int j = 0; j < s.n_decoders_cur; ++j)
                        if (!(st->decoders[j].completed || st->decoders[j].failed)) all_done = false;
                    if (all_done) s.phase = Phase::RANK;
                    else any_live = true;
                }
                const int64_t ts1 = time_us();
                for (int si : act) S[si].state->t_sample_us += (ts1 - ts0) / (int64_t) act.size();
                if (!any_live) break;

                if (!kv_copies.empty()) {           // the parent histories of every beam that changed parent in this step
                    if (!e.kv_copy_prefix_batch(kv_copies)) return -7;
                    kv_copies.clear();
                }
                // next-token rows of every live sequence
                rows.clear();
                lrows.clear();
                std::vector<std::pair<int, int>> owner;     // (stream, decoder) per row
                for (int si : act) {
                    Stream & s = S[si];
                    if (s.phase != Phase::STEPPING) continue;
                    const int n_past = (int) s.prompt.size() + i;
                    for (int j = 0; j < s.n_decoders_cur; ++j) {
                        whisper_decoder & d = s.state->decoders[j];
                        if (d.failed || d.completed) continue;
                        d.i_batch = (int) rows.size();
                        rows.push_back({d.sequence.tokens.back().id, n_past, d.kv.p, s.state->cross_base});
                        lrows.push_back((int) rows.size() - 1);
                        owner.emplace_back(si, j);
                    }
                }
                const int64_t td0 = time_us();
                dbg_host += td0 - ts0;
                dbg_steps++;
                bool ok = e.decode(rows, lrows, ctx.batch_cross.layer_stride);
                dbg_dec_submit += time_us() - td0;
                // selection for the NEXT iteration: device-path rows in one kernel; nothing is selected from the last step's logits
                // (the reference's loop ends there too, so no uniform may be drawn from the decoders' generators)
                if (ok && i + 1 < n_max) {
                    std::vector<SelItem> sel;
                    for (size_t r = 0; r < owner.size(); ++r)
                        if (S[owner[r].first].device_path) sel.push_back({owner[r].first, owner[r].second, (int) r});
                    ok = device_select(sel);
                }
                const int64_t td1 = time_us();
                dbg_sel += td1 - td0;
                {   // host-path rows: one D2H for the lot, then whisper_process_logits per decoder on the host threads
                    std::vector<int> host_rows;
                    for (size_t r = 0; r < owner.size(); ++r)
                        if (!S[owner[r].first].device_path) host_rows.push_back((int) r);
                    if (ok && !host_rows.empty()) {
                        const int r_lo = host_rows.front(), r_hi = host_rows.back();
                        logits_rows_host.resize((size_t) (r_hi - r_lo + 1) * hp.n_vocab);
                        ok = e.fetch_logits_rows(r_lo, r_hi - r_lo + 1, logits_rows_host.data());
                        if (ok)
                            parallel_for((int) host_rows.size(), [&](int q) {
                                const int r = host_rows[q];
                                Stream & s = S[owner[r].first];
                                whisper_decoder & d = s.state->decoders[owner[r].second];
                                process_logits_host(&ctx, vocab, hp.n_audio_ctx, *s.state, d, s.params, static_bits,
                                                    logits_rows_host.data() + (size_t) (r - r_lo) * hp.n_vocab, s.t_cur);
                            }, host_threads);     // the user's logits_filter_callback runs on these threads: honour n_threads
                    }
                }
                const int64_t td2 = time_us();
                // timing buckets as the reference: per-call, by batch width
                for (int si : act) {
                    Stream & s = S[si];
                    if (s.phase != Phase::STEPPING) continue;
                    int n_rows = 0;
                    for (const auto & o : owner)
                        if (o.first == si) n_rows++;
                    const int64_t share = (td1 - td0) / std::max<int64_t>(1, (int64_t) owner.size()) * n_rows;
                    if (n_rows == 1) {
                        s.state->t_decode_us += share;
                        s.state->n_decode++;
                    } else {
                        s.state->t_batchd_us += share;
                        s.state->n_batchd += n_rows;
                    }
                    s.state->t_sample_us += (td2 - td1) / std::max<int64_t>(1, (int64_t) act.size());
                    s.state->n_sample += n_rows;
                    if (!ok) {
                        wlog(GGML_LOG_LEVEL_ERROR, "%s: failed to decode\n", __func__);
                        fail_stream(s, -9);
                    } else if (s.params.abort_callback && s.params.abort_callback(s.params.abort_callback_user_data)) {
                        fail_stream(s, -9);
                    }
                }
                if (!ok) break;
            }
            // streams that ran out of steps without all decoders finishing
            for (int si : act)
                if (S[si].phase == Phase::STEPPING) S[si].phase = Phase::RANK;

            // B4. rank sequences, decide on fallback (src/whisper.cpp:7546-7606)
            for (int si : act) {
                Stream & s = S[si];
                if (s.phase != Phase::RANK) continue;
                whisper_state * st = s.state;
                const auto & p = s.params;
                double best_score = -INFINITY;
                for (int j = 0; j < s.n_decoders_cur; ++j) {
                    whisper_decoder & d = st->decoders[j];
                    if (d.failed) continue;
                    d.sequence.tokens.resize(d.sequence.result_len);
                    sequence_score(p, d.sequence);
                    if (d.sequence.result_len > 32 && d.sequence.entropy < p.entropy_thold) {
                        d.failed = true;
                        st->n_fail_h++;
                        continue;
                    }
                    if (best_score < d.sequence.score) {
                        best_score = d.sequence.score;
                        s.best_decoder_id = j;
                    }
                }
                bool success = true;
                if (s.it != (int) s.temperatures.size() - 1) {
                    const whisper_decoder & d = st->decoders[s.best_decoder_id];
                    if (d.failed || (d.sequence.avg_logprobs < p.logprob_thold && st->no_speech_prob < p.no_speech_thold)) {
                        success = false;
                        st->n_fail_p++;
                    }
                }
                if (success) {
                    stream_finish_window(ctx, s);
                } else {
                    s.it++;
                    s.phase = Phase::PROMPT;
                }
            }
        }
    }
    if (dbg) {
        fprintf(stderr, "run_streams: %d streams total %.1f ms | mel %.1f | encode %.1f | prompt %.1f | %d steps: host %.1f, "
                "decode submit %.1f, decode+select (incl. wait) %.1f ms\n", n_streams, (time_us() - dbg_t0) / 1e3, dbg_mel / 1e3,
                dbg_enc / 1e3, dbg_prompt / 1e3, dbg_steps, dbg_host / 1e3, dbg_dec_submit / 1e3, dbg_sel / 1e3);
    }
    int rc = 0;
    for (int i = 0; i < n_streams; ++i) {
        specs[i].rc = S[i].rc;
        if (S[i].rc != 0 && rc == 0) rc = S[i].rc;
    }
    if (cuda_failed() && rc == 0) rc = -6;
    return rc;
}

}  // namespace wb

// Host-only hook (needs no device): the token-level timestamp heuristic and the max_len re-wrapping of ONE segment, with every
// input given explicitly.  tok_state = {t_beg, t_last, tid_last} in / out.  Returns the number of segments (>= 1) after
// wrapping, or -1.  seg_t[2k], seg_t[2k+1], seg_ntok[k] describe segment k; tokens are updated in place, in order.
extern "C" WB200_API int whisper_b200_token_timestamps(
    const char * const * token_texts, int n_vocab, int token_eot, int token_beg, const float * pcm, int n_samples, long long seg_t0,
    long long seg_t1, whisper_token_data * tokens, int n_tokens, float thold_pt, float thold_ptsum, long long * tok_state, int max_len,
    int split_on_word, long long * seg_t, int * seg_ntok, int seg_cap) {
    if (!token_texts || !tokens || !tok_state || n_tokens < 0 || n_vocab <= 0) return -1;
    wb::Vocab vocab;
    vocab.n_vocab = n_vocab;
    vocab.token_eot = token_eot;
    vocab.token_beg = token_beg;
    vocab.id_to_token.assign(token_texts, token_texts + n_vocab);
    whisper_state st;
    wb::envelope_abs_mean(pcm, n_samples, 32, st.energy);
    st.t_beg = tok_state[0];
    st.t_last = tok_state[1];
    st.tid_last = (whisper_token) tok_state[2];
    st.result_all.push_back({(int64_t) seg_t0, (int64_t) seg_t1, "", 0.0f, {}, false});
    st.result_all.back().tokens.assign(tokens, tokens + n_tokens);
    wb::assign_token_times(vocab, st, 0, thold_pt, thold_ptsum);
    int n_seg = 1;
    if (max_len > 0) n_seg = wb::split_last_segment(vocab, st, max_len, split_on_word != 0);
    tok_state[0] = st.t_beg;
    tok_state[1] = st.t_last;
    tok_state[2] = st.tid_last;
    int k = 0;
    for (size_t i = 0; i < st.result_all.size(); ++i) {
        const auto & seg = st.result_all[i];
        if ((int) i < seg_cap && seg_t && seg_ntok) {
            seg_t[2 * i] = seg.t0;
            seg_t[2 * i + 1] = seg.t1;
            seg_ntok[i] = (int) seg.tokens.size();
        }
        for (const auto & t : seg.tokens)
            if (k < n_tokens) tokens[k++] = t;
    }
    return n_seg;
}

// Host-only hook (needs no device): the host sampling path's restatement of whisper_process_logits + the greedy
// whisper_sample_token on ONE explicit logits row and decoder state (token history, has_ts, seek_delta), for a vocabulary given
// as its n_vocab token strings and special-token ids = {eot, sot, translate, transcribe, solm, prev, nosp, not, beg}.
extern "C" WB200_API int whisper_b200_process_logits(const char * const * token_texts, int n_vocab, const int * special, int n_audio_ctx,
                                                     struct whisper_full_params params, float temperature, const float * logits_row,
                                                     const whisper_token * hist, int n_hist, int has_ts, int seek_delta, float * logits_out,
                                                     float * logprobs_out, float * probs_out, struct whisper_token_data * tok_out,
                                                     int topk_k, unsigned topk_seed, struct whisper_token_data * topk_out) {
    if (!token_texts || !special || !logits_row || n_vocab <= 0 || n_audio_ctx <= 0) return -1;
    wb::Vocab vocab;
    vocab.n_vocab = n_vocab;
    vocab.id_to_token.assign(token_texts, token_texts + n_vocab);
    for (int i = 0; i < n_vocab; ++i) vocab.token_to_id[vocab.id_to_token[i]] = i;
    vocab.token_eot = special[0]; vocab.token_sot = special[1]; vocab.token_translate = special[2]; vocab.token_transcribe = special[3];
    vocab.token_solm = special[4]; vocab.token_prev = special[5]; vocab.token_nosp = special[6]; vocab.token_not = special[7];
    vocab.token_beg = special[8];
    whisper_state st;
    whisper_decoder & dec = st.decoders[0];
    dec.has_ts = has_ts != 0;
    dec.seek_delta = seek_delta;
    for (int i = 0; i < n_hist; ++i) {
        whisper_token_data td = {};
        td.id = hist[i];
        dec.sequence.tokens.push_back(td);
    }
    std::vector<uint32_t> static_bits;
    wb::build_static_suppress(vocab, params, static_bits);
    wb::process_logits_host(nullptr, vocab, n_audio_ctx, st, dec, params, static_bits, logits_row, temperature);
    if (logits_out) memcpy(logits_out, dec.logits.data(), (size_t) n_vocab * sizeof(float));
    if (logprobs_out) memcpy(logprobs_out, dec.logprobs.data(), (size_t) n_vocab * sizeof(float));
    if (probs_out) memcpy(probs_out, dec.probs.data(), (size_t) n_vocab * sizeof(float));
    if (tok_out) *tok_out = wb::sample_token_host(vocab, dec, true);
    if (topk_out && topk_k > 0) {          // the sampled draws of the beam / best_of path, decoder RNG seeded as given
        dec.rng = std::mt19937(topk_seed);
        const auto toks = wb::sample_token_topk_host(vocab, dec, topk_k);
        for (int i = 0; i < topk_k; ++i) topk_out[i] = toks[i];
    }
    return 0;
}

// Host-only hook: whisper_sequence_score restated (csrc/full.cu <- src/whisper.cpp:6595-6641) on a sequence given by its token
// ids and log-probabilities; out = {sum_logprobs, avg_logprobs, entropy, score}.
extern "C" WB200_API int whisper_b200_sequence_score(struct whisper_full_params params, const float * plog, const whisper_token * ids, int n,
                                                     int result_len, double * out) {
    if (!plog || !ids || !out || n < 0) return -1;
    whisper_sequence seq = {};
    for (int i = 0; i < n; ++i) {
        whisper_token_data td = {};
        td.id = ids[i];
        td.plog = plog[i];
        seq.tokens.push_back(td);
    }
    seq.result_len = result_len;
    wb::sequence_score(params, seq);
    out[0] = seq.sum_logprobs; out[1] = seq.avg_logprobs; out[2] = seq.entropy; out[3] = seq.score;
    return 0;
}
