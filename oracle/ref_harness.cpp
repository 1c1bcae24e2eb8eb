// TEST INFRASTRUCTURE ONLY -- never linked into, imported or called by the product path.
//
// Harness translation unit for building the UNMODIFIED reference CPU implementation
// (whisper.cpp + ggml-cpu, from the sources where they lie under /root/reference) into
// oracle/_ref/libwhisper_ref_<arch>.so.  It textually includes the reference's
// src/whisper.cpp (it is NOT copied into this repo) so the file-static internals the
// public C API does not expose -- the mel buffer, the encoder output, the cross K/V cache --
// can be read back for parity checks.  Everything in include/whisper.h of the reference is
// exported by the resulting library as well, so the ctypes bindings used against our own
// libwhisper.so can be pointed at the reference unchanged.
//
// Build: oracle/build_ref.sh  (g++ on the reference's own sources; the reference's cmake
// build system is not run).

#include REF_WHISPER_CPP   // -DREF_WHISPER_CPP='"/root/reference/src/whisper.cpp"'

#include <cstring>

extern "C" {

// mel of the default state after whisper_pcm_to_mel: [n_mel][n_len] f32
// (reference: whisper_state::mel, src/whisper.cpp:414-420, filled at 3170-3260)
__attribute__((visibility("default")))
int ref_mel_dims(struct whisper_context * ctx, int * n_len, int * n_len_org, int * n_mel) {
    if (!ctx || !ctx->state) return -1;
    *n_len     = ctx->state->mel.n_len;
    *n_len_org = ctx->state->mel.n_len_org;
    *n_mel     = ctx->state->mel.n_mel;
    return 0;
}

__attribute__((visibility("default")))
int ref_mel_copy(struct whisper_context * ctx, float * out) {
    if (!ctx || !ctx->state) return -1;
    const auto & m = ctx->state->mel;
    memcpy(out, m.data.data(), m.data.size()*sizeof(float));
    return 0;
}

// direct entry to log_mel_spectrogram with explicit filters (for 128-bin sweeps on 80-bin models)
__attribute__((visibility("default")))
int ref_log_mel(const float * samples, int n_samples, int n_mel, const float * filters_data, int n_threads,
                float * out, int out_cap, int * n_len, int * n_len_org) {
    whisper_state st;
    whisper_filters f;
    f.n_mel = n_mel;
    f.n_fft = 1 + WHISPER_N_FFT/2;
    f.data.assign(filters_data, filters_data + (size_t) n_mel*f.n_fft);
    whisper_mel mel;
    if (!log_mel_spectrogram(st, samples, n_samples, WHISPER_SAMPLE_RATE, WHISPER_N_FFT, WHISPER_HOP_LENGTH,
                             n_mel, n_threads, f, false, mel)) {
        return -1;
    }
    *n_len = mel.n_len;
    *n_len_org = mel.n_len_org;
    if ((size_t) out_cap < mel.data.size()) return -2;
    memcpy(out, mel.data.data(), mel.data.size()*sizeof(float));
    return 0;
}

// encoder output after whisper_encode: [n_audio_ctx][n_audio_state] f32
// (reference: whisper_state::embd_enc, src/whisper.cpp:2241-2251)
__attribute__((visibility("default")))
int ref_embd_enc_copy(struct whisper_context * ctx, float * out, int n_floats) {
    if (!ctx || !ctx->state || !ctx->state->embd_enc) return -1;
    ggml_tensor * t = ctx->state->embd_enc;
    if ((int64_t) n_floats != ggml_nelements(t)) return -2;
    ggml_backend_tensor_get(t, out, 0, ggml_nbytes(t));
    return 0;
}

// cross K/V cache (F16 in the reference) converted to f32.  Layout of the reference buffer:
// flash_attn: K,V = [n_text_layer][n_ctx_pad][n_state]; otherwise K = [layer][n_ctx][n_state],
// V = [layer][n_state][n_ctx] (transposed) (src/whisper.cpp:2300-2339).  Raw copy; caller interprets.
__attribute__((visibility("default")))
int64_t ref_kv_cross_copy(struct whisper_context * ctx, int which /*0=k 1=v*/, float * out, int64_t cap) {
    if (!ctx || !ctx->state) return -1;
    ggml_tensor * t = which == 0 ? ctx->state->kv_cross.k : ctx->state->kv_cross.v;
    const int64_t n = ggml_nelements(t);
    if (!out) return n;
    if (cap < n) return -2;
    if (t->type == GGML_TYPE_F16) {
        std::vector<ggml_fp16_t> tmp(n);
        ggml_backend_tensor_get(t, tmp.data(), 0, ggml_nbytes(t));
        ggml_fp16_to_fp32_row(tmp.data(), out, n);
    } else {
        ggml_backend_tensor_get(t, out, 0, ggml_nbytes(t));
    }
    return n;
}

// The reference's token-level timestamp heuristic and max_len re-wrapping on ONE externally supplied segment
// (whisper_exp_compute_token_level_timestamps / whisper_wrap_segment / get_signal_energy, src/whisper.cpp:8425-8660, 6077-6130).
// tok_state = {t_beg, t_last, tid_last} in / out.  Same contract as whisper_b200_token_timestamps of the product.
__attribute__((visibility("default")))
int ref_token_timestamps(struct whisper_context * ctx, const float * pcm, int n_samples, long long seg_t0, long long seg_t1,
                         whisper_token_data * tokens, int n_tokens, float thold_pt, float thold_ptsum, long long * tok_state,
                         int max_len, int split_on_word, long long * seg_t, int * seg_ntok, int seg_cap) {
    if (!ctx || !ctx->state || !tokens || !tok_state) return -1;
    whisper_state & st = *ctx->state;
    st.energy   = get_signal_energy(pcm, n_samples, 32);
    st.t_beg    = tok_state[0];
    st.t_last   = tok_state[1];
    st.tid_last = (whisper_token) tok_state[2];
    st.result_all.clear();
    st.result_all.push_back({ (int64_t) seg_t0, (int64_t) seg_t1, "", 0.0f, {}, false });
    st.result_all.back().tokens.assign(tokens, tokens + n_tokens);
    whisper_exp_compute_token_level_timestamps(*ctx, st, 0, thold_pt, thold_ptsum);
    int n_seg = 1;
    if (max_len > 0) n_seg = whisper_wrap_segment(*ctx, st, max_len, split_on_word != 0);
    tok_state[0] = st.t_beg;
    tok_state[1] = st.t_last;
    tok_state[2] = st.tid_last;
    int k = 0;
    for (size_t i = 0; i < st.result_all.size(); ++i) {
        const auto & seg = st.result_all[i];
        if ((int) i < seg_cap && seg_t && seg_ntok) {
            seg_t[2*i] = seg.t0; seg_t[2*i + 1] = seg.t1; seg_ntok[i] = (int) seg.tokens.size();
        }
        for (const auto & t : seg.tokens) if (k < n_tokens) tokens[k++] = t;
    }
    return n_seg;
}

// Run the reference's own logit rules + greedy sampler on an externally supplied logits row.
// Used to pin the oracle restatement of whisper_process_logits / whisper_sample_token
// (src/whisper.cpp:6177-6445, 6460-6517).  `hist` are the tokens sampled so far.
__attribute__((visibility("default")))
int ref_process_logits(struct whisper_context * ctx, struct whisper_full_params params, float temperature,
                       const float * logits_row, const whisper_token * hist, int n_hist,
                       int has_ts, int seek_delta,
                       float * logits_out, float * logprobs_out, float * probs_out, whisper_token_data * tok_out) {
    if (!ctx || !ctx->state) return -1;
    auto & state = *ctx->state;
    const int n_vocab = ctx->vocab.n_vocab;
    state.logits.assign(logits_row, logits_row + n_vocab);
    whisper_decoder & dec = state.decoders[0];
    dec.i_batch = 0;
    dec.has_ts = has_ts != 0;
    dec.seek_delta = seek_delta;
    dec.sequence.tokens.clear();
    for (int i = 0; i < n_hist; ++i) {
        whisper_token_data td = {};
        td.id = hist[i];
        dec.sequence.tokens.push_back(td);
    }
    whisper_process_logits(*ctx, state, dec, params, temperature);
    if (logits_out)   memcpy(logits_out,   dec.logits.data(),   n_vocab*sizeof(float));
    if (logprobs_out) memcpy(logprobs_out, dec.logprobs.data(), n_vocab*sizeof(float));
    if (probs_out)    memcpy(probs_out,    dec.probs.data(),    n_vocab*sizeof(float));
    if (tok_out)      *tok_out = whisper_sample_token(*ctx, dec, true);
    return 0;
}

// k draws of whisper_sample_token_topk ("beam search" / best_of sampling, src/whisper.cpp:6519-6592) from the distribution left
// in decoder 0 by the last ref_process_logits call, with the decoder's mt19937 re-seeded first; and whisper_sequence_score
// (6595-6641) of a sequence given by its token log-probabilities.
__attribute__((visibility("default")))
int ref_sample_topk(struct whisper_context * ctx, int k, unsigned seed, whisper_token_data * out) {
    if (!ctx || !ctx->state || !out || k <= 0) return -1;
    whisper_decoder & dec = ctx->state->decoders[0];
    if ((int) dec.probs.size() != ctx->vocab.n_vocab) return -2;
    dec.rng = std::mt19937(seed);
    const auto toks = whisper_sample_token_topk(*ctx, dec, k);
    for (int i = 0; i < k; ++i) out[i] = toks[i];
    return 0;
}

__attribute__((visibility("default")))
int ref_sequence_score(struct whisper_full_params params, const float * plog, const whisper_token * ids, int n, int result_len,
                       double * out /* sum_logprobs, avg_logprobs, entropy, score */) {
    whisper_sequence seq = {};
    for (int i = 0; i < n; ++i) {
        whisper_token_data td = {};
        td.id = ids[i];
        td.plog = plog[i];
        seq.tokens.push_back(td);
    }
    seq.result_len = result_len;
    whisper_sequence_score(params, seq);
    out[0] = seq.sum_logprobs; out[1] = seq.avg_logprobs; out[2] = seq.entropy; out[3] = seq.score;
    return 0;
}

__attribute__((visibility("default")))
float ref_no_speech_prob(struct whisper_context * ctx) {
    return ctx && ctx->state ? ctx->state->no_speech_prob : -1.0f;
}

// ---- VAD ------------------------------------------------------------------------------------------------------------------
// whisper_vad_segments_from_probs (src/whisper.cpp:5209-5420) on explicit probabilities: it reads only probs and n_window of
// the context, so a bare context carries them.  seg_out[2i], [2i+1] = start, end (centiseconds).
__attribute__((visibility("default")))
int ref_vad_segments_from_probs(const float * probs, int n_probs, int n_window, struct whisper_vad_params params,
                                long long * seg_out, int cap) {
    whisper_vad_context v;
    v.n_window = n_window;
    v.probs.assign(probs, probs + n_probs);
    whisper_vad_segments * s = whisper_vad_segments_from_probs(&v, params);
    if (!s) return -1;
    const int n = (int) s->data.size();
    for (int i = 0; i < n && i < cap; ++i) {
        seg_out[2*i]     = s->data[i].start;
        seg_out[2*i + 1] = s->data[i].end;
    }
    whisper_vad_free_segments(s);
    return n;
}

// the audio filter of whisper_full (whisper_vad, src/whisper.cpp:6643-6825): filtered samples and the time mapping table
__attribute__((visibility("default")))
int ref_vad_filter(struct whisper_context * ctx, struct whisper_full_params params, const float * samples, int n_samples,
                   float * out, int cap, long long * table, int cap_pairs, int * n_pairs) {
    if (!ctx || !ctx->state) return -1;
    std::vector<float> filtered;
    if (!whisper_vad(ctx, ctx->state, params, samples, n_samples, filtered)) return -2;
    const auto & tab = ctx->state->vad_mapping_table;
    *n_pairs = (int) tab.size();
    for (int i = 0; i < (int) tab.size() && i < cap_pairs; ++i) {
        table[2*i]     = tab[i].processed_time;
        table[2*i + 1] = tab[i].original_time;
    }
    for (int i = 0; i < (int) filtered.size() && i < cap; ++i) out[i] = filtered[i];
    return (int) filtered.size();
}

// map_processed_to_original_time (src/whisper.cpp:7947-7989) on an explicit table
__attribute__((visibility("default")))
long long ref_vad_map_time(const long long * table, int n_pairs, long long t) {
    std::vector<vad_time_mapping> tab(n_pairs);
    for (int i = 0; i < n_pairs; ++i) tab[i] = {table[2*i], table[2*i + 1]};
    return map_processed_to_original_time(t, tab);
}

} // extern "C"
