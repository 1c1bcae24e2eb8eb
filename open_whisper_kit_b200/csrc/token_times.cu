// [EXPERIMENTAL in the reference] token-level timestamps and max_len segment splitting (params.token_timestamps, thold_pt,
// thold_ptsum, max_len, split_on_word), host side.
//
// What the reference computes (src/whisper.cpp:8398-8660 for the times, 6045-6130 for the splitting) is a four-stage
// heuristic over one segment; this file implements those stages as separate passes over two plain arrays of token
// boundaries (begin[], end[]; -1 = not known yet) and writes them back to the whisper_token_data records at the end:
//
//   1. anchors   a token whose best timestamp candidate (tid, with probability share pt and timestamp mass ptsum above the
//                thresholds) moves time forward pins the boundary between itself and its predecessor;
//   2. spread    every run of tokens between two pinned boundaries shares the interval in proportion to a per-character
//                "spoken weight" of the token texts;
//   3. repair    boundaries are made monotone;
//   4. snap      each text token's edges slide to where the |PCM| envelope crosses half of its local mean.
//
// The arithmetic (float envelope sums in sample order, double interpolation truncated to 10 ms ticks) is kept exactly as the
// reference's so that the resulting t0 / t1 / vlen are identical (tests/test_token_timestamps_host.py drives both).
#include "token_times.h"

#include <math.h>

#include <algorithm>
#include <atomic>
#include <thread>

namespace wb {

// mean |x| over a centred window of 2 * half_width + 1 samples, truncated at the ends of the signal but always divided by the
// full window length (src/whisper.cpp:8425-8442).  Samples are independent: blocks of 64 Ki go to the host threads.
void envelope_abs_mean(const float * pcm, int n_samples, int half_width, std::vector<float> & env) {
    env.assign((size_t) std::max(n_samples, 0), 0.0f);
    if (n_samples <= 0) return;
    const int block = 1 << 16;
    const int n_blocks = (n_samples + block - 1) / block;
    const float width = (float) (2 * half_width + 1);
    auto do_block = [&](int b) {
        const int stop = std::min(n_samples, (b + 1) * block);
        for (int i = b * block; i < stop; ++i) {
            const int lo = std::max(0, i - half_width), hi = std::min(n_samples - 1, i + half_width);
            float acc = 0;
            for (int k = lo; k <= hi; ++k) acc += fabs(pcm[k]);
            env[i] = acc / width;
        }
    };
    const int hw_threads = (int) std::thread::hardware_concurrency();
    const int n_thr = std::max(1, std::min(n_blocks, std::min(hw_threads > 0 ? hw_threads : 4, 32)));
    if (n_thr == 1) {
        for (int b = 0; b < n_blocks; ++b) do_block(b);
        return;
    }
    std::atomic<int> next{0};
    std::vector<std::thread> pool;
    auto worker = [&]() {
        for (int b = next.fetch_add(1); b < n_blocks; b = next.fetch_add(1)) do_block(b);
    };
    for (int t = 1; t < n_thr; ++t) pool.emplace_back(worker);
    worker();
    for (auto & th : pool) th.join();
}

namespace {

// how long a token takes to say, in arbitrary units: digits and sentence punctuation are slow, blanks are almost free
// (src/whisper.cpp:8400-8422)
float spoken_weight(const std::string & text) {
    static const struct Table {
        float w[256];
        Table() {
            for (float & x : w) x = 1.00f;
            w[(unsigned char) ' '] = 0.01f;
            w[(unsigned char) ','] = 2.00f;
            for (const char c : {'.', '!', '?'}) w[(unsigned char) c] = 3.00f;
            for (char c = '0'; c <= '9'; ++c) w[(unsigned char) c] = 3.00f;
        }
    } table;
    float total = 0.0f;
    for (const unsigned char c : text) total += table.w[c];
    return total;
}

// 10 ms ticks relative to the start of the segment <-> index into the 16 kHz envelope (src/whisper.cpp:8444-8454)
struct TickClock {
    int64_t origin;
    int n_samples;
    int sample_of(int64_t tick) const {
        const int s = (int) (((tick - origin) * WHISPER_SAMPLE_RATE) / 100);
        return std::max(0, std::min(n_samples - 1, s));
    }
    int64_t tick_of(int sample) const { return (100ll * sample) / WHISPER_SAMPLE_RATE + origin; }
};

struct Boundaries {
    std::vector<int64_t> begin, end;
};

// stage 2: tokens lo..hi share [begin[lo], end[hi]] in proportion to their spoken weights
void spread_run(const std::vector<whisper_token_data> & tokens, Boundaries & b, int lo, int hi) {
    double weight_sum = 0.0;
    for (int j = lo; j <= hi; ++j) weight_sum += tokens[j].vlen;
    const double span = (double) (b.end[hi] - b.begin[lo]);
    for (int j = lo; j < hi; ++j) {
        const double cut = b.begin[j] + span * tokens[j].vlen / weight_sum;
        b.end[j] = (int64_t) cut;
        b.begin[j + 1] = (int64_t) cut;
    }
}

// stage 4 for one token: the edge that sits in a loud stretch moves outwards to where the envelope falls under the threshold
// (but never across the neighbour), the edge that sits in a quiet stretch moves inwards to where it rises above it
void snap_to_envelope(const std::vector<float> & env, const TickClock & clk, Boundaries & b, int j, int n) {
    const int n_samples = clk.n_samples;
    const int reach = WHISPER_SAMPLE_RATE / 8;
    int first = clk.sample_of(b.begin[j]), last = clk.sample_of(b.end[j]);
    const int w0 = std::max(first - reach, 0), w1 = std::min(last + reach, n_samples);
    float level = 0.0f;
    for (int k = w0; k < w1; ++k) level += env[k];
    const float thold = 0.5 * level / (w1 - w0);

    int k = first;
    if (env[k] > thold && j > 0) {
        while (k > 0 && env[k] > thold) --k;
        b.begin[j] = clk.tick_of(k);
        if (b.begin[j] < b.end[j - 1]) b.begin[j] = b.end[j - 1];
        else first = k;
    } else {
        while (env[k] < thold && k < last) ++k;
        first = k;
        b.begin[j] = clk.tick_of(k);
    }
    k = last;
    if (env[k] > thold) {
        while (k < n_samples - 1 && env[k] > thold) ++k;
        b.end[j] = clk.tick_of(k);
        if (j < n - 1 && b.end[j] > b.begin[j + 1]) b.end[j] = b.begin[j + 1];
    } else {
        while (env[k] < thold && k > first) --k;
        b.end[j] = clk.tick_of(k);
    }
}

int count_utf8_chars(const std::string & s) {             // code points, not bytes (src/whisper.cpp:6052-6062)
    int n = 0;
    for (const unsigned char c : s) n += (c & 0xC0) != 0x80;
    return n;
}

}  // namespace

void assign_token_times(const Vocab & vocab, whisper_state & state, int i_segment, float thold_pt, float thold_ptsum) {
    whisper_segment & seg = state.result_all[i_segment];
    std::vector<whisper_token_data> & tokens = seg.tokens;
    const std::vector<float> & env = state.energy;
    if (env.empty()) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: no signal data available\n", __func__);
        return;
    }
    const int n = (int) tokens.size();
    if (n == 0) return;
    if (n == 1) {
        tokens[0].t0 = seg.t0;
        tokens[0].t1 = seg.t1;
        return;
    }
    Boundaries b;
    b.begin.resize(n);
    b.end.resize(n);
    for (int j = 0; j < n; ++j) {
        b.begin[j] = tokens[j].t0;
        b.end[j] = tokens[j].t1;
    }

    // stage 1: anchors.  The clock {t_beg, t_last, tid_last} runs on across the segments of one whisper_full call.
    if (tokens[0].id == vocab.token_beg) {
        b.begin[0] = b.end[0] = b.begin[1] = seg.t0;
        state.t_beg = state.t_last = seg.t0;
        state.tid_last = vocab.token_beg;
    } else {
        b.begin[0] = state.t_last;
    }
    for (int j = 0; j < n; ++j) {
        whisper_token_data & tk = tokens[j];
        tk.vlen = spoken_weight(vocab.id_to_token[tk.id]);
        const int64_t candidate = state.t_beg + 2 * (tk.tid - vocab.token_beg);
        const bool trusted = tk.pt > thold_pt && tk.ptsum > thold_ptsum;
        if (trusted && tk.tid > state.tid_last && candidate <= seg.t1) {
            if (j > 0) b.end[j - 1] = candidate;
            b.begin[j] = candidate;
            state.tid_last = tk.tid;
        }
    }
    b.end[n - 2] = seg.t1;
    b.begin[n - 1] = b.end[n - 1] = seg.t1;
    state.t_last = seg.t1;

    // stage 2: spread the unpinned runs
    for (int lo = 0; lo < n;) {
        int hi = lo;
        while (hi < n && b.end[hi] < 0) ++hi;
        if (hi == n) hi = n - 1;
        if (hi > lo) spread_run(tokens, b, lo, hi);
        lo = hi + 1;
    }

    // stage 3: monotone boundaries
    for (int j = 0; j + 1 < n; ++j) {
        if (b.end[j] < 0) b.begin[j + 1] = b.end[j];
        if (j > 0 && b.end[j - 1] > b.begin[j]) {
            b.begin[j] = b.end[j - 1];
            b.end[j] = std::max(b.begin[j], b.end[j]);
        }
    }

    // stage 4: text tokens only (special tokens keep their boundaries)
    const TickClock clk = {seg.t0, (int) env.size()};
    for (int j = 0; j < n; ++j)
        if (tokens[j].id < vocab.token_eot) snap_to_envelope(env, clk, b, j, n);

    for (int j = 0; j < n; ++j) {
        tokens[j].t0 = b.begin[j];
        tokens[j].t1 = b.end[j];
    }
}

// Splits the most recent segment into pieces of at most max_len characters (src/whisper.cpp:6077-6130).  First the cut
// positions are found on the token list (greedy: a piece is closed in front of the text token that would overflow it, never in
// front of its own first token, and with split_on_word only in front of a token that starts a word), then the pieces are
// materialised: a piece runs from its first token's t0 to the next piece's; only the last keeps speaker_turn_next.
int split_last_segment(const Vocab & vocab, whisper_state & state, int max_len, bool split_on_word) {
    const whisper_segment whole = state.result_all.back();
    const int n = (int) whole.tokens.size();
    std::vector<int> starts(1, 0);
    {
        int used = 0;
        for (int i = 0; i < n; ++i) {
            const whisper_token id = whole.tokens[i].id;
            if (id >= vocab.token_eot) continue;
            const std::string & txt = vocab.id_to_token[id];
            const int len = count_utf8_chars(txt);
            const bool may_cut_here = !split_on_word || (!txt.empty() && txt[0] == ' ');
            if (used + len > max_len && i > starts.back() && may_cut_here) {
                starts.push_back(i);
                used = 0;
            }
            used += len;
        }
    }
    const int n_pieces = (int) starts.size();
    state.result_all.pop_back();
    for (int k = 0; k < n_pieces; ++k) {
        const int lo = starts[k], hi = k + 1 < n_pieces ? starts[k + 1] : n;
        whisper_segment piece = {};
        piece.t0 = k == 0 ? whole.t0 : whole.tokens[lo].t0;
        piece.t1 = k + 1 < n_pieces ? whole.tokens[hi].t0 : whole.t1;
        piece.no_speech_prob = k == 0 ? whole.no_speech_prob : 0.0f;       // later pieces are value-initialised in the reference
        piece.speaker_turn_next = k + 1 < n_pieces ? false : whole.speaker_turn_next;
        piece.tokens.assign(whole.tokens.begin() + lo, whole.tokens.begin() + hi);
        for (const auto & tk : piece.tokens)
            if (tk.id < vocab.token_eot) piece.text += vocab.id_to_token[tk.id];
        state.result_all.push_back(std::move(piece));
    }
    return n_pieces;
}

}  // namespace wb
