// See grammar.h.  Each function cites the reference routine it restates (src/whisper.cpp).
#include "grammar.h"

#include <stdlib.h>

#include <utility>

namespace wb {

namespace {

struct Candidate {
    whisper_token id;
    const uint32_t * code_points;
    whisper_partial_utf8 partial_utf8;
};

using Stack = std::vector<const whisper_grammar_element *>;
using Rules = std::vector<std::vector<whisper_grammar_element>>;

// Decodes a UTF-8 string which may start / end with a partial sequence; always ends with a 0 code point
// (decode_utf8, 5498-5552)
std::pair<std::vector<uint32_t>, whisper_partial_utf8> decode_utf8(const char * src, whisper_partial_utf8 partial_start) {
    static const int lookup[] = {1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 2, 2, 3, 4};
    const char * pos = src;
    std::vector<uint32_t> code_points;
    uint32_t value = partial_start.value;
    int n_remain = partial_start.n_remain;
    while (*pos != 0 && n_remain > 0) {             // continue the previous token's sequence
        const uint8_t next_byte = (uint8_t) *pos;
        if ((next_byte >> 6) != 2) {
            code_points.push_back(0);
            return {std::move(code_points), whisper_partial_utf8{0, -1}};
        }
        value = (value << 6) + (next_byte & 0x3F);
        ++pos;
        --n_remain;
    }
    if (partial_start.n_remain > 0 && n_remain == 0) code_points.push_back(value);
    while (*pos != 0) {
        const uint8_t first_byte = (uint8_t) *pos;
        n_remain = lookup[first_byte >> 4] - 1;
        if (n_remain < 0) {
            code_points.clear();
            code_points.push_back(0);
            return {std::move(code_points), whisper_partial_utf8{0, n_remain}};
        }
        const uint8_t mask = (uint8_t) ((1 << (7 - n_remain)) - 1);
        value = first_byte & mask;
        ++pos;
        while (*pos != 0 && n_remain > 0) {
            value = (value << 6) + ((uint8_t) *pos & 0x3F);
            ++pos;
            --n_remain;
        }
        if (n_remain == 0) code_points.push_back(value);
    }
    code_points.push_back(0);
    return {std::move(code_points), whisper_partial_utf8{value, n_remain}};
}

bool is_end_of_sequence(const whisper_grammar_element * pos) {        // 5555-5561
    return pos->type == WHISPER_GRETYPE_END || pos->type == WHISPER_GRETYPE_ALT;
}

// does chr satisfy the char range at pos; second = the element after the range (5565-5588)
std::pair<bool, const whisper_grammar_element *> match_char(const whisper_grammar_element * pos, uint32_t chr) {
    bool found = false;
    const bool is_positive_char = pos->type == WHISPER_GRETYPE_CHAR;
    if (!is_positive_char && pos->type != WHISPER_GRETYPE_CHAR_NOT) abort();      // reference: WHISPER_ASSERT
    do {
        if (pos[1].type == WHISPER_GRETYPE_CHAR_RNG_UPPER) {
            found = found || (pos->value <= chr && chr <= pos[1].value);
            pos += 2;
        } else {
            found = found || pos->value == chr;
            pos += 1;
        }
    } while (pos->type == WHISPER_GRETYPE_CHAR_ALT);
    return {found == is_positive_char, pos};
}

// could some continuation of the partial sequence satisfy the char range at pos (5592-5636)
bool match_partial_char(const whisper_grammar_element * pos, whisper_partial_utf8 partial_utf8) {
    const bool is_positive_char = pos->type == WHISPER_GRETYPE_CHAR;
    if (!is_positive_char && pos->type != WHISPER_GRETYPE_CHAR_NOT) abort();
    const uint32_t partial_value = partial_utf8.value;
    const int n_remain = partial_utf8.n_remain;
    if (n_remain < 0 || (n_remain == 1 && partial_value < 2)) return false;
    uint32_t low = partial_value << (n_remain * 6);
    const uint32_t high = low | ((1 << (n_remain * 6)) - 1);
    if (low == 0) {
        if (n_remain == 2) low = 1 << 11;
        else if (n_remain == 3) low = 1 << 16;
    }
    do {
        if (pos[1].type == WHISPER_GRETYPE_CHAR_RNG_UPPER) {
            if (pos->value <= high && low <= pos[1].value) return is_positive_char;
            pos += 2;
        } else {
            if (low <= pos->value && pos->value <= high) return is_positive_char;
            pos += 1;
        }
    } while (pos->type == WHISPER_GRETYPE_CHAR_ALT);
    return !is_positive_char;
}

// one pushdown stack -> the stacks it can become, all ending at a character range (5641-5692)
void advance_stack(const Rules & rules, const Stack & stack, std::vector<Stack> & new_stacks) {
    if (stack.empty()) {
        new_stacks.emplace_back();
        return;
    }
    const whisper_grammar_element * pos = stack.back();
    switch (pos->type) {
        case WHISPER_GRETYPE_RULE_REF: {
            const size_t rule_id = (size_t) pos->value;
            const whisper_grammar_element * subpos = rules[rule_id].data();
            do {
                Stack new_stack(stack.begin(), stack.end() - 1);
                if (!is_end_of_sequence(pos + 1)) new_stack.push_back(pos + 1);
                if (!is_end_of_sequence(subpos)) new_stack.push_back(subpos);
                advance_stack(rules, new_stack, new_stacks);
                while (!is_end_of_sequence(subpos)) subpos++;
                if (subpos->type == WHISPER_GRETYPE_ALT) subpos++;
                else break;
            } while (true);
            break;
        }
        case WHISPER_GRETYPE_CHAR:
        case WHISPER_GRETYPE_CHAR_NOT:
            new_stacks.push_back(stack);
            break;
        default:
            abort();      // a stack is never left on END / ALT / CHAR_ALT / CHAR_RNG_UPPER (reference: WHISPER_ASSERT(false))
    }
}

// the stacks after accepting chr (5698-5724)
std::vector<Stack> accept(const Rules & rules, const std::vector<Stack> & stacks, uint32_t chr) {
    std::vector<Stack> new_stacks;
    for (const auto & stack : stacks) {
        if (stack.empty()) continue;
        const auto match = match_char(stack.back(), chr);
        if (match.first) {
            const whisper_grammar_element * pos = match.second;
            Stack new_stack(stack.begin(), stack.end() - 1);
            if (!is_end_of_sequence(pos)) new_stack.push_back(pos);
            advance_stack(rules, new_stack, new_stacks);
        }
    }
    return new_stacks;
}

std::vector<Candidate> reject_candidates(const Rules & rules, const std::vector<Stack> & stacks, const std::vector<Candidate> & candidates);

// 5731-5780
std::vector<Candidate> reject_candidates_for_stack(const Rules & rules, const Stack & stack, const std::vector<Candidate> & candidates) {
    std::vector<Candidate> rejects;
    if (stack.empty()) {
        for (const auto & tok : candidates)
            if (*tok.code_points != 0 || tok.partial_utf8.n_remain != 0) rejects.push_back(tok);
        return rejects;
    }
    const whisper_grammar_element * stack_pos = stack.back();
    std::vector<Candidate> next_candidates;
    for (const auto & tok : candidates) {
        if (*tok.code_points == 0) {
            // end of the token's full code points: reject iff it ended in a partial sequence this position cannot take
            if (tok.partial_utf8.n_remain != 0 && !match_partial_char(stack_pos, tok.partial_utf8)) rejects.push_back(tok);
        } else if (match_char(stack_pos, *tok.code_points).first) {
            next_candidates.push_back({tok.id, tok.code_points + 1, tok.partial_utf8});
        } else {
            rejects.push_back(tok);
        }
    }
    const whisper_grammar_element * stack_pos_after = match_char(stack_pos, 0).second;
    Stack stack_after(stack.begin(), stack.end() - 1);
    if (!is_end_of_sequence(stack_pos_after)) stack_after.push_back(stack_pos_after);
    std::vector<Stack> next_stacks;
    advance_stack(rules, stack_after, next_stacks);
    const auto next_rejects = reject_candidates(rules, next_stacks, next_candidates);
    for (const auto & tok : next_rejects) rejects.push_back({tok.id, tok.code_points - 1, tok.partial_utf8});
    return rejects;
}

// 5782-5796
std::vector<Candidate> reject_candidates(const Rules & rules, const std::vector<Stack> & stacks, const std::vector<Candidate> & candidates) {
    if (candidates.empty() || stacks.empty()) return {};
    auto rejects = reject_candidates_for_stack(rules, stacks.front(), candidates);
    for (size_t i = 1; i < stacks.size(); ++i) rejects = reject_candidates_for_stack(rules, stacks[i], rejects);
    return rejects;
}

}  // namespace

// whisper_grammar_init, 5798-5836
whisper_grammar grammar_init(const whisper_grammar_element ** rules, size_t n_rules, size_t i_start_rule) {
    const whisper_grammar_element * pos;
    Rules vec_rules(n_rules);
    for (size_t i = 0; i < n_rules; i++) {
        for (pos = rules[i]; pos->type != WHISPER_GRETYPE_END; pos++) vec_rules[i].push_back(*pos);
        vec_rules[i].push_back({WHISPER_GRETYPE_END, 0});
    }
    std::vector<Stack> stacks;
    pos = rules[i_start_rule];
    do {
        Stack stack;
        if (!is_end_of_sequence(pos)) stack.push_back(pos);
        advance_stack(vec_rules, stack, stacks);
        while (!is_end_of_sequence(pos)) pos++;
        if (pos->type == WHISPER_GRETYPE_ALT) pos++;
        else break;
    } while (true);
    whisper_grammar g;
    g.rules = std::move(vec_rules);
    g.stacks = std::move(stacks);
    return g;
}

// whisper_suppress_invalid_grammar, 5838-5880
void grammar_suppress_invalid(const std::vector<std::string> & id_to_token, int token_eot, float penalty, std::vector<float> & logits,
                              const whisper_grammar & grammar) {
    if (grammar.rules.empty() || grammar.stacks.empty()) return;
    std::vector<std::pair<std::vector<uint32_t>, whisper_partial_utf8>> decoded;
    decoded.reserve((size_t) token_eot);        // the candidates point into these vectors: no reallocation below
    std::vector<Candidate> candidates;
    for (whisper_token id = 0; id < token_eot; ++id) {
        const std::string & text = id_to_token[id];
        if (!text.empty()) {
            decoded.push_back(decode_utf8(text.c_str(), grammar.partial_utf8));
            candidates.push_back({id, decoded.back().first.data(), decoded.back().second});
        }
    }
    const auto rejects = reject_candidates(grammar.rules, grammar.stacks, candidates);
    for (const auto & reject : rejects) logits[reject.id] -= penalty;
}

// whisper_grammar_accept_token, 5882-5904
void grammar_accept_token(const std::vector<std::string> & id_to_token, whisper_grammar & grammar, whisper_token token) {
    if (grammar.rules.empty() || grammar.stacks.empty()) return;
    const std::string & text = id_to_token[token];
    if (text.rfind("[_", 0) == 0) return;
    const auto decoded = decode_utf8(text.c_str(), grammar.partial_utf8);
    const auto & code_points = decoded.first;
    for (auto it = code_points.begin(), end = code_points.end() - 1; it != end; ++it) grammar.stacks = accept(grammar.rules, grammar.stacks, *it);
    grammar.partial_utf8 = decoded.second;
}

}  // namespace wb

#include "whisper_b200.h"

extern "C" WB200_API int whisper_b200_grammar_match(const struct whisper_grammar_element ** rules, size_t n_rules, size_t i_start_rule,
                                                    const char * text) {
    if (!rules || !text || i_start_rule >= n_rules) return 0;
    whisper_grammar g = wb::grammar_init(rules, n_rules, i_start_rule);
    // the candidate check of the sampling path: is this text rejected at the current position?
    std::vector<std::string> vocab(1, std::string(text));
    std::vector<float> logits(1, 0.0f);
    wb::grammar_suppress_invalid(vocab, 1, 1.0f, logits, g);
    if (logits[0] != 0.0f) return 0;
    wb::grammar_accept_token(vocab, g, 0);
    for (const auto & stack : g.stacks)
        if (stack.empty()) return 1;
    return g.stacks.empty() ? 0 : 2;
}
