// Grammar-constrained sampling (whisper_full_params::grammar_rules): host-side restatement of the reference's pushdown
// automaton over token texts (src/whisper.cpp:769-782 and 5485-5905).  Runs on the host sampling path only: a call with grammar
// rules never takes the device selection kernel.
#pragma once

#include <stdint.h>

#include <string>
#include <vector>

#include "whisper.h"

struct whisper_partial_utf8 {
    uint32_t value = 0;   // bit value so far (unshifted)
    int n_remain = 0;     // bytes remaining; -1: invalid sequence
};

struct whisper_grammar {
    std::vector<std::vector<whisper_grammar_element>> rules;
    std::vector<std::vector<const whisper_grammar_element *>> stacks;
    whisper_partial_utf8 partial_utf8;      // left over from the previous token
};

namespace wb {

whisper_grammar grammar_init(const whisper_grammar_element ** rules, size_t n_rules, size_t i_start_rule);

// logits[id] -= penalty for every text token (id < eot) whose text cannot continue the grammar
void grammar_suppress_invalid(const std::vector<std::string> & id_to_token, int token_eot, float penalty, std::vector<float> & logits,
                              const whisper_grammar & grammar);

// advance the automaton over the text of an accepted token ("[_...]" specials are skipped)
void grammar_accept_token(const std::vector<std::string> & id_to_token, whisper_grammar & grammar, whisper_token token);

}  // namespace wb
