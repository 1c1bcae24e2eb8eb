"""Host logic of grammar-constrained sampling (csrc/grammar.cu <- reference src/whisper.cpp:5485-5905) through the host-only
hook whisper_b200_grammar_match: the pushdown automaton against regular expressions for the same languages.  No device needed."""
import ctypes as C
import random
import re

import regex

import pytest

import open_whisper_kit_b200 as pkg

END, ALT, RULE_REF, CHAR, CHAR_NOT, RNG_UPPER, CHAR_ALT = range(7)


class Element(C.Structure):
    _fields_ = [("type", C.c_int), ("value", C.c_uint32)]


def compile_rules(rules):
    arrays = [(Element * len(r))(*[Element(t, v) for t, v in r]) for r in rules]
    ptrs = (C.POINTER(Element) * len(arrays))(*[C.cast(a, C.POINTER(Element)) for a in arrays])
    return arrays, ptrs


def letters(rule_id):       # [a-z]+ as a right-recursive rule
    return [(CHAR, ord("a")), (RNG_UPPER, ord("z")), (RULE_REF, rule_id), (ALT, 0), (CHAR, ord("a")), (RNG_UPPER, ord("z")), (END, 0)]


GRAMMARS = {
    # root ::= word+ ; word ::= " " [a-z]+
    "words": ([[(RULE_REF, 1), (RULE_REF, 0), (ALT, 0), (RULE_REF, 1), (END, 0)], [(CHAR, ord(" ")), (RULE_REF, 2), (END, 0)], letters(2)],
              r"( [a-z]+)+", " abcxyz", r"( [a-z]+)*( [a-z]*)?"),
    # root ::= "yes" | "no" | "maybe" [!?]
    "choice": ([[(CHAR, ord("y")), (CHAR, ord("e")), (CHAR, ord("s")), (ALT, 0), (CHAR, ord("n")), (CHAR, ord("o")), (ALT, 0),
                 (CHAR, ord("m")), (CHAR, ord("a")), (CHAR, ord("y")), (CHAR, ord("b")), (CHAR, ord("e")),
                 (CHAR, ord("!")), (CHAR_ALT, ord("?")), (END, 0)]],
               r"yes|no|maybe[!?]", "yesnomaybe!?", None),
    # root ::= [^0-9]+ "é"   (negated range, multi-byte character)
    "negated": ([[(RULE_REF, 1), (CHAR, 0xE9), (END, 0)],
                 [(CHAR_NOT, ord("0")), (RNG_UPPER, ord("9")), (RULE_REF, 1), (ALT, 0), (CHAR_NOT, ord("0")), (RNG_UPPER, ord("9")), (END, 0)]],
                r"[^0-9]+é", "ab1é9 ", None),
}


def is_prefix(pattern, text):
    """Can `text` be extended to a sentence of the regular language?  (partial matching of the `regex` module)"""
    return regex.fullmatch(pattern, text, partial=True) is not None


@pytest.mark.parametrize("name", sorted(GRAMMARS))
def test_grammar_automaton_matches_regular_expression(name):
    lib = pkg.load()
    rules, pattern, alphabet, _ = GRAMMARS[name]
    keep, ptrs = compile_rules(rules)
    rng = random.Random(sum(map(ord, name)))
    n_complete = n_prefix = n_reject = 0
    samples = {"", " a", " ab cd", "yes", "no", "maybe", "maybe!", "maybe?", "maybes", "aé", "a1é", "é", "xyzé", "ab"}
    while len(samples) < 400:
        samples.add("".join(rng.choice(alphabet) for _ in range(rng.randint(1, 7))))
    for text in sorted(samples):
        if not text:
            continue
        got = lib.whisper_b200_grammar_match(C.cast(ptrs, C.c_void_p), len(rules), 0, text.encode("utf-8"))
        complete = re.fullmatch(pattern, text) is not None
        prefix = complete or is_prefix(pattern, text)
        if complete:
            assert got == 1 or (got == 2 and name == "words"), (text, got)      # " ab" is complete AND extensible: either stack
            n_complete += 1
        elif prefix:
            assert got == 2, (text, got)
            n_prefix += 1
        else:
            assert got == 0, (text, got)
            n_reject += 1
    assert n_complete >= 3 and n_prefix >= 3 and n_reject >= 20
    del keep
