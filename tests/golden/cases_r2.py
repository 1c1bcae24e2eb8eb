"""Round-2 parity cases (shared by tests/golden/make_golden_r2.py, which runs them on the reference, and by
tests/test_gpu_parity_r2.py, which runs them on the product): call parameters, synthetic PCM, model files."""
import os

import numpy as np

from open_whisper_kit_b200 import capi, modelgen

CASES = {
    # BASELINE config 3 (large-v3, 32 + 32 layers): greedy, two windows, 49 steps each
    "large-v3/synth2/nots48": dict(arch="large-v3", windows=2, seed=7, n_processors=2, strategy="greedy", no_timestamps=True,
                                   max_tokens=48, temperature_inc=0.0, gaps=True),
    # ... and one window in timestamp mode over the whole 220-token budget
    "large-v3/synth1/ts": dict(arch="large-v3", windows=1, seed=9, n_processors=1, strategy="greedy", no_timestamps=False,
                               max_tokens=0, temperature_inc=0.0),
    # BASELINE config 4 (large-v3-turbo, beam 5 + timestamps)
    "large-v3-turbo/synth2/beam5": dict(arch="large-v3-turbo", windows=2, seed=7, n_processors=2, strategy="beam", beam_size=5,
                                        no_timestamps=False, max_tokens=0, temperature_inc=0.0),
    "tiny/synth2/beam5": dict(arch="tiny", windows=2, seed=7, n_processors=2, strategy="beam", beam_size=5, no_timestamps=False,
                              max_tokens=0, temperature_inc=0.0),
    # temperature ladder: every temperature below the last is rejected by a logprob threshold no random model can meet
    "base.en/synth2/fallback": dict(arch="base.en", windows=2, seed=7, n_processors=2, strategy="greedy", best_of=3,
                                    no_timestamps=False, max_tokens=24, temperature_inc=0.4, logprob_thold=-0.5),
    # prompt carry-over across the windows of one stream (no_context = false)
    "tiny/synth3/carry": dict(arch="tiny", windows=3, seed=21, n_processors=1, strategy="greedy", no_timestamps=False,
                              max_tokens=40, temperature_inc=0.0, no_context=False),
    # initial prompt, offset / duration, translate, another language
    "tiny/synth2/prompt_offset_translate": dict(arch="tiny", windows=2, seed=33, n_processors=1, strategy="greedy",
                                                no_timestamps=False, max_tokens=32, temperature_inc=0.0,
                                                initial_prompt=" hello world", offset_ms=5000, duration_ms=40000, translate=True,
                                                language="de"),
}


def model_path(arch, cache="/tmp/models"):
    os.makedirs(cache, exist_ok=True)
    p = os.path.join(cache, f"{arch}-1.bin")
    if not os.path.exists(p):
        modelgen.write_model(p, arch, ftype=1)
    return p


def pcm_for(case):
    return np.concatenate([modelgen.synth_pcm(480000, seed=case["seed"], stream=i) for i in range(case["windows"])])


def params_for(w, case, keep):
    """whisper_full_params of a case; `keep` holds the byte strings the struct points into."""
    beam = case["strategy"] == "beam"
    p = w.default_params(capi.BEAM_SEARCH if beam else capi.GREEDY)
    if beam:
        p.beam_search.beam_size = case["beam_size"]
        p.greedy.best_of = case.get("best_of", 5)
    else:
        p.greedy.best_of = case.get("best_of", 1)
    p.temperature_inc = case["temperature_inc"]
    p.no_timestamps = case["no_timestamps"]
    p.max_tokens = case["max_tokens"]
    p.print_progress = False
    p.n_threads = case.get("n_threads", 8)
    if "logprob_thold" in case:
        p.logprob_thold = case["logprob_thold"]
    if "no_context" in case:
        p.no_context = case["no_context"]
    if "initial_prompt" in case:
        keep.append(case["initial_prompt"].encode())
        p.initial_prompt = keep[-1]
    p.offset_ms = case.get("offset_ms", 0)
    p.duration_ms = case.get("duration_ms", 0)
    p.translate = case.get("translate", False)
    keep.append(case.get("language", "en").encode())
    p.language = keep[-1]
    return p


