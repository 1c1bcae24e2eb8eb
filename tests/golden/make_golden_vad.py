"""Generates tests/golden/golden_vad.json from the live reference library (oracle/_ref, built from /root/reference):
speech probabilities of the Silero VAD on tests/golden/jfk.wav and on a synthetic multi-burst signal, the segments the
reference derives from them for several parameter sets, and the audio filter's time table.  Run in the build container:

    python tests/golden/make_golden_vad.py
"""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from open_whisper_kit_b200 import capi  # noqa: E402
from oracle import reflib  # noqa: E402
sys.path.insert(0, HERE)
from vad_cases import PARAM_SETS, read_wav, synthetic_bursts, vad_params  # noqa: E402

VAD_MODEL = os.path.join(HERE, "silero-v6.2.0-ggml.bin")


def main():
    lib, variant = reflib.load()
    assert lib is not None, "build oracle/_ref first (make -f oracle/Makefile.ref)"
    cp = lib.whisper_vad_default_context_params()
    vctx = lib.whisper_vad_init_from_file_with_params(VAD_MODEL.encode(), cp)
    assert vctx
    out = {"variant": variant, "cases": {}}
    signals = {"jfk": read_wav(os.path.join(HERE, "jfk.wav")), "bursts": synthetic_bursts()}
    for name, pcm in signals.items():
        pcm = np.ascontiguousarray(pcm, dtype=np.float32)
        assert lib.whisper_vad_detect_speech(vctx, pcm.ctypes.data_as(C.POINTER(C.c_float)), len(pcm))
        n = lib.whisper_vad_n_probs(vctx)
        probs = np.ctypeslib.as_array(lib.whisper_vad_probs(vctx), shape=(n,)).copy()
        case = {"n_samples": int(len(pcm)), "probs": [float(x) for x in probs], "segments": {}}
        for pname, kw in PARAM_SETS.items():
            segs = lib.whisper_vad_segments_from_probs(vctx, vad_params(lib, **kw))
            k = lib.whisper_vad_segments_n_segments(segs)
            case["segments"][pname] = [[lib.whisper_vad_segments_get_segment_t0(segs, i), lib.whisper_vad_segments_get_segment_t1(segs, i)]
                                       for i in range(k)]
            lib.whisper_vad_free_segments(segs)
        out["cases"][name] = case
        print(name, n, {k: len(v) for k, v in case["segments"].items()})
    lib.whisper_vad_free(vctx)
    with open(os.path.join(HERE, "golden_vad.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
