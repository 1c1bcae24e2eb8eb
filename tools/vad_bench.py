"""VAD throughput: whisper_vad_detect_speech on long audio, product (GPU) vs the compiled reference (host cores).
Prints one JSON line.  Usage: python tools/vad_bench.py [minutes_gpu] [seconds_ref]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import open_whisper_kit_b200 as pkg  # noqa: E402
import vad_cases  # noqa: E402
from open_whisper_kit_b200 import api  # noqa: E402
from oracle import reflib  # noqa: E402

MODEL = os.path.join(ROOT, "tests", "golden", "silero-v6.2.0-ggml.bin")


def main():
    minutes = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    ref_s = float(sys.argv[2]) if len(sys.argv) > 2 else 120.0
    base = vad_cases.synthetic_bursts()
    pcm = np.tile(base, int(np.ceil(minutes * 60 * 16000 / len(base))))[: int(minutes * 60 * 16000)]
    lib = pkg.load()
    out = {"audio_s": len(pcm) / 16000}
    with api.Vad(lib, MODEL) as v:
        v.detect(pcm[:16000 * 30])
        ts = []
        for _ in range(3):
            t0 = time.time()
            probs = v.detect(pcm)
            ts.append(time.time() - t0)
        out["b200"] = {"s": min(ts), "audio_s_per_s": out["audio_s"] / min(ts), "chunks": len(probs),
                       "us_per_chunk": 1e6 * min(ts) / len(probs), "segments": len(v.segments_from_probs())}
    ref, variant = reflib.load()
    if ref is not None:
        sub = pcm[: int(ref_s * 16000)]
        with api.Vad(ref, MODEL) as v:
            t0 = time.time()
            rp = v.detect(sub)
            dt = time.time() - t0
        out["reference"] = {"s": dt, "audio_s": len(sub) / 16000, "audio_s_per_s": len(sub) / 16000 / dt, "threads": 4,
                            "variant": variant, "max_abs_prob_diff_on_sample": float(np.abs(rp - probs[: len(rp)]).max())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
