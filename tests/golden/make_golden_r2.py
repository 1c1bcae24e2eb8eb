"""Round-2 golden fixtures: the BASELINE configurations and whisper_full features that had no parity test.

Generated from the UNMODIFIED reference (oracle/_ref, built by oracle/Makefile.ref from /root/reference); run in the build
container (the GPU box has no reference sources and no time for a large-v3 CPU run):

    python tests/golden/make_golden_r2.py [--cases a,b,...] [--variant v4|v3]

Output: tests/golden/golden_r2.json (committed).  Per case: the call parameters, the reference's segments
[t0, t1, [token ids], [p], [plog]] and, for the greedy no_timestamps cases, the reference's own top-2 logit gap and runner-up
at every step (teacher-forced through whisper_decode + the harness hook around whisper_process_logits), which is what tells a
parity test where the reference's decision margin is below the logits tolerance of the spec.
With --variant v3 the same cases run on the reference's AVX2 build and are stored under "<case>@v3" -- the two builds of the
reference itself are the noise floor of any token comparison on random-init weights.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from open_whisper_kit_b200 import api, capi, modelgen  # noqa: E402
from oracle import reflib  # noqa: E402

FP = C.POINTER(C.c_float)

from cases_r2 import CASES, model_path, params_for, pcm_for  # noqa: E402,F401


def run_case(ref, case):
    w = api.Whisper(ref, model_path(case["arch"]), use_gpu=False, flash_attn=False)
    keep = []
    nproc = case["n_processors"]
    case = dict(case, n_threads=max(1, 8 // nproc))
    p = params_for(w, case, keep)
    t = time.time()
    rc, segs = w.full(p, pcm_for(case), n_processors=nproc)
    out = dict(case)
    out["rc"] = rc
    out["segments"] = [[int(s.t0), int(s.t1), [int(x) for x in s.tokens], [round(float(t.p), 6) for t in s.token_data],
                        [round(float(t.plog), 5) for t in s.token_data]] for s in segs]
    out["reference_cpu_seconds"] = round(time.time() - t, 1)
    if case.get("gaps"):
        out["steps"] = run_gaps(ref, w, case, out["segments"])
    w.close()
    return out


def run_gaps(ref, w, case, segments):
    """Top-2 gap / runner-up of the reference's processed logits at every greedy step (one window per chunk)."""
    keep = []
    p = params_for(w, dict(case, n_threads=8), keep)
    n_vocab = ref.whisper_n_vocab(w.ctx)
    sot, tnot = ref.whisper_token_sot(w.ctx), ref.whisper_token_not(w.ctx)
    prompt = [sot]
    if ref.whisper_is_multilingual(w.ctx):
        prompt += [ref.whisper_token_lang(w.ctx, 0), ref.whisper_token_transcribe(w.ctx)]
    prompt.append(tnot)
    out = []
    lo = np.empty(n_vocab, np.float32)
    for wi, seg in enumerate(segments):
        toks = seg[2]
        pcm = modelgen.synth_pcm(480000, seed=case["seed"], stream=wi)
        assert w.pcm_to_mel(pcm, 8) == 0 and w.encode(0, 8) == 0
        rc, lg = w.decode(prompt, 0, 8)
        gaps, runner = [], []
        for k, tok in enumerate(toks):
            hist = (C.c_int32 * max(1, k))(*toks[:k])
            td = capi.whisper_token_data()
            rc = ref.ref_process_logits(w.ctx, p, 0.0, lg.ctypes.data_as(FP), hist, k, 0, 3000, lo.ctypes.data_as(FP), None, None,
                                        C.byref(td))
            assert rc == 0 and td.id == tok, (wi, k, td.id, tok)
            top2 = np.argpartition(-lo, 2)[:2]
            top2 = top2[np.argsort(-lo[top2])]
            gaps.append(round(float(lo[top2[0]] - lo[top2[1]]), 5))
            runner.append(int(top2[1]))
            if k + 1 < len(toks):
                rc, lg = w.decode([tok], len(prompt) + k, 8)
                assert rc == 0
        out.append({"gaps": gaps, "runner_up": runner})
        print(f"   window {wi}: min gap {min(gaps):.5f}", flush=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default=",".join(CASES))
    ap.add_argument("--variant", default=None, help="v3: run on the reference's AVX2 build and store as <case>@v3")
    args = ap.parse_args()
    if args.variant:
        lib = capi.load_library(os.path.join(ROOT, "oracle", "_ref", f"libwhisper_ref_{args.variant}.so"))
        capi.bind(lib, reflib.REF_PROTOTYPES)
        cb = capi.LOG_CB(lambda level, text, ud: None)
        lib._quiet_cb = cb
        lib.whisper_log_set(C.cast(cb, C.c_void_p), None)
        ref, variant = lib, args.variant
    else:
        ref, variant = reflib.load()
    assert ref is not None, "build oracle/_ref first (make -f oracle/Makefile.ref)"
    path = os.path.join(HERE, "golden_r2.json")
    golden = json.load(open(path)) if os.path.exists(path) else {}
    for name in args.cases.split(","):
        case = dict(CASES[name])
        if args.variant:
            case.pop("gaps", None)
        print("running", name, "on build", variant, flush=True)
        res = run_case(ref, case)
        res["reference_build"] = variant
        golden[name + ("@" + args.variant if args.variant else "")] = res
        n_tok = sum(len(s[2]) for s in res["segments"])
        print("   rc", res["rc"], "segments", len(res["segments"]), "tokens", n_tok, "cpu s", res["reference_cpu_seconds"], flush=True)
        json.dump(golden, open(path, "w"), indent=0, sort_keys=True)


if __name__ == "__main__":
    main()
