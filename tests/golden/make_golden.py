"""Generates the golden fixtures in this directory from the UNMODIFIED reference (oracle/_ref, built by
oracle/Makefile.ref from /root/reference).  Run in the build container (the reference sources are not present on
the GPU box):

    python tests/golden/make_golden.py [--cases tiny,base]

Outputs (committed):
  golden_tokens.json   greedy token sequences + segment times of the reference CPU path
                       (whisper_full / whisper_full_parallel, temperature fallback off, src/whisper.cpp:6827-7929)
  golden_tensors.npz   strided samples of the reference's mel, encoder output and prompt logits
                       (ref_mel_copy / ref_embd_enc_copy / whisper_get_logits on the same inputs)
The model files are NOT stored: they are regenerated bit-identically from the seed by
open_whisper_kit_b200.modelgen.write_model (numpy default_rng).
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
from open_whisper_kit_b200 import api, modelgen  # noqa: E402
from oracle import reflib  # noqa: E402

FP = C.POINTER(C.c_float)


def pcm_for(spec):
    if spec["kind"] == "jfk":
        return api.read_wav_f32(os.path.join(HERE, "jfk.wav"))
    return np.concatenate([modelgen.synth_pcm(480000, seed=spec["seed"], stream=i) for i in range(spec["windows"])])


TOKEN_CASES = {
    # config 1 of BASELINE.json: tiny.en on samples/jfk.wav, greedy (cli flags -bs 1 -bo 1 -nf)
    "tiny.en/jfk/ts/fa0": dict(arch="tiny.en", ftype=1, flash_attn=False, no_timestamps=False, n_processors=1, pcm=dict(kind="jfk")),
    "tiny.en/jfk/nots/fa0": dict(arch="tiny.en", ftype=1, flash_attn=False, no_timestamps=True, n_processors=1, pcm=dict(kind="jfk")),
    "tiny.en/jfk/ts/fa1": dict(arch="tiny.en", ftype=1, flash_attn=True, no_timestamps=False, n_processors=1, pcm=dict(kind="jfk")),
    "tiny.en/synth4/ts/fa0": dict(arch="tiny.en", ftype=1, flash_attn=False, no_timestamps=False, n_processors=4,
                                  pcm=dict(kind="synth", seed=7, windows=4)),
    "tiny/synth2/ts/fa0": dict(arch="tiny", ftype=1, flash_attn=False, no_timestamps=False, n_processors=2,
                               pcm=dict(kind="synth", seed=11, windows=2)),
    # config 2 of BASELINE.json: base.en, 16 x 30 s synthetic windows, greedy, both timestamp modes
    "base.en/synth16/ts/fa0": dict(arch="base.en", ftype=1, flash_attn=False, no_timestamps=False, n_processors=16,
                                   pcm=dict(kind="synth", seed=7, windows=16)),
    "base.en/synth16/nots/fa0": dict(arch="base.en", ftype=1, flash_attn=False, no_timestamps=True, n_processors=16,
                                     pcm=dict(kind="synth", seed=7, windows=16)),
}


def model_path(arch, ftype, cache="/tmp/models"):
    os.makedirs(cache, exist_ok=True)
    p = os.path.join(cache, f"{arch}-{ftype}.bin")
    if not os.path.exists(p):
        modelgen.write_model(p, arch, ftype=ftype)
    return p


def run_tokens(ref, case):
    w = api.Whisper(ref, model_path(case["arch"], case["ftype"]), use_gpu=False, flash_attn=case["flash_attn"])
    nproc = case["n_processors"]
    p = w.greedy_params(no_timestamps=case["no_timestamps"], n_threads=max(1, 8 // nproc))
    pcm = pcm_for(case["pcm"])
    t = time.time()
    rc, segs = w.full(p, pcm, n_processors=nproc)
    dt = time.time() - t
    out = dict(case)
    out["rc"] = rc
    out["segments"] = [[int(s.t0), int(s.t1), [int(x) for x in s.tokens]] for s in segs]
    out["reference_cpu_seconds"] = round(dt, 2)
    w.close()
    return out


def run_tensors(ref):
    out = {}
    pcm = api.read_wav_f32(os.path.join(HERE, "jfk.wav"))
    # (an all-F32 file, ftype=0, cannot be used: the reference's im2col asserts F16 conv kernels,
    #  ggml/src/ggml-cpu/ops.cpp:6130, so the F16 file with flash_attn=false is the golden mode)
    for arch, ftype, fa in (("tiny.en", 1, False), ("tiny.en", 1, True), ("tiny", 1, False)):
        key = f"{arch}/f{ftype}/fa{int(fa)}"
        w = api.Whisper(ref, model_path(arch, ftype), use_gpu=False, flash_attn=fa)
        assert w.pcm_to_mel(pcm, 4) == 0
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        ref.ref_mel_dims(w.ctx, C.byref(a), C.byref(b), C.byref(c))
        mel = np.empty((c.value, a.value), np.float32)
        ref.ref_mel_copy(w.ctx, mel.ctypes.data_as(FP))
        assert w.encode(0, 8) == 0
        d = modelgen.ARCHS[arch][2]
        enc = np.empty((1500, d), np.float32)
        assert ref.ref_embd_enc_copy(w.ctx, enc.ctypes.data_as(FP), enc.size) == 0
        sot = ref.whisper_token_sot(w.ctx)
        rc, lg = w.decode([sot], 0, 8)
        assert rc == 0
        out[key + "/mel_sub"] = mel[:, :1100:5].copy()
        out[key + "/mel_sum"] = np.array([mel.astype(np.float64).sum(), np.abs(mel.astype(np.float64)).sum()])
        out[key + "/enc_sub"] = enc[::25, ::3].copy()
        out[key + "/enc_sum"] = np.array([enc.astype(np.float64).sum(), np.abs(enc.astype(np.float64)).sum()])
        out[key + "/logits_sub"] = lg[::17].copy()
        top = np.argsort(-lg)[:16]
        out[key + "/logits_top_ids"] = top.astype(np.int32)
        out[key + "/logits_top_vals"] = lg[top].copy()
        w.close()
    return out


def run_gaps(ref, case):
    """Top-2 gap of the reference's processed logits at every greedy step (no_timestamps cases: one window per chunk).

    Random-init weights give Gaussian-extreme logits: over thousands of steps some top-2 gaps are far below any 16-bit
    implementation's logit error (the reference's own AVX2 and AVX-512 builds diverge on these inputs), so the parity
    test needs to know where the reference's choice was a near-tie.  Teacher-forced through the public API:
    whisper_decode for the raw logits, ref_process_logits (harness hook around whisper_process_logits) for the rules.
    """
    from open_whisper_kit_b200 import capi
    assert case["no_timestamps"]
    w = api.Whisper(ref, model_path(case["arch"], case["ftype"]), use_gpu=False, flash_attn=case["flash_attn"])
    p = w.greedy_params(no_timestamps=True, n_threads=8)
    n_vocab = ref.whisper_n_vocab(w.ctx)
    sot, tnot = ref.whisper_token_sot(w.ctx), ref.whisper_token_not(w.ctx)
    out = []
    lo = np.empty(n_vocab, np.float32)
    for wi, seg in enumerate(case["segments"]):
        toks = seg[2]
        if case["pcm"]["kind"] == "synth":
            pcm = modelgen.synth_pcm(480000, seed=case["pcm"]["seed"], stream=wi)
        else:
            pcm = pcm_for(case["pcm"])          # one chunk = the whole file
        assert w.pcm_to_mel(pcm, 8) == 0 and w.encode(0, 8) == 0
        rc, lg = w.decode([sot, tnot], 0, 8)
        gaps, runner = [], []
        for k, tok in enumerate(toks):
            hist = (C.c_int32 * max(1, k))(*toks[:k])
            td = capi.whisper_token_data()
            rc = ref.ref_process_logits(w.ctx, p, 0.0, lg.ctypes.data_as(FP), hist, k, 0, 3000, lo.ctypes.data_as(FP),
                                        None, None, C.byref(td))
            assert rc == 0 and td.id == tok, (wi, k, td.id, tok)
            top2 = np.argpartition(-lo, 2)[:2]
            top2 = top2[np.argsort(-lo[top2])]
            gaps.append(round(float(lo[top2[0]] - lo[top2[1]]), 5))
            runner.append(int(top2[1]))
            if k + 1 < len(toks):
                rc, lg = w.decode([tok], 2 + k, 8)
                assert rc == 0
        out.append({"gaps": gaps, "runner_up": runner})
        print(f"   window {wi}: min gap {min(gaps):.5f}", flush=True)
    w.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="tiny,base")
    ap.add_argument("--skip-tensors", action="store_true")
    ap.add_argument("--gaps", action="store_true", help="(re)compute the per-step top-2 gaps of the no_timestamps batch case")
    args = ap.parse_args()
    ref, variant = reflib.load()
    assert ref is not None, "build oracle/_ref first (make -f oracle/Makefile.ref)"
    path = os.path.join(HERE, "golden_tokens.json")
    golden = json.load(open(path)) if os.path.exists(path) else {}
    for name, case in TOKEN_CASES.items():
        if not any(name.startswith(c) for c in args.cases.split(",")) or (args.gaps and name in golden):
            continue
        print("running", name, flush=True)
        golden[name] = run_tokens(ref, case)
        n_tok = sum(len(s[2]) for s in golden[name]["segments"])
        print("   rc", golden[name]["rc"], "segments", len(golden[name]["segments"]), "tokens", n_tok,
              "cpu s", golden[name]["reference_cpu_seconds"], flush=True)
        json.dump(golden, open(path, "w"), indent=0, sort_keys=True)
    if args.gaps:
        for name in ("base.en/synth16/nots/fa0", "tiny.en/jfk/nots/fa0"):
            if "steps" in golden[name] or not any(name.startswith(c) for c in args.cases.split(",")):
                continue
            print("margins of", name, flush=True)
            golden[name]["steps"] = run_gaps(ref, golden[name])
            json.dump(golden, open(path, "w"), indent=0, sort_keys=True)
    if not args.skip_tensors:
        np.savez_compressed(os.path.join(HERE, "golden_tensors.npz"), **run_tensors(ref))
    print("reference variant", variant)


if __name__ == "__main__":
    main()
