"""Development aid: stage-by-stage comparison of the CUDA path with the compiled reference (oracle/_ref) on a GPU box.

Prints statistics instead of asserting, so one gpurun call shows where a divergence starts.
"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import open_whisper_kit_b200 as pkg  # noqa: E402
from open_whisper_kit_b200 import api, capi, modelgen  # noqa: E402
from oracle import reflib  # noqa: E402

FP = C.POINTER(C.c_float)


def stats(name, a, b):
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    print(f"  {name}: max|d|={d.max():.4e} mean|d|={d.mean():.4e} ref_rms={np.sqrt((b.astype(np.float64)**2).mean()):.4f} "
          f"argmax_d={np.unravel_index(d.argmax(), d.shape)}", flush=True)


def main():
    arch = sys.argv[1] if len(sys.argv) > 1 else "tiny.en"
    ftype = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    lib = pkg.load()
    ref, variant = reflib.load()
    print("reference variant:", variant, "dtype env:", os.environ.get("WHISPER_B200_DTYPE", "f16"))
    os.makedirs("/tmp/models", exist_ok=True)
    path = f"/tmp/models/{arch}-{ftype}.bin"
    if not os.path.exists(path):
        modelgen.write_model(path, arch, ftype=ftype)
    pcm = api.read_wav_f32(os.path.join(ROOT, "tests", "golden", "jfk.wav"))
    n_state = modelgen.ARCHS[arch][2]
    for fa in (False, True):
        print(f"== {arch} ftype={ftype} flash_attn={fa}")
        w = api.Whisper(lib, path, flash_attn=fa)
        r = api.Whisper(ref, path, use_gpu=False, flash_attn=fa)
        # mel
        assert w.pcm_to_mel(pcm) == 0 and r.pcm_to_mel(pcm, 4) == 0
        n_len, n_mel = C.c_int(), C.c_int()
        lib.whisper_b200_get_mel(w.ctx, None, None, 0, C.byref(n_len), C.byref(n_mel))
        mel = np.empty((n_mel.value, n_len.value), np.float32)
        assert lib.whisper_b200_get_mel(w.ctx, None, mel.ctypes.data_as(FP), mel.size, C.byref(n_len), C.byref(n_mel)) == 0
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        ref.ref_mel_dims(r.ctx, C.byref(a), C.byref(b), C.byref(c))
        rmel = np.empty((c.value, a.value), np.float32)
        ref.ref_mel_copy(r.ctx, rmel.ctypes.data_as(FP))
        print("  mel shapes", mel.shape, rmel.shape, "n_len_org", lib.whisper_n_len(w.ctx), b.value)
        stats("mel", mel, rmel)
        # encoder
        t = time.time(); assert w.encode(0) == 0; t1 = time.time() - t
        t = time.time(); assert r.encode(0, 8) == 0; t2 = time.time() - t
        enc = np.empty((1500, n_state), np.float32)
        assert lib.whisper_b200_get_encoder_output(w.ctx, enc.ctypes.data_as(FP), enc.size) == 0
        renc = np.empty((1500, n_state), np.float32)
        assert ref.ref_embd_enc_copy(r.ctx, renc.ctypes.data_as(FP), renc.size) == 0
        print(f"  encode: cuda {t1*1e3:.1f} ms, reference cpu {t2*1e3:.1f} ms")
        stats("embd_enc", enc, renc)
        # decoder logits: prompt pass then a few single-token steps
        sot = lib.whisper_token_sot(w.ctx)
        prompt = [sot] if not lib.whisper_is_multilingual(w.ctx) else [sot, lib.whisper_token_lang(w.ctx, 0), lib.whisper_token_transcribe(w.ctx)]
        rc1, lg = w.decode(prompt, 0)
        rc2, rlg = r.decode(prompt, 0, 8)
        assert rc1 == 0 and rc2 == 0
        stats("logits(prompt)", lg, rlg)
        srt = np.sort(rlg)[::-1]
        print(f"  ref top-2 gap {srt[0]-srt[1]:.4f}; argmax ours {lg.argmax()} ref {rlg.argmax()}")
        toks = list(prompt)
        for step in range(3):
            nxt = int(rlg.argmax())
            rc1, lg = w.decode([nxt], len(toks))
            rc2, rlg = r.decode([nxt], len(toks), 8)
            toks.append(nxt)
            stats(f"logits(step {step})", lg, rlg)
        # greedy tokens
        for nots in (True, False):
            p = w.greedy_params(no_timestamps=nots)
            t = time.time(); rc, segs = w.full(p, pcm); t1 = time.time() - t
            mine = [(s.t0, s.t1, s.tokens) for s in segs]
            p2 = r.greedy_params(no_timestamps=nots, n_threads=8)
            t = time.time(); rc2, segs2 = r.full(p2, pcm); t2 = time.time() - t
            theirs = [(s.t0, s.t1, s.tokens) for s in segs2]
            ta = [x for s in mine for x in s[2]]
            tb = [x for s in theirs for x in s[2]]
            n_same = 0
            for x, y in zip(ta, tb):
                if x != y:
                    break
                n_same += 1
            print(f"  full(no_timestamps={nots}): rc {rc}/{rc2} segs {len(mine)}/{len(theirs)} tokens {len(ta)}/{len(tb)} "
                  f"identical={ta == tb} common_prefix={n_same} seg_times_equal={[(a[0], a[1]) for a in mine] == [(a[0], a[1]) for a in theirs]} "
                  f"cuda {t1:.2f}s ref {t2:.2f}s")
            if ta != tb:
                print("   ours  :", ta[max(0, n_same - 3):n_same + 5])
                print("   theirs:", tb[max(0, n_same - 3):n_same + 5])
        w.close()
        r.close()


if __name__ == "__main__":
    main()
