"""Development aid: BASELINE.json config 4 shape on one GPU -- large-v3-turbo, beam 5 with timestamps, N windows."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import open_whisper_kit_b200 as pkg  # noqa: E402
from open_whisper_kit_b200 import capi, modelgen  # noqa: E402

n_win = int(sys.argv[1]) if len(sys.argv) > 1 else 15
lib = pkg.load()
os.makedirs("/tmp/models", exist_ok=True)
path = "/tmp/models/large-v3-turbo-1.bin"
if not os.path.exists(path):
    modelgen.write_model(path, "large-v3-turbo")
cb = capi.LOG_CB(lambda level, text, ud: None)
lib.whisper_log_set(C.cast(cb, C.c_void_p), None)
ctx = lib.whisper_init_from_file_with_params(path.encode(), lib.whisper_context_default_params())
assert ctx
pcm = np.concatenate([modelgen.synth_pcm(480000, seed=7, stream=i) for i in range(n_win)])
for name, strategy in (("greedy+timestamps", 0), ("beam5+timestamps", 1)):
    p = lib.whisper_full_default_params(strategy)
    p.print_progress = False
    p.temperature_inc = 0.0
    p.language = b"en"
    if strategy == 0:
        p.greedy.best_of = 1
    for rep in range(2):
        t = time.time()
        rc = lib.whisper_full_parallel(ctx, p, pcm.ctypes.data_as(C.POINTER(C.c_float)), len(pcm), n_win)
        dt = time.time() - t
        n_seg = lib.whisper_full_n_segments(ctx)
        n_tok = sum(lib.whisper_full_n_tokens(ctx, i) for i in range(n_seg))
        ids = [lib.whisper_full_get_token_id(ctx, i, j) for i in range(n_seg) for j in range(lib.whisper_full_n_tokens(ctx, i))]
        chk = hash(tuple(ids)) & 0xffffffff
        print(f"{name} rep {rep}: rc {rc} {dt*1e3:.1f} ms -> {30.0*n_win/dt:.1f}x real time; {n_seg} segments, {n_tok} tokens, checksum {chk:08x}", flush=True)
lib.whisper_free(ctx)
