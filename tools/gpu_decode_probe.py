"""Development aid: host/device time split of the batched whisper_full loop (WHISPER_B200_DEBUG_TIMING)."""
import ctypes as C
import os
import sys
import time

os.environ["WHISPER_B200_DEBUG_TIMING"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import open_whisper_kit_b200 as pkg  # noqa: E402
from open_whisper_kit_b200 import capi, modelgen  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "large-v3"
n_win = int(sys.argv[2]) if len(sys.argv) > 2 else 64
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
max_tokens = int(sys.argv[4]) if len(sys.argv) > 4 else 0
lib = pkg.load()
os.makedirs("/tmp/models", exist_ok=True)
path = f"/tmp/models/{arch}-1.bin"
if not os.path.exists(path):
    t = time.time()
    modelgen.write_model(path, arch)
    print("model written in %.1f s" % (time.time() - t), flush=True)
cb = capi.LOG_CB(lambda level, text, ud: None)
lib.whisper_log_set(C.cast(cb, C.c_void_p), None)
ctx = lib.whisper_init_from_file_with_params(path.encode(), lib.whisper_context_default_params())
assert ctx
pcm = np.concatenate([modelgen.synth_pcm(480000, seed=7, stream=i) for i in range(n_win)])
p = lib.whisper_full_default_params(0)
p.greedy.best_of = 1
p.temperature_inc = 0.0
p.no_timestamps = True
p.print_progress = False
p.max_tokens = max_tokens
for r in range(reps):
    t = time.time()
    rc = lib.whisper_full_parallel(ctx, p, pcm.ctypes.data_as(C.POINTER(C.c_float)), len(pcm), n_win)
    dt = time.time() - t
    print(f"rep {r}: rc {rc} {dt*1e3:.1f} ms -> {30.0*n_win/dt:.1f}x real time; launches {lib.whisper_b200_kernel_launches(ctx)}", flush=True)
lib.whisper_free(ctx)
