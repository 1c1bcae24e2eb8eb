"""Quick device-timed probe of the single kernels (mel GB/s, GEMM TFLOP/s).  Not the bench; a development aid."""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import open_whisper_kit_b200 as pkg  # noqa: E402
from open_whisper_kit_b200 import modelgen  # noqa: E402

lib = pkg.load(strict_api=False)
FP = C.POINTER(C.c_float)
peaks = {"hbm_gbs": 6549.4, "bf16_tflops": 1681.8}
try:
    peaks.update(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))))
except OSError:
    pass

out = {}
for n_mel in (80, 128):
    filt = modelgen.mel_filters(n_mel)
    for n_streams, n_samples in ((64, 480000), (120, 480000 * 1)):
        ms = lib.whisper_b200_kernel_log_mel_bench(n_streams, n_samples, filt.ctypes.data_as(FP), n_mel, 20, 1)
        secs = n_streams * n_samples / 16000.0
        byts = secs * (16000 * 4 + 100 * n_mel * 4)
        gbs = byts / (ms * 1e-3) / 1e9 if ms > 0 else -1
        print(f"mel n_mel={n_mel} streams={n_streams}x{n_samples}: {ms:.4f} ms  {gbs:.1f} GB/s "
              f"({gbs / peaks['hbm_gbs']:.3f} of measured HBM)", flush=True)
        out[f"mel_{n_mel}_{n_streams}"] = {"ms": ms, "gbs": gbs}

for dtype in (0, 1):
    for (M, N, K, gelu) in ((96000, 3840, 1280, 0), (96000, 1280, 1280, 0), (96000, 5120, 1280, 1),
                            (96000, 1280, 5120, 0), (24000, 1536, 512, 0), (24000, 2048, 512, 1),
                            (8192, 8192, 8192, 0)):
        ms = lib.whisper_b200_kernel_gemm_bench(dtype, M, N, K, gelu, 10)
        tf = 2.0 * M * N * K / (ms * 1e-3) / 1e12 if ms > 0 else -1
        print(f"gemm dtype={dtype} {M}x{N}x{K} gelu={gelu}: {ms:.4f} ms  {tf:.1f} TFLOP/s "
              f"({tf / peaks['bf16_tflops']:.3f} of measured burst)", flush=True)
        out[f"gemm_{dtype}_{M}_{N}_{K}_{gelu}"] = {"ms": ms, "tflops": tf}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
