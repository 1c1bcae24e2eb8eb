"""Small single-kernel workloads for `ncu --set full` captures (one launch group per kernel of interest)."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import open_whisper_kit_b200 as pkg  # noqa: E402
from open_whisper_kit_b200 import modelgen  # noqa: E402

lib = pkg.load()
which = sys.argv[1]
FP = C.POINTER(C.c_float)
if which == "cross":
    print(lib.whisper_b200_kernel_step_bench(1, 0, 64, 1280, 0, 2))
elif which == "mel":
    filt = modelgen.mel_filters(128)
    print(lib.whisper_b200_kernel_log_mel_bench(64, 480000, filt.ctypes.data_as(FP), 128, 2, 0))
elif which == "gemm":
    print(lib.whisper_b200_kernel_gemm_bench(0, 48000, 3840, 1280, 0, 2))
elif which == "skinny":
    print(lib.whisper_b200_kernel_gemm_bench(0, 64, 5120, 1280, 0, 2))
