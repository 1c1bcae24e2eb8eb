"""Golden vectors for the ggml block formats, produced by the UNMODIFIED reference (oracle/_ref):
quantize_row_{q4_0,q4_1,q5_0,q5_1,q8_0}_ref and dequantize_row_* of ggml/src/ggml-quants.c on the fixed input of
tests/test_oracle_pinning.py::_quant_inputs, and dequantize_row_* (incl. q2_K..q6_K) on the seeded random blocks of
_random_blocks.  Run in the build container: python tests/golden/make_golden_quant.py"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from open_whisper_kit_b200 import modelgen  # noqa: E402
from oracle import reflib  # noqa: E402
from test_oracle_pinning import _quant_inputs  # noqa: E402

FP = C.POINTER(C.c_float)
ref, variant = reflib.load()
assert ref is not None, "build oracle/_ref first (make -C oracle -f Makefile.ref)"
x = _quant_inputs()
out = {}
for qtype, (_, _, bb) in modelgen.QUANT_TYPES.items():
    raw = np.zeros(len(x) // 32 * bb, np.uint8)
    getattr(ref, f"quantize_row_{qtype}_ref")(x.ctypes.data_as(FP), raw.ctypes.data_as(C.c_void_p), C.c_int64(len(x)))
    y = np.empty_like(x)
    getattr(ref, f"dequantize_row_{qtype}")(raw.ctypes.data_as(C.c_void_p), y.ctypes.data_as(FP), C.c_int64(len(x)))
    out[f"{qtype}/raw"], out[f"{qtype}/deq"] = raw, y
from test_oracle_pinning import _random_blocks  # noqa: E402
for qtype in list(modelgen.QUANT_TYPES) + list(modelgen.KQUANT_TYPES):
    raw, n_el = _random_blocks(qtype)
    y = np.empty(raw.shape[0] * n_el, np.float32)
    getattr(ref, f"dequantize_row_{qtype}")(raw.ctypes.data_as(C.c_void_p), y.ctypes.data_as(FP), C.c_int64(y.size))
    out[f"{qtype}/rand_deq16"] = y.astype(np.float16)           # what the loader must store: the f32 expansion rounded once
np.savez_compressed(os.path.join(HERE, "golden_quant.npz"), **out)
print("wrote golden_quant.npz from reference build", variant)
