"""TEST INFRASTRUCTURE ONLY -- builds and binds oracle/mel_oracle.c (CPU restatement of the reference log-mel).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_LIB = os.path.join(_BUILD, "libmel_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "mel_oracle.c")
    os.makedirs(_BUILD, exist_ok=True)
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        # same optimisation level / FMA availability as the reference build (oracle/Makefile.ref, VARIANT=v3)
        subprocess.run(["gcc", "-O3", "-march=x86-64-v3", "-fPIC", "-shared", "-o", _LIB, src, "-lm"], check=True)
    return _LIB


def _load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        fp = C.POINTER(C.c_float)
        lib.oracle_log_mel.restype = C.c_int
        lib.oracle_log_mel.argtypes = [fp, C.c_int, fp, C.c_int, fp, C.c_long, C.POINTER(C.c_int)]
        lib.oracle_log_mel_f64.restype = C.c_int
        lib.oracle_log_mel_f64.argtypes = [fp, C.c_int, fp, C.c_int, C.POINTER(C.c_double), C.c_long]
        _lib = lib
    return _lib


def log_mel(pcm, filters):
    """-> (mel [n_mel][n_len] float32, n_len_org); restates reference src/whisper.cpp:3170-3260."""
    lib = _load()
    pcm = np.ascontiguousarray(pcm, dtype=np.float32)
    filters = np.ascontiguousarray(filters, dtype=np.float32)
    n_mel = filters.shape[0]
    fp = C.POINTER(C.c_float)
    org = C.c_int(0)
    n_len = lib.oracle_log_mel(pcm.ctypes.data_as(fp), len(pcm), filters.ctypes.data_as(fp), n_mel, None, 0,
                               C.byref(org))
    out = np.empty((n_mel, n_len), dtype=np.float32)
    rc = lib.oracle_log_mel(pcm.ctypes.data_as(fp), len(pcm), filters.ctypes.data_as(fp), n_mel,
                            out.ctypes.data_as(fp), out.size, C.byref(org))
    assert rc == n_len
    return out, org.value


def log_mel_f64(pcm, filters):
    """Mathematically exact (float64, direct DFT) log-mel; small inputs only."""
    lib = _load()
    pcm = np.ascontiguousarray(pcm, dtype=np.float32)
    filters = np.ascontiguousarray(filters, dtype=np.float32)
    n_mel = filters.shape[0]
    n_len = (len(pcm) + 480000) // 160
    out = np.empty((n_mel, n_len), dtype=np.float64)
    fp = C.POINTER(C.c_float)
    rc = lib.oracle_log_mel_f64(pcm.ctypes.data_as(fp), len(pcm), filters.ctypes.data_as(fp), n_mel,
                                out.ctypes.data_as(C.POINTER(C.c_double)), out.size)
    assert rc == n_len
    return out
