"""Development aid: the encoder's GEMM shapes through whisper_b200_kernel_gemm_bench (device-resident operands, 16-bit output),
TFLOP/s per shape with and without the GELU epilogue.  M = 48000 rows = one 32-window encoder chunk of large-v3."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import open_whisper_kit_b200 as pkg  # noqa: E402
lib = pkg.load()
M = 48000
for name, N, K in (("qkv", 3840, 1280), ("attn out", 1280, 1280), ("mlp up", 5120, 1280), ("mlp down", 1280, 5120), ("cross kv", 2560, 1280)):
    for gelu in (0, 1):
        ms = lib.whisper_b200_kernel_gemm_bench(0, M, N, K, gelu, 20)
        print(f"{name:9s} M={M} N={N} K={K} gelu={gelu}: {ms:.3f} ms = {2.0 * M * N * K / ms / 1e9:.0f} TFLOP/s", flush=True)
