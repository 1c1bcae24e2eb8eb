"""Development aid: per-kernel microseconds of the decoder-step kernels at full clocks (back-to-back launches)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import open_whisper_kit_b200 as pkg  # noqa: E402

lib = pkg.load()
R, d = 64, 1280
print("layernorm      us", lib.whisper_b200_kernel_step_bench(0, 0, R, d, 0, 2000))
x = lib.whisper_b200_kernel_step_bench(1, 0, R, d, 0, 200)
print("cross_attn     us", x, "GB/s", R * 1500 * 2 * d * 2 / (x * 1e-6) / 1e9)
print("self_attn@100  us", lib.whisper_b200_kernel_step_bench(2, 0, R, d, 100, 1000))
print("self_attn@220  us", lib.whisper_b200_kernel_step_bench(2, 0, R, d, 220, 1000))
print("kv_append      us", lib.whisper_b200_kernel_step_bench(3, 0, R, d, 5, 2000))
for (M, N, K) in ((64, 3840, 1280), (64, 1280, 1280), (64, 5120, 1280), (64, 1280, 5120), (64, 51866, 1280), (16, 1536, 512), (16, 512, 512)):
    ms = lib.whisper_b200_kernel_gemm_bench(0, M, N, K, 0, 500)
    print(f"skinny gemm {M}x{N}x{K}: {ms*1e3:.2f} us  weights {N*K*2/(ms*1e-3)/1e9:.0f} GB/s")
