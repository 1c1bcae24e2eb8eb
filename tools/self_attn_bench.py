"""Development aid: back-to-back self-attention launches (whisper_b200_kernel_step_bench which=2) at several positions.
WHISPER_B200_SELF_MMA=0 selects the CUDA-core kernel."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import open_whisper_kit_b200 as pkg  # noqa: E402
lib = pkg.load()
for pos in (0, 16, 64, 113, 219, 447):
    us = lib.whisper_b200_kernel_step_bench(2, 0, 64, 1280, pos, 200)
    print(f"self-attention R=64 d=1280 pos={pos}: {us:.2f} us per launch (SELF_MMA={os.environ.get('WHISPER_B200_SELF_MMA', '1')})")
