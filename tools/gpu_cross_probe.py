"""Development aid: decode timing of the unfused path with the two load flavours of the cross-attention kernel."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for chain, ld in (("0", "0"), ("1", "0"), ("2", "0")):
    env = dict(os.environ, WHISPER_B200_CHAIN=chain, WHISPER_B200_CROSS_LD=ld, WHISPER_B200_CHAIN_UNITS="4")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gpu_decode_probe.py"), "large-v3", "64", "2", "60"], env=env,
                       capture_output=True, text=True)
    last = [l for l in (r.stdout + r.stderr).splitlines() if "run_streams" in l]
    print(f"chain={chain} cross_ld={ld}: {last[-1] if last else r.stderr[-400:]}", flush=True)
