"""Development aid: batched decode timing with the chain kernel on and off (WHISPER_B200_CHAIN)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
arch = sys.argv[1] if len(sys.argv) > 1 else "large-v3"
configs = [x.split(",") for x in (sys.argv[2] if len(sys.argv) > 2 else "0,2,0;1,2,1").split(";")]
for chain, units, trace in configs:
    env = dict(os.environ, WHISPER_B200_CHAIN=chain, WHISPER_B200_CHAIN_UNITS=units)
    if trace == "1":
        env["WHISPER_B200_CHAIN_TRACE"] = "1"
    print(f"== chain={chain} min_units={units} trace={trace}", flush=True)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gpu_decode_probe.py"), arch, "64", "2", "60"], env=env,
                       capture_output=True, text=True)
    print(r.stdout[-1500:], r.stderr[-2500:], flush=True)
