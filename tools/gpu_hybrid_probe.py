"""Development aid: hybrid chain mode under a few settings."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for chain, units, coop in (("0", "4", "1"), ("2", "4", "1"), ("2", "4", "0"), ("2", "2", "0"), ("2", "8", "0"), ("1", "4", "0")):
    env = dict(os.environ, WHISPER_B200_CHAIN=chain, WHISPER_B200_CHAIN_UNITS=units, WHISPER_B200_CHAIN_COOP=coop)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gpu_decode_probe.py"), "large-v3", "64", "2", "60"], env=env,
                       capture_output=True, text=True)
    last = [l for l in (r.stdout + r.stderr).splitlines() if "run_streams" in l]
    print(f"chain={chain} units={units} coop={coop}: {last[-1][-75:] if last else (r.stdout + r.stderr)[-400:]}", flush=True)
