set -u
O=gpurun_out
run() { # name env...
  local n=$1; shift
  env "$@" timeout 300 python tools/gpu_decode_probe.py large-v3 64 3 60 > $O/r3_t11_$n.log 2>&1; echo "$n rc=$?"; grep "run_streams" $O/r3_t11_$n.log | tail -2
}
python - <<'PY'
import torch
p=torch.cuda.get_device_properties(0)
print("L2", p.L2_cache_size)
PY
run persist0_pf4 WHISPER_B200_L2_PERSIST_MB=0 WHISPER_B200_CROSS_PF_CHUNKS=4
run persistmax_pf4 WHISPER_B200_CROSS_PF_CHUNKS=4
run persistmax_pf6 WHISPER_B200_CROSS_PF_CHUNKS=6
run persist48_pf4 WHISPER_B200_L2_PERSIST_MB=48 WHISPER_B200_CROSS_PF_CHUNKS=4
run persist64_pf3 WHISPER_B200_L2_PERSIST_MB=64 WHISPER_B200_CROSS_PF_CHUNKS=3
run persistmax_pf0 WHISPER_B200_CROSS_PF_CHUNKS=0
