set -u
O=gpurun_out
run() { local n=$1; shift
  env "$@" timeout 300 python tools/gpu_decode_probe.py large-v3 64 3 0 > $O/r4_ks_$n.log 2>&1; echo "$n rc=$?"; grep "^rep" $O/r4_ks_$n.log | tail -2
}
run ks8
run ks4 WHISPER_B200_TCS_KS_SMALL=4
run ks5 WHISPER_B200_TCS_KS_SMALL=5
run ks8b
