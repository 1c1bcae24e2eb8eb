set -u
O=gpurun_out
WHISPER_B200_TCS_TRACE=21423:14 timeout 300 python tools/gpu_decode_probe.py large-v3 8 2 125 > $O/r4_tcs_trace_r8.log 2>&1; echo "rc=$?"
grep -E "^rep|run_streams" $O/r4_tcs_trace_r8.log | tail -3
timeout 300 python tools/gpu_decode_probe.py large-v3 8 3 0 > $O/r4_probe_r8.log 2>&1; grep -E "^rep|run_streams" $O/r4_probe_r8.log | tail -2
