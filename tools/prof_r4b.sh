#!/bin/bash
# live per-CTA time stamps of the decoder-step GEMMs (two layers at a mid-sequence step) -> where a layer's time goes
set -u
OUT=gpurun_out
WHISPER_B200_TCS_TRACE=21630:14 timeout 300 python tools/gpu_decode_probe.py large-v3 64 1 125 > $OUT/r4_tcs_trace.log 2>&1; echo "rc=$?"
grep -c tcs_trace $OUT/r4_tcs_trace.log
