#!/bin/bash
# Round-2 (third session) baseline pass, run under gpurun on ONE B200: the GPU test suite, the bench line, then one
# `ncu --set full` capture of the decoder self-attention kernel at a mid-sequence position (skip = steps x 32 layers).
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/r4_gputest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/r4_gputest.log
timeout 600 python bench.py > $OUT/r4_bench_n1.json 2> $OUT/r4_bench_n1.err; echo "bench rc=$?"; head -c 600 $OUT/r4_bench_n1.json; echo
DEC="python tools/gpu_decode_probe.py large-v3 64 1 120"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cross_attn_kernel -s 3400 -c 1 -f -o $OUT/r4_self_attn $DEC > $OUT/r4_ncu_self_attn.log 2>&1
echo "ncu self_attn rc=$?"
ls -la $OUT/*.ncu-rep
