#!/bin/bash
# Refresh of the launch list and of the cross-attention capture after the last kernel changes of the round (one B200).
set -u
OUT=gpurun_out
DEC="python tools/gpu_decode_probe.py large-v3 64 1 6"
$DEC > $OUT/r4_prof_plain_dec.log 2>&1 || { echo "plain decode probe failed"; exit 1; }
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/r4_launches_decode.csv $DEC > $OUT/r4_ncu_ll.log 2>&1; echo "launch list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cross_attn_bulk -s 40 -c 1 -f -o $OUT/r4_cross_attn $DEC > $OUT/r4_ncu_cross_attn.log 2>&1; echo "ncu cross rc=$?"
