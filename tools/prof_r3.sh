#!/bin/bash
# Round-2 (second session) profiling pass, run under gpurun on ONE B200: a plain run first, then the ncu launch list of decode steps
# and one `ncu --set full` capture per changed kernel.  Reports land in gpurun_out/; tools/ncu_traffic.py turns their raw pages into
# profiles/r3_ncu_summary.txt and profiles/ncu_traffic.json.
set -u
OUT=gpurun_out
DEC="python tools/gpu_decode_probe.py large-v3 64 1 6"
$DEC > $OUT/r3_prof_plain_dec.log 2>&1 || { echo "plain decode probe failed"; tail -5 $OUT/r3_prof_plain_dec.log; exit 1; }
tail -2 $OUT/r3_prof_plain_dec.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/r3_launches_decode.csv $DEC > $OUT/r3_ncu_ll.log 2>&1
echo "launch list rc=$?"
full() {   # name regex skip cmd...
    local name=$1 re=$2 skip=$3; shift 3
    ncu --set full --clock-control none --import-source on -k regex:$re -s $skip -c 1 -f -o $OUT/r3_$name "$@" > $OUT/r3_ncu_$name.log 2>&1
    echo "ncu $name rc=$?"
}
full cross_attn cross_attn_bulk 40 $DEC
full enc_attn_tc enc_attn_tc 3 $DEC
full tc_gemm2 tc_gemm2 20 $DEC
full tc_skinny tc_skinny 200 $DEC
ls -la $OUT/r3_*.ncu-rep
