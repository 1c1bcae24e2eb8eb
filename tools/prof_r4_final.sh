#!/bin/bash
# Final pass of the round on ONE B200: GPU test suite, bench line, then (each only after its plain run exited 0) the ncu launch list of
# decode steps and one `ncu --set full` capture per kernel changed this session.  tools/ncu_traffic.py turns the raw pages into
# profiles/r4_ncu_summary.txt and profiles/ncu_traffic.json.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/r4_gputest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/r4_gputest.log
timeout 600 python bench.py > $OUT/r4_bench_n1.json 2> $OUT/r4_bench_n1.err; echo "bench rc=$?"; head -c 400 $OUT/r4_bench_n1.json; echo
DEC="python tools/gpu_decode_probe.py large-v3 64 1 6"
$DEC > $OUT/r4_prof_plain_dec.log 2>&1 || { echo "plain decode probe failed"; tail -5 $OUT/r4_prof_plain_dec.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/r4_launches_decode.csv $DEC > $OUT/r4_ncu_ll.log 2>&1
echo "launch list rc=$?"
full() {   # name regex skip cmd...
    local name=$1 re=$2 skip=$3; shift 3
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$re -s $skip -c 1 -f -o $OUT/r4_$name "$@" > $OUT/r4_ncu_$name.log 2>&1
    echo "ncu $name rc=$?"
}
full enc_attn_tc enc_attn_tc 3 $DEC
full tc_gemm2 tc_gemm2 20 $DEC
full self_attn_mma self_attn_mma 3400 python tools/gpu_decode_probe.py large-v3 64 1 120
ls -la $OUT/r4_*.ncu-rep
