#!/bin/bash
# Round-2 profiling pass (run under gpurun on ONE B200): plain runs first, then ncu launch lists and one --set full capture per
# kernel of interest.  Reports land in gpurun_out/; tools/ncu_traffic.py turns their raw pages into profiles/*.txt / ncu_traffic.json.
set -u
OUT=gpurun_out
DEC="python tools/gpu_decode_probe.py large-v3 64 1 6"
BEAM="python tools/gpu_config4_probe.py 15"
$DEC > $OUT/r2_prof_plain_dec.log 2>&1 || { echo "plain decode probe failed"; tail -5 $OUT/r2_prof_plain_dec.log; exit 1; }
$BEAM > $OUT/r2_prof_plain_beam.log 2>&1 || { echo "plain beam probe failed"; tail -5 $OUT/r2_prof_plain_beam.log; exit 1; }
tail -2 $OUT/r2_prof_plain_dec.log; tail -2 $OUT/r2_prof_plain_beam.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $OUT/r2_launches_decode.csv $DEC > $OUT/r2_ncu_ll.log 2>&1
echo "launch list rc=$?"
full() {   # name regex skip cmd...
    local name=$1 re=$2 skip=$3; shift 3
    ncu --set full --clock-control none --import-source on -k regex:$re -s $skip -c 1 -f -o $OUT/r2_$name "$@" > $OUT/r2_ncu_$name.log 2>&1
    echo "ncu $name rc=$?"
}
full cross_attn cross_attn_bulk 40 $DEC
full mel mel_kernel 0 $DEC
full layernorm_vec layernorm_vec 10 $DEC
full enc_attn_tc enc_attn_tc 3 $DEC
full tc_gemm "tc_gemm_kernel" 20 $DEC
full tc_skinny tc_skinny 200 $DEC
full sample_greedy sample_kernel 2 $DEC
full sample_draws sample_kernel 30 $BEAM
full cross_attn_nq5 cross_attn_bulk 60 $BEAM
ls -la $OUT/*.ncu-rep | tail -12
