# Round-1 closing profile pass (run under gpurun): launch lists of the bench step + ncu --set full of the kernels that changed
set -x
cd $GRAFT_REPO_ROOT
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench_s1.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 236830 -c 1500 --csv --log-file gpurun_out/r1_launches_encoder.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 250000 -c 1100 --csv --log-file gpurun_out/r1_launches_decode.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l2.log 2>&1
tail -2 gpurun_out/ncu_l1.log gpurun_out/ncu_l2.log
for k in cross skinny; do python tools/gpu_ncu_targets.py $k > gpurun_out/plain_$k.log 2>&1 || exit 1; done
ncu --set full --clock-control none --import-source on -k regex:cross_attn -c 1 -f -o gpurun_out/r1_cross_attn python tools/gpu_ncu_targets.py cross > gpurun_out/ncu_cross.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_skinny_kernel -c 1 -f -o gpurun_out/r1_tc_skinny python tools/gpu_ncu_targets.py skinny > gpurun_out/ncu_skinny.log 2>&1
python tools/gpu_decode_probe.py large-v3 64 1 8 > gpurun_out/plain_probe8.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:sample_greedy -s 5 -c 1 -f -o gpurun_out/r1_sample python tools/gpu_decode_probe.py large-v3 64 1 8 > gpurun_out/ncu_sample.log 2>&1
ls -la gpurun_out/*.ncu-rep
