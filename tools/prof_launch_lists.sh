set -x
cd $GRAFT_REPO_ROOT
# 1. plain run of the exact command profiled below
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench_s1.log 2>&1 || exit 1
# 2. launch lists (one pass per kernel): encoder part of the timed step, then ~3 decode steps
ncu --metrics gpu__time_duration.sum --clock-control none -s 236830 -c 1500 --csv --log-file gpurun_out/r1_launches_encoder.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 250000 -c 1100 --csv --log-file gpurun_out/r1_launches_decode.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l2.log 2>&1
tail -2 gpurun_out/ncu_l1.log gpurun_out/ncu_l2.log
