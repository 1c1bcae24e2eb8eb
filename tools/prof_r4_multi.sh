#!/bin/bash
# BASELINE configs 4 (large-v3-turbo, "beam search" 5 + timestamps, 120 windows in total) and 5 (log-mel sweep) on N GPUs of one box:
#   gpurun --gpus N -- bash tools/prof_r4_multi.sh N
set -u
N=${1:-2}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 bench.py --gpus $N --config turbo-beam5 --steps 3 --warmup 3 > $O/r4_bench_turbo_beam5_n$N.json 2> $O/r4_bench_turbo_beam5_n$N.err; echo "turbo-beam5 rc=$?"; tail -c 700 $O/r4_bench_turbo_beam5_n$N.json | head -c 400; echo
timeout 900 $TR --master-port 29512 bench.py --gpus $N --config mel-sweep --steps 5 --warmup 3 > $O/r4_bench_mel_sweep_n$N.json 2> $O/r4_bench_mel_sweep_n$N.err; echo "mel-sweep rc=$?"; head -c 300 $O/r4_bench_mel_sweep_n$N.json; echo
