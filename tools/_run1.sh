set -u
O=gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-fp8-line > $O/r3_t9_bench.log 2> $O/r3_t9_bench.err; echo "bench rc=$?"
WHISPER_B200_GEMM_2CTA=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-fp8-line > $O/r3_t9_bench_1cta.log 2> $O/r3_t9_bench_1cta.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 3 --warmup 3 --windows 8 --no-cpu-baseline --no-fp8-line > $O/r3_t9_bench_w8.log 2> $O/r3_t9_bench_w8.err; echo "bench rc=$?"
WHISPER_B200_CROSS_PF_CHUNKS=0 timeout 600 python bench.py --steps 3 --warmup 3 --windows 8 --no-cpu-baseline --no-fp8-line > $O/r3_t9_bench_w8_nopf.log 2> $O/r3_t9_bench_w8_nopf.err; echo "bench rc=$?"
