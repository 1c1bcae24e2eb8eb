# ncu --set full captures of the top kernels (one small workload each, see tools/gpu_ncu_targets.py); run under gpurun
set -x
cd $GRAFT_REPO_ROOT
for k in cross mel skinny gemm; do
  python tools/gpu_ncu_targets.py $k > gpurun_out/plain_$k.log 2>&1 || exit 1
done
ncu --set full --clock-control none --import-source on -k regex:cross_attn_kernel -c 1 -f -o gpurun_out/r1_cross_attn python tools/gpu_ncu_targets.py cross > gpurun_out/ncu_cross.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mel_kernel -c 1 -f -o gpurun_out/r1_mel python tools/gpu_ncu_targets.py mel > gpurun_out/ncu_mel.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:skinny_gemm_kernel -c 1 -f -o gpurun_out/r1_skinny python tools/gpu_ncu_targets.py skinny > gpurun_out/ncu_skinny.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -c 1 -f -o gpurun_out/r1_tc_gemm python tools/gpu_ncu_targets.py gemm > gpurun_out/ncu_gemm.log 2>&1
ls -la gpurun_out/*.ncu-rep
