#!/bin/bash
# Round-2 profiling pass 2 (GPU box): ncu --set full of the kernels added late in the round - the 16 -> 16 channel kernels
# and rds_fwd inside a cfg3 training step, the 2 x 2 cluster GEMM inside a cfg2 step, the fused inference kernels inside
# an eval forward + decode of a 448-line batch - plus launch lists of the cfg3 step and the decode batch.
set -x
cd "$(dirname "$0")/.."
WORKLOAD=train_cfg3 WARM=2 STEPS=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --csv --log-file gpurun_out/r02b_launches_cfg3.csv python tools/one_step.py > gpurun_out/r02b_launches_cfg3.log 2>&1
DECODE=1 WARM=2 STEPS=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --csv --log-file gpurun_out/r02b_launches_decode.csv python tools/one_step.py > gpurun_out/r02b_launches_decode.log 2>&1
WORKLOAD=train_cfg3 WARM=2 STEPS=1 timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:"conv16_fwd_kernel|conv16_wgrad_kernel|rds_fwd_kernel" -s 6 -c 4 \
  -o gpurun_out/r02b_c16 -f python tools/one_step.py > gpurun_out/r02b_c16.log 2>&1
WARM=2 STEPS=1 timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:"tc_gemm_x3_kernel" -s 24 -c 6 \
  -o gpurun_out/r02b_gemm -f python tools/one_step.py > gpurun_out/r02b_gemm.log 2>&1
DECODE=1 WARM=2 STEPS=1 timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:"tc_conv_fwd|tc_gemm_x3_persist_kernel|bilstm_fwd_cluster_kernel" -s 30 -c 14 \
  -o gpurun_out/r02b_decode -f python tools/one_step.py > gpurun_out/r02b_decode.log 2>&1
for n in c16 gemm decode; do ncu -i gpurun_out/r02b_$n.ncu-rep --page raw --csv > gpurun_out/r02b_${n}_raw.csv 2>/dev/null; done
ls -la gpurun_out/r02b_*
