#!/bin/bash
# Round-2 profiling pass (GPU box): launch list of three eager cfg2 training steps + ncu --set full of the top kernels.
# Outputs land in gpurun_out/; tools/ncu_summary.py turns them into the tables under profiles/.
set -x
cd "$(dirname "$0")/.."
WARM=2 STEPS=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --csv --log-file gpurun_out/r02_launches.csv python tools/one_step.py > gpurun_out/r02_launches.log 2>&1
WARM=2 STEPS=1 timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"bilstm_fwd_cluster_kernel|bilstm_bwd_kernel|tc_conv_fwd_kernel|tc_gemm_x3_persist_kernel|tc_conv_fwd_persist_kernel" -s 20 -c 12 \
  -o gpurun_out/r02_top_kernels -f python tools/one_step.py > gpurun_out/r02_top_kernels.log 2>&1
ncu -i gpurun_out/r02_top_kernels.ncu-rep --page raw --csv > gpurun_out/r02_top_kernels_raw.csv 2>/dev/null
ls -la gpurun_out/r02_*
