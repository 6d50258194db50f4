#!/bin/bash
# Round-2 captures after the fp16 / W8 one-row projections moved to the tensor-core GEMV: cfg2 launch list and ncu --set full of
# the headline's dominant kernel (gemv_mma_kernel<0,*,4>, one token row) inside the real decode step.
NV='--nvtx --nvtx-include profile/'
L='--metrics gpu__time_duration.sum --clock-control none --csv'
F='--set full --clock-control none --import-source on'
S='python tools/ncu_summary.py'
ncu $NV $L --log-file gpurun_out/r02c_cfg2_decode_launches.csv python tools/ncu_decode.py --workload cfg2 --steps 2 > gpurun_out/r02c_p1.log 2>&1
$S launches gpurun_out/r02c_cfg2_decode_launches.csv gpurun_out/r02c_cfg2_decode_launches.txt
ncu $NV $F -k regex:gemv_mma -c 5 -o gpurun_out/r02c_gemv_mma_fp16_m1 python tools/ncu_decode.py --workload cfg2 --steps 1 > gpurun_out/r02c_p2.log 2>&1
$S full gpurun_out/r02c_gemv_mma_fp16_m1.ncu-rep gpurun_out/r02c_gemv_mma_fp16_m1_full.txt
python tools/ncu_hotspots.py gpurun_out/r02c_gemv_mma_fp16_m1.ncu-rep >> gpurun_out/r02c_gemv_mma_fp16_m1_full.txt
ls -la gpurun_out/r02c_*
