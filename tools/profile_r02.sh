#!/bin/bash
# Round-2 ncu captures (run on the GPU box through gpurun; outputs under gpurun_out/, summarised on the box by
# tools/ncu_summary.py so only the text summaries and the small reports travel back).  One GPU; never a multi-rank command.
set -x
mkdir -p gpurun_out
NV='--nvtx --nvtx-include profile/'
L='--metrics gpu__time_duration.sum --clock-control none --csv'
F='--set full --clock-control none --import-source on'
S='python tools/ncu_summary.py'
ncu $NV $L --log-file gpurun_out/r02_cfg2_decode_launches.csv python tools/ncu_decode.py --workload cfg2 --steps 2 > gpurun_out/r02_p1.log 2>&1
ncu $NV $L --log-file gpurun_out/r02_sq_decode_launches.csv python tools/ncu_decode.py --workload sq --steps 2 > gpurun_out/r02_p2.log 2>&1
ncu $NV $L --log-file gpurun_out/r02_cfg3_int8kv_decode_launches.csv python tools/ncu_decode.py --workload cfg3_int8kv --steps 2 > gpurun_out/r02_p2b.log 2>&1
ncu $NV $L --log-file gpurun_out/r02_cfg4_context_2layers_launches.csv python tools/ncu_decode.py --prefill --layers 2 > gpurun_out/r02_p3.log 2>&1
for w in cfg2_decode sq_decode cfg3_int8kv_decode cfg4_context_2layers; do $S launches gpurun_out/r02_${w}_launches.csv gpurun_out/r02_${w}_launches.txt; done
ncu $NV $F -k regex:gemv_kernel -c 4 -o gpurun_out/r02_gemv_fp16_m1 python tools/ncu_decode.py --workload cfg2 --steps 1 > gpurun_out/r02_p4.log 2>&1
ncu $NV $F -k regex:gemv_kernel -c 4 -o gpurun_out/r02_gemv_sq_m1 python tools/ncu_decode.py --workload sq --steps 1 > gpurun_out/r02_p5.log 2>&1
ncu $NV $F -k regex:decode_step -c 1 -o gpurun_out/r02_decode_step_cfg2 python tools/ncu_decode.py --workload cfg2 --fused 1 --steps 1 > gpurun_out/r02_p6.log 2>&1
ncu $NV $F -k regex:decode_step -c 1 -o gpurun_out/r02_decode_step_sq python tools/ncu_decode.py --workload sq --fused 1 --steps 1 > gpurun_out/r02_p7.log 2>&1
for r in gemv_fp16_m1 gemv_sq_m1 decode_step_cfg2 decode_step_sq; do $S full gpurun_out/r02_$r.ncu-rep gpurun_out/r02_${r}_full.txt; done
rm -f gpurun_out/r02_decode_step_cfg2.ncu-rep gpurun_out/r02_decode_step_sq.ncu-rep gpurun_out/r02_gemv_sq_m1.ncu-rep
ls -la gpurun_out/r02_*
# tensor-core decode GEMV (gemv_mma.cu, cp.async ring form): the kernel alone on the LLaMA-7B shapes
ncu $F -k regex:gemv_mma -s 8 -c 1 -o gpurun_out/mma4_w8m8_qkv python tools/gemv_mma_bench.py w8 8 qkv > gpurun_out/mma4_a.log 2>&1
ncu $F -k regex:gemv_mma -s 8 -c 1 -o gpurun_out/mma4_w8m8_down python tools/gemv_mma_bench.py w8 8 down > gpurun_out/mma4_b.log 2>&1
ncu $F -k regex:gemv_mma -s 8 -c 1 -o gpurun_out/mma4_w4m1_gu python tools/gemv_mma_bench.py w4 1 gate_up > gpurun_out/mma4_c.log 2>&1
for r in w8m8_qkv w8m8_down w4m1_gu; do $S full gpurun_out/mma4_$r.ncu-rep gpurun_out/r02_gemv_mma_${r}_full.txt; python tools/ncu_hotspots.py gpurun_out/mma4_$r.ncu-rep >> gpurun_out/r02_gemv_mma_${r}_full.txt; done
# launch lists of the tree with the new GEMV
ncu $NV $L --log-file gpurun_out/r02b_cfg3_int8kv_decode_launches.csv python tools/ncu_decode.py --workload cfg3_int8kv --steps 2 > gpurun_out/r02_p8.log 2>&1
ncu $NV $L --log-file gpurun_out/r02b_cfg5_decode_launches.csv python tools/ncu_decode.py --workload cfg5 --steps 2 > gpurun_out/r02_p9.log 2>&1
for w in cfg3_int8kv_decode cfg5_decode; do $S launches gpurun_out/r02b_${w}_launches.csv gpurun_out/r02b_${w}_launches.txt; done
