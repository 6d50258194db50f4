NV='--nvtx --nvtx-include profile/'
L='--metrics gpu__time_duration.sum --clock-control none --csv'
S='python tools/ncu_summary.py'
python bench.py > gpurun_out/r02_bench_final.log 2>&1
ncu $NV $L --log-file gpurun_out/r02b_cfg3_int8kv_decode_launches.csv python tools/ncu_decode.py --workload cfg3_int8kv --steps 2 > gpurun_out/r02_p8.log 2>&1
ncu $NV $L --log-file gpurun_out/r02b_cfg5_decode_launches.csv python tools/ncu_decode.py --workload cfg5 --steps 2 > gpurun_out/r02_p9.log 2>&1
for w in cfg3_int8kv_decode cfg5_decode; do $S launches gpurun_out/r02b_${w}_launches.csv gpurun_out/r02b_${w}_launches.txt; done
tail -c 300 gpurun_out/r02_bench_final.log
