# Final validation of the round-2 tree on one B200: GPU tests, smoke, cfg4 launch list, default bench.
NV='--nvtx --nvtx-include profile/'
L='--metrics gpu__time_duration.sum --clock-control none --csv'
python -m pytest tests -m gpu -x -q 2>&1 | grep -v DEBUG | tail -3 > gpurun_out/r02_final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/r02_final_tests.log 2>&1
ncu $NV $L --log-file gpurun_out/r02d_cfg4_context_2layers_launches.csv python tools/ncu_decode.py --prefill --layers 2 > gpurun_out/r02d_p.log 2>&1
python tools/ncu_summary.py launches gpurun_out/r02d_cfg4_context_2layers_launches.csv gpurun_out/r02d_cfg4_context_2layers_launches.txt
python bench.py > gpurun_out/r02_bench_final4.log 2>&1
cat gpurun_out/r02_final_tests.log | tail -4
tail -c 200 gpurun_out/r02_bench_final4.log
