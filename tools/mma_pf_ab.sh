# A/B: next-projection L2 request window (TB_MMA_PF) on the tensor-core GEMV
for pf in 0 1; do
  for w in cfg5 cfg3_int8kv; do
    TB_MMA_PF=$pf python bench.py --workload $w --only-headline --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/mmapf_${w}_$pf.log 2>&1
    python - <<PY
import json
for l in open("gpurun_out/mmapf_${w}_$pf.log"):
    if l.startswith("{"):
        d=json.loads(l); print("$w MMA_PF=$pf", d["value"], d["decode_step"]["ms"], d["roofline"]["us_per_launch"], d["roofline"]["frac"])
PY
  done
  TB_GEMV_MMA_MIN_M=1 TB_MMA_PF=$pf python bench.py --workload cfg2 --only-headline --no-cpu-baseline --steps 3 --warmup 3 2>&1 | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('cfg2 mma PF=$pf', d['value'], d['decode_step']['ms'], d['roofline']['us_per_launch'])"
done
