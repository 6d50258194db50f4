# A/B: FMA GEMV (default for <= 4 rows, except int4) vs the tensor-core GEMV (TB_GEMV_MMA_MIN_M=1) on one-row decode steps
for mm in 0 1; do
  for w in cfg2 sq w8_b1; do
    TB_GEMV_MMA_MIN_M=$mm python bench.py --workload $w --only-headline --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/minm_${w}_$mm.log 2>&1
    python - <<PY
import json
for l in open("gpurun_out/minm_${w}_$mm.log"):
    if l.startswith("{"):
        d=json.loads(l); print("$w MMA_MIN_M=$mm", d["value"], d["decode_step"]["ms"], d["roofline"]["us_per_launch"], d["roofline"]["frac"])
PY
  done
done
