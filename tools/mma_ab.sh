# A/B of the decode schedule at 5..8 rows: RMSNorm fused into the tensor-core GEMV prologue (TB_FUSE_NORM_ROWS=8) or run once
# as its own PDL-chained kernel (default 4); workloads that use gemv_mma_kernel
for rows in 4 8; do
  for w in cfg3_int8kv cfg5_b8 cfg5 cfg3; do
    TB_FUSE_NORM_ROWS=$rows python bench.py --workload $w --only-headline --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/mma_${w}_norm$rows.log 2>&1
    python - <<PY
import json
for l in open("gpurun_out/mma_${w}_norm$rows.log"):
    if l.startswith("{"):
        d=json.loads(l); print("$w FUSE_NORM_ROWS=$rows", d["value"], d["decode_step"]["ms"], d["decode_step"]["kernels"], d["roofline"]["us_per_launch"], d["roofline"]["frac"])
PY
  done
done
