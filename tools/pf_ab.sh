# A/B of the two L2 request schemes of the decode GEMV (TB_GEMV_SELF_PF rows, TB_PF_MB / TB_PF_ATTN_MB next-weights windows)
for cfg in "0 0 0" "1 0 0" "2 0 0" "0 12 12" "1 12 12" "2 12 12" "0 8 8" "0 16 16" "1 6 6"; do
  set -- $cfg
  for w in cfg2 sq; do
    TB_GEMV_SELF_PF=$1 TB_PF_MB=$2 TB_PF_ATTN_MB=$3 python bench.py --workload $w --only-headline --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/pf_${w}_$1_$2_$3.log 2>&1
    python - <<PY
import json
for l in open("gpurun_out/pf_${w}_$1_$2_$3.log"):
    if l.startswith("{"):
        d=json.loads(l); print("$w SELF=$1 PF=$2 ATTN=$3", d["value"], d["decode_step"]["ms"], d["roofline"]["us_per_launch"])
PY
  done
done
