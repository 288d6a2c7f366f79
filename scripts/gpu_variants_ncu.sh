#!/bin/bash
# Per library variant (scripts/build_variants.sh): the tile kernel's duration, executed warp instructions, issue utilisation and
# registers from one ncu pass over a C4 frame, next to the device-timed bench line.  usage (under gpurun): bash scripts/gpu_variants_ncu.sh <tag> [workload]
TAG="${1:-var}"; W="${2:-c4}"
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
for so in swiftshader_b200/csrc/variants/*.so; do
  name=$(basename $so .so); name=${name#libswcuda_}
  SWCU_LIB=$PWD/$so timeout 300 python bench.py --workload $W --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$name', d['config']['workload'][:2], 'ms', round(d['ms_per_step'],4), {k: round(v,4) for k,v in d['kernels_ms'].items() if v>0.02})"
  SWCU_LIB=$PWD/$so timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio \
    --clock-control none -k regex:'^k_tile$' -s 4 -c 1 --csv --log-file "$OUT/ncu_$name.csv" python bench.py --workload $W --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
  python - "$OUT/ncu_$name.csv" "$name" <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
if rows:
    h = rows[0]
    print("   ", sys.argv[2], {r[h.index("Metric Name")].split("__")[-1][:28]: r[h.index("Metric Value")] for r in rows[1:]})
PY
done 2>&1 | tee "$OUT/variants.log"
