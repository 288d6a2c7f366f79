#!/bin/bash
# Short GPU-box visit: parity tests, one bench line per listed workload, and one `ncu --set full` capture (with source
# page exported as CSV) of the tile + setup kernels of the first workload.
# usage (under gpurun, from the repo root): bash scripts/gpu_prof.sh <tag> [workloads, default "c4"] [noprof]
set -u
TAG="${1:-prof}"
WLS="${2:-c4}"
NOPROF="${3:-}"
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/smi.txt" 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest_gpu.log"
tail -4 "$OUT/pytest_gpu.log"
for w in $WLS; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline > "$OUT/bench_$w.json" 2> "$OUT/bench_$w.err"
  python - "$OUT/bench_$w.json" <<'EOF'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(j["config"]["workload"], "ms/step", round(j["ms_per_step"], 4), "e2e", j["e2e"], "roofline", j["roofline"], "kernels", j.get("kernels_ms"))
except Exception as e:
    print("bench parse failed", e)
EOF
done
if [ -z "$NOPROF" ]; then
  W1=$(echo $WLS | cut -d' ' -f1)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^(k_tile|k_setup|k_setup_1x)$' -s 4 -c 2 -f -o "$OUT/prof_$W1" \
    python bench.py --workload $W1 --steps 2 --warmup 3 --no-cpu-baseline > "$OUT/ncu_full.log" 2>&1
  ncu -i "$OUT/prof_$W1.ncu-rep" --page source --print-source cuda,sass --csv > "$OUT/source_$W1.csv" 2>/dev/null
  ncu -i "$OUT/prof_$W1.ncu-rep" --page raw --csv > "$OUT/raw_$W1.csv" 2>/dev/null
fi
ls -la "$OUT"
