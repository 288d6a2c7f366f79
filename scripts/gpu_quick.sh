#!/bin/bash
# Quick GPU visit: parity suite + C4/C5 bench lines.  usage (under gpurun): bash scripts/gpu_quick.sh <tag> [nopytest]
TAG="${1:-q}"
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
if [ "${2:-}" != "nopytest" ]; then
  timeout 1400 python -m pytest tests -m gpu -q > "$OUT/pytest_gpu.log" 2>&1; echo "exit $?" >> "$OUT/pytest_gpu.log"
  grep -E "^FAILED|^ERROR|passed|failed" "$OUT/pytest_gpu.log" | cut -c1-250 | tail -25
fi
for w in c4 c5; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline > "$OUT/bench_$w.json" 2> "$OUT/bench_$w.err"
  python - "$OUT/bench_$w.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["config"]["workload"], "ms/step", round(d["ms_per_step"], 4), "e2e ms", round(d["e2e"]["ms_per_step"], 4), "frac", round(d["roofline"]["frac"], 4),
          {k: round(v, 4) for k, v in d["kernels_ms"].items()})
except Exception as e:
    print("bench failed:", e)
PY
  tail -3 "$OUT/bench_$w.err"
done
