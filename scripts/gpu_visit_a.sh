#!/bin/bash
# Visit A of the last session: parity + bench lines of the write-only path, setup launch-bound variants, full captures of the C1 / C3 tile kernel.
TAG="${1:-r3a}"
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
timeout 900 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?" >> "$OUT/pytest_gpu.log"
tail -4 "$OUT/pytest_gpu.log"
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["config"]["workload"], "ms/step", round(d["ms_per_step"], 4), "sust", round((d.get("sustained") or {}).get("ms_per_step", 0), 4), "e2e ms", round(d["e2e"]["ms_per_step"], 4), "frac", round(d["roofline"]["frac"], 4),
          {k: round(v, 4) for k, v in d["kernels_ms"].items()})
except Exception as e:
    print("bench failed:", e)
PY
}
for w in c1 c2 c3 c4 c5; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline > "$OUT/bench_$w.json" 2> "$OUT/bench_$w.err"; show "$OUT/bench_$w.json"; tail -2 "$OUT/bench_$w.err"
done
WORKLOADS="c4" bash scripts/gpu_variants.sh 2>&1 | tee "$OUT/variants.log"
for w in c1 c3; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^k_tile$' -s 4 -c 1 -f -o "$OUT/prof_$w" \
    python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > "$OUT/ncu_full_$w.log" 2>&1
  ncu -i "$OUT/prof_$w.ncu-rep" --page source --print-source cuda,sass --csv > "$OUT/source_$w.csv" 2>/dev/null
  ncu -i "$OUT/prof_$w.ncu-rep" --page raw --csv > "$OUT/raw_$w.csv" 2>/dev/null
  rm -f "$OUT/prof_$w.ncu-rep"
done
ls -la "$OUT"
