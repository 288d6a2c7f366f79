#!/bin/bash
# One GPU-box visit: parity tests, bench lines for every workload, ncu launch list + one full capture of the tile kernel.
# usage (from the repo root, under gpurun): bash scripts/gpu_round.sh <tag> [quick]
set -u
TAG="${1:-r1}"
MODE="${2:-full}"
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/smi.txt" 2>&1
if [ "$MODE" != "quick" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1
  echo "pytest exit $?" >> "$OUT/pytest_gpu.log"
  tail -5 "$OUT/pytest_gpu.log"
fi
timeout 600 python bench.py --workload c4 > "$OUT/bench_c4.json" 2> "$OUT/bench_c4.err"; tail -c 3000 "$OUT/bench_c4.json"
for w in c1 c2 c3 c5; do
  timeout 600 python bench.py --workload $w > "$OUT/bench_$w.json" 2> "$OUT/bench_$w.err"
done
if [ "$MODE" != "quick" ]; then
  timeout 600 python bench.py --impl reference --workload c4 --steps 5 --warmup 1 > "$OUT/bench_ref_c4.json" 2> "$OUT/bench_ref_c4.err"
fi
# launch list (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_c4.csv" \
  python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline > "$OUT/ncu_launch.log" 2>&1
# full capture of the tile kernel and of setup (2 launches each, after warm-up)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'^(k_tile|k_setup|k_setup_1x)$' -s 4 -c 2 -f -o "$OUT/prof_c4" \
  python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline > "$OUT/ncu_full.log" 2>&1
ncu -i "$OUT/prof_c4.ncu-rep" --page source --print-source cuda,sass --csv > "$OUT/source_c4.csv" 2>/dev/null
ncu -i "$OUT/prof_c4.ncu-rep" --page raw --csv > "$OUT/raw_c4.csv" 2>/dev/null
rm -f "$OUT/prof_c4.ncu-rep"   # (gpurun brings back at most 64 MiB: the CSV exports are what profiles/summarize.py reads)
if [ "$MODE" != "quick" ]; then
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'^(k_tile|k_setup|k_setup_1x)$' -s 4 -c 2 -f -o "$OUT/prof_c5" \
  python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu-baseline > "$OUT/ncu_full_c5.log" 2>&1
ncu -i "$OUT/prof_c5.ncu-rep" --page source --print-source cuda,sass --csv > "$OUT/source_c5.csv" 2>/dev/null
ncu -i "$OUT/prof_c5.ncu-rep" --page raw --csv > "$OUT/raw_c5.csv" 2>/dev/null
rm -f "$OUT/prof_c5.ncu-rep"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_c5.csv" \
  python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu-baseline > "$OUT/ncu_launch_c5.log" 2>&1
fi
ls -la "$OUT"
