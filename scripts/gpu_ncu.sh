#!/bin/bash
# ncu full capture of selected kernels of one workload.  usage (under gpurun): bash scripts/gpu_ncu.sh <tag> <workload> <kernel regex> [skip] [count]
TAG="$1"; W="$2"; RX="$3"; SKIP="${4:-6}"; CNT="${5:-2}"
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RX" -s "$SKIP" -c "$CNT" -f -o "$OUT/prof_$W" python bench.py --workload "$W" --steps 2 --warmup 3 --no-cpu-baseline > "$OUT/ncu_full_$W.log" 2>&1
ncu -i "$OUT/prof_$W.ncu-rep" --page source --print-source cuda,sass --csv > "$OUT/source_$W.csv" 2>/dev/null
ncu -i "$OUT/prof_$W.ncu-rep" --page raw --csv > "$OUT/raw_$W.csv" 2>/dev/null
rm -f "$OUT/prof_$W.ncu-rep"
python profiles/ncu_raw.py "$OUT/raw_$W.csv" | head -60
