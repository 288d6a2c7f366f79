#!/bin/bash
# bench.py at N GPUs under torchrun.  usage (under gpurun --gpus N): bash scripts/gpu_scale.sh <tag> <N> [extra bench args]
TAG="$1"; N="$2"; shift 2
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus $N "$@" > "$OUT/bench_n$N.json" 2> "$OUT/bench_n$N.err"
echo "exit $?"
python - "$OUT/bench_n$N.json" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    def show(tag, x):
        print(tag, x["config"]["workload"][:2], "N", x["config"]["bands"], "ms/step", round(x["ms_per_step"], 4), "hash_ok", x.get("frame_hash_ok"),
              "e2e ms", round(x["e2e"]["ms_per_step"], 4) if "e2e" in x else None, "sustained", round(x["sustained"]["ms_per_step"], 4) if "sustained" in x else None,
              {k: round(v, 4) for k, v in x["kernels_ms"].items()})
    show("primary", d)
    if "secondary" in d:
        show("secondary", d["secondary"])
except Exception as e:
    print("bench failed:", e)
PY
grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$" "$OUT/bench_n$N.err" | tail -8
