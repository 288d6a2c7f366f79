#!/bin/bash
# Benchmarks every library variant built by scripts/build_variants.sh (C4 and C5, device-timed ms per step and kernel times).
for so in swiftshader_b200/csrc/variants/*.so; do
  for w in ${WORKLOADS:-c4 c5}; do
    SWCU_LIB=$PWD/$so timeout 300 python bench.py --workload $w --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$so'.split('_')[-1], d['config']['workload'][:2], 'ms', round(d['ms_per_step'],4), {k: round(v,4) for k,v in d['kernels_ms'].items() if v>0.02})"
  done
done
