#!/bin/bash
# SASS evidence of the TMA / mbarrier path in the built library (no GPU needed): counts of the Blackwell bulk-tensor instructions per
# kernel family.  usage: bash profiles/sass_evidence.sh <tag>   -> profiles/<tag>/sass_tma.txt
TAG="$1"; OUT="$(dirname "$0")/$TAG"; mkdir -p "$OUT"
SO="$(dirname "$0")/../swiftshader_b200/csrc/libswcuda.so"
{
  echo "cuobjdump -sass $(basename $SO)  ($(sha256sum $SO | cut -c1-16), built from kernels.cuh $(sha256sum $(dirname $SO)/kernels.cuh | cut -c1-16))"
  cuobjdump -sass "$SO" > /tmp/_sass.txt
  for m in UTMALDG UTMASTG UTMACMDFLUSH "SYNCS.*TRYWAIT" "SYNCS.ARRIVE.TRANS64" "ATOMS.OR" "MATCH.ANY" "REDUX"; do
    echo "$m: $(grep -cE "$m" /tmp/_sass.txt)"
  done
  echo
  echo "per kernel (instructions; UTMALDG / UTMASTG):"
  awk '/Function :/ {name=$3} /^ +\/\*[0-9a-f]+\*\/ / {n[name]++; if ($0 ~ /UTMALDG/) l[name]++; if ($0 ~ /UTMASTG/) s[name]++} END {for (k in n) printf "%6d %3d %3d %s\n", n[k], l[k], s[k], k}' /tmp/_sass.txt | sort -rn | head -60
} > "$OUT/sass_tma.txt"
head -12 "$OUT/sass_tma.txt"
