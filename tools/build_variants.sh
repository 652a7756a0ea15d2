#!/bin/bash
# Build liboptb variants with different -D switches into build_variants/ (git-ignored via *.so; travels with gpurun).
#   tools/build_variants.sh name1 "-DA=1 -DB=2" name2 "..." ...
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
OUT="$ROOT/build_variants"
mkdir -p "$OUT"
cd "$ROOT/optable_b200/csrc"
pids=()
while [ $# -ge 2 ]; do
  name="$1"; flags="$2"; shift 2
  ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared $flags -o "$OUT/liboptb_$name.so" optb.cu \
      && echo "built $name ($flags)" ) &
  pids+=($!)
  if [ ${#pids[@]} -ge 4 ]; then wait "${pids[0]}"; pids=("${pids[@]:1}"); fi
done
wait
