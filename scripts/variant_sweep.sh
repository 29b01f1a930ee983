#!/bin/bash
# usage: scripts/variant_sweep.sh OUTFILE CONFIG KERNELS [variant ...]  -- kernel_bench for the default build and each variant
out=$1; cfg=$2; only=$3; shift 3
: > "$out"
echo "== default" >> "$out"
python scripts/kernel_bench.py --config "$cfg" --only "$only" >> "$out" 2>&1
for v in "$@"; do
  echo "== $v" >> "$out"
  BG_LIB_PATH=breedgym_b200/_variants/$v.so python scripts/kernel_bench.py --config "$cfg" --only "$only" >> "$out" 2>&1
done
