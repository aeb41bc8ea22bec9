#!/bin/bash
# usage: tools/sweep_variants.sh [script.py]  -- runs a perf script once per library in build_variants/ (dev tool, GPU box)
script=${1:-tools/sweep_one.py}
for so in build_variants/libdfl_*.so; do
  echo "== $so"
  DFL_B200_LIB=$so python $script 2>&1 | tail -3
done
