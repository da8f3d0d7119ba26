#!/bin/bash
# Developer tool: A/B the CUDA library builds under build/ab/*.so on spread-out / prior-uniform ensembles.
cd "$(dirname "$0")/.."
names=${@:-$(ls build/ab/*.so | xargs -n1 basename | sed 's/\.so$//')}
for n in $names; do
  echo "== $n"
  MAGPROP_B200_LIB=$PWD/build/ab/$n.so PRIOR=1 python tools/gpu_spread.py 2>&1 | tail -5
done
