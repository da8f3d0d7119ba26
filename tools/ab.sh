#!/bin/bash
# Developer tool: A/B the CUDA library builds under build/ab/*.so on the headline workload.
#   gpurun -- 'bash tools/ab.sh [names...]'
cd "$(dirname "$0")/.."
names=${@:-$(ls build/ab/*.so | xargs -n1 basename | sed 's/\.so$//')}
for n in $names; do
  MAGPROP_B200_LIB=$PWD/build/ab/$n.so python bench.py --steps 10 --warmup 3 --no-extra --no-cpu > /tmp/ab_$n.json 2>/tmp/ab_$n.err || { echo "$n FAILED"; tail -3 /tmp/ab_$n.err; continue; }
  python - "$n" <<'PY'
import json,sys
d=json.load(open(f"/tmp/ab_{sys.argv[1]}.json"))
print("%-12s value %.4g  e2e %.4g  ms/step %.3f  mean_rhs %.1f nonfinite %d"%(sys.argv[1],d["value"],d["e2e"]["value"],d["ms_per_step"],d["roofline"]["mean_rhs_per_eval"],d["nonfinite_lnprob"]))
PY
done
