#!/bin/bash
# Developer tool: A/B library builds on the headline workload with the ensemble count matched to each build's
# resident capacity:  tools/ab_waves.sh name:ensembles ...
cd "$(dirname "$0")/.."
for spec in "$@"; do
  n=${spec%%:*}; e=${spec##*:}
  MAGPROP_B200_LIB=$PWD/build/ab/$n.so python bench.py --steps 10 --warmup 3 --no-extra --no-cpu --ensembles $e > /tmp/ab_$n.json 2>/tmp/ab_$n.err || { echo "$n FAILED"; tail -3 /tmp/ab_$n.err; continue; }
  python - "$n" "$e" <<'PY'
import json,sys
d=json.load(open(f"/tmp/ab_{sys.argv[1]}.json"))
print("%-10s ensembles %5s value %.4g  e2e %.4g  ms/step %.3f  mean_rhs %.1f"%(sys.argv[1],sys.argv[2],d["value"],d["e2e"]["value"],d["ms_per_step"],d["roofline"]["mean_rhs_per_eval"]))
PY
done
