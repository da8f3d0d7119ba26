#!/bin/bash
# Developer tool: the standard GPU cycle (run through gpurun): parity tests, bench, one full ncu capture.
#   gpurun --timeout 1200 -- 'bash tools/gpu_cycle.sh TAG [full]'
TAG=${1:-dev}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("value %.4g e2e %.4g ms/step %.3f frac %.3f mean_rhs %.1f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["roofline"]["frac"],d["roofline"]["mean_rhs_per_eval"]))
for k,v in d["extra"].items(): print(" ",k,{a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items()})
PY
if [ "$2" = "full" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-extra --no-cpu > gpurun_out/ncu_l_$TAG.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:advance_kernel -c 1 -s 2 -o gpurun_out/full_$TAG python bench.py --steps 1 --warmup 1 --no-extra --no-cpu > gpurun_out/ncu_f_$TAG.log 2>&1
fi
