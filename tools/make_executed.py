"""profiles/r02_executed.json from the ncu counter pass of tools/profile_round2.sh (developer tool):
executed FP64 flop (2*DFMA + DMUL + DADD thread instructions, predicated on) per right-hand-side evaluation of the
explicit integrator, which bench.py multiplies by the live RHS rate for its roofline block.
    python tools/make_executed.py gpurun_out/r02_flops_headline.csv gpurun_out/r02_ncu_b2.log"""
import collections, csv, json, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
line = [l for l in open(sys.argv[2]) if l.startswith("{")][-1]
bench = json.loads(line)
W = bench["config"]["walkers_per_step_per_gpu"]                    # walkers of one step (3 datasets)
mean_rhs = bench["roofline"]["mean_rhs_per_eval"]
agg = collections.OrderedDict()
for r in rows:
    agg.setdefault((int(r[0]), r[4]), {})[r[12]] = float(r[14])
tot = collections.Counter(); per_kernel = collections.defaultdict(collections.Counter)
for (i, name), m in agg.items():
    kind = "advance_explicit" if "advance_kernel<0" in name or "advance_kernel<(bool)0" in name else (
        "advance_implicit" if "advance_kernel" in name else ("setup" if "setup" in name else "reduce"))
    for k, v in m.items():
        per_kernel[kind][k.replace("smsp__sass_thread_inst_executed_op_", "").replace("_pred_on.sum", "").replace(".sum", "")] += v
adv = per_kernel["advance_explicit"]
flop = 2 * adv["dfma"] + adv["dmul"] + adv["dadd"]
n_rhs = W * mean_rhs
out = {
    "source": "profiles/r02_flops_headline.csv (ncu --metrics smsp__sass_thread_inst_executed_op_{dfma,dmul,dadd}_pred_on.sum over one "
              "bench step: %d walkers x %.1f RHS evaluations)" % (W, mean_rhs),
    "flop_per_rhs": flop / n_rhs,
    "dfma_per_rhs": adv["dfma"] / n_rhs, "dmul_per_rhs": adv["dmul"] / n_rhs, "dadd_per_rhs": adv["dadd"] / n_rhs,
    "fp64_inst_per_rhs": (adv["dfma"] + adv["dmul"] + adv["dadd"]) / n_rhs,
    "warp_inst_per_rhs": adv["smsp__inst_executed"] * 32 / n_rhs,
    "advance_explicit_ms_per_step": adv["gpu__time_duration"] / 1e6,
    "fp64_pipe_active_pct": None,
    "stage_ms_per_step": {k: v["gpu__time_duration"] / 1e6 for k, v in per_kernel.items()},
    "stage_flop_per_eval": {k: (2 * v["dfma"] + v["dmul"] + v["dadd"]) / W for k, v in per_kernel.items()},
    "dram_bytes_per_launch": (adv["dram__bytes_read"] + adv["dram__bytes_write"]) / 3,
}
pipe = [m["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"] for (i, n), m in agg.items() if "advance_kernel<0" in n or "advance_kernel<(bool)0" in n]
out["fp64_pipe_active_pct"] = sum(pipe) / len(pipe)
json.dump(out, open("profiles/r02_executed.json", "w"), indent=1)
print(json.dumps(out, indent=1))
