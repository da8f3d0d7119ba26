"""Digest an .ncu-rep (one kernel) into the numbers DESIGN.md / profiles/ quote (developer tool).

    python tools/ncu_digest.py gpurun_out/prof.ncu-rep [walkers rhs_per_walker] [--hot N] [--csv out.csv]

Prints launch stats, pipe/issue utilisation, stall mix per issued instruction, the SASS opcode mix
(per RHS evaluation when walkers/rhs are given) and optionally the N hottest SASS lines by samples.
"""
import collections
import csv
import io
import subprocess
import sys


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    rep = args[0]
    per = None
    if len(args) >= 3:
        per = (int(args[1]) / 32.0) * float(args[2])          # warp-level RHS evaluations
    hot = int(sys.argv[sys.argv.index("--hot") + 1]) if "--hot" in sys.argv else 0
    out_csv = sys.argv[sys.argv.index("--csv") + 1] if "--csv" in sys.argv else None
    out_json = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
    raw = ncu_csv(rep, "raw")
    hdr, units, row = raw[0], raw[1], raw[2]
    get = lambda k: row[hdr.index(k)] if k in hdr else "n/a"
    keys = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "sm__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "smsp__thread_inst_executed_per_inst_executed.ratio",
            "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
            "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
            "smsp__sass_inst_executed_op_global_ld.sum", "l1tex__t_sectors_lookup_hit.sum", "l1tex__t_sectors_lookup_miss.sum"]
    lines = []
    for k in keys:
        v = get(k)
        if v == "n/a":
            for h in hdr:
                if h.endswith(k):
                    v = row[hdr.index(h)]
        lines.append((k, v, units[hdr.index(k)] if k in hdr else ""))
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
            lines.append((h.replace("smsp__average_warps_issue_stalled_", "stall/issue ").replace("_per_issue_active.ratio", ""),
                          row[hdr.index(h)], ""))
    for k, v, u in lines:
        print(f"{k:75s} {v} {u}")
    src = ncu_csv(rep, "source", ("--print-source", "sass"))
    sh = src[1]
    ia, isrc, ismp = sh.index("Instructions Executed"), sh.index("Source"), sh.index("# Samples")
    by, smp, tot, tots = collections.Counter(), collections.Counter(), 0, 0
    rows = []
    for r in src[2:]:
        if len(r) <= ia:
            continue
        n, s = int(r[ia]), r[isrc].strip()
        toks = s.split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        by[op] += n; smp[op] += int(r[ismp]); tot += n; tots += int(r[ismp])
        rows.append((int(r[ismp]), n, s))
    print(f"\ntotal warp instructions {tot}  samples {tots}" + (f"  per warp-RHS {tot / per:.1f}" if per else ""))
    fp64 = sum(by[o] for o in ("DFMA", "DMUL", "DADD", "DSETP"))
    print(f"FP64-pipe instructions (DFMA+DMUL+DADD+DSETP) {fp64} = {100.0 * fp64 / tot:.1f}%" + (f"  per warp-RHS {fp64 / per:.1f}" if per else ""))
    for op, n in by.most_common(24):
        print(f"  {op:10s} {n:12d} {100.0 * n / tot:5.1f}%" + (f"  per-RHS {n / per:6.2f}" if per else "") + f"  samples {100.0 * smp[op] / max(tots, 1):5.1f}%")
    if out_json and per:
        import json
        val = {k: v for k, v, _ in lines}
        dram = (float(val["dram__bytes_read.sum"]) * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}[units[hdr.index("dram__bytes_read.sum")]]
                + float(val["dram__bytes_write.sum"]) * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}[units[hdr.index("dram__bytes_write.sum")]])
        with open(out_json, "w") as f:
            json.dump({"source": f"{out_csv or rep} (ncu --set full, {args[1]} walkers x {args[2]} RHS evaluations)",
                       "flop_per_rhs": round((2 * by["DFMA"] + by["DMUL"] + by["DADD"]) / per, 1),
                       "fp64_inst_per_rhs": round(fp64 / per, 1),
                       "fp64_pipe_active_pct": round(float(val["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]), 1),
                       "issue_active_pct": round(float(val["smsp__issue_active.avg.pct_of_peak_sustained_active"]), 1),
                       "warp_inst_per_rhs": round(tot / per, 1), "dram_bytes_per_launch": int(dram)}, f, indent=1)
    if hot:
        print("\nhottest SASS lines by stall samples")
        for s_, n, txt in sorted(rows, reverse=True)[:hot]:
            print(f"  {s_:6d} {n:10d}  {txt}")
    if out_csv:
        with open(out_csv, "w", newline="") as f:
            wr = csv.writer(f)
            wr.writerow(["metric", "value", "unit"])
            for k, v, u in lines:
                wr.writerow([k, v, u])
            wr.writerow(["total_warp_instructions", tot, "inst"])
            wr.writerow(["fp64_pipe_warp_instructions", fp64, "inst"])
            for op, n in by.most_common(24):
                wr.writerow([f"sass_op_{op}", n, "inst"])


if __name__ == "__main__":
    main()
