"""The header of a profiles/*_digest.txt: selected metrics of every kernel in an .ncu-rep (developer tool).
    python tools/ncu_summary.py rep.ncu-rep          (follow with tools/ncu_lines.py for the per-line part)"""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
WANT = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum", "smsp__sass_inst_executed_op_global_ld.sum",
        "l1tex__t_sectors_lookup_hit.sum", "l1tex__t_sectors_lookup_miss.sum"]
STALL = "smsp__average_warp_latency_issue_stalled_"     # (older ncu) / warps_issue_stalled ratios below
for r in data:
    row = dict(zip(hdr, r)); u = dict(zip(hdr, units))
    for k in WANT:
        if k in row: print(f"{k:75s} {row[k]} {u.get(k, '')}")
    pref = "smsp__average_warps_issue_stalled_"
    for k in sorted(row):
        if k.startswith(pref) and k.endswith("_per_issue_active.ratio"):
            print(f"stall/issue {k[len(pref):-len('_per_issue_active.ratio')]:40s} {row[k]}")
    print()
