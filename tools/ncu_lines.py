"""Per-source-line digest of an .ncu-rep captured with --import-source on (developer tool):
    python tools/ncu_lines.py rep.ncu-rep [top]
Lists, for the source lines with the most issued (warp) instructions: the line, warp instructions, average
active threads per instruction, and the share of all issued instructions -- i.e. where a divergent launch
spends its issue slots."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; fname = ""
lines = []
for r in rows:
    if len(r) == 2 and r[0] in ("File Path", "File Name"):
        fname = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        try:
            wi = int(r[hdr.index("Instructions Executed")]); ti = int(r[hdr.index("Predicated-On Thread Instructions Executed")])
        except ValueError:
            continue
        if wi: lines.append((wi, ti, fname, int(r[0]), r[1].strip()))
tot_w = sum(l[0] for l in lines); tot_t = sum(l[1] for l in lines)
print(f"total warp instructions {tot_w:,}  active threads per instruction {tot_t / tot_w:.2f}")
for wi, ti, f, ln, src in sorted(lines, reverse=True)[:top]:
    print(f"{100 * wi / tot_w:5.1f}%  {wi:>12,}  thr {ti / wi:5.1f}  {f}:{ln}  {src[:110]}")

# ---- by region of magprop_core.cuh (line ranges of the functions on the integrator's path)
REGIONS = [("bits/table_locate/poly/ld2", 108, 233), ("exp_c", 313, 339), ("pow_m17_fast/cold", 341, 387), ("pow_m17_seeded(1)", 388, 438),
           ("rsqrt_pos(2)/rcp_pos", 656, 696), ("exp_small", 697, 720), ("exp_small10+rcp_pos2", 721, 751), ("spin_g", 790, 850),
           ("dense_eval", 875, 886), ("controller", 970, 986), ("breakup_sliding", 1014, 1048), ("locate_kink", 1049, 1091),
           ("step_spin_chain", 1092, 1196), ("disc_stages<N>", 1197, 1238), ("disc_stages_dp5", 1239, 1277),
           ("integrator_step", 1278, 1316), ("spin_fJ", 1326, 1376), ("radau_step", 1377, 1506), ("load/drain_nodes", 1606, 1655),
           ("luminosity stage", 596, 655)]
agg = collections.OrderedDict((n, [0, 0]) for n, _, _ in REGIONS)
other = collections.Counter(); other_t = collections.Counter()
for wi, ti, f, ln, src in lines:
    hit = False
    if f == "magprop_core.cuh":
        for n, a, b in REGIONS:
            if a <= ln <= b:
                agg[n][0] += wi; agg[n][1] += ti; hit = True; break
    if not hit:
        other[f] += wi; other_t[f] += ti
print("\nby region:")
for n, (wi, ti) in agg.items():
    if wi: print(f"{100 * wi / tot_w:5.1f}%  {wi:>13,}  thr {ti / wi:5.1f}  {n}")
for f in other:
    print(f"{100 * other[f] / tot_w:5.1f}%  {other[f]:>13,}  thr {other_t[f] / other[f]:5.1f}  ({f})")
