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

# ---- by region of magprop_core.cuh: every region runs from the line of its marker to the next marker's
import os, re
MARKS = [("bits/table_locate/poly/ld2", r"^MP_HD int64_t dbits"), ("disc tables: series/asymptotics", r"^MP_HD double disc_S_series|^// ---- S\(u\)"),
         ("exp_c", r"^MP_HD double exp_c"), ("pow_m17_fast/cold", r"^MP_HD double pow_m17_fast"), ("pow_m17_seeded(1)", r"^MP_HD double pow_m17_seeded\("),
         ("rcp/rsqrt_fast", r"^MP_HD double rcp_fast"), ("walker setup / cold rhs", r"^struct Walker|^MP_HD void walker_setup"),
         ("luminosity stage", r"^MP_HD Lum luminosity"), ("rsqrt_pos(2)/rcp_pos", r"^MP_HD double rsqrt_pos\("), ("exp_small", r"^MP_HD double exp_small\("),
         ("exp_small10+rcp_pos2", r"^MP_HD double exp_small10"), ("spin_f", r"^struct StageDisc"), ("spin_g", r"^MP_HD double spin_g\(const Spec& sp, const Walker& w, const StageDisc& d, double y, unsigned& side, unsigned& regime"),
         ("step control", r"^struct StepControl"), ("Integrator/dense_eval/init", r"^struct Integrator"), ("DP tableau", r"^MP_CONST_QUALIFIER double kDP"),
         ("breakup_sliding", r"^static bool breakup_sliding_block\("), ("locate_kink", r"^static double locate_kink"), ("step_spin_chain", r"^MP_HD int step_spin_chain"),
         ("disc_stages<N>", r"^MP_HD void disc_stages\("), ("disc_stages_dp5", r"^MP_HD void disc_stages_dp5"), ("integrator_step", r"^MP_HD bool integrator_step"),
         ("radau (spin_fJ, radau_step)", r"^struct RadauC"), ("records / load / drain_nodes", r"^struct WalkerRec|^// ---- one walker, in three stages"),
         ("reduce_rows", r"^MP_HD double reduce_rows")]
core = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "magprop_b200", "csrc", "magprop_core.cuh")
src_lines = open(core).read().splitlines()
starts = []
for n, pat in MARKS:
    hit = next((i + 1 for i, l in enumerate(src_lines) if re.search(pat, l) and not l.rstrip().endswith(";")), None)   # (not a declaration)
    if hit: starts.append((hit, n))
starts.sort()
REGIONS = [(n, a, (starts[i + 1][0] - 1 if i + 1 < len(starts) else len(src_lines))) for i, (a, n) in enumerate(starts)]
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
