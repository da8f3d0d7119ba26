"""Which source lines issue the FP64 multiplies / adds / FMAs (developer tool):
    python tools/ncu_ops.py rep.ncu-rep [OPCODE ...]      default: DMUL DADD DFMA DSETP"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; ops = sys.argv[2:] or ["DMUL", "DADD", "DFMA", "DSETP"]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; fname = ""; cur = None
agg = {o: collections.Counter() for o in ops}; src = {}
tot = collections.Counter()
for r in rows:
    if len(r) == 2 and r[0] in ("File Path", "File Name"):
        fname = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r; ie = hdr.index("Instructions Executed"); continue
    if not hdr or len(r) != len(hdr): continue
    if r[0].isdigit():
        cur = (fname, int(r[0])); src[cur] = r[1].strip(); continue
    if r[0] == "" and cur:
        sass = r[3].strip(); op = sass.split()[0].split(".")[0] if sass else ""
        if sass.startswith("@"): op = sass.split()[1].split(".")[0]
        try: n = int(r[ie])
        except ValueError: continue
        tot[op] += n
        if op in agg: agg[op][cur] += n
allw = sum(tot.values())
print("warp instructions", allw, " ".join(f"{o} {tot[o]} ({100*tot[o]/allw:.1f}%)" for o in ops))
for o in ops:
    print(f"\n== {o}")
    for (f, ln), n in agg[o].most_common(28):
        print(f"{100*n/tot[o]:5.1f}%  {n:>11,}  {f}:{ln}  {src[(f, ln)][:105]}")
