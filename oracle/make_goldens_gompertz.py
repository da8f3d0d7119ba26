"""TEST INFRASTRUCTURE (build container only).  Golden vectors for the comparison model of ``code/figure_5.py:222-363``:
the reference's OWN lines are read from /root/reference at run time, dedented and executed (the script itself cannot be
imported: it plots at module level and matplotlib is absent), for the script's four parameter sets (``:201-204``) over a
shortened duration ``j``; inputs + outputs go to tests/golden/gompertz.npz and the oracle is asserted against them.

    python oracle/make_goldens_gompertz.py
"""
import os
import sys
import textwrap

import numpy as np

REF = "/root/reference/code/figure_5.py"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import gompertz_oracle as GO    # noqa: E402

lines = open(REF).read().split("\n")
consts = "\n".join(lines[7:22])                                  # figure_5.py:8-22 (G ... GM)
body = textwrap.dedent("\n".join(lines[221:363]))                # :222-363 "# === Ben's model === #" ... Ltot_bg
assert "Ben's model" in body and "Ltot_bg" in body and "omass" in consts
grbs = {"Humped": [1.0, 5.0, 1.0e-3, 100.0, 1.0, 1.0e-6], "Classic": [1.0, 5.0, 1.0e-4, 1000.0, 1.0, 1.0e-6],
        "Sloped": [10.0, 5.0, 1.0e-4, 1000.0, 1.0, 1.0e-6], "Stuttering": [5.0, 5.0, 1.0e-2, 500.0, 1.0, 1.0e-6]}   # :201-204
J = 20000
out = {"names": np.array(list(grbs)), "pars": np.array(list(grbs.values())), "n_steps": J}
for name, p in grbs.items():
    ns = {"np": np}
    exec(consts, ns)
    ns["j"] = float(J)
    ns["B"], ns["P"], ns["MdiscI"], ns["RdiscI"], ns["epsilon"], ns["delta"] = p
    with np.errstate(all="ignore"):
        exec(body, ns)
    ref = np.array([ns["t"], ns["Ltot_bg"], ns["Lprop_bg"], ns["Ldip_bg"]])
    t, Ltot, Lp, Ld = GO.curves(p, J)
    mine = np.array([t, Ltot, Lp, Ld])
    same = (mine == ref) | (np.isnan(mine) & np.isnan(ref))
    print(name, "oracle == reference loop:", bool(same.all()), " Ltot range", np.nanmin(ref[1]), np.nanmax(ref[1]))
    assert same.all(), name
    out[f"{name}_curves"] = ref[:, ::20]
    out[f"{name}_last"] = ref[:, -1]
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "gompertz.npz"), **out)
print("wrote tests/golden/gompertz.npz")
