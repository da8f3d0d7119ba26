"""TEST INFRASTRUCTURE (build container only).  Golden vectors for the data-preparation row (f2):
runs the UNMODIFIED reference ``code/kcorr.py:k_correction`` and ``code/clean_data.py``'s loadtxt call on
the reference's own short-GRB files and stores inputs + outputs in tests/golden/dataprep.npz.

astropy is not installed, and ``kcorr.py`` imports ``astropy.cosmology.WMAP9`` at module level, so a stub
module is planted in ``sys.modules`` for the import only; the luminosity distance handed to
``k_correction`` is therefore this repo's restatement (pinned separately against astropy's documented
WMAP9 comoving distances in tests/test_dataprep.py).

    python oracle/make_goldens_dataprep.py
"""
import os
import sys
import types

import numpy as np
import pandas as pd

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

stub = types.ModuleType("astropy"); cosmo = types.ModuleType("astropy.cosmology"); cosmo.WMAP9 = object()
stub.cosmology = cosmo
sys.modules["astropy"] = stub; sys.modules["astropy.cosmology"] = cosmo
sys.path.insert(0, os.path.join(REF, "code"))
import kcorr as ref_kcorr                                    # noqa: E402
from scipy.stats.mstats import gmean                          # noqa: E402

from magprop_b200.dataprep import luminosity_distance_cm      # noqa: E402

props = pd.read_csv(os.path.join(REF, "data", "kcorr_sgrbs.csv"), index_col="GRB")
out = {"grbs": [], "n_rows": []}
for grb in props.index.tolist():
    raw = np.loadtxt(os.path.join(REF, "data", "SGRBS", f"{grb}_raw.txt"), comments=["!", "NO", "READ"])   # clean_data.py:25
    df = pd.DataFrame(data={"t": raw[:, 0], "tpos": raw[:, 1], "tneg": raw[:, 2], "flux": raw[:, 3],
                            "fluxpos": raw[:, 4], "fluxneg": raw[:, 5]})
    out["grbs"].append(str(grb)); out["n_rows"].append(len(df))
    if str(grb) not in ("061210", "080123", "051227"):       # keep the fixture small: three bursts in full
        continue
    g, s, z = (float(props[c][grb]) for c in ("Gamma", "sigma", "z"))
    dl = luminosity_distance_cm(z)
    k = ref_kcorr.k_correction(df, g, s, z, dl)
    k["Lum50err"] = gmean([k["Lum50pos"].values, np.abs(k["Lum50neg"].values)])            # kcorr.py:107-108
    out[f"{grb}_in"] = raw
    out[f"{grb}_props"] = np.array([g, s, z, dl])
    for c in ("t", "tpos", "tneg", "Lum50", "Lum50pos", "Lum50neg", "Lum50err"):
        out[f"{grb}_{c}"] = k[c].values
out["grbs"] = np.array(out["grbs"]); out["n_rows"] = np.array(out["n_rows"])
out["Gamma"] = props["Gamma"].values; out["sigma"] = props["sigma"].values; out["z"] = props["z"].values
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "dataprep.npz"), **out)
print("wrote dataprep.npz", dict(zip(out["grbs"], out["n_rows"])))
