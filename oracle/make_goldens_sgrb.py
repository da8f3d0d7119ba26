"""TEST INFRASTRUCTURE (build container only).  BASELINE configs[2] on the REAL short-GRB sample:
the 15 bursts of the reference's ``data/kcorr_sgrbs.csv`` cleaned (``code/clean_data.py:24-30``) and
k-corrected by the UNMODIFIED reference ``code/kcorr.py:k_correction`` (``:12-55``, ``:107-108``), and the
UNMODIFIED reference ``magnetar.lnprob(pars, data, "S", custom_lims=...)`` (``magnetar/mcmc_eqns.py:87-119``)
evaluated on 32 walkers per burst.  Output: tests/golden/sgrb_sample.npz.

    python oracle/make_goldens_sgrb.py

* astropy is absent and ``kcorr.py`` imports ``astropy.cosmology.WMAP9`` at module level, so a stub module
  is planted for the import only; the luminosity distance is this repo's WMAP9 restatement (pinned to
  astropy's documented distances in tests/test_dataprep.py).
* The packaged ``lnlike`` passes its parameters to the model un-logged (``magnetar/mcmc_eqns.py:22-34``)
  while the shipped ``mcmc_limits.csv`` holds log10 bounds, so the fits use a custom
  limits file in linear space -- the ``custom_lims`` argument the reference provides for exactly that.
* GRB 060614 has rest-frame times beyond the grid's 1e6 s: the reference raises ``ValueError`` from
  ``interp1d`` (``magnetar/funcs.py:214``).  That is recorded (``raises``), and the burst is ALSO stored cut
  at t <= 1e6 s (what a user has to do to fit it) with goldens on the cut data.
* Walkers per burst: the 24 best of 192 log-uniform prior draws by chi-square (so the model is of the
  data's order of magnitude and the chi-square is sensitive to it) + 7 arbitrary prior draws + 1 outside.
"""
import os
import sys
import tempfile
import types
import warnings
from multiprocessing import Pool

import numpy as np
import pandas as pd

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

stub = types.ModuleType("astropy"); cosmo = types.ModuleType("astropy.cosmology"); cosmo.WMAP9 = object()
stub.cosmology = cosmo
sys.modules["astropy"] = stub; sys.modules["astropy.cosmology"] = cosmo
sys.path.insert(0, os.path.join(REF, "code"))
sys.path.insert(0, REF)
import kcorr as ref_kcorr                                     # noqa: E402  (reference, unmodified)
import magnetar.mcmc_eqns as ref_pkg_mc                       # noqa: E402  (reference, unmodified)
from scipy.stats.mstats import gmean                           # noqa: E402

from magprop_b200.dataprep import clean_raw, k_correct_grb, luminosity_distance_cm   # noqa: E402
from oracle import magprop_oracle as O                         # noqa: E402

LIMS = pd.DataFrame({"pars": ["B", "P", "MdiscI", "RdiscI", "epsilon", "delta", "dipeff", "propeff", "f_beam"],
                     "lower": [1e-3, 0.69, 1e-5, 50.0, 0.1, 1e-5, 0.01, 0.01, 1.0],
                     "upper": [10.0, 10.0, 1e-1, 2000.0, 1000.0, 50.0, 1.0, 1.0, 600.0]})
N_SCAN, N_BEST, N_OTHER = 192, 24, 7
T_MAX = 1.0e6


def _scan(a):
    th, t, y, e = a
    return O.lnprob(th, t, y, e, O.packaged_spec("S"), LIMS["lower"].values[:6], LIMS["upper"].values[:6])


def _ref(a):
    th, t, y, e, lim_path = a
    os.chdir(REF)
    frame = pd.DataFrame({"t": t, "Lum50": y, "Lum50err": e})
    try:
        return float(ref_pkg_mc.lnprob(th, frame, "S", custom_lims=lim_path)), 0
    except ValueError:          # interp1d: a data time outside the grid (magnetar/funcs.py:214)
        return np.nan, 1
    except TypeError:           # odeint flagged: the packaged lnlike has no 'flag' branch, `y - "flag"` raises
        return np.nan, 2        # (magnetar/mcmc_eqns.py:22-37)


def _tight(a):
    th, t, y, e = a
    return O.lnprob(th, t, y, e, O.packaged_spec("S"), LIMS["lower"].values[:6], LIMS["upper"].values[:6], tight=True)


def main():
    props = pd.read_csv(os.path.join(REF, "data", "kcorr_sgrbs.csv"), index_col="GRB")
    rng = np.random.RandomState(20170613)
    lo, hi = LIMS["lower"].values[:6], LIMS["upper"].values[:6]
    with tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False) as fh:
        LIMS.to_csv(fh, index=False)
        lim_path = fh.name
    pool = Pool(os.cpu_count())
    out = {"grbs": np.array([str(g) for g in props.index]), "lims_lower": LIMS["lower"].values,
           "lims_upper": LIMS["upper"].values, "Gamma": props["Gamma"].values, "sigma": props["sigma"].values,
           "z": props["z"].values}
    raises = []
    for grb in out["grbs"]:
        raw_path = os.path.join(REF, "data", "SGRBS", f"{grb}_raw.txt")
        raw = np.loadtxt(raw_path, comments=["!", "NO", "READ"])                     # clean_data.py:25
        df = pd.DataFrame(data={"t": raw[:, 0], "tpos": raw[:, 1], "tneg": raw[:, 2], "flux": raw[:, 3],
                                "fluxpos": raw[:, 4], "fluxneg": raw[:, 5]})
        g, s, z = (float(props[c][grb]) for c in ("Gamma", "sigma", "z"))
        dl = luminosity_distance_cm(z)
        k = ref_kcorr.k_correction(df, g, s, z, dl)                                  # reference arithmetic
        k["Lum50err"] = gmean([k["Lum50pos"].values, np.abs(k["Lum50neg"].values)])  # kcorr.py:107-108
        # this repo's data preparation reproduces the reference's on every burst
        mine = k_correct_grb(clean_raw(raw_path), g, s, z)
        for c in ("t", "Lum50", "Lum50pos", "Lum50neg"):
            assert np.array_equal(mine[c], k[c].values), (grb, c)
        assert np.allclose(mine["Lum50err"], k["Lum50err"].values, rtol=4e-16, atol=0), grb
        t, y, e = k["t"].values, k["Lum50"].values, k["Lum50err"].values
        out[f"{grb}_raw"] = raw
        out[f"{grb}_t"], out[f"{grb}_Lum50"], out[f"{grb}_Lum50err"] = t, y, e
        # does the reference accept the burst as it is?
        probe = np.array([1.0, 5.0, 1e-3, 100.0, 1.0, 1.0])
        _, code = _ref((probe, t, y, e, lim_path))
        raised = code == 1
        raises.append(raised)
        keep = t <= T_MAX if raised else np.ones(t.size, bool)
        tc, yc, ec = t[keep], y[keep], e[keep]
        out[f"{grb}_keep"] = keep
        # walkers
        scan = np.exp(rng.uniform(np.log(lo), np.log(hi), size=(N_SCAN, 6)))
        lp = np.array(pool.map(_scan, [(th, tc, yc, ec) for th in scan], chunksize=2))
        order = np.argsort(-np.where(np.isfinite(lp), lp, -np.inf))
        best = scan[order[:N_BEST]]
        other = np.exp(rng.uniform(np.log(lo), np.log(hi), size=(N_OTHER, 6)))
        outside = best[0].copy(); outside[1] = 0.5                                    # P below the 0.69 ms floor
        theta = np.vstack([best, other, outside[None, :]])
        ref = pool.map(_ref, [(th, tc, yc, ec, lim_path) for th in theta])
        assert not any(r[1] == 1 for r in ref), grb
        ref_flag = np.array([r[1] == 2 for r in ref])            # the reference raises TypeError for these walkers
        ref_lnp = np.array([-np.inf if r[1] == 2 else r[0] for r in ref])
        tight = np.array(pool.map(_tight, [(th, tc, yc, ec) for th in theta]))
        orc = np.array([O.lnprob(th, tc, yc, ec, O.packaged_spec("S"), lo, hi) for th in theta])
        fin = np.isfinite(ref_lnp)
        assert (np.isfinite(orc) == fin).all(), grb
        err = np.abs(orc[fin] - ref_lnp[fin]) / np.abs(ref_lnp[fin])
        assert err.max() < 1e-12, (grb, err.max())
        et = np.abs(tight[fin] - ref_lnp[fin]) / np.abs(ref_lnp[fin])
        out[f"{grb}_theta"], out[f"{grb}_ref_lnprob"], out[f"{grb}_tight_lnprob"] = theta, ref_lnp, tight
        out[f"{grb}_ref_flagged"] = ref_flag
        print(f"{grb}: D={t.size} (kept {keep.sum()}) raises={raised}  t=[{t.min():.3g}, {t.max():.3g}]  "
              f"best lnprob {ref_lnp[0]:.4g}  flagged {ref_flag.sum()}  -inf {np.isinf(ref_lnp).sum()}  oracle-vs-ref {err.max():.1e}  "
              f"default-vs-tight max {et.max():.1e}")
    out["raises"] = np.array(raises)
    os.unlink(lim_path)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sgrb_sample.npz"), **out)
    print("wrote tests/golden/sgrb_sample.npz; ValueError bursts:", [g for g, r in zip(out["grbs"], raises) if r])


if __name__ == "__main__":
    main()
