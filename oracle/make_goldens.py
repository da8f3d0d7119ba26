"""Generate tests/golden/*.npz from the UNMODIFIED reference and pin the oracle.

Run in the build container only (needs /root/reference):

    python oracle/make_goldens.py

What it does
  1. imports the reference's two copies of the hot path
       code/synthetic_datasets/{funcs,mcmc_eqns}.py   ("script")
       magnetar/{funcs,mcmc_eqns}.py                  ("packaged")
  2. converts the reference's own hot-path fixtures
       tests/test_data/odes_integrated_by_odeint.csv, model_light_curve.csv
     to .npz (reference golden vectors)
  3. evaluates the reference on seeded parameter batches and stores its outputs
     (state, light curves, model-at-data, lnprior, lnprob), plus the same
     quantities from the *tight* oracle (rtol 1e-13) used to localise LSODA's
     own truncation error
  4. asserts oracle/magprop_oracle.py reproduces every reference number it just
     stored (<= 4 ulp-level relative difference) -- this is what "parity
     pinned" in the oracle header refers to.

The only reference lines restated (not imported) are
code/synthetic_datasets/mcmc_eqns.py:16-25,66-81, because ``mod == 'flag'`` on
an ndarray raises under numpy >= 2.
"""
import os
import sys
import tempfile
import warnings
from multiprocessing import Pool

sys.dont_write_bytecode = True
REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(REF, "code", "synthetic_datasets"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

import numpy as np
import pandas as pd

warnings.filterwarnings("ignore")
import funcs as ref_script            # noqa: E402  (reference, script variant)
import mcmc_eqns as ref_script_mc     # noqa: E402
import magnetar.funcs as ref_pkg      # noqa: E402  (reference, packaged variant)
import magnetar.mcmc_eqns as ref_pkg_mc  # noqa: E402

from oracle import magprop_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
SUB = 20     # curve goldens keep every 20th grid node (+ the last one)


def ref_script_lnprob(theta, x, y, yerr):
    """mcmc_eqns.py:52-81 with :22 restated for numpy>=2; returns (lnp, flagged)."""
    lp = ref_script_mc.lnprior(theta)
    if not np.isfinite(lp):
        return -np.inf, False
    arr = np.array(theta)
    arr[2:] = 10.0 ** arr[2:]
    mod = ref_script.model_lum(arr, xdata=x)
    if isinstance(mod, str):
        return -np.inf, True
    ll = -0.5 * np.sum(((y - mod) / yerr) ** 2.0)
    if not np.isfinite(ll):
        return -np.inf, False
    return ll + lp, False


def _task_script(a):
    theta, x, y, yerr = a
    ref, flagged = ref_script_lnprob(theta, x, y, yerr)
    orc = O.lnprob(theta, x, y, yerr, O.script_spec(), O.SCRIPT_LOWER, O.SCRIPT_UPPER)
    tight = O.lnprob(theta, x, y, yerr, O.script_spec(), O.SCRIPT_LOWER, O.SCRIPT_UPPER, tight=True)
    return ref, flagged, orc, tight


def _task_curve(a):
    variant, pars = a
    if variant == "script":
        ref = ref_script.model_lum(pars)
        spec = O.script_spec()
    else:
        ref = ref_pkg.model_lc(pars)
        spec = O.packaged_spec()
    grid = spec.grid()
    s_def, ok, _ = O.integrate(pars, spec)
    s_tight, ok2, _ = O.integrate(pars, spec, tight=True)
    if isinstance(ref, str):
        return None
    lt = O.luminosity(s_tight, pars, spec)
    ld = O.luminosity(s_def, pars, spec)
    return (ref, s_def, s_tight, np.array(ld) / 1e50, np.array(lt) / 1e50)


def sub(a):
    idx = np.unique(np.append(np.arange(0, a.shape[-1], SUB), a.shape[-1] - 1))
    return a[..., idx], idx


def relerr(a, b):
    a = np.asarray(a, float); b = np.asarray(b, float)
    den = np.maximum(np.abs(a), np.abs(b))
    with np.errstate(all="ignore"):
        r = np.where(den > 0, np.abs(a - b) / den, 0.0)
    r[np.isinf(a) & np.isinf(b) & (a == b)] = 0.0
    return r


def main():
    os.makedirs(OUT, exist_ok=True)
    pool = Pool(os.cpu_count())
    rng = np.random.RandomState(O.SYNTH_SEED)

    # ---- 2. reference fixtures ------------------------------------------------
    f1 = pd.read_csv(os.path.join(REF, "tests/test_data/odes_integrated_by_odeint.csv"), index_col=False)
    f2 = pd.read_csv(os.path.join(REF, "tests/test_data/model_light_curve.csv"), index_col=False)
    np.savez_compressed(
        os.path.join(OUT, "reference_fixtures.npz"),
        odes_t=f1["t"].values, odes_Mdisc=f1["Mdisc"].values, odes_omega=f1["omega"].values,
        odes_pars=np.array([1.0, 1.0, 1e-3, 100.0, 1.0, 10.0]),      # tests/test_funcs.py:37-43 (B,P,M,R,eps,delta)
        lc_t=f2["t"].values, lc_Ltot=f2["Ltot"].values, lc_Lprop=f2["Lprop"].values, lc_Ldip=f2["Ldip"].values,
        lc_pars=np.array([1.0, 5.0, 1e-3, 100.0, 0.1, 1.0]),         # tests/test_funcs.py:55
    )
    # the oracle must satisfy the reference's own assertions on them (np.isclose defaults)
    spec = O.packaged_spec()
    soln, ok, _ = O.integrate([1.0, 1.0, 1e-3, 100.0, 1.0, 10.0], spec, grid=f1["t"].values)
    assert ok and np.isclose(soln[:, 0], f1["Mdisc"]).all() and np.isclose(soln[:, 1], f1["omega"]).all()
    lc = O.model([1.0, 5.0, 1e-3, 100.0, 0.1, 1.0], spec)
    for i, c in ((0, "t"), (1, "Ltot"), (2, "Lprop"), (3, "Ldip")):
        assert np.isclose(lc[i], f2[c]).all(), c
    print("oracle satisfies reference fixtures:",
          relerr(soln[:, 1], f1["omega"]).max(), relerr(lc[1], f2["Ltot"]).max())
    assert O.init_conds(0.001, 1.0) == tuple(ref_pkg.init_conds(0.001, 1.0))   # tests/test_funcs.py:12-25
    assert O.init_conds(0.001, 1.0) == ref_script.init_conds(0.001, 1.0)

    # ---- 3a. curves -------------------------------------------------------------
    names = list(O.SYNTH_TRUTHS)
    extra_log = rng.uniform(O.SCRIPT_LOWER, O.SCRIPT_UPPER, size=(24, 6))
    extra = extra_log.copy(); extra[:, 2:] = 10.0 ** extra[:, 2:]
    for variant in ("script", "packaged"):
        plist = [O.SYNTH_TRUTHS[n] for n in names] + list(extra)
        res = pool.map(_task_curve, [(variant, p) for p in plist])
        keep = [i for i, r in enumerate(res) if r is not None]
        pars = np.array([plist[i] for i in keep])
        ref = np.array([res[i][0] for i in keep])           # (N,4,G)
        s_def = np.array([res[i][1] for i in keep])         # (N,G,2)
        s_tight = np.array([res[i][2] for i in keep])
        l_def = np.array([res[i][3] for i in keep])         # (N,3,G): Ltot,Lprop,Ldip /1e50
        l_tight = np.array([res[i][4] for i in keep])
        # oracle == reference (curves come from the same LSODA run)
        e = relerr(ref[:, 1:], l_def).max()
        print(f"[{variant}] oracle vs reference curves: max rel {e:.2e} over {len(keep)} parameter sets")
        assert e < 1e-12
        r_s, idx = sub(ref)
        np.savez_compressed(
            os.path.join(OUT, f"curves_{variant}.npz"),
            pars=pars, node_index=idx,
            ref_curves=r_s,                                  # reference (t, Ltot, Lprop, Ldip)/1e50
            state_default=sub(np.moveaxis(s_def, 1, 2))[0],  # (N,2,Gs)
            state_tight=sub(np.moveaxis(s_tight, 1, 2))[0],
            lum_tight=sub(l_tight)[0],
            n_named=len(names), names=np.array(names),
        )

    # ---- 3b. synthetic datasets + script lnprob ---------------------------------
    data = {}
    for n in names:
        curves = ref_script.model_lum(O.SYNTH_TRUTHS[n])
        data[n] = O.synth_dataset(n, curves)
    thetas, owner = [], []
    for n in names:
        ball = O.SYNTH_TRUTHS_LOG[n] + 1e-4 * rng.randn(48, 6)          # synth_mcmc.py:175-176
        wide = O.SYNTH_TRUTHS_LOG[n] + 0.05 * rng.randn(48, 6)          # burnt-in posterior-ish spread
        uni = rng.uniform(O.SCRIPT_LOWER, O.SCRIPT_UPPER, size=(96, 6))  # prior box, full stiffness range
        edge = np.array([O.SCRIPT_LOWER, O.SCRIPT_UPPER,
                         np.nextafter(O.SCRIPT_LOWER, -np.inf), np.nextafter(O.SCRIPT_UPPER, np.inf),
                         np.where(np.arange(6) == 3, np.nan, O.SYNTH_TRUTHS_LOG[n])])
        for th in np.vstack([ball, wide, uni, edge]):
            thetas.append(th); owner.append(names.index(n))
    thetas = np.array(thetas); owner = np.array(owner)
    res = pool.map(_task_script, [(th, *data[names[o]]) for th, o in zip(thetas, owner)], chunksize=4)
    ref_lnp = np.array([r[0] for r in res]); flagged = np.array([r[1] for r in res])
    orc_lnp = np.array([r[2] for r in res]); tight_lnp = np.array([r[3] for r in res])
    e = relerr(ref_lnp, orc_lnp)
    print(f"[script] oracle vs reference lnprob: max rel {e.max():.2e}; flagged {flagged.sum()} / {len(res)}; "
          f"-inf {np.isinf(ref_lnp).sum()}")
    assert e.max() < 1e-12 and (np.isinf(ref_lnp) == np.isinf(orc_lnp)).all()
    fin = np.isfinite(ref_lnp) & np.isfinite(tight_lnp)
    et = relerr(ref_lnp[fin], tight_lnp[fin])
    print(f"[script] default-vs-tight lnprob: median {np.median(et):.2e} p99 {np.percentile(et, 99):.2e} max {et.max():.2e}")
    lnprior = np.array([ref_script_mc.lnprior(th) for th in thetas])
    np.savez_compressed(
        os.path.join(OUT, "lnprob_script.npz"),
        names=np.array(names), theta=thetas, dataset=owner,
        ref_lnprob=ref_lnp, ref_flagged=flagged, tight_lnprob=tight_lnp, ref_lnprior=lnprior,
        **{f"{n}_{c}": data[n][i] for n in names for i, c in enumerate(("x", "y", "yerr"))},
    )

    # ---- 3c. packaged variant: model at data times on the "S" grid, 6..9 parameters
    os.chdir(REF)  # magnetar/mcmc_eqns.py:55 reads a cwd-relative CSV
    tS = np.sort(10.0 ** rng.uniform(-2.5, 5.5, size=120))
    truthS = np.array([2.0, 2.0, 5e-3, 300.0, 3.0, 2.0])
    yS = ref_pkg.model_lc(truthS, xdata=tS, GRBtype="S")
    yerrS = 0.2 * yS
    yS = yS + rng.normal(0.0, yerrS)
    frame = pd.DataFrame({"t": tS, "Lum50": yS, "Lum50err": yerrS})
    lims = pd.DataFrame({"pars": ["B", "P", "MdiscI", "RdiscI", "epsilon", "delta", "dipeff", "propeff", "f_beam"],
                         "lower": [1e-3, 0.69, 1e-5, 50.0, 0.1, 1e-3, 0.01, 0.01, 1.0],
                         "upper": [10.0, 10.0, 1e-1, 2000.0, 1000.0, 50.0, 1.0, 1.0, 600.0]})
    with tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False) as fh:
        lims.to_csv(fh, index=False)
        lim_path = fh.name
    lo9, hi9 = lims["lower"].values, lims["upper"].values
    pk_theta, pk_ref, pk_orc, pk_model = [], [], [], []
    for ndim in (6, 7, 8, 9):
        lo, hi = O.prior_bounds("packaged", ndim, lo9, hi9)
        draws = np.exp(rng.uniform(np.log(lo), np.log(hi), size=(10, ndim)))
        draws[0, :6] = truthS
        draws[-1, 0] = hi[0] * 1.5          # out of prior
        for th in draws:
            r = ref_pkg_mc.lnprob(th, frame, "S", custom_lims=lim_path)
            spec = O.packaged_spec("S")
            o = O.lnprob(th, tS, yS, yerrS, spec, lo, hi)
            pk_theta.append(np.pad(th, (0, 9 - ndim), constant_values=np.nan))
            pk_ref.append(r); pk_orc.append(o)
    pk_ref = np.array(pk_ref); pk_orc = np.array(pk_orc)
    e = relerr(pk_ref, pk_orc)
    print(f"[packaged] oracle vs reference lnprob (6..9 pars, S grid): max rel {e.max():.2e}")
    assert e.max() < 1e-12
    mod_truth = ref_pkg.model_lc(truthS, xdata=tS, GRBtype="S")
    # default CSV limits: log-space bounds fed raw to a linear-space model
    dflt_theta = np.array([[1.0, 5.0, -2.0, 2.0, 0.5, 0.3], [3.0, 1.0, -1.5, 3.0, 2.0, -4.0], [1.0, 5.0, -2.0, 2.0, 0.5, 2.0]])
    dflt_ref = np.array([ref_pkg_mc.lnprob(th, frame, "S") for th in dflt_theta])
    print("[packaged] default-limits lnprob:", dflt_ref, " (-0.5*sum((y/yerr)^2) =", -0.5 * np.sum((yS / yerrS) ** 2), ")")
    os.unlink(lim_path)
    np.savez_compressed(
        os.path.join(OUT, "lnprob_packaged.npz"),
        t=tS, Lum50=yS, Lum50err=yerrS, truth=truthS, model_at_truth=mod_truth,
        lims_lower=lo9, lims_upper=hi9, theta=np.array(pk_theta), ref_lnprob=pk_ref,
        default_theta=dflt_theta, default_ref_lnprob=dflt_ref,
    )
    pool.close()
    print("goldens written to", OUT)


if __name__ == "__main__":
    main()
