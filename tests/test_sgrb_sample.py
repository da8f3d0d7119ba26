"""BASELINE configs[2] on the REAL short-GRB sample (tests/golden/sgrb_sample.npz, written by
oracle/make_goldens_sgrb.py from the reference's own data files, k-correction and ``magnetar.lnprob``).

CPU: the data preparation reproduces the reference's k-corrected sample from the raw rows on all 15 bursts;
the oracle reproduces the stored reference lnprob; GRB 060614's out-of-grid times are an error.
GPU: ``magnetar.lnprob_batch`` (packaged model, "S" grid, custom limits) against the reference values on
32 walkers per burst, incl. the 1944-point burst; the per-burst error statistics against the default- and
the converged-tolerance oracle are reported (warnings summary + gpurun_out/parity_report.json).
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, parity_stats, relerr, report

TOL_TIGHT = 5e-7
TOL_REF = 1e-6


@pytest.fixture(scope="module")
def sgrb():
    return np.load(os.path.join(GOLDEN, "sgrb_sample.npz"))


def frame_of(s, grb, cut=True):
    keep = s[f"{grb}_keep"] if cut else np.ones(s[f"{grb}_t"].size, bool)
    return {"t": s[f"{grb}_t"][keep], "Lum50": s[f"{grb}_Lum50"][keep], "Lum50err": s[f"{grb}_Lum50err"][keep]}


def write_limits(s, path):
    names = ["B", "P", "MdiscI", "RdiscI", "epsilon", "delta", "dipeff", "propeff", "f_beam"]
    with open(path, "w") as fh:
        fh.write("pars,lower,upper\n")
        for n, lo, hi in zip(names, s["lims_lower"], s["lims_upper"]):
            fh.write(f"{n},{float(lo)!r},{float(hi)!r}\n")
    return str(path)


def test_sample_shape(sgrb):
    grbs = [str(g) for g in sgrb["grbs"]]
    assert len(grbs) == 15
    assert [sgrb[f"{g}_t"].size for g in grbs] == [253, 80, 33, 1944, 19, 8, 112, 36, 52, 214, 410, 240, 172, 151, 63]
    assert [g for g, r in zip(grbs, sgrb["raises"]) if r] == ["060614"]
    assert sgrb["060614_t"].max() > 1.0e6 and sgrb["060614_keep"].sum() == 1941
    for g in grbs:
        assert sgrb[f"{g}_theta"].shape == (32, 6) and np.isneginf(sgrb[f"{g}_ref_lnprob"][-1])


def test_dataprep_reproduces_reference_kcorrection_on_all_bursts(sgrb):
    """code/clean_data.py + code/kcorr.py on every burst: rest-frame times and luminosities bit for bit, the
    geometric-mean error to 2 ulp (gmean's exp(mean(log)) vs this repo's)."""
    from magprop_b200.dataprep import k_correct_grb
    from magprop_b200.dataprep.clean_data import COLUMNS
    for i, g in enumerate(sgrb["grbs"]):
        raw = sgrb[f"{g}_raw"]
        cols = {name: raw[:, k] for k, name in enumerate(COLUMNS)}
        k = k_correct_grb(cols, float(sgrb["Gamma"][i]), float(sgrb["sigma"][i]), float(sgrb["z"][i]))
        assert np.array_equal(k["t"], sgrb[f"{g}_t"]) and np.array_equal(k["Lum50"], sgrb[f"{g}_Lum50"])
        assert np.allclose(k["Lum50err"], sgrb[f"{g}_Lum50err"], rtol=4e-16, atol=0.0)


def test_oracle_reproduces_reference_lnprob_on_the_sample(sgrb):
    from oracle import magprop_oracle as O
    lo, hi = sgrb["lims_lower"][:6], sgrb["lims_upper"][:6]
    for g in ("061210", "061006", "060614"):                 # 8, 19 and 1941 points
        f = frame_of(sgrb, g)
        for w in (0, 31):
            got = O.lnprob(sgrb[f"{g}_theta"][w], f["t"], f["Lum50"], f["Lum50err"], O.packaged_spec("S"), lo, hi)
            want = sgrb[f"{g}_ref_lnprob"][w]
            assert (np.isneginf(got) and np.isneginf(want)) or abs(got - want) <= 1e-12 * abs(want)


def test_out_of_grid_burst_is_an_error_on_the_host(sgrb, hostsim):
    """GRB 060614 uncut: interp1d raises in the reference (magnetar/funcs.py:214); the node program refuses."""
    import ctypes as C
    from magprop_b200 import _capi as A
    from magprop_b200.engine import time_grid
    grid = time_grid("S")
    f = frame_of(sgrb, "060614", cut=False)
    D = f["t"].size
    scratch_i = np.zeros(2 * D + 4, np.int32)
    scratch_d = np.zeros(D)
    n = C.c_int(0)
    rc = hostsim.hs_node_program(A.ptr(grid), grid.size, A.ptr(f["t"]), A.ptr(f["Lum50"]), A.ptr(f["Lum50err"]), D,
                                 C.byref(n), A.ptr(scratch_i), A.ptr(np.zeros(D, np.int32)), A.ptr(scratch_d),
                                 A.ptr(np.zeros(D)), A.ptr(np.zeros(D, np.int32)))
    assert rc == A.MP_ERR_DATA_RANGE


@pytest.mark.gpu
def test_gpu_lnprob_on_all_15_bursts(built, sgrb, tmp_path):
    from magprop_b200 import _capi as A
    from magprop_b200 import magnetar
    lims = write_limits(sgrb, tmp_path / "lims.csv")
    all_got, all_ref, all_tight = [], [], []
    for g in sgrb["grbs"]:
        g = str(g)
        f = frame_of(sgrb, g)
        theta, ref, tight, flagged = sgrb[f"{g}_theta"], sgrb[f"{g}_ref_lnprob"], sgrb[f"{g}_tight_lnprob"], sgrb[f"{g}_ref_flagged"]
        got = magnetar.lnprob_batch(theta, f, "S", custom_lims=lims)
        assert not np.isnan(got).any()
        assert np.isneginf(got[-1])                                         # outside the prior
        assert magnetar.lnprob(theta[0], f, "S", custom_lims=lims) == got[0]
        ok = np.isfinite(ref) & np.isfinite(tight) & np.isfinite(got) & ~flagged
        assert ok.sum() >= np.isfinite(ref).sum() - 1, g                     # at most one integrator-budget difference
        assert relerr(got[ok], tight[ok]).max() < TOL_TIGHT, g
        slack = TOL_REF * np.abs(ref[ok]) + 1.5 * np.abs(ref[ok] - tight[ok])
        assert (np.abs(got[ok] - ref[ok]) <= slack).all(), g
        all_got.append(got[ok]); all_ref.append(ref[ok]); all_tight.append(tight[ok])
        if g in ("060614", "100212A"):
            report(f"sgrb_{g}_D{f['t'].size}", parity_stats(got[ok], ref[ok], tight[ok]))
    report("sgrb_all_15_bursts", parity_stats(np.concatenate(all_got), np.concatenate(all_ref), np.concatenate(all_tight)))
    # the uncut burst: the reference raises ValueError (interp1d), and so does the drop-in
    with pytest.raises(ValueError):
        magnetar.lnprob_batch(sgrb["060614_theta"], frame_of(sgrb, "060614", cut=False), "S", custom_lims=lims)
