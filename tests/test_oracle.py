"""The oracle against the committed golden vectors (generated from the
unmodified reference by oracle/make_goldens.py) and the reference's own
fixtures.  CPU only."""
import numpy as np
import pytest

from conftest import relerr
from oracle import magprop_oracle as O


def test_init_conds_exact():
    # tests/test_funcs.py:12-25
    Mdisc, omega = O.init_conds(0.001, 1.0)
    assert Mdisc == 0.001 * 1.99e33 and omega == (2.0 * np.pi) / (1.0e-3 * 1.0)


def test_reference_fixture_odes(golden):
    # tests/test_funcs.py:28-48 (np.isclose defaults), packaged odes through odeint
    f = golden["reference_fixtures"]
    soln, ok, _ = O.integrate(f["odes_pars"], O.packaged_spec(), grid=f["odes_t"])
    assert ok
    assert np.isclose(soln[:, 0], f["odes_Mdisc"]).all() and np.isclose(soln[:, 1], f["odes_omega"]).all()
    assert relerr(soln[:, 1], f["odes_omega"]).max() < 1e-7


def test_reference_fixture_light_curve(golden):
    # tests/test_funcs.py:51-63
    f = golden["reference_fixtures"]
    lc = O.model(f["lc_pars"], O.packaged_spec())
    assert np.isclose(lc[0], f["lc_t"]).all() and np.isclose(lc[1], f["lc_Ltot"]).all()
    assert np.isclose(lc[2], f["lc_Lprop"]).all() and np.isclose(lc[3], f["lc_Ldip"]).all()
    assert (f["lc_Lprop"] == 0.0).all() and (lc[2] == 0.0).all()      # magnetar/funcs.py:193
    assert relerr(lc[1], f["lc_Ltot"]).max() < 1e-7


@pytest.mark.parametrize("variant", ["script", "packaged"])
def test_curves_match_reference(golden, variant):
    g = golden[f"curves_{variant}"]
    spec = O.script_spec() if variant == "script" else O.packaged_spec()
    for i in range(int(g["n_named"])):
        lc = O.model(g["pars"][i], spec)
        assert relerr(lc[:, g["node_index"]], g["ref_curves"][i]).max() < 1e-12


def test_script_lnprob_matches_reference(golden):
    g = golden["lnprob_script"]
    spec = O.script_spec()
    pick = np.r_[0:6, 96:104, 192:197]          # ball, prior-uniform, edge rows of the first dataset
    name = str(g["names"][0])
    x, y, yerr = g[f"{name}_x"], g[f"{name}_y"], g[f"{name}_yerr"]
    for i in pick:
        got = O.lnprob(g["theta"][i], x, y, yerr, spec, O.SCRIPT_LOWER, O.SCRIPT_UPPER)
        want = g["ref_lnprob"][i]
        assert (np.isneginf(got) and np.isneginf(want)) or relerr(got, want) < 1e-12
    # prior decisions, every stored walker, bit-exact
    lp = np.array([O.lnprior(th, O.SCRIPT_LOWER, O.SCRIPT_UPPER) for th in g["theta"]])
    assert (lp == g["ref_lnprior"]).all()


def test_packaged_lnprob_matches_reference(golden):
    g = golden["lnprob_packaged"]
    spec = O.packaged_spec("S")
    for th, want in list(zip(g["theta"], g["ref_lnprob"]))[::5]:
        th = th[~np.isnan(th)]
        lo, hi = O.prior_bounds("packaged", len(th), g["lims_lower"], g["lims_upper"])
        got = O.lnprob(th, g["t"], g["Lum50"], g["Lum50err"], spec, lo, hi)
        assert (np.isneginf(got) and np.isneginf(want)) or relerr(got, want) < 1e-12


def test_interp_out_of_range_raises():
    grid = O.script_spec().grid()
    with pytest.raises(ValueError):
        O.interp_linear(grid, np.ones_like(grid), [0.5])
    with pytest.raises(ValueError):
        O.interp_linear(grid, np.ones_like(grid), [1.0e6 * (1 + 1e-12)])


def test_bad_grbtype():
    with pytest.raises(ValueError):
        O.packaged_spec("X")
