"""The drop-in modules (same names/signatures/returns as the reference's
magnetar package and code/synthetic_datasets modules), written like the
reference's own tests/test_funcs.py."""
import numpy as np
import pytest

from conftest import relerr
from oracle import magprop_oracle as O

pytestmark = pytest.mark.gpu


def test_init_conds(built):
    # tests/test_funcs.py:12-25
    from magprop_b200.magnetar import init_conds
    Msol = 1.99e33
    MdiscI, P = 0.001, 1.0
    assert ((init_conds(MdiscI, P)[0] == MdiscI * Msol) & (init_conds(MdiscI, P)[1] == (2.0 * np.pi) / (1.0e-3 * P)))
    from magprop_b200.synthetic.funcs import init_conds as ic
    assert ic(MdiscI, P) == (MdiscI * Msol, (2.0 * np.pi) / (1.0e-3 * P))


def test_model_light_curve(built, golden):
    # tests/test_funcs.py:51-63
    from magprop_b200.magnetar import model_lc
    f = golden["reference_fixtures"]
    t, Ltot, Lprop, Ldip = model_lc([1.0, 5.0, 0.001, 100.0, 0.1, 1.0])
    assert np.isclose(t, f["lc_t"]).all() and np.isclose(Ltot, f["lc_Ltot"]).all()
    assert np.isclose(Lprop, f["lc_Lprop"]).all() and np.isclose(Ldip, f["lc_Ldip"]).all()
    assert (t == np.logspace(0.0, 6.0, num=10001, base=10.0)).all()
    with pytest.raises(ValueError):
        model_lc([1.0, 5.0, 0.001, 100.0, 0.1, 1.0], GRBtype="X")       # magnetar/funcs.py:138-141


def test_odes_and_ODEs(built):
    from magprop_b200.magnetar import odes, init_conds
    from magprop_b200.synthetic.funcs import ODEs
    y0 = init_conds(0.001, 1.0)
    got = odes(y0, 1.0, 1.0, 0.001, 100.0, 1.0, 10.0)
    want = O.rhs(y0, 1.0, 1.0, 0.001, 100.0, 1.0, 10.0, 1.0, 0.1, 1.0, 0.9, inertia_factor=0.8, mdot_factor=1.0)
    assert isinstance(got, np.ndarray) and relerr(got, want).max() < 1e-11
    got = ODEs(y0, 3.0, 1.0, 0.001, 100.0, 1.0, 10.0, 10.0, 0.1, 1.0, 0.9)
    want = O.rhs(y0, 3.0, 1.0, 0.001, 100.0, 1.0, 10.0, 10.0, 0.1, 1.0, 0.9)
    assert isinstance(got, tuple) and relerr(got, want).max() < 1e-11


def test_model_lum_returns(built, golden):
    from magprop_b200.synthetic.funcs import model_lum, tarr
    g = golden["curves_script"]
    out = model_lum(g["pars"][0])
    assert out.shape == (4, 10001) and (out[0] == tarr).all()
    assert relerr(out[1:, g["node_index"]], g["lum_tight"][0]).max() < 5e-7
    x = tarr[[5, 700, 9000]]
    at = model_lum(g["pars"][0], xdata=x)
    assert at.shape == (3,) and relerr(at, out[1, [5, 700, 9000]]).max() < 1e-8
    with pytest.raises(ValueError):
        model_lum(g["pars"][0], xdata=[2.0e6])                               # interp1d bounds_error
    # kwargs reach the model (f_beam scales, n changes the propeller switch-on)
    assert relerr(model_lum(g["pars"][0], xdata=x, f_beam=2.0), 2.0 * at).max() < 1e-15
    assert relerr(model_lum(g["pars"][0], n=1.0)[1], out[1]).max() > 1e-3


def test_script_mcmc_eqns(built, golden, tmp_path):
    from magprop_b200.synthetic import mcmc_eqns as mc
    g = golden["lnprob_script"]
    x, y, yerr = g["Humped_x"], g["Humped_y"], g["Humped_yerr"]
    m = np.where(g["dataset"] == 0)[0]
    fbad = str(tmp_path / "bad.csv")
    for i in list(m[:4]) + list(m[-5:]):
        th = g["theta"][i]
        assert mc.lnprior(th) == g["ref_lnprior"][i]
        got = mc.lnprob(th, x, y, yerr, fbad)
        want = g["ref_lnprob"][i]
        assert (np.isneginf(got) and np.isneginf(want)) or relerr(got, want) < 2e-6
    th = g["theta"][m[0]]
    assert mc.lnlike(th, x, y, yerr) == pytest.approx(mc.lnprob(th, x, y, yerr, None), rel=1e-14)
    batch = mc.lnprob_batch(g["theta"][m], x, y, yerr)
    assert batch.shape == (m.size,) and not np.isnan(batch).any()
    fin = np.isfinite(g["ref_lnprob"][m])
    assert relerr(batch[fin], g["tight_lnprob"][m][fin]).max() < 5e-7


def test_packaged_mcmc_eqns(built, golden, tmp_path):
    import pandas as pd
    from magprop_b200 import magnetar
    g = golden["lnprob_packaged"]
    data = pd.DataFrame({"t": g["t"], "Lum50": g["Lum50"], "Lum50err": g["Lum50err"]})
    lims = tmp_path / "lims.csv"
    pd.DataFrame({"pars": list("abcdefghi"), "lower": g["lims_lower"], "upper": g["lims_upper"]}).to_csv(lims, index=False)
    for th, want in list(zip(g["theta"], g["ref_lnprob"]))[::3]:
        th = th[~np.isnan(th)]
        got = magnetar.lnprob(th, data, "S", custom_lims=str(lims))
        assert (np.isneginf(got) and np.isneginf(want)) or relerr(got, want) < 2e-6
    # default limits (packaged CSV): same constant the reference returns
    for th, want in zip(g["default_theta"], g["default_ref_lnprob"]):
        got = magnetar.lnprob(th, data, "S")
        assert (np.isneginf(got) and np.isneginf(want)) or relerr(got, want) < 1e-12
    assert magnetar.lnprior([1.0, 5.0, -2.0, 2.0, 0.5, 0.3]) == 0.0
    assert magnetar.lnprior([1.0, 5.0, -2.0, 2.0, 0.5, 2.0]) == -np.inf
    assert magnetar.lnprior([1.0, 5.0, -2.0, 2.0, 0.5, 0.3, 700.0]) == -np.inf      # 7-vector uses the f_beam row
    assert magnetar.lnprior([1.0, 5.0, -2.0, 2.0, 0.5, 0.3, 600.0]) == 0.0
    with pytest.raises(ValueError):
        magnetar.lnprior([1.0] * 6, custom_lims=str(tmp_path / "missing.csv"))
