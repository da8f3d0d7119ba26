"""Row f3: the model copy inlined in the paper-figure scripts (I = 4/5 M R^2, 3*Mdisc, swept n) as a
ModelSpec.  The oracle's figure_spec is pinned bit-for-bit against figure_1.py / figure_4.py's own
``odes`` (oracle/make_goldens_figures.py -> tests/golden/figure_rhs.npz)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, relerr
from magprop_b200 import _capi as A
from oracle import magprop_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden", "figure_rhs.npz")

FIG4 = {"Humped": [1.0, 5.0, 0.001, 100.0, 0.1, 1.0, 10.0], "Classic": [1.0, 5.0, 0.001, 1000.0, 0.1, 1.0, 10.0],      # figure_4.py:114-117
        "Sloped": [1.0, 1.0, 0.001, 100.0, 10.0, 10.0, 1.0], "Stuttering": [1.0, 5.0, 1.e-5, 100.0, 0.1, 100.0, 10.0]}


def test_oracle_figure_rhs_is_the_figure_scripts_rhs():
    g = np.load(GOLD)
    for key in ("figure_1", "figure_4"):
        for row in g[key]:
            y, t, pars, n, want = row[:2], row[2], row[3:8], row[8], row[9:11]
            got = O._rhs_for(O.figure_spec(n=n))(y, t, *pars)
            assert got[0] == want[0] and got[1] == want[1]


def test_oracle_figure3_rhs_both_torques():
    g = np.load(GOLD)
    for row in g["figure_3"]:
        y, t, pars, want, bucc = row[:2], row[2], row[3:8], row[9:11], bool(row[11])
        got = O._rhs_for(O.figure_spec(n=10.0, bucciantini=bucc))(y, t, *pars)
        assert got[0] == want[0] and got[1] == want[1]


def test_core_bucciantini_curve_vs_oracle(hostsim):
    """The per-walker core (host build) with the Bucciantini torque vs the oracle at figure_3.py's parameters."""
    import ctypes as C
    from magprop_b200.engine import time_grid
    from test_host_logic import _hs_curves
    p = np.array([[1.0, 5.0, 0.001, 1000.0, 0.1, 1.0]])                     # figure_3.py:173-178
    out, state, st, _ = _hs_curves(hostsim, A.figure_model_spec(n=10.0, bucciantini=True), time_grid(None), p, stride=50)
    assert st[0] == 0
    tight = O.model(p[0], O.figure_spec(n=10.0, bucciantini=True), tight=True)
    idx = np.arange(0, 10001, 50)
    assert relerr(out[0, 0], tight[1][idx]).max() < 5e-7 and relerr(out[0, 2], tight[3][idx]).max() < 5e-7
    plain = O.model(p[0], O.figure_spec(n=10.0), tight=True)
    assert relerr(tight[1][idx], plain[1][idx]).max() > 0.1                  # the torque model matters


@pytest.mark.gpu
def test_figure3_rhs_on_device(built):
    from magprop_b200.engine import rhs_batch
    r = np.load(GOLD)["figure_3"]
    for bucc in (False, True):
        q = r[r[:, 11] == float(bucc)]
        got = rhs_batch(A.figure_model_spec(n=10.0, bucciantini=bucc), q[:, :2], q[:, 2], q[:, 3:8], [10.0, 0.1, 1.0, 0.9])
        assert relerr(got, q[:, 9:11]).max() < 1e-11


@pytest.mark.gpu
def test_figure3_bucciantini_curves_vs_oracle(built):
    from magprop_b200.engine import Likelihood, time_grid
    from test_gpu_parity import assert_curves_close
    rng = np.random.RandomState(8)
    pars = np.array([[1.0, 5.0, 0.001, 1000.0, 0.1, 1.0], [1.0, 5.0, 0.001, 100.0, 0.1, 1.0], [2.0, 2.0, 1e-4, 300.0, 1.0, 10.0]])
    lk = Likelihood(A.figure_model_spec(n=10.0, bucciantini=True), time_grid(None))
    out, st = lk.curves(pars, node_stride=25)
    lk.close()
    idx = np.arange(0, 10001, 25)
    for i, p in enumerate(pars):
        assert st[i] == 0
        tight = O.model(p, O.figure_spec(n=10.0, bucciantini=True), tight=True)
        assert_curves_close(out[i], tight[1:][:, idx])


@pytest.mark.gpu
def test_figure_rhs_on_device(built):
    from magprop_b200.engine import rhs_batch
    g = np.load(GOLD)
    rows = np.concatenate([g["figure_1"], g["figure_4"]])
    for n in (1.0, 10.0, 50.0):
        r = rows[rows[:, 8] == n]
        got = rhs_batch(A.figure_model_spec(n=n), r[:, :2], r[:, 2], r[:, 3:8], [n, 0.1, 1.0, 0.9])
        assert relerr(got, r[:, 9:11]).max() < 1e-11


@pytest.mark.gpu
def test_figure4_light_curves_vs_oracle(built):
    from magprop_b200.engine import Likelihood, time_grid
    from test_gpu_parity import assert_curves_close
    for name, p in FIG4.items():
        n = p[6]
        lk = Likelihood(A.figure_model_spec(n=n), time_grid(None))
        out, st = lk.curves(np.array([p[:6]]), node_stride=25)
        lk.close()
        assert st[0] == 0
        tight = O.model(np.array(p[:6]), O.figure_spec(n=n), tight=True)
        idx = np.r_[np.arange(0, 10001, 25)]
        if idx[-1] != 10000:
            idx = np.r_[idx, 10000]
        assert_curves_close(out[0], tight[1:][:, idx])


def test_gompertz_oracle_reproduces_the_reference_loop():
    """The comparison model of figure 5 (figure_5.py:222-363): the oracle against vectors made by executing the
    reference's own loop (oracle/make_goldens_gompertz.py)."""
    from oracle import gompertz_oracle as GO
    g = np.load(os.path.join(GOLDEN, "gompertz.npz"))
    n = int(g["n_steps"])
    for name, p in zip(g["names"], g["pars"]):
        t, Ltot, Lp, Ld = GO.curves(p, n)
        mine = np.array([t, Ltot, Lp, Ld])
        assert np.array_equal(mine[:, ::20], g[f"{name}_curves"], equal_nan=True)
        assert np.array_equal(mine[:, -1], g[f"{name}_last"], equal_nan=True)


@pytest.mark.gpu
def test_gompertz_comparator_on_the_device(built):
    """The CUDA comparator against the reference loop's vectors (CUDA's pow/exp differ from libm by <= 2 ulp per call,
    and the model is an explicit Euler recurrence: 1e-9 after 20 000 steps), its stride, and a full 1e6-step run."""
    from magprop_b200.engine import gompertz_curves
    g = np.load(os.path.join(GOLDEN, "gompertz.npz"))
    n = int(g["n_steps"])
    t, out = gompertz_curves(g["pars"], n_steps=n, stride=20)
    for w, name in enumerate(g["names"]):
        ref = g[f"{name}_curves"]
        assert np.array_equal(t, ref[0])
        assert relerr(out[w], ref[1:] / 1.0e50).max() < 1e-9
    t1, full = gompertz_curves(g["pars"][:1], n_steps=2000, stride=1)
    assert np.array_equal(full[0][:, ::20], gompertz_curves(g["pars"][:1], n_steps=2000, stride=20)[1][0])
    tl, long_run = gompertz_curves(g["pars"], n_steps=10 ** 6, stride=1000)
    assert long_run.shape == (4, 3, 1000) and np.isfinite(long_run).all() and tl[-1] == 999001.0
    assert (long_run[:, 0] >= long_run[:, 1]).all() and (long_run[:, 0, -1] < long_run[:, 0, 0]).all()
