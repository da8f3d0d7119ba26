"""Host-side preparation (node program, grids, limits) and the per-walker core
compiled for the host, against the goldens.  CPU only.

The host build of magprop_core.cuh is a debugging aid for the GPU-less build
container; the GPU parity tests (test_gpu_parity.py) are the parity tests
proper and go through the C ABI."""
import ctypes as C

import numpy as np
import pytest
from scipy.special import hyp1f1

from conftest import relerr
from magprop_b200 import _capi as A
from magprop_b200.engine import time_grid
from oracle import magprop_oracle as O

LNPROB_TOL_TIGHT = 5e-7     # vs the converged (rtol 1e-13) oracle
CURVE_TOL_TIGHT = 5e-7


def test_time_grid_matches_reference_bits():
    assert (time_grid(None) == np.logspace(0.0, 6.0, num=10001, base=10.0)).all()
    assert (time_grid("L") == time_grid(None)).all()
    assert (time_grid("S") == np.logspace(-3.0, 6.0, num=10001, base=10.0)).all()
    with pytest.raises(ValueError):
        time_grid("X")


def test_node_program(hostsim):
    grid = time_grid(None)
    rng = np.random.RandomState(3)
    on = grid[[0, 17, 17, 5000, 10000]]                    # on-node data incl. duplicates and both ends
    off = np.sort(rng.uniform(1.0, 1e6, size=7))
    t = np.concatenate([off, on])
    perm = rng.permutation(t.size); t = t[perm]
    D = t.size
    y = np.ones(D); e = np.ones(D)
    nn = C.c_int(0)
    ngi = np.zeros(2 * D, np.int32); lo = np.zeros(D, np.int32); dx = np.zeros(D); Dx = np.zeros(D); order = np.zeros(D, np.int32)
    rc = hostsim.hs_node_program(A.ptr(grid), grid.size, A.ptr(t), A.ptr(y), A.ptr(e), D, C.byref(nn), A.ptr(ngi),
                                 A.ptr(lo), A.ptr(dx), A.ptr(Dx), A.ptr(order))
    assert rc == 0
    nodes = ngi[:nn.value]
    assert (np.diff(nodes) > 0).all()
    ts = t[order]
    assert (np.diff(ts) >= 0).all()
    for i in range(D):
        j = nodes[lo[i]]
        assert grid[j] <= ts[i]
        if dx[i] == 0.0:
            assert ts[i] == grid[j]
        else:
            assert ts[i] < grid[j + 1] and nodes[lo[i] + 1] == j + 1
            assert dx[i] == ts[i] - grid[j] and Dx[i] == grid[j + 1] - grid[j]
    # interpolation weights reproduce np.interp on an arbitrary curve
    curve = np.sin(np.log(grid)) + 2.0
    mine = np.where(dx == 0, curve[nodes[lo]], (curve[np.minimum(nodes[lo] + 1, 10000)] - curve[nodes[lo]]) / Dx * dx + curve[nodes[lo]])
    assert np.allclose(mine, np.interp(ts, grid, curve), rtol=1e-15, atol=0)
    # out of range / NaN data are rejected (interp1d bounds_error)
    for bad in (0.999, 1.0000001e6, np.nan):
        tb = t.copy(); tb[0] = bad
        assert hostsim.hs_node_program(A.ptr(grid), grid.size, A.ptr(tb), A.ptr(y), A.ptr(e), D, C.byref(nn), A.ptr(ngi),
                                       A.ptr(lo), A.ptr(dx), A.ptr(Dx), A.ptr(order)) == A.MP_ERR_DATA_RANGE


def test_disc_mass_kernel_function(hostsim):
    """S(u) = -(3/2) u^(-2/3) 1F1(1;1/3;-u) over table, series and asymptotic ranges."""
    u = np.concatenate([10 ** np.random.RandomState(0).uniform(-6, 8, 4000), 2.0 ** np.arange(-12, 26)])
    got = np.array([hostsim.hs_disc_S(float(v)) for v in u])
    want = -1.5 * u ** (-2.0 / 3.0) * hyp1f1(1.0, 1.0 / 3.0, -u)
    scale = np.maximum(np.abs(want), u ** (-5.0 / 3.0) * 0.05)     # near S's zero crossing compare absolutely
    assert (np.abs(got - want) / scale).max() < 5e-13              # scipy's 1F1 itself is ~1e-14
    # S solves S' = -S + u^(-5/3)
    h = 1e-4
    for v in (0.02, 0.9, 7.0, 300.0):
        d = (hostsim.hs_disc_S(v * (1 + h)) - hostsim.hs_disc_S(v * (1 - h))) / (2 * v * h)
        assert abs(d + hostsim.hs_disc_S(v) - v ** (-5.0 / 3.0)) < 1e-6 * v ** (-5.0 / 3.0) + 1e-7 * abs(d)


def test_table_coordinate_from_the_mantissa(hostsim):
    """Both tables' local coordinate is read off the mantissa bits of u; it must be, bit for bit, the definition the
    tables were generated with -- (u - centre of the sub-interval) * 2^(NSUB_LOG2 + 1 - e) -- and the row index the
    sub-interval's number."""
    rng = np.random.RandomState(3)
    u = np.concatenate([2.0 ** rng.uniform(-10, 22, 20000), np.ldexp(1.0 + rng.randint(0, 512, 2000) / 512.0, rng.randint(-10, 22, 2000)),
                        np.nextafter(2.0 ** np.arange(-9, 22), 0.0), 2.0 ** np.arange(-10, 22)])
    s, row, k = C.c_double(), C.c_int(), C.c_int()
    hostsim.hs_table_coord.argtypes = [C.c_double, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    for fast in (0, 1):
        for v in u:
            assert hostsim.hs_table_coord(float(v), fast, C.byref(s), C.byref(row), C.byref(k)) == 1
            m, e = np.frexp(v); m, e = 2.0 * m, e - 1                      # v = m 2^e, 1 <= m < 2
            nsub = 1 << k.value
            j = int(np.floor((m - 1.0) * nsub))
            centre = np.ldexp(1.0 + (j + 0.5) / nsub, int(e))
            want = np.ldexp(v - centre, k.value + 1 - int(e))
            assert s.value == want and -1.0 <= s.value < 1.0, (v, s.value, want)
            assert row.value == (int(e) + 10) * nsub + j
        for v in (0.0, -1.0, 2.0 ** -11, 2.0 ** 22, np.inf, np.nan):     # outside: a loadable row all the same
            assert hostsim.hs_table_coord(float(v), fast, C.byref(s), C.byref(row), C.byref(k)) == 0 and row.value == 0


def test_step_size_controller_in_the_log_domain(hostsim):
    """The controller is evaluated in log2 (no divisions): it must be Hairer's dopri5 rule -- accepted:
    h_new = h / clamp(err^0.17 / facold^0.04 / 0.9, 0.1, 5), facold <- max(err, 1e-4); rejected:
    h_new = h / min(err^0.17 / 0.9, 5) -- to single precision."""
    hostsim.hs_step_scale.restype = C.c_double
    hostsim.hs_step_scale.argtypes = [C.c_double, C.c_double, C.c_float, C.c_int, C.POINTER(C.c_float)]
    rng = np.random.RandomState(4)
    lf = C.c_float()
    for _ in range(2000):
        sk = 10.0 ** rng.uniform(-18, -4); err = 10.0 ** rng.uniform(-6, 3); facold = max(10.0 ** rng.uniform(-6, 1), 1e-4)
        got = hostsim.hs_step_scale(err * sk, sk, float(np.log2(facold)), 1, C.byref(lf))
        want = 1.0 / min(5.0, max(0.1, err ** 0.17 / facold ** 0.04 / 0.9))
        assert abs(got - want) < 2e-5 * want
        assert abs(lf.value - np.log2(max(err, 1e-4))) < 1e-4
        got = hostsim.hs_step_scale(err * sk, sk, float(np.log2(facold)), 0, C.byref(lf))
        assert abs(got - 1.0 / min(5.0, err ** 0.17 / 0.9)) < 2e-5 * got
    # a vanishing error grows by the full factor 10, a NaN error shrinks by the full factor 5
    assert abs(hostsim.hs_step_scale(0.0, 1e-9, -13.0, 1, C.byref(lf)) - 10.0) < 1e-4
    assert abs(hostsim.hs_step_scale(float("nan"), 1e-9, -13.0, 1, C.byref(lf)) - 0.2) < 1e-6
    assert abs(hostsim.hs_step_scale(float("nan"), 1e-9, -13.0, 0, C.byref(lf)) - 0.2) < 1e-6


def _hs_curves(hostsim, spec, grid, pars, stride=1):
    pars = np.ascontiguousarray(np.atleast_2d(pars), dtype=np.float64)
    W = pars.shape[0]
    Gs = len(range(0, grid.size, stride)) + (0 if (grid.size - 1) % stride == 0 else 1)
    out = np.zeros((W, 3, Gs)); state = np.zeros((W, 2, Gs)); st = np.zeros(W, np.int32); nr = np.zeros(W, np.int32)
    gs = hostsim.hs_curves(C.byref(spec), A.ptr(grid), grid.size, A.ptr(pars), W, pars.shape[1], stride, A.ptr(out),
                           A.ptr(state), A.ptr(st), A.ptr(nr))
    assert gs == Gs
    return out, state, st, nr


@pytest.mark.parametrize("variant", ["script", "packaged"])
def test_core_curves_vs_goldens(hostsim, golden, variant):
    g = golden[f"curves_{variant}"]
    spec = A.script_model_spec(unlog=False) if variant == "script" else A.packaged_model_spec()
    out, state, st, _ = _hs_curves(hostsim, spec, time_grid(None), g["pars"], stride=20)
    assert (st == 0).all()
    assert out.shape[2] == g["node_index"].size
    assert relerr(state[:, 0], g["state_tight"][:, 0]).max() < 1e-10       # disc mass: closed form vs LSODA(1e-13)
    assert relerr(state[:, 1], g["state_tight"][:, 1]).max() < 1e-7        # spin
    assert relerr(out, g["lum_tight"]).max() < CURVE_TOL_TIGHT
    # against the reference itself (default odeint tolerances): 1e-6, widened only where the
    # reference's own truncation error (reference vs converged solution) exceeds it
    ref = g["ref_curves"][:, 1:]
    slack = 1e-6 * np.abs(ref) + 1.5 * np.abs(ref - g["lum_tight"])
    assert (np.abs(out - ref) <= slack).all()
    frac_within = (relerr(out, ref) < 1e-6).mean()
    assert frac_within > 0.97
    if variant == "packaged":
        assert (out[:, 1] == 0.0).all()                                   # Lprop == 0 (magnetar/funcs.py:193)


def test_core_reference_fixtures(hostsim, golden):
    f = golden["reference_fixtures"]
    out, state, st, _ = _hs_curves(hostsim, A.packaged_model_spec(), time_grid(None), [f["odes_pars"], f["lc_pars"]])
    assert (st == 0).all()
    # tests/test_funcs.py:46-48 and :58-63, np.isclose defaults
    assert np.isclose(state[0, 0], f["odes_Mdisc"]).all() and np.isclose(state[0, 1], f["odes_omega"]).all()
    assert np.isclose(out[1, 0], f["lc_Ltot"]).all() and np.isclose(out[1, 1], f["lc_Lprop"]).all()
    assert np.isclose(out[1, 2], f["lc_Ldip"]).all()
    assert relerr(state[0, 1], f["odes_omega"]).max() < 1e-6 and relerr(out[1, 0], f["lc_Ltot"]).max() < 1e-6


def test_core_lnprob_vs_goldens(hostsim, golden):
    g = golden["lnprob_script"]
    grid = time_grid(None)
    spec = A.script_model_spec()
    prior = A.prior_spec(O.SCRIPT_LOWER, O.SCRIPT_UPPER)
    for di, name in enumerate(g["names"]):
        m = g["dataset"] == di
        th = np.ascontiguousarray(g["theta"][m])
        keep = np.ones(len(th), bool)
        x, y, ye = [np.ascontiguousarray(g[f"{name}_{c}"]) for c in ("x", "y", "yerr")]
        lnp = np.zeros(len(th)); st = np.zeros(len(th), np.int32); nr = np.zeros(len(th), np.int32)
        small = A.script_model_spec(max_steps=40000)    # keep the CPU run short: stiff outliers may flag
        rc = hostsim.hs_lnprob(C.byref(small), C.byref(prior), A.ptr(grid), grid.size, A.ptr(x), A.ptr(y), A.ptr(ye),
                               x.size, A.ptr(th), len(th), 6, A.ptr(lnp), A.ptr(st), A.ptr(nr), None)
        assert rc == 0
        ref, tight, flagged = g["ref_lnprob"][m], g["tight_lnprob"][m], g["ref_flagged"][m]
        # prior decisions are bit-exact, including the inclusive edges, nextafter-outside and NaN rows
        assert ((st & A.WALKER_PRIOR_REJECT) != 0).tolist() == np.isneginf(g["ref_lnprior"][m]).tolist()
        ok = np.isfinite(ref) & ((st & A.WALKER_INTEGRATOR_FAIL) == 0)
        assert ok.sum() > 180
        assert relerr(lnp[ok], tight[ok]).max() < LNPROB_TOL_TIGHT
        slack = 1e-6 * np.abs(ref[ok]) + 1.5 * np.abs(ref[ok] - tight[ok])
        assert (np.abs(lnp[ok] - ref[ok]) <= slack).all()
        assert not np.isnan(lnp).any()


def test_core_model_at_data_S_grid(hostsim, golden):
    g = golden["lnprob_packaged"]
    grid = time_grid("S")
    t = np.ascontiguousarray(g["t"])
    out = np.zeros((1, t.size)); st = np.zeros(1, np.int32)
    spec = A.packaged_model_spec()
    truth = np.ascontiguousarray(g["truth"])
    assert hostsim.hs_model_at(C.byref(spec), A.ptr(grid), grid.size, A.ptr(t), t.size, A.ptr(truth), 1, 6,
                               A.ptr(out), A.ptr(st)) == 0
    assert st[0] == 0
    assert relerr(out[0], g["model_at_truth"]).max() < 1e-6


def test_core_steps_land_on_the_cap_kink(hostsim):
    """The explicit integrator ends a step on the time at which the Alfven radius reaches the light-cylinder
    cap (funcs.py:109-110) instead of stepping across the kink, for each of the four synthetic truths."""
    from scipy.optimize import brentq
    grid = time_grid(None)
    spec = A.script_model_spec(unlog=False)
    for name, p in O.SYNTH_TRUTHS.items():
        pars = np.ascontiguousarray(np.asarray(p, float))
        out = np.zeros((4000, 5))
        n = hostsim.hs_trace(C.byref(spec), A.ptr(grid), grid.size, A.ptr(pars), C.c_double(1e6), 0, A.ptr(out), 4000)
        rows = out[:n]
        acc = rows[rows[:, 3] > 0]
        t_end, om_end = acc[:, 0] + acc[:, 1], acc[:, 2]
        assert n < 140 and (n - len(acc)) <= 4                       # trial steps, discarded trials
        # margin Rm - k*Rlc along the accepted solution (oracle formulas), its sign change and root
        soln = O.integrate(pars, O.script_spec(), tight=True)[0]
        Md, om = soln[:, 0], soln[:, 1]
        B, P, MdiscI, RdiscI, eps, delta = pars
        tvisc = RdiscI * 1e5 / (0.1 * 1.0 * 1e7)
        mu = 1e15 * B * O.R_NS ** 3
        Rm = mu ** (4 / 7) * O.GM ** (-1 / 7) * (3 * Md / tvisc) ** (-2 / 7)
        margin = Rm - 0.9 * O.C_LIGHT / om
        i = np.where(np.sign(margin[1:]) != np.sign(margin[:-1]))[0]
        assert i.size == 1
        j = i[0]
        t_kink = grid[j] + (grid[j + 1] - grid[j]) * margin[j] / (margin[j] - margin[j + 1])
        # the state is interpolated linearly when the kink is located, so the landing is good to ~1e-3 of the
        # step (h/t ~ 0.1 here; a chance hit within 5e-4 of t has probability ~1e-2)
        assert np.abs(t_end / t_kink - 1.0).min() < 5e-4, name
