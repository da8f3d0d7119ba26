"""CPU oracle for the magprop likelihood hot path.  TEST INFRASTRUCTURE ONLY.

This module is a CPU restatement of the reference's algorithm (SciPy LSODA
``odeint`` + NumPy luminosity stage + linear interpolation + chi-square).  It
exists to CHECK the CUDA path.  Only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import
it; nothing under ``magprop_b200/`` does, and the product has no CPU path.

Parity status: PINNED.  ``oracle/make_goldens.py`` imports the unmodified
reference from ``/root/reference`` (in the build container) and asserts this
restatement reproduces it (state trajectories, light curves, lnprior / lnlike /
lnprob) and the reference's own two hot-path fixtures
(``tests/test_data/odes_integrated_by_odeint.csv``,
``tests/test_data/model_light_curve.csv``); the resulting vectors are committed
under ``tests/golden/`` and re-checked by ``tests/test_oracle.py``.

Third-party arithmetic: ``scipy.integrate.odeint`` (ODEPACK LSODA; reference
pins scipy==1.3.0 in requirements.txt:9, this image has 1.18.1) and NumPy's
``interp`` (what ``scipy.interpolate.interp1d(kind="linear")`` evaluates for
1-D float64 input).  Both are present in the image on the build box and the
GPU box, so the oracle calls them directly rather than restating LSODA.

The reference carries two divergent copies of the path:
  * "script"   : code/synthetic_datasets/funcs.py + mcmc_eqns.py
  * "packaged" : magnetar/funcs.py + mcmc_eqns.py + mcmc_limits.csv
They are expressed here by one ``ModelSpec`` so the arithmetic is stated once.
Operation order inside ``rhs`` / ``luminosity`` follows the reference line by
line so LSODA sees bit-identical derivatives.
"""
from __future__ import annotations

import dataclasses
import math
import os
from typing import Optional, Sequence

import numpy as np
from scipy.integrate import odeint

# --- constants: code/synthetic_datasets/funcs.py:12-18, magnetar/funcs.py:7-13
G_NEWTON = 6.674e-8
C_LIGHT = 3.0e10
R_NS = 1.0e6
MSOL = 1.99e33
M_NS = 1.4 * MSOL
GM = G_NEWTON * M_NS
N_GRID = 10001

SUCCESS_MESSAGE = "Integration successful."  # funcs.py:172, magnetar/funcs.py:153


@dataclasses.dataclass(frozen=True)
class ModelSpec:
    """Every knob in which the reference's copies of the model differ."""

    name: str
    inertia_factor: float      # I = f*M*R^2: 0.35 funcs.py:17 | 4/5 magnetar/funcs.py:12
    mdot_factor: float         # (f*Mdisc/tvisc)^(-2/7): 3 funcs.py:105 | 1 magnetar/funcs.py:64
    # knobs seen by the RHS handed to odeint
    rhs_n: float
    rhs_alpha: float
    rhs_cs7: float
    rhs_k: float
    # knobs seen by the luminosity stage (the packaged model_lc does not forward
    # its kwargs to the RHS, magnetar/funcs.py:150-151, so the two sets differ)
    lum_n: float
    lum_alpha: float
    lum_cs7: float
    lum_k: float
    dipeff: float
    propeff: float
    f_beam: float
    breakup_rhs: float         # rot_param > x  => Nacc = 0 in RHS (0.27 both)
    breakup_lum: float         # same test in luminosity stage: 0.27 funcs.py:206 | 0.0 magnetar/funcs.py:193
    lprop_binding_term: bool   # subtract (GM/Rm)*eta2*Mdisc/tvisc: funcs.py:222-223 yes | magnetar/funcs.py:206 no
    grid_lo: float             # log10 of first grid node
    grid_hi: float
    unlog_from: int            # lnlike un-logs theta[unlog_from:6] (script: 2, mcmc_eqns.py:17; packaged: 6 = none)
    mdot_sum_prop_first: bool = False  # Mdotfb-Mdotprop-Mdotacc (magnetar/funcs.py:88) vs Mdotfb-Mdotacc-Mdotprop (funcs.py:129)
    bucciantini: bool = False          # RHS dipole torque (-2/3)(mu^2 w^3/c^3)(Rlc/Rm)^3, figure_3.py:142-143

    @property
    def inertia(self) -> float:
        return self.inertia_factor * M_NS * R_NS ** 2.0

    def grid(self) -> np.ndarray:
        return np.logspace(self.grid_lo, self.grid_hi, num=N_GRID, base=10.0)


def script_spec(n=10.0, alpha=0.1, cs7=1.0, k=0.9, dipeff=1.0, propeff=1.0,
                f_beam=1.0) -> ModelSpec:
    """model_lum's defaults and forwarding: code/synthetic_datasets/funcs.py:146-170."""
    return ModelSpec("script", 0.35, 3.0, n, alpha, cs7, k, n, alpha, cs7, k,
                     dipeff, propeff, f_beam, 0.27, 0.27, True, 0.0, 6.0, 2)


def figure_spec(n=10.0, alpha=0.1, cs7=1.0, k=0.9, bucciantini=False) -> ModelSpec:
    """The copy of the model inlined in the paper-figure scripts: code/figure_4.py:14 (I = 4/5 M R^2),
    :139-140 (3*Mdisc), :158-166 (0.27), :173 (binding term); figure_1.py:126 sweeps n."""
    return ModelSpec("figure", 4.0 / 5.0, 3.0, n, alpha, cs7, k, n, alpha, cs7, k,
                     1.0, 1.0, 1.0, 0.27, 0.27, True, 0.0, 6.0, 6,
                     bucciantini=bucciantini,
                     mdot_sum_prop_first=True)    # figure_1.py:88: Mdotfb - Mdotprop - Mdotacc


def packaged_spec(GRBtype=None, dipeff=0.05, propeff=0.4, f_beam=1.0, n=1.0,
                  alpha=0.1, cs7=1.0, k=0.9) -> ModelSpec:
    """model_lc: magnetar/funcs.py:105-220 (RHS always sees odes' own defaults)."""
    if GRBtype is None or GRBtype == "L":
        lo = 0.0
    elif GRBtype == "S":
        lo = -3.0
    else:
        raise ValueError(
            "Please provide a valid value for GRBtype.\nOptions are: L, S, or None.")
    return ModelSpec("packaged", 4.0 / 5.0, 1.0, 1.0, 0.1, 1.0, 0.9, n, alpha,
                     cs7, k, dipeff, propeff, f_beam, 0.27, 0.0, False, lo, 6.0, 6, True)


# --------------------------------------------------------------------------- a1
def init_conds(MdiscI, P):
    """funcs.py:51-71 / magnetar/funcs.py:17-29."""
    return MdiscI * MSOL, (2.0 * np.pi) / (1.0e-3 * P)


def _binding_energy():
    x = GM / (R_NS * (C_LIGHT ** 2.0))
    return 0.6 * M_NS * (C_LIGHT ** 2.0) * (x / (1.0 - 0.5 * x))


MOD_W = _binding_energy()


# --------------------------------------------------------------------------- a2
def rhs(y, t, B, MdiscI, RdiscI, epsilon, delta, n, alpha, cs7, k,
        inertia_factor=0.35, mdot_factor=3.0, breakup=0.27, prop_first=False, bucciantini=False):
    """Coupled RHS, funcs.py:75-142 / magnetar/funcs.py:33-101 (scalar form)."""
    Mdisc, omega = y
    inertia = inertia_factor * M_NS * R_NS ** 2.0
    Rdisc = RdiscI * 1.0e5
    tvisc = Rdisc / (alpha * cs7 * 1.0e7)
    mu = 1.0e15 * B * (R_NS ** 3.0)
    M0 = delta * MdiscI * MSOL
    tfb = epsilon * tvisc

    Rm = ((mu ** (4.0 / 7.0)) * (GM ** (-1.0 / 7.0))
          * ((mdot_factor * Mdisc) / tvisc) ** (-2.0 / 7.0))
    Rc = (GM / (omega ** 2.0)) ** (1.0 / 3.0)
    Rlc = C_LIGHT / omega
    if Rm >= (k * Rlc):
        Rm = k * Rlc
    w = (Rm / Rc) ** (3.0 / 2.0)
    rot_param = (0.5 * inertia * (omega ** 2.0)) / MOD_W

    Ndip = (-1.0 * (mu ** 2.0) * (omega ** 3.0)) / (6.0 * (C_LIGHT ** 3.0))
    if bucciantini:                                            # figure_3.py:142-143
        Ndip = ((-2.0 / 3.0) * (((mu ** 2.0) * (omega ** 3.0)) / (C_LIGHT ** 3.0)) * ((Rlc / Rm) ** 3.0))
    eta2 = 0.5 * (1.0 + np.tanh(n * (w - 1.0)))
    eta1 = 1.0 - eta2
    Mdotprop = eta2 * (Mdisc / tvisc)
    Mdotacc = eta1 * (Mdisc / tvisc)
    Mdotfb = (M0 / tfb) * ((t + tfb) / tfb) ** (-5.0 / 3.0)
    if prop_first:
        Mdotdisc = Mdotfb - Mdotprop - Mdotacc
    else:
        Mdotdisc = Mdotfb - Mdotacc - Mdotprop

    if rot_param > breakup:
        Nacc = 0.0
    elif Rm >= R_NS:
        Nacc = ((GM * Rm) ** 0.5) * (Mdotacc - Mdotprop)
    else:
        Nacc = ((GM * R_NS) ** 0.5) * (Mdotacc - Mdotprop)
    omegadot = (Nacc + Ndip) / inertia
    return Mdotdisc, omegadot


def _rhs_for(spec: ModelSpec):
    def f(y, t, B, MdiscI, RdiscI, epsilon, delta):
        return rhs(y, t, B, MdiscI, RdiscI, epsilon, delta, spec.rhs_n,
                   spec.rhs_alpha, spec.rhs_cs7, spec.rhs_k,
                   spec.inertia_factor, spec.mdot_factor, spec.breakup_rhs,
                   spec.mdot_sum_prop_first, spec.bucciantini)
    return f


# --------------------------------------------------------------------------- a3/a4
def integrate(pars, spec: ModelSpec, tight: bool = False, grid=None):
    """LSODA over the grid: funcs.py:168-173 / magnetar/funcs.py:150-154.

    Returns (soln[G,2], ok, info).  ``tight`` integrates the same system at
    rtol=atol-relative 1e-13 with an effectively unlimited step budget; it is
    the converged answer used to localise the default-tolerance oracle's own
    truncation error, not a reference behaviour.
    """
    B, P, MdiscI, RdiscI, epsilon, delta = [float(v) for v in pars]
    y0 = init_conds(MdiscI, P)
    tarr = spec.grid() if grid is None else grid
    kw = {}
    if tight:
        kw = dict(rtol=1e-13, atol=[1.0, 1e-12], mxstep=5_000_000)
    # LSODA prints Fortran warnings to fd 1; the reference silences them
    # (funcs.py:24-47,168).  Silence by fd redirection here as well.
    with _quiet_fd1():
        soln, info = odeint(_rhs_for(spec), y0, tarr,
                            args=(B, MdiscI, RdiscI, epsilon, delta),
                            full_output=True, **kw)
    return soln, info["message"] == SUCCESS_MESSAGE, info


class _quiet_fd1:
    def __enter__(self):
        import sys
        try:
            sys.stdout.flush()
        except Exception:
            pass
        self._saved = os.dup(1)
        self._null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self._null, 1)

    def __exit__(self, *exc):
        os.dup2(self._saved, 1)
        os.close(self._saved)
        os.close(self._null)
        return False


# --------------------------------------------------------------------------- a5
def luminosity(soln, pars, spec: ModelSpec, dipeff=None, propeff=None, f_beam=None):
    """Luminosity stage on solution arrays: funcs.py:175-229 / magnetar/funcs.py:157-210.

    Returns (Ltot, Lprop, Ldip) in erg/s (NOT yet divided by 1e50).
    """
    B, P, MdiscI, RdiscI, epsilon, delta = [float(v) for v in pars]
    dipeff = spec.dipeff if dipeff is None else dipeff
    propeff = spec.propeff if propeff is None else propeff
    f_beam = spec.f_beam if f_beam is None else f_beam
    inertia = spec.inertia
    Mdisc = np.array(soln[:, 0])
    omega = np.array(soln[:, 1])
    n, alpha, cs7, k = spec.lum_n, spec.lum_alpha, spec.lum_cs7, spec.lum_k

    Rdisc = RdiscI * 1.0e5
    tvisc = Rdisc / (alpha * cs7 * 1.0e7)
    mu = 1.0e15 * B * (R_NS ** 3.0)

    with np.errstate(all="ignore"):
        Rm = ((mu ** (4.0 / 7.0)) * (GM ** (-1.0 / 7.0))
              * ((spec.mdot_factor * Mdisc) / tvisc) ** (-2.0 / 7.0))
        Rc = (GM / (omega ** 2.0)) ** (1.0 / 3.0)
        Rlc = C_LIGHT / omega
        Rm = np.where(Rm >= (k * Rlc), (k * Rlc), Rm)
        w = (Rm / Rc) ** (3.0 / 2.0)
        rot_param = (0.5 * inertia * (omega ** 2.0)) / MOD_W
        eta2 = 0.5 * (1.0 + np.tanh(n * (w - 1.0)))
        eta1 = 1.0 - eta2
        Mdotprop = eta2 * (Mdisc / tvisc)
        Mdotacc = eta1 * (Mdisc / tvisc)

        # the reference's Python loop (funcs.py:204-212) as selects
        lever = np.where(Rm >= R_NS, (GM * Rm) ** 0.5, (GM * R_NS) ** 0.5)
        Nacc = np.where(rot_param > spec.breakup_lum, 0.0, lever * (Mdotacc - Mdotprop))
        # (a NaN rot_param / Rm falls through both selects to the reference's
        # final else-branch, exactly as the comparisons in its loop do)

        if spec.lprop_binding_term:
            Ldip = dipeff * (((mu ** 2.0) * (omega ** 4.0)) / (6.0 * (C_LIGHT ** 3.0)))
        else:
            Ndip = (-1.0 * (mu ** 2.0) * (omega ** 3.0)) / (6.0 * (C_LIGHT ** 3.0))
            Ldip = dipeff * (-1.0 * Ndip * omega)
        Ldip = np.where(Ldip <= 0.0, 0.0, Ldip)
        Ldip = np.where(np.isfinite(Ldip), Ldip, 0.0)

        if spec.lprop_binding_term:
            Lprop = propeff * ((-1.0 * Nacc * omega) - ((GM / Rm) * eta2 * (Mdisc / tvisc)))
        else:
            Lprop = propeff * (-1.0 * Nacc * omega)
        Lprop = np.where(Lprop <= 0.0, 0.0, Lprop)
        Lprop = np.where(np.isfinite(Lprop), Lprop, 0.0)

        Ltot = f_beam * (Ldip + Lprop)
    return Ltot, Lprop, Ldip


# --------------------------------------------------------------------------- a6/a7
FLAG = "flag"


def interp_linear(grid, values, xdata):
    """interp1d(grid, values)(xdata), kind='linear', bounds_error=True
    (funcs.py:233-234 / magnetar/funcs.py:214-215)."""
    x = np.asarray(xdata, dtype=np.float64)
    if x.size and (x.min() < grid[0] or x.max() > grid[-1]):
        lo = "below" if x.min() < grid[0] else "above"
        raise ValueError(f"A value in x_new is {lo} the interpolation range.")
    return np.interp(x, grid, values)


def model(pars, spec: ModelSpec, xdata=None, tight=False, **eff):
    """model_lum / model_lc: returns (4,G) curves, (D,) luminosities, or 'flag'."""
    grid = spec.grid()
    soln, ok, _ = integrate(pars, spec, tight=tight, grid=grid)
    if not ok:
        return FLAG
    Ltot, Lprop, Ldip = luminosity(soln, pars, spec, **eff)
    if xdata is None:
        return np.array([grid, Ltot / 1.0e50, Lprop / 1.0e50, Ldip / 1.0e50])
    return interp_linear(grid, Ltot, xdata) / 1.0e50


# --------------------------------------------------------------------------- a9-a11
SCRIPT_UPPER = np.array([10.0, 10.0, -2.0, np.log10(2000.0), 2.0, 3.0])   # mcmc_eqns.py:40
SCRIPT_LOWER = np.array([1.0e-3, 0.69, -6.0, np.log10(50.0), -2.0, -1.0])  # mcmc_eqns.py:41

# magnetar/mcmc_limits.csv:2-10, parsed as the decimal literals written there
PACKAGED_LOWER = np.array([0.001, 0.69, -3.0, 1.6989700043360187, -1.0, -5.0, 0.01, 0.01, 1.0])
PACKAGED_UPPER = np.array([10.0, 10.0, -1.0, 3.3010299956639813, 3.0, 1.6989700043360187, 1.0, 1.0, 600.0])


def prior_bounds(variant: str, ndim: int, lower=None, upper=None):
    """Bounds vectors as lnprior slices them (magnetar/mcmc_eqns.py:62-79)."""
    if variant == "script":
        return SCRIPT_LOWER.copy(), SCRIPT_UPPER.copy()
    lo = PACKAGED_LOWER if lower is None else np.asarray(lower, float)
    hi = PACKAGED_UPPER if upper is None else np.asarray(upper, float)
    if ndim == 7:
        return np.append(lo[:6], lo[-1]), np.append(hi[:6], hi[-1])
    return lo[:ndim].copy(), hi[:ndim].copy()


def lnprior(theta, lower, upper):
    """Top-hat, inclusive bounds, NaN rejects (mcmc_eqns.py:43-49)."""
    theta = np.asarray(theta, float)
    if np.all(theta <= upper) and np.all(theta >= lower):
        return 0.0
    return -np.inf


def lnlike(theta, x, y, yerr, spec: ModelSpec, tight=False):
    """mcmc_eqns.py:5-25 (script) / magnetar/mcmc_eqns.py:6-37 (packaged)."""
    arr = np.array(theta, dtype=float)
    eff = {}
    if spec.name == "script":
        arr[spec.unlog_from:] = 10.0 ** arr[spec.unlog_from:]
    else:
        if len(arr) == 7:
            eff = dict(f_beam=arr[6])
        elif len(arr) == 8:
            eff = dict(dipeff=arr[6], propeff=arr[7])
        elif len(arr) == 9:
            eff = dict(dipeff=arr[6], propeff=arr[7], f_beam=arr[8])
    mod = model(arr[:6], spec, xdata=x, tight=tight, **eff)
    if isinstance(mod, str):
        return -np.inf
    return -0.5 * np.sum(((np.asarray(y) - mod) / np.asarray(yerr)) ** 2.0)


def lnprob(theta, x, y, yerr, spec: ModelSpec, lower, upper, tight=False):
    """mcmc_eqns.py:52-81 / magnetar/mcmc_eqns.py:87-119 (non-finite ll -> -inf)."""
    lp = lnprior(theta, lower, upper)
    if not np.isfinite(lp):
        return -np.inf
    ll = lnlike(theta, x, y, yerr, spec, tight=tight)
    if not np.isfinite(ll):
        return -np.inf
    return ll + lp


def _lnprob_task(args):
    return lnprob(*args)


def lnprob_batch(thetas, x, y, yerr, spec, lower, upper, tight=False, pool=None):
    """The reference's pool.map(lnprob, walkers) pattern (synth_mcmc.py:178-185)."""
    tasks = [(th, x, y, yerr, spec, lower, upper, tight) for th in np.asarray(thetas)]
    if pool is None:
        return np.array([_lnprob_task(t) for t in tasks])
    return np.array(pool.map(_lnprob_task, tasks, chunksize=max(1, len(tasks) // (8 * (pool._processes or 1)))))


# --------------------------------------------------------------------------- synthetic data recipe
SYNTH_TRUTHS = {  # generate_data.py:10-15
    "Humped": np.array([1.0, 5.0, 1.0e-3, 100.0, 0.1, 1.0]),
    "Classic": np.array([1.0, 5.0, 1.0e-3, 1000.0, 0.1, 1.0]),
    "Sloped": np.array([1.0, 1.0, 1.0e-3, 100.0, 10.0, 10.0]),
    "Stuttering": np.array([1.0, 5.0, 1.0e-5, 100.0, 0.1, 100.0]),
}
SYNTH_TRUTHS_LOG = {  # synth_mcmc.py:16-21
    "Humped": np.array([1.0, 5.0, -3.0, 2.0, -1.0, 0.0]),
    "Classic": np.array([1.0, 5.0, -3.0, 3.0, -1.0, 0.0]),
    "Sloped": np.array([1.0, 1.0, -3.0, 2.0, 1.0, 1.0]),
    "Stuttering": np.array([1.0, 5.0, -5.0, 2.0, -1.0, 2.0]),
}
SYNTH_SEED = 20170613  # the reference does not seed; fixed here


def synth_dataset(name, curves, seed=SYNTH_SEED, npts=50):
    """generate_data.py:58-71 applied to a (4,G) curve array."""
    rng = np.random.RandomState(seed + sum(map(ord, name)))
    inx = np.sort(rng.randint(low=0, high=curves.shape[1], size=npts))
    x = curves[0, inx].copy()
    y = curves[1, inx].copy()
    yerr = 0.25 * y
    y = y + rng.normal(loc=0.0, scale=yerr, size=len(yerr))
    return x, y, yerr
