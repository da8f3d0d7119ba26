"""Drop-in for the reference's ``code/synthetic_datasets/mcmc_eqns.py``
(``lnlike``, ``lnprior``, ``lnprob``) plus ``lnprob_batch`` for emcee's
``vectorize=True`` protocol.  Prior test, ODE solve, luminosity, interpolation
and chi-square all happen in one CUDA launch per call."""
import numpy as np

from .. import _cache
from .. import _capi as A

upper = np.array([10.0, 10.0, -2.0, np.log10(2000.0), 2.0, 3.0])      # mcmc_eqns.py:40
lower = np.array([1.0e-3, 0.69, -6.0, np.log10(50.0), -2.0, -1.0])    # mcmc_eqns.py:41

DEVICE = 0


def _lik(x, y, yerr, with_prior=True):
    spec = A.script_model_spec()
    if with_prior:
        return _cache.get(spec, None, x, y, yerr, lower, upper, device=DEVICE)
    return _cache.get(spec, None, x, y, yerr, device=DEVICE)


def lnlike(pars, x, y, yerr):
    """mcmc_eqns.py:5-25: -0.5*chi2 of the model at log-space ``pars``; -inf when
    the integration flags."""
    lnp, status, _ = _lik(x, y, yerr, with_prior=False).lnprob(np.asarray(pars, float).reshape(1, -1), return_info=True)
    return float(lnp[0])


def lnprior(pars):
    """mcmc_eqns.py:28-49: 0.0 inside the inclusive box, -inf outside (NaN rejects).
    The comparisons run on the device, the same code the lnprob kernel uses."""
    lk = _cache.get(A.script_model_spec(), None, None, None, None, lower, upper, device=DEVICE)
    _, status, _ = lk.lnprob(np.asarray(pars, float).reshape(1, -1), return_info=True)
    return -np.inf if (status[0] & A.WALKER_PRIOR_REJECT) else 0.0


def _write_bad(fbad, rows):
    with open(fbad, "a") as f:
        for pars in rows:
            f.write(", ".join(f"{k}" for k in pars) + "\n")      # mcmc_eqns.py:73-78


def lnprob_batch(coords, x, y, yerr, fbad=None):
    """lnprob for every row of ``coords`` [W, 6] in ONE launch -> array [W].

    Usable as ``emcee.EnsembleSampler(nwalkers, 6, lnprob_batch, args=(x, y, yerr,
    fbad), vectorize=True)``.  Never returns NaN; -inf means certain rejection."""
    coords = np.atleast_2d(np.asarray(coords, dtype=np.float64))
    lnp, status, _ = _lik(x, y, yerr).lnprob(coords, return_info=True)
    if fbad is not None:
        bad = (status & (A.WALKER_INTEGRATOR_FAIL | A.WALKER_NONFINITE_LNLIKE)) != 0
        if bad.any():
            _write_bad(fbad, coords[bad])
    return lnp


def lnprob(pars, x, y, yerr, fbad):
    """mcmc_eqns.py:52-81 for one walker."""
    return float(lnprob_batch(np.asarray(pars, float).reshape(1, -1), x, y, yerr, fbad)[0])
