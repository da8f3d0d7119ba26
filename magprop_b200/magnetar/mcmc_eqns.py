"""Drop-in for the reference's ``magnetar/mcmc_eqns.py`` (``lnlike``,
``lnprior``, ``lnprob``) plus ``lnprob_batch`` for emcee's ``vectorize=True``."""
import csv
import os

import numpy as np

from .. import _cache
from .. import _capi as A

DEVICE = 0
_DEFAULT_LIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mcmc_limits.csv")
_lims_cache = {}

__all__ = ["lnlike", "lnprior", "lnprob", "lnprob_batch"]


def _read_limits(path):
    """lower/upper columns of a limits CSV (magnetar/mcmc_eqns.py:54-60).  The
    reference re-reads a cwd-relative file on every call (:55); here the packaged
    copy is found from any cwd and parsed once per (path, mtime)."""
    if path is None:
        path = _DEFAULT_LIMS
    try:
        key = (os.path.abspath(path), os.path.getmtime(path))
    except OSError:
        raise ValueError("Please provide a valid file path.")
    if key not in _lims_cache:
        lo, hi = [], []
        try:
            with open(path, newline="") as fh:
                for row in csv.DictReader(fh):
                    lo.append(float(row["lower"]))
                    hi.append(float(row["upper"]))
        except (KeyError, ValueError):
            raise ValueError("Please provide a valid file path.")
        _lims_cache[key] = (np.array(lo), np.array(hi))
    return _lims_cache[key]


def _bounds(ndim, custom_lims):
    lo, hi = _read_limits(custom_lims)
    if ndim == 7:                                   # magnetar/mcmc_eqns.py:64-75
        return np.append(lo[:6], lo[-1]), np.append(hi[:6], hi[-1])
    return lo[:ndim], hi[:ndim]                    # :78-79


def _columns(data):
    return (np.asarray(data["t"], float), np.asarray(data["Lum50"], float), np.asarray(data["Lum50err"], float))


def lnlike(pars, data, GRBtype):
    """magnetar/mcmc_eqns.py:6-37: 6/7/8/9-parameter dispatch, parameters passed raw."""
    x, y, yerr = _columns(data)
    lk = _cache.get(A.packaged_model_spec(), GRBtype, x, y, yerr, device=DEVICE)
    lnp = lk.lnprob(np.asarray(pars, float).reshape(1, -1))
    return float(lnp[0])


def lnprior(pars, custom_lims=None):
    """magnetar/mcmc_eqns.py:40-84."""
    pars = np.asarray(pars, float).ravel()
    lo, hi = _bounds(pars.size, custom_lims)
    lk = _cache.get(A.packaged_model_spec(), None, None, None, None, lo, hi, device=DEVICE)
    _, status, _ = lk.lnprob(pars.reshape(1, -1), return_info=True)
    return -np.inf if (status[0] & A.WALKER_PRIOR_REJECT) else 0.0


def lnprob_batch(coords, data, GRBtype, custom_lims=None):
    """lnprob for every row of coords [W, ndim] in one launch (vectorize=True)."""
    coords = np.atleast_2d(np.asarray(coords, dtype=np.float64))
    x, y, yerr = _columns(data)
    lo, hi = _bounds(coords.shape[1], custom_lims)
    lk = _cache.get(A.packaged_model_spec(), GRBtype, x, y, yerr, lo, hi, device=DEVICE)
    return lk.lnprob(coords)


def lnprob(pars, data, GRBtype, custom_lims=None):
    """magnetar/mcmc_eqns.py:87-119."""
    return float(lnprob_batch(np.asarray(pars, float).reshape(1, -1), data, GRBtype, custom_lims)[0])
