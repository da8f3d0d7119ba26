"""Goodness-of-fit numbers of a finished fit: the interface of the reference's
``magnetar/fit_stats.py`` (``redchisq`` :6-33, ``aicc`` :36-62) on top of ONE weighted
sum of squared residuals.

Both statistics are functions of ``chi2 = sum(((ydata - ymod)/sd)^2)``:

    redchisq = chi2 / (N - 1 - deg)                (chi2 itself when ``deg`` is None)
    aicc     = -chi2 + 2k + 2k(k+1)/(N - k - 1)

so the arithmetic lives in ``_chi2`` and the two public functions only dress it.  For a
fit held on the GPU the same ``chi2`` comes out of the likelihood kernel
(``lnlike = -chi2/2``, ``Likelihood.lnprob`` without a prior): ``from_lnlike`` turns that
number into both statistics without touching the data again.
"""
import numpy as np


def _chi2(ydata, ymod, sd=None) -> float:
    """Sum of squared residuals, weighted by ``sd`` when given (NumPy's pairwise ``np.sum``,
    as the reference uses, so the value is the reference's to the last bit)."""
    resid = np.asarray(ydata, dtype=np.float64) - np.asarray(ymod, dtype=np.float64)
    if sd is not None:
        resid = resid / np.asarray(sd, dtype=np.float64)
    return np.sum(resid ** 2.0)


def redchisq(ydata, ymod, deg=None, sd=None):
    """Chi-square of ``ymod`` against ``ydata``; divided by the degrees of freedom
    ``ydata.size - 1 - deg`` when the number of fitted parameters ``deg`` is given."""
    chi2 = _chi2(ydata, ymod, sd)
    if deg is None:
        return chi2
    return chi2 / (np.size(ydata) - 1.0 - deg)


def aicc(ydata, ymod, yerr, Npars):
    """Corrected Akaike information criterion.  The three arrays must have one length;
    the reference reports which pair disagrees before raising, and so does this."""
    sizes = {"ydata": np.size(ydata), "ymod": np.size(ymod), "yerr": np.size(yerr)}
    if len(set(sizes.values())) != 1:
        for a, b in (("ydata", "ymod"), ("ydata", "yerr"), ("ymod", "yerr")):
            print(f"{a}.size == {b}.size:", sizes[a] == sizes[b])
        raise ValueError("ydata, ymod and yerr should all be the same length")
    return _aicc_from_chi2(_chi2(ydata, ymod, yerr), sizes["ydata"], Npars)


def _aicc_from_chi2(chi2, n, k):
    return -1.0 * chi2 + 2.0 * k + ((2.0 * k) * (k + 1.0)) / (n - k - 1.0)


def from_lnlike(lnlike, n_data, n_pars):
    """(reduced chi-square, AICc) from the likelihood kernel's ``lnlike = -chi2/2``."""
    chi2 = -2.0 * float(lnlike)
    return chi2 / (n_data - 1.0 - n_pars), _aicc_from_chi2(chi2, n_data, n_pars)
