"""Drop-in for the reference's ``magnetar/fit_stats.py`` (``redchisq``, ``aicc``).

Two scalar reductions over a finished fit; they are
not in the sampler loop, so they stay NumPy on the host exactly as the reference
has them -- the model values they consume come from the CUDA path."""
import numpy as np


def redchisq(ydata, ymod, deg=None, sd=None):
    """fit_stats.py:6-33: chi-square, divided by ``ydata.size - 1 - deg`` when ``deg`` is given."""
    ydata, ymod = np.asarray(ydata), np.asarray(ymod)
    if sd is not None:
        chisq = np.sum(((ydata - ymod) / np.asarray(sd)) ** 2.0)
    else:
        chisq = np.sum((ydata - ymod) ** 2.0)
    if deg is not None:
        nu = ydata.size - 1.0 - deg
        return chisq / nu
    return chisq


def aicc(ydata, ymod, yerr, Npars):
    """fit_stats.py:36-62: corrected Akaike information criterion; ``ValueError`` on a length mismatch."""
    ydata, ymod, yerr = np.asarray(ydata), np.asarray(ymod), np.asarray(yerr)
    cond1 = ydata.size == ymod.size
    cond2 = ydata.size == yerr.size
    cond3 = ymod.size == yerr.size
    if (not cond1) or (not cond2) or (not cond3):
        print("ydata.size == ymod.size:", cond1)
        print("ydata.size == yerr.size:", cond2)
        print("ymod.size == yerr.size:", cond3)
        raise ValueError("ydata, ymod and yerr should all be the same length")
    a = -1.0 * np.sum(((ydata - ymod) / yerr) ** 2.0)
    b = 2.0 * Npars
    c = ((2.0 * Npars) * (Npars + 1.0)) / (ydata.size - Npars - 1.0)
    return a + b + c
