"""Posterior summaries of a finished chain -- the numerical part of the reference's
``code/synthetic_datasets/plot_synth.py`` (``:150-223``); the plots themselves are out of scope.

``posterior_summary`` reproduces: pairwise correlation coefficients (``:150-157``), the 2.5/50/97.5
percentiles after un-logging columns 2.. (``:160-166``), the best-fit model at the data and on the
grid through the CUDA path (``:180,206``), reduced chi-square / AICc (``:181-191``) and the LaTeX
table row (``:193-203``).  Two reference defects are kept visible rather than silently changed:
the burn-in ``skip`` is computed but never applied (``:142-143``), and ``aicc`` is called with
(ymod, y) swapped (``:183``) -- symmetric in the statistic, so the value is the same."""
import numpy as np

from ..magnetar.fit_stats import aicc, redchisq
from .funcs import model_lum


def posterior_summary(samples, x, y, yerr, grb="", truths=None, n_burn=0):
    samples = np.array(samples, dtype=np.float64)
    Npars = samples.shape[1]
    stats = {"Nburn": n_burn}
    corrs = []
    for i in range(Npars):
        for j in range(i + 1, Npars):
            corrs.append(float(np.corrcoef(samples[:, i], samples[:, j])[0, 1]))
    stats["correlations"] = corrs
    samples[:, 2:] = 10.0 ** samples[:, 2:]                          # out of log-space
    trip = [(v[1], v[2] - v[1], v[1] - v[0]) for v in zip(*np.percentile(samples, [2.5, 50.0, 97.5], axis=0))]
    names = ["B", "P_i", "MdiscI", "RdiscI", "epsilon", "delta"]
    pars = [t[0] for t in trip]
    stats["pars"] = {n: tuple(float(v) for v in t) for n, t in zip(names, trip)}
    if truths is not None:
        stats["pars"]["truths"] = [float(v) for v in truths]
    ymod = model_lum(pars, xdata=x)
    if isinstance(ymod, str):
        raise RuntimeError("the median parameters flag in model_lum")
    y, yerr = np.asarray(y, float), np.asarray(yerr, float)
    stats["stats"] = {"aicc": float(aicc(ymod, y, yerr, Npars)),
                      "chi_square_red": float(redchisq(y, ymod, deg=Npars, sd=yerr))}
    cells = " & ".join("$%s^{+%s}_{-%s}$" % t for t in trip)
    stats["latex"] = "\n\n%s & %s & $%s$ \\\\ [2pt]" % (grb, cells, stats["stats"]["chi_square_red"])
    fit = model_lum(pars)                                            # [t, Ltot, Lprop, Ldip] on the grid
    return stats, pars, ymod, fit
