"""Posterior summaries of a finished chain -- the numerical part of the reference's
``code/synthetic_datasets/plot_synth.py`` (``:150-223``); the plots themselves are out of scope.

``posterior_summary`` reproduces: pairwise correlation coefficients (``:150-157``), the 2.5/50/97.5
percentiles after un-logging columns 2.. (``:160-166``), the best-fit model at the data and on the
grid through the CUDA path (``:180,206``), reduced chi-square / AICc (``:181-191``) and the LaTeX
table row (``:193-203``).  Two reference defects are kept visible rather than silently changed:
the burn-in ``skip`` is computed but never applied (``:142-143``), and ``aicc`` is called with
(ymod, y) swapped (``:183``) -- symmetric in the statistic, so the value is the same."""
import numpy as np

from ..magnetar.fit_stats import aicc, from_lnlike, redchisq
from .funcs import model_lum

NAMES = ["B", "P_i", "MdiscI", "RdiscI", "epsilon", "delta"]
PERCENTILES = (2.5, 50.0, 97.5)


def _latex_row(grb, trip, chisq_r):
    cells = " & ".join("$%s^{+%s}_{-%s}$" % t for t in trip)
    return "\n\n%s & %s & $%s$ \\\\ [2pt]" % (grb, cells, chisq_r)


def posterior_summary_device(chain, x, y, yerr, grb="", truths=None, n_burn=0):
    """``posterior_summary`` for a chain that is still on the GPU (``DeviceEnsemble.run(store=True)[0]``, any
    shape ``[..., Npars]``, a float64 CUDA tensor): the correlations and percentiles are reductions over the chain
    where it lies (``mp_chain_moments``, ``mp_chain_order_statistics``) and the fit statistics come from the
    likelihood kernel's chi-square -- a few dozen numbers cross PCIe instead of the chain.  Same numbers as the host
    version to rounding (the percentiles' two neighbouring order statistics are exact; their interpolation is
    NumPy's own, applied after un-logging as ``plot_synth.py:160-166`` does)."""
    from .. import _cache
    from .. import _capi as A
    from ..engine import chain_moments, chain_order_statistics
    from .mcmc_eqns import DEVICE
    flat = chain.reshape(-1, chain.shape[-1]).contiguous()
    n, Npars = flat.shape
    dev = flat.device.index or 0
    stats = {"Nburn": n_burn}
    _, cov = chain_moments(flat.data_ptr(), n, Npars, device=dev)
    sd = np.sqrt(np.diag(cov))
    stats["correlations"] = [float(cov[i, j] / (sd[i] * sd[j])) for i in range(Npars) for j in range(i + 1, Npars)]
    trip = []
    for col in range(Npars):
        idx = [q / 100.0 * (n - 1) for q in PERCENTILES]                  # np.percentile, method="linear"
        lo = [min(int(np.floor(v)), n - 1) for v in idx]
        ranks = sorted({r for k in lo for r in (k, min(k + 1, n - 1))})
        vals = dict(zip(ranks, chain_order_statistics(flat.data_ptr(), n, Npars, col, ranks, device=dev)))
        out = []
        for v, k in zip(idx, lo):
            pair = np.array([vals[k], vals[min(k + 1, n - 1)]])
            if col >= 2:
                pair = 10.0 ** pair                                        # out of log-space (plot_synth.py:160)
            out.append(float(np.percentile(pair, 100.0 * (v - k))))
        trip.append((out[1], out[2] - out[1], out[1] - out[0]))
    pars = [t[0] for t in trip]
    stats["pars"] = {nm: tuple(float(v) for v in t) for nm, t in zip(NAMES, trip)}
    if truths is not None:
        stats["pars"]["truths"] = [float(v) for v in truths]
    # fit statistics of the median parameters from the kernel's chi-square (lnlike = -chi2/2)
    x, y, yerr = (np.asarray(a, float) for a in (x, y, yerr))
    lk = _cache.get(A.script_model_spec(unlog=False), None, x, y, yerr, device=DEVICE)
    lnl, status, _ = lk.lnprob(np.asarray(pars, float).reshape(1, -1), return_info=True)
    if status[0] & A.WALKER_INTEGRATOR_FAIL:
        raise RuntimeError("the median parameters flag in model_lum")
    chisq_r, aic = from_lnlike(lnl[0], y.size, Npars)
    stats["stats"] = {"aicc": float(aic), "chi_square_red": float(chisq_r)}
    stats["latex"] = _latex_row(grb, trip, stats["stats"]["chi_square_red"])
    return stats, pars


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
    pars = [t[0] for t in trip]
    stats["pars"] = {n: tuple(float(v) for v in t) for n, t in zip(NAMES, trip)}
    if truths is not None:
        stats["pars"]["truths"] = [float(v) for v in truths]
    ymod = model_lum(pars, xdata=x)
    if isinstance(ymod, str):
        raise RuntimeError("the median parameters flag in model_lum")
    y, yerr = np.asarray(y, float), np.asarray(yerr, float)
    stats["stats"] = {"aicc": float(aicc(ymod, y, yerr, Npars)),
                      "chi_square_red": float(redchisq(y, ymod, deg=Npars, sd=yerr))}
    stats["latex"] = _latex_row(grb, trip, stats["stats"]["chi_square_red"])
    fit = model_lum(pars)                                            # [t, Ltot, Lprop, Ldip] on the grid
    return stats, pars, ymod, fit
