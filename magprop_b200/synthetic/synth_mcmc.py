"""The MCMC driver around the likelihood -- what the reference's
``code/synthetic_datasets/synth_mcmc.py`` does with emcee, on the device-resident ensemble.

    res = run("Humped", x, y, yerr, n_walk=50, n_step=500, seed=1)      # fused stretch move on the GPU
    write_outputs(dirname, "Humped", res)                                # the reference's file set

Covered: initial ball (``synth_mcmc.py:175-176``), the sampler run
(``:178-185``), ``{GRB}_chain.csv`` with its ``Npars, Nwalk, Nstep`` header (``:188-194``), the
per-parameter and ``_lnp`` files (``:197-213``), mean acceptance fraction and integrated
autocorrelation time (``:216-221``), ``{GRB}_info.json`` (``:223-226``; written with plain Python
numbers -- the reference's ``json.dump`` of a NumPy array raises).  Not covered: argparse CLI, trace plot.
"""
import json
import os

import numpy as np

truths = {                                                       # synth_mcmc.py:16-21 (log-space for 2..5)
    "Humped": np.array([1.0, 5.0, -3.0, 2.0, -1.0, 0.0]),
    "Classic": np.array([1.0, 5.0, -3.0, 3.0, -1.0, 0.0]),
    "Sloped": np.array([1.0, 1.0, -3.0, 2.0, 1.0, 1.0]),
    "Stuttering": np.array([1.0, 5.0, -5.0, 2.0, -1.0, 2.0]),
}


def initial_ball(grb, n_walk, n_pars=6, rng=None):
    """synth_mcmc.py:175-176: ``truths[grb] + 1e-4*randn(Npars)`` per walker, drawn walker by walker."""
    rng = rng or np.random
    p0 = np.array(truths[grb])
    return np.array([p0 + 1.0e-4 * rng.randn(n_pars) for _ in range(n_walk)])


class RunResult:
    """chain[walker, step, dim] and lnprobability[walker, step] as emcee's ``sampler.chain`` /
    ``sampler.lnprobability`` index them (synth_mcmc.py:193-194), plus the acceptance fractions."""

    def __init__(self, chain_swd, lnp_sw, acceptance_fraction, seed):
        self._chain = np.asarray(chain_swd)          # [step, walker, dim]
        self._lnp = np.asarray(lnp_sw)               # [step, walker]
        self.acceptance_fraction = np.asarray(acceptance_fraction)
        self.seed = seed

    @property
    def chain(self):
        return self._chain.transpose(1, 0, 2)

    @property
    def lnprobability(self):
        return self._lnp.T

    def get_chain(self):
        return self._chain

    def get_log_prob(self):
        return self._lnp

    def get_autocorr_time(self, **kw):
        return integrated_time(self._chain, **kw)


def run(grb, x, y, yerr, n_walk, n_step, seed=0, p0=None, device=0, dist=None, fbad=None):
    """One chain of ``n_step`` stretch-move steps for ``n_walk`` walkers, positions resident on the GPU
    (one fused launch per half-step; with ``dist`` the ensemble is sharded over the ranks).  ``fbad``: the
    reference's bad-parameter file (``synth_mcmc.py:159,181``, ``mcmc_eqns.py:72-79``) -- proposals whose likelihood
    was not finite are logged on the device and appended to it after the run (by rank 0)."""
    from .. import _capi as A
    from ..engine import Likelihood, time_grid
    from ..sampler import DeviceEnsemble
    from .mcmc_eqns import lower, upper

    rng = np.random.RandomState(seed)
    if p0 is None:
        p0 = initial_ball(grb, n_walk, rng=rng)
    p0 = np.asarray(p0, dtype=np.float64)
    lk = Likelihood(A.script_model_spec(), time_grid(None), x, y, yerr, lower, upper, device=device)
    try:
        ens = DeviceEnsemble.from_likelihood(lk, n_walk, p0.shape[1], a=2.0, seed=seed, dist=dist)
        ens.initialise(p0)
        chain, lnp = ens.run(n_step, store=True)
        ens.check_peers()
        res = RunResult(chain.cpu().numpy(), lnp.cpu().numpy(), ens.acceptance_fraction().cpu().numpy(), seed)
        bad, dropped = ens.drain_bad()
        res.bad_rows, res.bad_dropped = bad, dropped
        if fbad is not None and ens.rank == 0 and len(bad):
            from .mcmc_eqns import _write_bad
            _write_bad(fbad, bad)
        ens.close()
    finally:
        lk.close()
    return res


# ---- integrated autocorrelation time (emcee.autocorr, restated; emcee is not installed) -----------
def _next_pow_two(n):
    i = 1
    while i < n:
        i = i << 1
    return i


def function_1d(x):
    """Normalised autocorrelation function of a 1-D series via FFT."""
    x = np.atleast_1d(x)
    if len(x.shape) != 1:
        raise ValueError("invalid dimensions for 1D autocorrelation function")
    n = _next_pow_two(len(x))
    f = np.fft.fft(x - np.mean(x), n=2 * n)
    acf = np.fft.ifft(f * np.conjugate(f))[: len(x)].real
    acf /= acf[0]
    return acf


def _auto_window(taus, c):
    m = np.arange(len(taus)) < c * taus
    if np.any(m):
        return int(np.argmin(m))
    return len(taus) - 1


class AutocorrError(Exception):
    def __init__(self, tau, *args):
        self.tau = tau
        super().__init__(*args)


def integrated_time(x, c=5, tol=50, quiet=False):
    """Integrated autocorrelation time per dimension of a chain [step, walker, dim] (Sokal's automatic
    windowing, as emcee's ``integrated_time``: ACF averaged over walkers, window = first M with
    M >= c*tau(M); raises ``AutocorrError`` when the chain is shorter than ``tol*tau`` unless ``quiet``)."""
    x = np.atleast_1d(x)
    if len(x.shape) == 1:
        x = x[:, np.newaxis, np.newaxis]
    if len(x.shape) == 2:
        x = x[:, :, np.newaxis]
    if len(x.shape) != 3:
        raise ValueError("invalid dimensions")
    n_t, n_w, n_d = x.shape
    tau_est = np.empty(n_d)
    windows = np.empty(n_d, dtype=int)
    for d in range(n_d):
        f = np.zeros(n_t)
        for k in range(n_w):
            f += function_1d(x[:, k, d])
        f /= n_w
        taus = 2.0 * np.cumsum(f) - 1.0
        windows[d] = _auto_window(taus, c)
        tau_est[d] = taus[windows[d]]
    flag = tol * tau_est > n_t
    if np.any(flag):
        msg = ("The chain is shorter than {0} times the integrated autocorrelation time for {1} parameter(s). "
               "Use this estimate with caution and run a longer chain!\n").format(tol, np.sum(flag))
        msg += "N/{0} = {1:.0f};\ntau: {2}".format(tol, n_t / tol, tau_est)
        if not quiet:
            raise AutocorrError(tau_est, msg)
    return tau_est


# ---- the reference's output files ------------------------------------------------------------
def create_filenames(GRB, root="."):
    """synth_mcmc.py:71-105 (the dataset directory must exist; the bad-parameter file is truncated)."""
    data_dirname = os.path.join(root, "data", "synthetic_datasets", GRB)
    if not os.path.exists(data_dirname):
        raise FileNotFoundError("Please make sure your chosen dataset exists.")
    plot_dirname = os.path.join(root, "plots", "synthetic_datasets", GRB)
    os.makedirs(plot_dirname, exist_ok=True)
    fdata = os.path.join(data_dirname, f"{GRB}.csv")
    fchain = os.path.join(data_dirname, f"{GRB}_chain.csv")
    fbad = os.path.join(data_dirname, f"{GRB}_bad.csv")
    finfo = os.path.join(data_dirname, f"{GRB}_info.json")
    fplot = os.path.join(plot_dirname, f"{GRB}_trace.png")
    open(fbad, "w").close()
    return fdata, fchain, fbad, finfo, fplot, data_dirname


def write_outputs(fn, GRB, res, quiet_autocorr=True):
    """Write ``{GRB}_chain.csv``, ``{fn}_{k}.csv``, ``{fn}_lnp.csv`` and ``{GRB}_info.json`` in the
    reference's formats (synth_mcmc.py:188-226).  ``fn`` is the dataset directory; the per-parameter
    files are named ``f"{fn}_{k}.csv"`` exactly as the reference builds them (next to the directory).
    Returns the info dictionary."""
    chain, lnp = res.chain, res.lnprobability                 # [walker, step, dim], [walker, step]
    Nwalk, Nstep, Npars = chain.shape
    fchain = os.path.join(fn, f"{GRB}_chain.csv")
    with open(fchain, "w") as f:
        f.write(f"{Npars}, {Nwalk}, {Nstep}\n")
        rows = np.concatenate([chain.transpose(1, 0, 2).reshape(Nstep * Nwalk, Npars),
                               lnp.T.reshape(Nstep * Nwalk, 1)], axis=1)
        for r in rows:                                        # step-major, walker-minor (:190-194)
            f.write(", ".join(f"{v:.6f}" for v in r) + "\n")
    for k in range(Npars):                                    # one line per step (:197-204)
        with open(f"{fn}_{k}.csv", "w") as f:
            for j in range(Nstep):
                f.write(", ".join(f"{v:.6f}" for v in chain[:, j, k]) + "\n")
    with open(f"{fn}_lnp.csv", "w") as f:                     # (:207-213)
        for j in range(Nstep):
            f.write(", ".join(f"{v:.6f}" for v in lnp[:, j]) + "\n")
    tau = res.get_autocorr_time(quiet=quiet_autocorr)
    info = {"Npars": int(Npars), "Nwalk": int(Nwalk), "Nstep": int(Nstep), "seed": int(res.seed),
            "acceptance_fraction": float(np.mean(res.acceptance_fraction)), "tau": [float(t) for t in tau]}
    with open(os.path.join(fn, f"{GRB}_info.json"), "w") as f:
        json.dump(info, f)
    return info


def read_chain(fchain):
    """Inverse of the ``_chain.csv`` writer: (samples [Nstep*Nwalk, Npars], lnp, (Npars, Nwalk, Nstep)) --
    how plot_synth.py:139-143 reads it back."""
    with open(fchain) as f:
        Npars, Nwalk, Nstep = (int(v) for v in f.readline().split(","))
    arr = np.loadtxt(fchain, delimiter=",", skiprows=1)
    return arr[:, :Npars], arr[:, Npars], (Npars, Nwalk, Nstep)
