"""BASELINE configs[0] and configs[1] end to end (developer tool): the MCMC itself, not just lnprob.

configs[0]  Humped, 50 walkers x 500 steps          -- ours on the GPU (device-resident stretch move) AND the
            reference path on this box's host cores (oracle lnprob = the reference's odeint path, emcee's move
            restated in magprop_b200.sampler.EnsembleSampler, multiprocessing.Pool as synth_mcmc.py:178 does)
configs[1]  Classic / Sloped / Stuttering, 256 walkers x 2000 steps, ours on one GPU

Prints wall times and the posterior medians with their 2.5/97.5 percentiles next to the truths; the two
configs[0] runs are different random chains of the same posterior, so their medians are compared in units of
the posterior width.
    python tools/run_configs.py [--no-cpu]
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np

from magprop_b200.synthetic import synth_mcmc as S
from magprop_b200.sampler import EnsembleSampler


def _oracle_lnprob(theta, x, y, yerr):
    from oracle import magprop_oracle as O
    return O.lnprob(theta, x, y, yerr, O.script_spec(), O.SCRIPT_LOWER, O.SCRIPT_UPPER)


class PoolLnprob:
    def __init__(self, pool, x, y, yerr):
        self.pool, self.args = pool, (x, y, yerr)

    def __call__(self, coords):
        return np.array(self.pool.starmap(_oracle_lnprob, [(c, *self.args) for c in coords]))


def summary(name, chain, burn):
    flat = chain[burn:].reshape(-1, chain.shape[-1])
    lo, med, hi = np.percentile(flat, [2.5, 50, 97.5], axis=0)
    print(f"  {name}: median {np.round(med, 3).tolist()}\n  {'':{len(name)}}  2.5%   {np.round(lo, 3).tolist()}\n  {'':{len(name)}}  97.5%  {np.round(hi, 3).tolist()}")
    return med, lo, hi


def main():
    g = np.load(os.path.join(ROOT, "tests", "golden", "lnprob_script.npz"))
    data = {n: (g[f"{n}_x"], g[f"{n}_y"], g[f"{n}_yerr"]) for n in ("Humped", "Classic", "Sloped", "Stuttering")}
    print("configs[0]: Humped, 50 walkers x 500 steps; truth", S.truths["Humped"].tolist())
    S.run("Humped", *data["Humped"], n_walk=50, n_step=5, seed=1)            # warm-up (context, first launch)
    t0 = time.perf_counter()
    res = S.run("Humped", *data["Humped"], n_walk=50, n_step=500, seed=1)
    t_gpu = time.perf_counter() - t0
    print(f"  ours (1 GPU, fused stretch move): {t_gpu:.3f} s for {50 * 501} evaluations, acceptance {res.acceptance_fraction.mean():.3f}")
    med_g, lo_g, hi_g = summary("ours", res.get_chain(), 250)
    if "--no-cpu" not in sys.argv:
        from multiprocessing import get_context
        with get_context("fork").Pool(os.cpu_count()) as pool:
            smp = EnsembleSampler(50, 6, PoolLnprob(pool, *data["Humped"]), vectorize=True, seed=2)
            p0 = S.initial_ball("Humped", 50, rng=np.random.RandomState(2))
            t0 = time.perf_counter()
            smp.run_mcmc(p0, 500)
            t_cpu = time.perf_counter() - t0
        print(f"  reference path ({os.cpu_count()} host cores, Pool): {t_cpu:.1f} s, acceptance {smp.acceptance_fraction.mean():.3f}  -> {t_cpu / t_gpu:.0f}x")
        med_c, lo_c, hi_c = summary("ref ", smp.get_chain(), 250)
        width = 0.5 * ((hi_g - lo_g) + (hi_c - lo_c)) / 2 / 1.96
        print("  |median difference| / posterior sigma:", np.round(np.abs(med_g - med_c) / width, 2).tolist())
    print("configs[1]: 256 walkers x 2000 steps on one GPU")
    for n in ("Classic", "Sloped", "Stuttering"):
        t0 = time.perf_counter()
        res = S.run(n, *data[n], n_walk=256, n_step=2000, seed=3)
        dt = time.perf_counter() - t0
        tau = res.get_autocorr_time(quiet=True)
        print(f" {n}: {dt:.2f} s for {256 * 2001} evaluations ({256 * 2001 / dt:.3e} evals/s, {2000 / dt:.0f} steps/s), acceptance "
              f"{res.acceptance_fraction.mean():.3f}, tau {np.round(tau, 1).tolist()}; truth {S.truths[n].tolist()}")
        summary(n, res.get_chain(), 1000)


def concurrent_config1(data):
    """configs[1] with the three datasets' chains advancing side by side (one stream each)."""
    import torch
    from magprop_b200 import _capi as A
    from magprop_b200.engine import Likelihood, time_grid
    from magprop_b200.sampler import DeviceEnsemble, run_concurrently
    from magprop_b200.synthetic.mcmc_eqns import lower, upper
    liks, ens = [], []
    for i, n in enumerate(("Classic", "Sloped", "Stuttering")):
        lk = Likelihood(A.script_model_spec(), time_grid(None), *data[n], lower, upper)
        e = DeviceEnsemble.from_likelihood(lk, 256, 6, a=2.0, seed=3)
        e.initialise(S.initial_ball(n, 256, rng=np.random.RandomState(3)))
        liks.append(lk); ens.append(e)
    run_concurrently(ens, 5)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = run_concurrently(ens, 2000, store=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"configs[1], three datasets side by side: {dt:.2f} s for {3 * 256 * 2000} evaluations ({3 * 256 * 2000 / dt:.3e} evals/s)")
    for (chain, lnp), n in zip(res, ("Classic", "Sloped", "Stuttering")):
        print(f"  {n}: final mean lnprob {lnp[-1].mean().item():.2f}, median {np.round(np.median(chain[1000:].cpu().numpy().reshape(-1, 6), axis=0), 3).tolist()}")
    for lk in liks:
        lk.close()


if __name__ == "__main__":
    main()
    g_ = np.load(os.path.join(ROOT, "tests", "golden", "lnprob_script.npz"))
    concurrent_config1({n: (g_[f"{n}_x"], g_[f"{n}_y"], g_[f"{n}_yerr"]) for n in ("Classic", "Sloped", "Stuttering")})
