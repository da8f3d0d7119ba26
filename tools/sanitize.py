"""Small launches of every kernel for compute-sanitizer (developer tool):
    compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from magprop_b200 import _capi as A
from magprop_b200.engine import Likelihood, time_grid, chain_moments, chain_order_statistics, gompertz_curves
from magprop_b200.sampler import DeviceEnsemble
from magprop_b200.synthetic.mcmc_eqns import lower as LO, upper as HI
from magprop_b200.synthetic.synth_mcmc import truths as TR
g = np.load(os.path.join(ROOT, "tests/golden/lnprob_script.npz"))
s = np.load(os.path.join(ROOT, "tests/golden/sgrb_sample.npz"))
rng = np.random.RandomState(1)
lk = Likelihood(A.script_model_spec(), time_grid(None), g["Classic_x"], g["Classic_y"], g["Classic_yerr"], LO, HI)
th = rng.uniform(LO - 0.05, HI + 0.05, size=(12000, 6))                  # ordered launch, stiff queue, prior rejects
lnp, st, nr = lk.lnprob(th, return_info=True)
print("lnprob ordered:", np.isfinite(lnp).sum(), "finite; stiff", lk.last_stiff_count())
print("lnprob small:", lk.lnprob(th[:100]).shape)
pars = th[:64].copy(); pars[:, 2:] = 10.0 ** pars[:, 2:]
print("model_at_data:", lk.model_at_data(np.clip(pars, 1e-6, None)).shape, " curves:", lk.curves(np.clip(pars[:8], 1e-6, None), node_stride=50, with_state=True)[0].shape)
ens = DeviceEnsemble.from_likelihood(lk, 64, 6, seed=3)
ens.initialise(TR["Classic"] + 1e-3 * rng.randn(64, 6))
chain, _ = ens.run(4, store=True)
big = DeviceEnsemble.from_likelihood(lk, 20000, 6, seed=4)               # ordered move
big.initialise(np.clip(TR["Classic"] + 0.2 * rng.randn(20000, 6), LO, HI)); big.run(2)
torch.cuda.synchronize()
print("mcmc:", float(ens.acceptance_fraction().mean()), float(big.acceptance_fraction().mean()))
n = chain.shape[0] * chain.shape[1]
print("chain:", chain_moments(chain.data_ptr(), n, 6)[0][:2], chain_order_statistics(chain.data_ptr(), n, 6, 1, [0, n // 2, n - 1]))
grb = "100212A"; keep = s[f"{grb}_keep"]
l3 = Likelihood(A.packaged_model_spec(), time_grid("S"), s[f"{grb}_t"][keep], s[f"{grb}_Lum50"][keep], s[f"{grb}_Lum50err"][keep],
                s["lims_lower"][:6], s["lims_upper"][:6])
print("coop reduce (410 points):", np.isfinite(l3.lnprob(s[f"{grb}_theta"])).sum())
print("gompertz:", gompertz_curves([[1.0, 5.0, 1e-3, 100.0, 1.0, 1e-6]], n_steps=2000, stride=100)[1].shape)
lk.close(); l3.close()
