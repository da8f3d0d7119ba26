"""Time the lnprob kernel on the bench workload shape (developer A/B tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from magprop_b200 import _capi as A
from magprop_b200.engine import Likelihood, time_grid
from magprop_b200.synthetic.mcmc_eqns import lower as _LO, upper as _HI
from magprop_b200.synthetic.synth_mcmc import truths as _TR


class O:      # the constants these tools need, from the package (the oracle is test infrastructure)
    SCRIPT_LOWER, SCRIPT_UPPER, SYNTH_TRUTHS_LOG = _LO, _HI, _TR
g = np.load(os.path.join(ROOT, "tests/golden/lnprob_script.npz"))
W = int(os.environ.get("W", 262144))
names = os.environ.get("DS", "Classic,Sloped,Stuttering").split(",")
tot = 0.0; res = []
for name in names:
    lk = Likelihood(A.script_model_spec(), time_grid(None), g[f"{name}_x"], g[f"{name}_y"], g[f"{name}_yerr"], O.SCRIPT_LOWER, O.SCRIPT_UPPER)
    rng = np.random.RandomState(1)
    theta = O.SYNTH_TRUTHS_LOG[name] + 1e-4*rng.randn(W,6)
    d_th = torch.from_numpy(theta).cuda(); d_lnp = torch.empty(W, dtype=torch.float64, device="cuda"); d_nr = torch.empty(W, dtype=torch.int32, device="cuda")
    for _ in range(3): lk.lnprob_device(d_th.data_ptr(), W, 6, d_lnp.data_ptr(), 0, d_nr.data_ptr())
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): lk.lnprob_device(d_th.data_ptr(), W, 6, d_lnp.data_ptr(), 0, d_nr.data_ptr())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/5; tot += ms
    res.append("%s %.3f ms %.3e/s nrhs %.0f lnp0 %.10f" % (name, ms, W/ms*1e3, d_nr.float().mean().item(), d_lnp[0].item()))
    lk.close()
print(os.path.basename(os.environ.get("MAGPROP_B200_LIB", "default")), "| total %.3f ms  %.4e evals/s |" % (tot, W*len(names)/tot*1e3), " ; ".join(res))
