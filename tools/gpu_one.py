"""One ensemble shape, a few launches (developer tool for ncu captures):  CASE=ball|s0.01|s0.05|s0.2|prior  W=262144  DS=Classic"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from magprop_b200 import _capi as A
from magprop_b200.engine import Likelihood, time_grid
from magprop_b200.synthetic.mcmc_eqns import lower as LO, upper as HI
from magprop_b200.synthetic.synth_mcmc import truths as TR
g = np.load(os.path.join(ROOT, "tests/golden/lnprob_script.npz"))
name = os.environ.get("DS", "Classic"); W = int(os.environ.get("W", 1 << 18)); case = os.environ.get("CASE", "s0.2")
lk = Likelihood(A.script_model_spec(), time_grid(None), g[f"{name}_x"], g[f"{name}_y"], g[f"{name}_yerr"], LO, HI)
rng = np.random.RandomState(5)
if case == "prior":
    th = rng.uniform(LO, HI, size=(W, 6))
elif case == "ball":
    th = TR[name] + 1e-4 * rng.randn(W, 6)
else:
    th = np.clip(TR[name] + float(case[1:]) * rng.randn(W, 6), LO, HI)
d_th = torch.from_numpy(np.ascontiguousarray(th)).cuda(); d_l = torch.empty(W, dtype=torch.float64, device="cuda"); d_n = torch.empty(W, dtype=torch.int32, device="cuda")
for _ in range(int(os.environ.get("REPS", 3))):
    lk.lnprob_device(d_th.data_ptr(), W, 6, d_l.data_ptr(), 0, d_n.data_ptr())
torch.cuda.synchronize()
print(case, W, "mean_rhs", float(d_n.double().mean()), "stiff", lk.last_stiff_count())
