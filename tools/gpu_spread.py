"""Time lnprob on a posterior-like spread (developer tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from magprop_b200 import _capi as A
from magprop_b200.engine import Likelihood, time_grid
from oracle import magprop_oracle as O
g = np.load(os.path.join(ROOT, "tests/golden/lnprob_script.npz"))
W = int(os.environ.get("W", 262144)); name = os.environ.get("DS", "Classic"); sig = float(os.environ.get("SIG", 0.05))
lk = Likelihood(A.script_model_spec(), time_grid(None), g[f"{name}_x"], g[f"{name}_y"], g[f"{name}_yerr"], O.SCRIPT_LOWER, O.SCRIPT_UPPER)
rng = np.random.RandomState(99)
theta = np.clip(O.SYNTH_TRUTHS_LOG[name] + sig*rng.randn(W,6), O.SCRIPT_LOWER, O.SCRIPT_UPPER)
if os.environ.get("SORT"):
    theta = theta[np.lexsort((theta[:,4], theta[:,3]))]
d_th = torch.from_numpy(theta).cuda(); d_lnp = torch.empty(W, dtype=torch.float64, device="cuda"); d_nr = torch.empty(W, dtype=torch.int32, device="cuda")
for _ in range(3): lk.lnprob_device(d_th.data_ptr(), W, 6, d_lnp.data_ptr(), 0, d_nr.data_ptr())
torch.cuda.synchronize()
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): lk.lnprob_device(d_th.data_ptr(), W, 6, d_lnp.data_ptr(), 0, d_nr.data_ptr())
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)/5
print("%s sig=%g W=%d: %.3f ms %.3e evals/s nrhs mean %.0f max %d stiff %d" % (name, sig, W, ms, W/ms*1e3, d_nr.float().mean().item(), d_nr.max().item(), lk.last_stiff_count()))
