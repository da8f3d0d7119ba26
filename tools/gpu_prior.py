"""Prior-uniform ensemble: time, RHS-count distribution, failures (developer tool).
Run under `ncu --metrics gpu__time_duration.sum` to split the explicit and the stiff launch."""
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
W = int(os.environ.get("W", 65536)); name = os.environ.get("DS", "Classic")
lk = Likelihood(A.script_model_spec(), time_grid(None), g[f"{name}_x"], g[f"{name}_y"], g[f"{name}_yerr"], O.SCRIPT_LOWER, O.SCRIPT_UPPER)
rng = np.random.RandomState(99)
theta = rng.uniform(O.SCRIPT_LOWER, O.SCRIPT_UPPER, size=(W, 6))
if os.environ.get("SORT"):
    theta = theta[np.argsort(theta[:, 2] + theta[:, 5])]
d_th = torch.from_numpy(theta).cuda(); d_lnp = torch.empty(W, dtype=torch.float64, device="cuda")
d_nr = torch.empty(W, dtype=torch.int32, device="cuda"); d_st = torch.empty(W, dtype=torch.int32, device="cuda")
for _ in range(2): lk.lnprob_device(d_th.data_ptr(), W, 6, d_lnp.data_ptr(), d_st.data_ptr(), d_nr.data_ptr())
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): lk.lnprob_device(d_th.data_ptr(), W, 6, d_lnp.data_ptr(), d_st.data_ptr(), d_nr.data_ptr())
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
nr = d_nr.cpu().numpy(); st = d_st.cpu().numpy()
q = np.percentile(nr, [0, 25, 50, 75, 90, 99, 100]).astype(int)
print("prior-uniform %s W=%d: %.3f ms  %.3e evals/s  stiff hand-overs %d  fails %d" % (name, W, ms, W / ms * 1e3, lk.last_stiff_count(), int(((st & 2) != 0).sum())))
print("  n_rhs min/25/50/75/90/99/max", q.tolist(), "mean %.0f" % nr.mean(), " sum %.3e" % nr.sum())
# per-warp imbalance of the explicit launch: max over the 32 lanes vs mean
w = nr[: (W // 32) * 32].reshape(-1, 32)
print("  per-warp max-lane n_rhs mean %.0f  (lane mean %.0f)" % (w.max(1).mean(), w.mean()))
