"""First GPU contact: parity on the goldens + a crude timing sweep."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from magprop_b200 import _capi as A
from magprop_b200.engine import Likelihood, time_grid, fp64_peak_tflops
from magprop_b200.synthetic.mcmc_eqns import lower as _LO, upper as _HI
from magprop_b200.synthetic.synth_mcmc import truths as _TR


class O:      # the constants these tools need, from the package (the oracle is test infrastructure)
    SCRIPT_LOWER, SCRIPT_UPPER, SYNTH_TRUTHS_LOG = _LO, _HI, _TR
print("fp64 peak TFLOP/s:", fp64_peak_tflops(0))
g = np.load(os.path.join(ROOT, "tests/golden/lnprob_script.npz"))
grid = time_grid(None)
def rel(a,b):
    den = np.maximum(np.abs(a),np.abs(b)); return np.where(den>0, np.abs(a-b)/np.where(den>0,den,1), 0)
for di, name in enumerate(g["names"]):
    m = g["dataset"] == di
    th = g["theta"][m]
    lk = Likelihood(A.script_model_spec(), grid, g[f"{name}_x"], g[f"{name}_y"], g[f"{name}_yerr"], O.SCRIPT_LOWER, O.SCRIPT_UPPER)
    t0=time.time(); lnp, st, nr = lk.lnprob(th, return_info=True); dt=time.time()-t0
    ref, tight, fl = g["ref_lnprob"][m], g["tight_lnprob"][m], g["ref_flagged"][m]
    fin = np.isfinite(ref)
    print(name, "time %.3f"%dt, "inf-match", (np.isinf(ref)==np.isinf(lnp))[~fl].all(), "status", np.bincount(st),
          "vs tight max %.2e vs ref max %.2e" % (rel(lnp[fin],tight[fin]).max(), rel(lnp[fin],ref[fin]).max()), "nrhs max", nr.max())
    # throughput: ball around truth
    rng = np.random.RandomState(1)
    for W in (128, 4096, 65536, 1<<20):
        theta = O.SYNTH_TRUTHS_LOG[name] + 1e-4*rng.randn(W,6)
        d_th = torch.from_numpy(theta).cuda(); d_lnp = torch.empty(W, dtype=torch.float64, device="cuda"); d_nr = torch.empty(W, dtype=torch.int32, device="cuda")
        for _ in range(2): lk.lnprob_device(d_th.data_ptr(), W, 6, d_lnp.data_ptr(), 0, d_nr.data_ptr())
        torch.cuda.synchronize()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); lk.lnprob_device(d_th.data_ptr(), W, 6, d_lnp.data_ptr(), 0, d_nr.data_ptr()); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print("   W=%8d  %.3f ms  %.3e evals/s  mean nrhs %.0f" % (W, ms, W/ms*1e3, d_nr.float().mean().item()))
    lk.close()
