"""Time model-curve generation (BASELINE config 4 shape) on device buffers (developer tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from magprop_b200 import _capi as A
from magprop_b200.engine import Likelihood, time_grid
from magprop_b200.synthetic.mcmc_eqns import lower as _LO, upper as _HI
from magprop_b200.synthetic.synth_mcmc import truths as _TR


class O:      # the constants these tools need, from the package (the oracle is test infrastructure)
    SCRIPT_LOWER, SCRIPT_UPPER, SYNTH_TRUTHS_LOG = _LO, _HI, _TR

W = int(os.environ.get("W", 16384))
rng = np.random.RandomState(5)
lk = Likelihood(A.script_model_spec(), time_grid(None))
for label, theta in (("ball", O.SYNTH_TRUTHS_LOG["Humped"] + 0.05 * rng.randn(W, 6)),
                     ("prior", rng.uniform(O.SCRIPT_LOWER, O.SCRIPT_UPPER, size=(W, 6)))):
    pars = theta.copy(); pars[:, 2:] = 10.0 ** pars[:, 2:]
    d_p = torch.from_numpy(pars).cuda()
    for stride in (1, 10):
        Gs = lk.curve_nodes(stride)
        d_out = torch.empty((W, 3, Gs), dtype=torch.float64, device="cuda")
        d_st = torch.empty(W, dtype=torch.int32, device="cuda")
        for _ in range(2):
            lk.curves_device(d_p.data_ptr(), W, 6, stride, d_out.data_ptr(), 0, d_st.data_ptr())
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            lk.curves_device(d_p.data_ptr(), W, 6, stride, d_out.data_ptr(), 0, d_st.data_ptr())
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        gb = W * 3 * Gs * 8 / 1e9
        print("curves %-5s W=%d stride=%d Gs=%d: %.2f ms  %.3e curves/s  %.1f GB/s written  checksum %.10e  fails %d" % (
            label, W, stride, Gs, ms, W / ms * 1e3, gb / ms * 1e3, float(torch.nan_to_num(d_out).sum().item()),
            int((d_st & 2).ne(0).sum().item())), flush=True)
