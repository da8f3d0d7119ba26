"""Ball vs posterior-like spread vs prior-uniform ensembles on one dataset (developer tool): evaluations/s,
RHS statistics and the per-warp imbalance (max lane / mean lane RHS count)."""
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
name = os.environ.get("DS", "Classic"); W = int(os.environ.get("W", 1 << 18))
lk = Likelihood(A.script_model_spec(), time_grid(None), g[f"{name}_x"], g[f"{name}_y"], g[f"{name}_yerr"], O.SCRIPT_LOWER, O.SCRIPT_UPPER)
rng = np.random.RandomState(5)
truth = O.SYNTH_TRUTHS_LOG[name]
cases = {"ball 1e-4": truth + 1e-4 * rng.randn(W, 6),
         "spread 0.01": np.clip(truth + 0.01 * rng.randn(W, 6), O.SCRIPT_LOWER, O.SCRIPT_UPPER),
         "spread 0.05": np.clip(truth + 0.05 * rng.randn(W, 6), O.SCRIPT_LOWER, O.SCRIPT_UPPER),
         "spread 0.2": np.clip(truth + 0.2 * rng.randn(W, 6), O.SCRIPT_LOWER, O.SCRIPT_UPPER)}
if os.environ.get("PRIOR"):
    cases["prior uniform"] = rng.uniform(O.SCRIPT_LOWER, O.SCRIPT_UPPER, size=(W, 6))
key = os.environ.get("SORT")
for label, th in cases.items():
    if key == "md": th = th[np.argsort(th[:, 2] + th[:, 5])]
    if key == "r": th = th[np.argsort(th[:, 3])]
    if key == "mdr": th = th[np.lexsort((th[:, 3], np.round((th[:, 2] + th[:, 5]) * 4)))]
    md = th[:, 2] + th[:, 5]
    if key == "e_md": th = th[np.lexsort((md, np.round(th[:, 4] * 4)))]
    if key == "e_mdD": th = th[np.lexsort((-md, np.round(th[:, 4] * 4)))]
    if key == "eD_md": th = th[np.lexsort((md, -np.round(th[:, 4] * 4)))]
    if key == "eD_mdD": th = th[np.lexsort((-md, -np.round(th[:, 4] * 4)))]
    if key == "lpt": th = th[np.argsort(-(0.56 * th[:, 1] + 0.41 * th[:, 2] + 0.35 * th[:, 5] - 0.1 * th[:, 3]))]
    if key == "e_lpt": th = th[np.lexsort((-(0.56 * th[:, 1] + 0.41 * th[:, 2] + 0.35 * th[:, 5]), np.round(th[:, 4] * 4)))]
    if key == "e_mdZ":      # zig-zag: ascending in even bins, descending in odd ones
        eb = np.round(th[:, 4] * 4); th = th[np.lexsort((np.where(eb % 2 == 0, md, -md), eb))]
    if key == "md_e": th = th[np.lexsort((th[:, 4], np.round(md * 4)))]
    if key == "e2_md": th = th[np.lexsort((md, np.round(th[:, 4] * 2)))]
    if key == "md2_e": th = th[np.lexsort((th[:, 4], np.round(md * 2)))]
    if key == "e_r": th = th[np.lexsort((th[:, 3], np.round(th[:, 4] * 4)))]
    if key == "m": th = th[np.argsort(th[:, 2])]
    if key == "d": th = th[np.argsort(th[:, 5])]
    if key == "e_d": th = th[np.lexsort((th[:, 5], np.round(th[:, 4] * 4)))]
    if key == "e_m": th = th[np.lexsort((th[:, 2], np.round(th[:, 4] * 4)))]
    if key == "b": th = th[np.argsort(th[:, 0])]
    if key == "p": th = th[np.argsort(th[:, 1])]
    if key == "e": th = th[np.argsort(th[:, 4])]
    if key == "bp": th = th[np.lexsort((th[:, 1], np.round(np.log10(th[:, 0]) * 8)))]
    if key == "all": th = th[np.lexsort((th[:, 4], np.round(th[:, 3] * 4), np.round(th[:, 1] * 2), np.round(np.log10(th[:, 0]) * 4), np.round((th[:, 2] + th[:, 5]) * 2)))]
    d_th = torch.from_numpy(np.ascontiguousarray(th)).cuda(); d_l = torch.empty(W, dtype=torch.float64, device="cuda"); d_n = torch.empty(W, dtype=torch.int32, device="cuda")
    for _ in range(2): lk.lnprob_device(d_th.data_ptr(), W, 6, d_l.data_ptr(), 0, d_n.data_ptr())
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): lk.lnprob_device(d_th.data_ptr(), W, 6, d_l.data_ptr(), 0, d_n.data_ptr())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    nr = d_n.cpu().numpy().reshape(-1, 32)
    print(f"{label:12s} {ms:7.3f} ms  {W / ms * 1e3:.3e} evals/s  mean_rhs {nr.mean():.0f}  per-warp max {nr.max(1).mean():.0f}  stiff {lk.last_stiff_count()}")
