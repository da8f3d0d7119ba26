"""Config 3 shape (developer tool): independent fits of 15 short-GRB-like datasets -- packaged model,
"S" grid (1e-3..1e6 s), the real sample's points-per-burst -- sharded over the ranks with no communication.
The k-corrected sample itself lives under /root/reference/data (not on the GPU box), so the light curves
here are the model at a physical truth + 25 % noise on log-uniform rest-frame time stamps; what is
measured is lnprob evaluations/s per burst and for the whole sample.
    python tools/gpu_config3.py            (or under torchrun for N ranks)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from magprop_b200 import _capi as A
from magprop_b200.engine import Likelihood, time_grid

D_PER_GRB = [253, 80, 33, 1944, 19, 8, 112, 36, 52, 214, 410, 240, 172, 151, 63]
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dev = torch.cuda.current_device()
W = int(os.environ.get("W", 65536))
truth = np.array([2.0, 3.0, 3e-3, 300.0, 1.0, 5.0])
lower = np.array([1e-3, 0.69, 1e-5, 50.0, 0.1, 1e-5]); upper = np.array([10.0, 10.0, 1e-1, 2000.0, 1000.0, 50.0])
grid = time_grid("S")
rng = np.random.RandomState(3)
spec = A.packaged_model_spec()
tot_evals, tot_ms = 0, 0.0
for i, D in enumerate(D_PER_GRB):
    t = np.sort(10 ** rng.uniform(np.log10(0.011), np.log10(9e5), D))
    if i % world != rank:
        continue
    lk0 = Likelihood(spec, grid, t, np.ones(D), np.ones(D), device=dev)
    y = lk0.model_at_data(truth)[0]; lk0.close()
    yerr = 0.25 * y; y = y + rng.normal(0, yerr)
    lk = Likelihood(spec, grid, t, y, yerr, lower, upper, device=dev)
    th = truth * (1 + 1e-3 * rng.randn(W, 6))
    d_th = torch.from_numpy(th).cuda(); d_lnp = torch.empty(W, dtype=torch.float64, device="cuda"); d_nr = torch.empty(W, dtype=torch.int32, device="cuda")
    for _ in range(2): lk.lnprob_device(d_th.data_ptr(), W, 6, d_lnp.data_ptr(), 0, d_nr.data_ptr())
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): lk.lnprob_device(d_th.data_ptr(), W, 6, d_lnp.data_ptr(), 0, d_nr.data_ptr())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    tot_evals += W; tot_ms += ms
    print(f"rank {rank} GRB#{i:2d} D={D:5d} nodes={lk.D:5d}  {ms:8.3f} ms  {W / ms * 1e3:.3e} evals/s  mean_rhs {d_nr.double().mean().item():.0f} finite {torch.isfinite(d_lnp).float().mean().item():.3f}")
    lk.close()
print(f"rank {rank}: {tot_evals} evals in {tot_ms:.2f} ms -> {tot_evals / tot_ms * 1e3:.3e} evals/s")
