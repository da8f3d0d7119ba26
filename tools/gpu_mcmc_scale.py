"""Huge-ensemble stretch move (BASELINE configs[4] shape) under torchrun: ms per MCMC step and evaluations/s for
NWALK walkers sharded over the ranks (in-place NCCL all-gather of the moved rows after every half-step).
    NWALK=8388608 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/gpu_mcmc_scale.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from magprop_b200 import _capi as A
from magprop_b200.engine import Likelihood, time_grid
from magprop_b200.sampler import DeviceEnsemble
from magprop_b200.synthetic.mcmc_eqns import lower, upper
from magprop_b200.synthetic.synth_mcmc import truths

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
os.environ["NCCL_DEBUG"] = "NONE"
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
g = np.load(os.path.join(ROOT, "tests", "golden", "lnprob_script.npz"))
lk = Likelihood(A.script_model_spec(), time_grid(None), g["Humped_x"], g["Humped_y"], g["Humped_yerr"], lower, upper, device=local)
n = int(os.environ.get("NWALK", 1 << 23))
p0 = truths["Humped"] + 1e-2 * np.random.RandomState(4).randn(n, 6)
for _ in (0,):
    ens = DeviceEnsemble.from_likelihood(lk, n, 6, a=2.0, seed=17, dist=dist if world > 1 else None)
    ens.initialise(p0)
    ens.run(2)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ens.run(4); e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 4], device="cuda", dtype=torch.float64)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"NWALK={n} ranks={world} {t.item():.3f} ms/step  {n / t.item() * 1e3:.3e} evals/s  acceptance {ens.acceptance_fraction().mean().item():.3f}", flush=True)
    else:
        ens.acceptance_fraction()
lk.close()
if world > 1: dist.destroy_process_group()
