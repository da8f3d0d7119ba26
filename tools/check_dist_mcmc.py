"""Sharded stretch move over NCCL == single-GPU stretch move, bit for bit (developer check; also run by
tests/test_sampler.py when two GPUs are visible).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_dist_mcmc.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from magprop_b200 import _capi as A
from magprop_b200.engine import Likelihood, time_grid
from magprop_b200.sampler import DeviceEnsemble
from magprop_b200.synthetic.mcmc_eqns import lower, upper
from magprop_b200.synthetic.synth_mcmc import truths

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ.setdefault("NCCL_DEBUG", "NONE")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
g = np.load(os.path.join(ROOT, "tests", "golden", "lnprob_script.npz"))
lk = Likelihood(A.script_model_spec(), time_grid(None), g["Humped_x"], g["Humped_y"], g["Humped_yerr"], lower, upper, device=local)
n = int(os.environ.get("NWALK", 4096))
p0 = truths["Humped"] + 1e-2 * np.random.RandomState(4).randn(n, 6)
sharded = DeviceEnsemble.from_likelihood(lk, n, 6, a=2.0, seed=17, dist=dist)
sharded.initialise(p0)
sharded.run(6)
single = DeviceEnsemble.from_likelihood(lk, n, 6, a=2.0, seed=17, dist=None)
single.initialise(p0)
single.run(6)
torch.cuda.synchronize()
same = torch.equal(sharded.coords, single.coords) and torch.equal(sharded.lnp, single.lnp)
acc = float(sharded.acceptance_fraction().mean().item())
print(f"rank {rank}: sharded == single: {same}; acceptance {acc:.3f}; world {dist.get_world_size()}", flush=True)
ok = torch.tensor([1 if same else 0], device="cuda")
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
lk.close()
dist.destroy_process_group()
sys.exit(0 if int(ok.item()) == 1 else 1)
