"""Where a sharded stretch-move step spends its time (developer tool): per half-step, CUDA-event times of the move
kernels and of the exchange (peer barrier / all-gather + unpack) on every rank, for both exchanges, plus the same
per-rank work with no exchange at all (what the kernels cost by themselves at this ensemble size).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/dist_phases.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from magprop_b200 import _capi as A
from magprop_b200.engine import Likelihood, time_grid
from magprop_b200.sampler import DeviceEnsemble
from magprop_b200.synthetic.mcmc_eqns import lower, upper
from magprop_b200.synthetic.synth_mcmc import truths

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
g = np.load(os.path.join(ROOT, "tests", "golden", "lnprob_script.npz"))
lk = Likelihood(A.script_model_spec(), time_grid(None), g["Classic_x"], g["Classic_y"], g["Classic_yerr"], lower, upper, device=local)
per_rank = int(os.environ.get("PER_RANK", 1 << 18))
n = per_rank * world
nsteps = int(os.environ.get("STEPS", 8))
p0 = truths["Classic"] + 1e-4 * np.random.RandomState(7).randn(n, 6)
for exchange in ("peer", "allgather", "none"):
    ens = DeviceEnsemble.from_likelihood(lk, n, 6, a=2.0, seed=2017, dist=dist, exchange="peer" if exchange == "none" else exchange)
    ens.initialise(p0)
    ens.run(2)
    torch.cuda.synchronize(); dist.barrier()
    ev = []
    t_all0 = torch.cuda.Event(enable_timing=True); t_all1 = torch.cuda.Event(enable_timing=True)
    t_all0.record()
    for it in range(nsteps):
        for split in (0, 1):
            a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            a.record()
            ens.backend.half_step(ens, ens.step, split)
            b.record()
            if exchange == "peer":
                ens._epoch += 1
                ens.backend.barrier(ens, ens._epoch)
            elif exchange == "allgather":
                dist.all_gather_into_tensor(ens.gathered, ens.pack)
                ens.backend.unpack(ens, ens.step, split, ens.gathered)
            c.record()
            ev.append((a, b, c))
        ens.step += 1
    t_all1.record()
    torch.cuda.synchronize()
    move = np.array([a.elapsed_time(b) for a, b, c in ev]); exch = np.array([b.elapsed_time(c) for a, b, c in ev])
    tot = t_all0.elapsed_time(t_all1)
    t = torch.tensor([tot, move.sum(), exch.sum()], dtype=torch.float64, device="cuda")
    mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    mn = t.clone(); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"{exchange:9s} world {world} n {n}: ms/step {mx[0].item() / nsteps:.3f}  move kernels {mn[1].item() / nsteps:.3f}..{mx[1].item() / nsteps:.3f}"
              f"  exchange {mn[2].item() / nsteps:.3f}..{mx[2].item() / nsteps:.3f}  evals/s {n * nsteps / mx[0].item() * 1e3:.3e}", flush=True)
    if exchange != "none":
        ens.check_peers()
    dist.barrier()
    ens.close()
lk.close()
dist.destroy_process_group()
