"""Ensemble drivers: RNG known answers, host sampler sanity, the multi-rank
driver under gloo (world_size 2, CPU), and the fused device half-step (GPU)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, relerr
import stretch_ref as R
from magprop_b200.sampler import DeviceEnsemble, EnsembleSampler, rank_slice


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    f = lambda c, k: tuple(int(v) for v in R.philox4x32_10(*[np.uint64(x) for x in c], *k))
    assert f((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert f((0xffffffff,) * 4, (0xffffffff,) * 2) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert f((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)
    u = np.concatenate(R.draws(7, 3, np.arange(20000)))
    assert 0.0 < u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.01


def gauss_lnprob(q):
    return -0.5 * np.sum((np.asarray(q) / np.array([1.0, 2.0, 0.5])) ** 2, axis=1)


def test_host_sampler_recovers_gaussian():
    s = EnsembleSampler(40, 3, gauss_lnprob, vectorize=True, seed=1)
    p0 = 0.1 * np.random.RandomState(0).randn(40, 3)
    s.run_mcmc(p0, 1500)
    flat = s.get_chain()[500:].reshape(-1, 3)
    assert np.allclose(flat.std(axis=0), [1.0, 2.0, 0.5], rtol=0.15)
    assert 0.3 < s.acceptance_fraction.mean() < 0.9
    assert s.get_log_prob().shape == (1500, 40)
    with pytest.raises(ValueError):
        EnsembleSampler(5, 3, gauss_lnprob)
    bad = EnsembleSampler(8, 3, lambda q: np.full(len(q), np.nan), seed=0)
    with pytest.raises(ValueError):
        bad.run_mcmc(np.zeros((8, 3)), 1)


def test_rank_slice():
    assert rank_slice(8, 0, 2) == (0, 4) and rank_slice(8, 1, 2) == (4, 8)
    with pytest.raises(ValueError):
        rank_slice(7, 0, 2)


def _cpu_half_step(coords, lnp, active, complement, a, seed, step, accepted):
    c, l, acc = coords.numpy(), lnp.numpy(), accepted.numpy()
    R.half_step(c, l, active.numpy(), complement.numpy(), a, seed, step, gauss_lnprob, acc)


def _run_ensemble(dist, nsteps=25):
    rng = np.random.RandomState(5)
    p0 = rng.randn(32, 3)
    ens = DeviceEnsemble(_cpu_half_step, 32, 3, a=2.0, seed=11, device="cpu", dist=dist)
    ens.set_state(p0, gauss_lnprob(p0))
    chain, lps = ens.run(nsteps, store=True)
    return chain.numpy(), lps.numpy(), ens.acceptance_fraction().numpy()


def _gloo_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    chain, lps, acc = _run_ensemble(dist)
    if rank == 0:
        np.savez(out, chain=chain, lps=lps, acc=acc)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_ensemble_matches_single_rank(tmp_path):
    """world_size-2 gloo run == single-process run, bit for bit (counter-based RNG,
    half split over ranks, all-gather of the updated rows each half-step)."""
    import torch.multiprocessing as mp
    single = _run_ensemble(None)
    out = str(tmp_path / "r0.npz")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(2, port, out), nprocs=2, join=True)
    two = np.load(out)
    assert (two["chain"] == single[0]).all() and (two["lps"] == single[1]).all()
    assert (two["acc"] == single[2]).all()
    assert 0.2 < single[2].mean() < 0.95


@pytest.mark.gpu
def test_device_half_step_matches_numpy_double(built, golden):
    import torch
    from magprop_b200 import _capi as A
    from magprop_b200.engine import Likelihood, time_grid
    from oracle import magprop_oracle as O
    g = golden["lnprob_script"]
    lk = Likelihood(A.script_model_spec(), time_grid(None), g["Humped_x"], g["Humped_y"], g["Humped_yerr"],
                    O.SCRIPT_LOWER, O.SCRIPT_UPPER)
    rng = np.random.RandomState(3)
    n = 64
    p0 = O.SYNTH_TRUTHS_LOG["Humped"] + 1e-2 * rng.randn(n, 6)
    ens = DeviceEnsemble.from_likelihood(lk, n, 6, a=2.0, seed=99)
    ens.initialise(p0)
    torch.cuda.synchronize()
    lnp0 = ens.lnp.cpu().numpy()
    assert relerr(lnp0, lk.lnprob(p0)).max() == 0.0
    # numpy double driven by the SAME device likelihood
    coords, lnp, acc = p0.copy(), lnp0.copy(), np.zeros(n, np.int32)
    half = n // 2
    idx = np.arange(n)
    nsteps = 6
    chain, lps = ens.run(nsteps, store=True)
    torch.cuda.synchronize()
    for it in range(nsteps):
        for split in (0, 1):
            active = idx[:half] if split == 0 else idx[half:]
            comp = idx[half:] if split == 0 else idx[:half]
            R.half_step(coords, lnp, active, comp, 2.0, 99, 2 * it + split, lk.lnprob, acc)
        assert (chain[it].cpu().numpy() == coords).all()
        assert (lps[it].cpu().numpy() == lnp).all()
    assert (ens.accepted.cpu().numpy() == acc).all()
    assert 0 < acc.sum() < n * nsteps
    assert np.isfinite(lnp).all()
    lk.close()


@pytest.mark.gpu
def test_run_concurrently_matches_sequential_runs(built, golden):
    """Three datasets' ensembles advanced side by side (one stream each) give the chains of three separate runs."""
    import torch
    from magprop_b200 import _capi as A
    from magprop_b200.engine import Likelihood, time_grid
    from magprop_b200.sampler import run_concurrently
    from oracle import magprop_oracle as O
    g = golden["lnprob_script"]
    names = ("Classic", "Sloped", "Stuttering")

    def make():
        liks, ens = [], []
        for k, n in enumerate(names):
            lk = Likelihood(A.script_model_spec(), time_grid(None), g[f"{n}_x"], g[f"{n}_y"], g[f"{n}_yerr"],
                            O.SCRIPT_LOWER, O.SCRIPT_UPPER)
            e = DeviceEnsemble.from_likelihood(lk, 64, 6, a=2.0, seed=5 + k)
            e.initialise(O.SYNTH_TRUTHS_LOG[n] + 1e-3 * np.random.RandomState(k).randn(64, 6))
            liks.append(lk); ens.append(e)
        return liks, ens

    liks, ens = make()
    seq = [e.run(8, store=True) for e in ens]
    torch.cuda.synchronize()
    for lk in liks:
        lk.close()
    liks, ens = make()
    par = run_concurrently(ens, 8, store=True)
    torch.cuda.synchronize()
    for (c0, l0), (c1, l1), e in zip(seq, par, ens):
        assert torch.equal(c0, c1) and torch.equal(l0, l1) and e.step == 8
    for lk in liks:
        lk.close()


@pytest.mark.gpu
def test_nccl_sharded_ensemble_matches_single_gpu(built):
    """Two ranks over NCCL, each moving half of every half-ensemble and all-gathering the updates, reproduce the
    single-GPU chain bit for bit (counter-based RNG, no cross-rank state).  Needs two visible GPUs."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(root, "tools", "check_dist_mcmc.py")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("sharded == single: True") == 2
