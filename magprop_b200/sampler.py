"""Ensemble (stretch-move) drivers for the batched likelihood.

The reference hands its ``lnprob`` to emcee (``synth_mcmc.py:178-185``: default
``StretchMove(a=2)``, ``Pool``-parallel).  emcee is not installed in this image,
so this module carries the minimum needed to run and measure the path:

``EnsembleSampler``     host-side stretch move with emcee's call pattern; the
                        log-probability is any callable, normally
                        ``mcmc_eqns.lnprob_batch`` with ``vectorize=True``
                        (one kernel launch per half-step).
``DeviceEnsemble``      positions stay on the GPU; proposal + likelihood +
                        accept are ONE fused launch per half-step
                        (``mp_stretch_half_step``).  With ``torch.distributed``
                        initialised the active half is split over the ranks and
                        the updated rows are all-gathered each half-step (NCCL
                        over NVLink on GPUs; gloo in the CPU tests).

Move semantics (Goodman & Weare 2010, as emcee's RedBlueMove implements them):
for each half S with complement C:  z = ((a-1)u+1)^2/a,  q = c_j - (c_j - s) z
with j uniform in C (with replacement),  accept iff
(ndim-1) ln z + lp(q) - lp(s) > ln u'.   C includes the updates made by the
first half-step of the same step.
"""
from __future__ import annotations

import numpy as np


class EnsembleSampler:
    """Minimal host-side stand-in for ``emcee.EnsembleSampler`` (stretch move only)."""

    def __init__(self, nwalkers, ndim, log_prob_fn, args=(), kwargs=None, vectorize=True, a=2.0, seed=None):
        if nwalkers % 2 or nwalkers < 2 * ndim:
            raise ValueError("nwalkers must be even and at least 2*ndim")
        self.nwalkers, self.ndim, self.a = nwalkers, ndim, float(a)
        self.log_prob_fn, self.args, self.kwargs = log_prob_fn, tuple(args), dict(kwargs or {})
        self.vectorize = vectorize
        self.random = np.random.RandomState(seed)
        self.chain = None          # [nsteps, nwalkers, ndim]
        self.log_prob = None       # [nsteps, nwalkers]
        self.naccepted = np.zeros(nwalkers, dtype=np.int64)
        self.iteration = 0

    def compute_log_prob(self, coords):
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        if not np.isfinite(coords).all():
            raise ValueError("At least one parameter value was infinite or NaN")
        if self.vectorize:
            lp = np.asarray(self.log_prob_fn(coords, *self.args, **self.kwargs), dtype=np.float64)
        else:
            lp = np.array([self.log_prob_fn(c, *self.args, **self.kwargs) for c in coords], dtype=np.float64)
        if np.isnan(lp).any():
            raise ValueError("Probability function returned NaN")
        return lp

    def run_mcmc(self, p0, nsteps):
        coords = np.array(p0, dtype=np.float64).reshape(self.nwalkers, self.ndim)
        lp = self.compute_log_prob(coords)
        chain = np.empty((nsteps, self.nwalkers, self.ndim))
        lps = np.empty((nsteps, self.nwalkers))
        for it in range(nsteps):
            inds = np.arange(self.nwalkers) % 2
            self.random.shuffle(inds)
            for split in (0, 1):
                S = np.where(inds == split)[0]
                Cset = np.where(inds != split)[0]
                ns = S.size
                zz = ((self.a - 1.0) * self.random.rand(ns) + 1.0) ** 2.0 / self.a
                rint = self.random.randint(Cset.size, size=ns)
                c = coords[Cset[rint]]
                q = c - (c - coords[S]) * zz[:, None]
                new_lp = self.compute_log_prob(q)
                lnpdiff = (self.ndim - 1.0) * np.log(zz) + new_lp - lp[S]
                accept = lnpdiff > np.log(self.random.rand(ns))
                coords[S[accept]] = q[accept]
                lp[S[accept]] = new_lp[accept]
                self.naccepted[S[accept]] += 1
            chain[it] = coords
            lps[it] = lp
            self.iteration += 1
        self.chain = chain if self.chain is None else np.concatenate([self.chain, chain])
        self.log_prob = lps if self.log_prob is None else np.concatenate([self.log_prob, lps])
        return coords, lp

    def get_chain(self):
        return self.chain

    def get_log_prob(self):
        return self.log_prob

    @property
    def acceptance_fraction(self):
        return self.naccepted / max(1, self.iteration)


def rank_slice(n_items: int, rank: int, world: int):
    """Contiguous equal share of n_items for `rank`; n_items must divide evenly."""
    if n_items % world:
        raise ValueError(f"half-ensemble of {n_items} walkers does not split evenly over {world} ranks")
    m = n_items // world
    return rank * m, (rank + 1) * m


class DeviceEnsemble:
    """Ensemble whose positions live on the device; fixed halves [0, n/2) and [n/2, n).

    half_step(coords, lnp, active, complement, a, seed, step, accepted) must perform one
    stretch-move half-step in place for the walkers listed in `active`.  On a GPU it is
    ``Likelihood.stretch_half_step`` (see ``from_likelihood``); tests inject a CPU double
    so the sharding/all-gather logic runs under gloo.
    """

    def __init__(self, half_step, nwalkers, ndim, a=2.0, seed=0, device="cuda", dist=None):
        import torch
        if nwalkers % 2 or nwalkers < 2 * ndim:
            raise ValueError("nwalkers must be even and at least 2*ndim")
        self.torch = torch
        self.half_step, self.nwalkers, self.ndim, self.a, self.seed = half_step, nwalkers, ndim, float(a), int(seed)
        self.device = torch.device(device)
        self.dist = dist if (dist is not None and dist.is_initialized() and dist.get_world_size() > 1) else None
        # gloo (the CPU tests) needs a separate send buffer; NCCL gathers in place
        self.inplace_gather = bool(self.dist) and self.dist.get_backend() == "nccl"
        self.rank = self.dist.get_rank() if self.dist else 0
        self.world = self.dist.get_world_size() if self.dist else 1
        half = nwalkers // 2
        lo, hi = rank_slice(half, self.rank, self.world)
        idx = torch.arange(nwalkers, dtype=torch.int32, device=self.device)
        self.halves = (idx[:half].contiguous(), idx[half:].contiguous())
        self.mine = (self.halves[0][lo:hi].contiguous(), self.halves[1][lo:hi].contiguous())
        self.my_rows = ((lo, hi), (half + lo, half + hi))
        self.coords = torch.empty((nwalkers, ndim), dtype=torch.float64, device=self.device)
        self.lnp = torch.empty(nwalkers, dtype=torch.float64, device=self.device)
        self.accepted = torch.zeros(nwalkers, dtype=torch.int32, device=self.device)
        self.step = 0

    @classmethod
    def from_likelihood(cls, lik, nwalkers, ndim, a=2.0, seed=0, dist=None):
        import torch
        device = torch.device("cuda", lik.device)

        def half_step(coords, lnp, active, complement, a_, seed_, step_, accepted):
            lik.stretch_half_step(coords.data_ptr(), lnp.data_ptr(), coords.shape[0], coords.shape[1],
                                  active.data_ptr(), active.numel(), complement.data_ptr(), complement.numel(),
                                  a_, seed_, step_, accepted.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)

        ens = cls(half_step, nwalkers, ndim, a=a, seed=seed, device=device, dist=dist)
        ens._lik = lik
        return ens

    def set_state(self, coords, lnp):
        t = self.torch
        self.coords.copy_(t.as_tensor(np.asarray(coords), dtype=t.float64).to(self.device)
                          if not isinstance(coords, t.Tensor) else coords)
        self.lnp.copy_(t.as_tensor(np.asarray(lnp), dtype=t.float64).to(self.device)
                       if not isinstance(lnp, t.Tensor) else lnp)

    def initialise(self, p0):
        """Positions from p0 and their lnprob from the likelihood (GPU ensembles only)."""
        self.set_state(p0, np.zeros(self.nwalkers))
        self._lik.lnprob_device(self.coords.data_ptr(), self.nwalkers, self.ndim, self.lnp.data_ptr(), 0, 0,
                                self.torch.cuda.current_stream().cuda_stream)

    def _gather(self, split):
        """All-gather this rank's updated rows of half `split` into the replicated arrays."""
        if not self.dist:
            return
        half = self.nwalkers // 2
        lo, hi = self.my_rows[split]
        base = split * half
        if self.inplace_gather:
            # NCCL's in-place all-gather: this rank's rows already sit at their place in the output
            self.dist.all_gather_into_tensor(self.coords[base:base + half], self.coords[lo:hi])
            self.dist.all_gather_into_tensor(self.lnp[base:base + half], self.lnp[lo:hi])
        else:
            send_c = self.coords[lo:hi].clone()
            send_l = self.lnp[lo:hi].clone()
            self.dist.all_gather_into_tensor(self.coords[base:base + half], send_c)
            self.dist.all_gather_into_tensor(self.lnp[base:base + half], send_l)

    def run(self, nsteps, store=False):
        """nsteps stretch-move steps (2 half-steps each).  Returns the chain
        [nsteps, nwalkers, ndim] and lnprob [nsteps, nwalkers] when store=True."""
        t = self.torch
        chain = t.empty((nsteps, self.nwalkers, self.ndim), dtype=t.float64, device=self.device) if store else None
        lps = t.empty((nsteps, self.nwalkers), dtype=t.float64, device=self.device) if store else None
        for it in range(nsteps):
            for split in (0, 1):
                self.half_step(self.coords, self.lnp, self.mine[split], self.halves[1 - split], self.a, self.seed,
                               2 * self.step + split, self.accepted)
                self._gather(split)
            self.step += 1
            if store:
                chain[it].copy_(self.coords)
                lps[it].copy_(self.lnp)
        return (chain, lps) if store else None

    def acceptance_fraction(self):
        acc = self.accepted.clone()
        if self.dist:
            self.dist.all_reduce(acc)
        return acc.double() / max(1, self.step)


def run_concurrently(ensembles, nsteps, store=False):
    """Advance several independent device ensembles (e.g. one per dataset / burst) ``nsteps`` steps each, every
    ensemble on its own CUDA stream with the launches interleaved step by step, so that small ensembles --
    which are latency-bound, a few warps each -- share the GPU instead of queueing behind one another.
    Returns a list of (chain, lnprob) per ensemble when ``store`` (else None)."""
    import torch
    streams = [torch.cuda.Stream(device=e.device) for e in ensembles]
    for e, st in zip(ensembles, streams):
        st.wait_stream(torch.cuda.current_stream(e.device))
    out = []
    if store:
        for e in ensembles:
            out.append((torch.empty((nsteps, e.nwalkers, e.ndim), dtype=torch.float64, device=e.device),
                        torch.empty((nsteps, e.nwalkers), dtype=torch.float64, device=e.device)))
    for it in range(nsteps):
        for k, (e, st) in enumerate(zip(ensembles, streams)):
            with torch.cuda.stream(st):
                for split in (0, 1):
                    e.half_step(e.coords, e.lnp, e.mine[split], e.halves[1 - split], e.a, e.seed, 2 * e.step + split,
                                e.accepted)
                    e._gather(split)
                e.step += 1
                if store:
                    out[k][0][it].copy_(e.coords)
                    out[k][1][it].copy_(e.lnp)
    for e, st in zip(ensembles, streams):
        torch.cuda.current_stream(e.device).wait_stream(st)
    return out if store else None
