"""NumPy restatement of the device stretch-move half-step (test double).

Same counter-based RNG (Philox4x32-10, Salmon et al. 2011), same draw layout and
the same arithmetic order as ``propose()`` and ``deliver()`` in magprop_kernels.cu, with the
log-probability supplied by the caller.  Used (a) to check the kernel's RNG and
accept/reject logic on the GPU and (b) as the CPU stand-in for the kernel in the
gloo tests of the multi-rank driver."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over the counter words (uint64 arrays holding 32-bit values)."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & MASK for c in (c0, c1, c2, c3)]
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
        n1 = p1 & MASK
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
        n3 = p0 & MASK
        c0, c1, c2, c3 = n0 & MASK, n1, n2 & MASK, n3
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def u01(hi, lo):
    v = ((hi << np.uint64(32)) | lo) >> np.uint64(11)
    return (v.astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def draws(seed, step, walkers):
    """(u_z, u_partner, u_accept) for each walker index."""
    walkers = np.asarray(walkers, dtype=np.uint64)
    s_lo, s_hi = step & 0xFFFFFFFF, (step >> 32) & 0xFFFFFFFF
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    z = np.zeros_like(walkers)
    a = philox4x32_10(z + np.uint64(s_lo), z + np.uint64(s_hi), walkers, z, k0, k1)
    b = philox4x32_10(z + np.uint64(s_lo), z + np.uint64(s_hi), walkers, z + np.uint64(1), k0, k1)
    return u01(a[0], a[1]), u01(a[2], a[3]), u01(b[0], b[1])


def half_step(coords, lnp, active, complement, a, seed, step, lnprob_fn, accepted=None):
    """In-place half-step on NumPy arrays; returns the proposals and the accept mask."""
    active = np.asarray(active)
    complement = np.asarray(complement)
    ndim = coords.shape[1]
    uz, up, ua = draws(seed, step, active)
    zr = (a - 1.0) * uz + 1.0
    z = zr * zr / a
    pj = np.minimum((up * complement.size).astype(np.int64), complement.size - 1)
    c = coords[complement[pj]]
    x = coords[active]
    q = c + -((c + -x) * z[:, None])
    lp_new = np.asarray(lnprob_fn(q), dtype=np.float64)
    lnpdiff = ((ndim - 1.0) * np.log(z) + lp_new) + -lnp[active]
    acc = lnpdiff > np.log(ua)
    coords[active[acc]] = q[acc]
    lnp[active[acc]] = lp_new[acc]
    if accepted is not None:
        accepted[active[acc]] += 1
    return q, acc


# ---- the ensemble order: keyed permutation of [0, n) (perm_at / make_split_perm in magprop_kernels.cu) ----
def _mix32(x):
    x = x & MASK
    x ^= x >> np.uint64(16); x = (x * np.uint64(0x7feb352d)) & MASK
    x ^= x >> np.uint64(15); x = (x * np.uint64(0x846ca68b)) & MASK
    x ^= x >> np.uint64(16)
    return x


def ensemble_order(n, seed, step, randomize=True):
    """order[g] = walker at position g of the ensemble order of MCMC step `step`; halves are g < n/2, g >= n/2."""
    g = np.arange(n, dtype=np.uint64)
    if not randomize:
        return g.astype(np.int64)
    hb = 1
    while (1 << (2 * hb)) < n:
        hb += 1
    mask = np.uint64((1 << hb) - 1)
    z = np.zeros(1, dtype=np.uint64)
    k = philox4x32_10(z + np.uint64(step & 0xFFFFFFFF), z + np.uint64((step >> 32) & 0xFFFFFFFF), z, z + np.uint64(2),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    k0, k1 = np.uint64(int(k[0][0])), np.uint64(int(k[1][0]))
    x = g.copy()
    todo = np.ones(n, dtype=bool)
    while todo.any():
        xx = x[todo]
        L, R = xx >> np.uint64(hb), xx & mask
        for r in range(6):
            F = _mix32(((R + np.uint64((r * 0x9E3779B9) & 0xFFFFFFFF)) & MASK) ^ (k1 if r & 1 else k0)) & mask
            L, R = R, L ^ F
        xx = (L << np.uint64(hb)) | R
        x[todo] = xx
        todo[todo] = xx >= np.uint64(n)
    return x.astype(np.int64)


class NumpyBackend:
    """CPU stand-in for ``magprop_b200.sampler.CudaBackend``: the same half-step on NumPy views of the
    ensemble's (CPU) tensors, with the log-probability supplied by the caller."""

    peer_capable = False

    def __init__(self, lnprob_fn):
        self.lnprob_fn = lnprob_fn

    def _sets(self, ens, step, split):
        n, half = ens.nwalkers, ens.nwalkers // 2
        order = ensemble_order(n, ens.seed, step, ens.randomize_split)
        m = half // ens.world
        pos0 = split * half + ens.rank * m
        return order, order[pos0:pos0 + m], order[(1 - split) * half:(1 - split) * half + half]

    def half_step(self, ens, step, split):
        coords, lnp, acc = ens.coords.numpy(), ens.lnp.numpy(), ens.accepted.numpy()
        _, active, comp = self._sets(ens, step, split)
        half_step(coords, lnp, active, comp, ens.a, ens.seed, 2 * step + split, self.lnprob_fn, acc)
        if ens.pack is not None:
            p = ens.pack.numpy()
            p[:, :ens.ndim] = coords[active]
            p[:, ens.ndim] = lnp[active]

    def unpack(self, ens, step, split, gathered):
        order, _, _ = self._sets(ens, step, split)
        half = ens.nwalkers // 2
        rows = order[split * half:split * half + half]
        g = gathered.numpy()
        ens.coords.numpy()[rows] = g[:, :ens.ndim]
        ens.lnp.numpy()[rows] = g[:, ens.ndim]

    def lnprob_all(self, ens):
        ens.lnp.numpy()[:] = self.lnprob_fn(ens.coords.numpy())
