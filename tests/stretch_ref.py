"""NumPy restatement of the device stretch-move half-step (test double).

Same counter-based RNG (Philox4x32-10, Salmon et al. 2011), same draw layout and
the same arithmetic order as ``stretch_kernel`` in magprop_kernels.cu, with the
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
