"""Drop-in for the reference's ``magnetar/funcs.py`` (``init_conds``, ``odes``,
``model_lc`` and the module constants); arithmetic runs in the CUDA library."""
import numpy as np

from .. import _cache
from .. import _capi as A
from ..engine import rhs_batch, time_grid

# Global constants (magnetar/funcs.py:7-13)
G = 6.674e-8
c = 3.0e10
R = 1.0e6
Msol = 1.99e33
M = 1.4 * Msol
I = (4.0 / 5.0) * M * (R ** 2.0)
GM = G * M

DEVICE = 0

__all__ = ["G", "c", "R", "Msol", "M", "I", "GM", "init_conds", "odes", "model_lc"]


def init_conds(MdiscI, P):
    """magnetar/funcs.py:17-29: array([Mdisc0 [g], omega0 [1/s]])."""
    Mdisc0 = MdiscI * Msol
    omega0 = (2.0 * np.pi) / (1.0e-3 * P)
    return np.array([Mdisc0, omega0])


def odes(y, t, B, MdiscI, RdiscI, epsilon, delta, n=1.0, alpha=0.1, cs7=1.0, k=0.9):
    """magnetar/funcs.py:33-101: array([dMdisc/dt, domega/dt]), evaluated on the GPU."""
    spec = A.packaged_model_spec()
    out = rhs_batch(spec, np.asarray(y, float).reshape(1, 2), [t], [[B, MdiscI, RdiscI, epsilon, delta]],
                    [n, alpha, cs7, k], device=DEVICE)
    return out[0].copy()


def model_lc(pars, xdata=None, GRBtype=None, dipeff=0.05, propeff=0.4, f_beam=1.0, n=1.0, alpha=0.1, cs7=1.0,
             k=0.9):
    """magnetar/funcs.py:105-220.  (4, 10001) array [t, Ltot, Lprop, Ldip]/1e50, or
    the luminosity at ``xdata``, or 'flag'.  As in the reference, n/alpha/cs7/k reach
    only the luminosity stage (:150-151) and Lprop is identically 0 (:193)."""
    grid = time_grid(GRBtype)
    spec = A.packaged_model_spec(dipeff=dipeff, propeff=propeff, f_beam=f_beam, n=n, alpha=alpha, cs7=cs7, k=k)
    p = np.asarray(pars, dtype=np.float64)
    batched = p.ndim == 2
    p2 = np.atleast_2d(p)
    if p2.shape[1] != 6:
        raise ValueError("not enough values to unpack (expected 6)" if p2.shape[1] < 6
                         else "too many values to unpack (expected 6)")
    if xdata is None:
        lk = _cache.get(spec, GRBtype, device=DEVICE)
        out, status = lk.curves(p2, node_stride=1)
        res = [("flag" if (s & A.WALKER_INTEGRATOR_FAIL) else np.vstack([grid[None, :], o])) for o, s in zip(out, status)]
    else:
        x = np.asarray(xdata, dtype=np.float64)
        dummy = np.ones_like(x)
        lk = _cache.get(spec, GRBtype, x, dummy, dummy, device=DEVICE)
        out, status = lk.model_at_data(p2, return_status=True)
        res = [("flag" if (s & A.WALKER_INTEGRATOR_FAIL) else o) for o, s in zip(out, status)]
    return res if batched else res[0]
