"""Drop-in for the reference's ``code/synthetic_datasets/funcs.py``.

Same names, argument meaning and return conventions (``init_conds``, ``ODEs``,
``model_lum``, the module constants and ``tarr``); the arithmetic runs in the
CUDA library (one walker per thread) instead of SciPy's odeint.  ``pars`` may
also be a 2-D array [W, 6]: then the batched result is returned.
"""
import numpy as np

from .. import _cache
from .. import _capi as A
from ..engine import rhs_batch

# Global constants (funcs.py:12-19)
G = 6.674e-8
c = 3.0e10
R = 1.0e6
Msol = 1.99e33
M = 1.4 * Msol
I = 0.35 * M * R ** 2.0
GM = G * M
tarr = np.logspace(0.0, 6.0, num=10001, base=10.0)

DEVICE = 0


def init_conds(MdiscI, P_i):
    """funcs.py:51-71: (Mdisc0 [g], omega0 [1/s]) as a tuple."""
    Mdisc0 = MdiscI * Msol
    omega0 = (2.0 * np.pi) / (1.0e-3 * P_i)
    return Mdisc0, omega0


def ODEs(y, t, B, MdiscI, RdiscI, epsilon, delta, n, alpha, cs7, k):
    """funcs.py:75-142: (dMdisc/dt, domega/dt), evaluated on the GPU."""
    spec = A.script_model_spec(n=n, alpha=alpha, cs7=cs7, k=k)
    out = rhs_batch(spec, np.asarray(y, float).reshape(1, 2), [t], [[B, MdiscI, RdiscI, epsilon, delta]],
                    [n, alpha, cs7, k], device=DEVICE)
    return out[0, 0], out[0, 1]


def model_lum(pars, xdata=None, n=10.0, alpha=0.1, cs7=1.0, k=0.9, dipeff=1.0, propeff=1.0, f_beam=1.0):
    """funcs.py:146-236.  Returns the (4, 10001) array [t, Ltot, Lprop, Ldip]/1e50,
    or the luminosity at ``xdata``, or the string 'flag' when the integration failed."""
    spec = A.script_model_spec(n=n, alpha=alpha, cs7=cs7, k=k, dipeff=dipeff, propeff=propeff, f_beam=f_beam,
                               unlog=False)
    p = np.asarray(pars, dtype=np.float64)
    batched = p.ndim == 2
    p2 = np.atleast_2d(p)
    if p2.shape[1] != 6:
        raise ValueError("not enough values to unpack (expected 6)" if p2.shape[1] < 6
                         else "too many values to unpack (expected 6)")
    if xdata is None:
        lk = _cache.get(spec, None, device=DEVICE)
        out, status = lk.curves(p2, node_stride=1)
        res = [("flag" if (s & A.WALKER_INTEGRATOR_FAIL) else np.vstack([tarr[None, :], o])) for o, s in zip(out, status)]
    else:
        x = np.asarray(xdata, dtype=np.float64)
        dummy = np.ones_like(x)
        lk = _cache.get(spec, None, x, dummy, dummy, device=DEVICE)
        out, status = lk.model_at_data(p2, return_status=True)
        res = [("flag" if (s & A.WALKER_INTEGRATOR_FAIL) else o) for o, s in zip(out, status)]
    return res if batched else res[0]
