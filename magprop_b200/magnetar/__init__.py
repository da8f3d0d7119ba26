"""Drop-in for the reference's ``magnetar`` package namespace (magnetar/__init__.py:1-3)."""
from .fit_stats import *   # noqa: F401,F403
from .funcs import *       # noqa: F401,F403
from .mcmc_eqns import *   # noqa: F401,F403
