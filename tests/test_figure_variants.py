"""Row f3: the model copy inlined in the paper-figure scripts (I = 4/5 M R^2, 3*Mdisc, swept n) as a
ModelSpec.  The oracle's figure_spec is pinned bit-for-bit against figure_1.py / figure_4.py's own
``odes`` (oracle/make_goldens_figures.py -> tests/golden/figure_rhs.npz)."""
import os

import numpy as np
import pytest

from conftest import relerr
from magprop_b200 import _capi as A
from oracle import magprop_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden", "figure_rhs.npz")

FIG4 = {"Humped": [1.0, 5.0, 0.001, 100.0, 0.1, 1.0, 10.0], "Classic": [1.0, 5.0, 0.001, 1000.0, 0.1, 1.0, 10.0],      # figure_4.py:114-117
        "Sloped": [1.0, 1.0, 0.001, 100.0, 10.0, 10.0, 1.0], "Stuttering": [1.0, 5.0, 1.e-5, 100.0, 0.1, 100.0, 10.0]}


def test_oracle_figure_rhs_is_the_figure_scripts_rhs():
    g = np.load(GOLD)
    for key in ("figure_1", "figure_4"):
        for row in g[key]:
            y, t, pars, n, want = row[:2], row[2], row[3:8], row[8], row[9:11]
            got = O._rhs_for(O.figure_spec(n=n))(y, t, *pars)
            assert got[0] == want[0] and got[1] == want[1]


@pytest.mark.gpu
def test_figure_rhs_on_device(built):
    from magprop_b200.engine import rhs_batch
    g = np.load(GOLD)
    rows = np.concatenate([g["figure_1"], g["figure_4"]])
    for n in (1.0, 10.0, 50.0):
        r = rows[rows[:, 8] == n]
        got = rhs_batch(A.figure_model_spec(n=n), r[:, :2], r[:, 2], r[:, 3:8], [n, 0.1, 1.0, 0.9])
        assert relerr(got, r[:, 9:11]).max() < 1e-11


@pytest.mark.gpu
def test_figure4_light_curves_vs_oracle(built):
    from magprop_b200.engine import Likelihood, time_grid
    from test_gpu_parity import assert_curves_close
    for name, p in FIG4.items():
        n = p[6]
        lk = Likelihood(A.figure_model_spec(n=n), time_grid(None))
        out, st = lk.curves(np.array([p[:6]]), node_stride=25)
        lk.close()
        assert st[0] == 0
        tight = O.model(np.array(p[:6]), O.figure_spec(n=n), tight=True)
        idx = np.r_[np.arange(0, 10001, 25)]
        if idx[-1] != 10000:
            idx = np.r_[idx, 10000]
        assert_curves_close(out[0], tight[1:][:, idx])
