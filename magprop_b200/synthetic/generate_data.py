"""Synthetic GRB datasets -- the recipe of the reference's
``code/synthetic_datasets/generate_data.py`` with the light curve from the CUDA path.

    x, y, yerr = generate("Humped", seed=20170613)

``generate_data.py:58-67``: ``model = model_lum(pars)``; 50 random grid nodes (sorted, duplicates
allowed); ``yerr = 0.25*y``; ``y += N(0, yerr)``.  The reference never seeds NumPy; pass ``seed`` to
make a dataset reproducible (bench.py and the goldens use 20170613)."""
import os

import numpy as np

from .funcs import model_lum

GRBs = {                                                         # generate_data.py:10-15
    "Humped": np.array([1.0, 5.0, 1.0e-3, 100.0, 0.1, 1.0]),
    "Classic": np.array([1.0, 5.0, 1.0e-3, 1000.0, 0.1, 1.0]),
    "Sloped": np.array([1.0, 1.0, 1.0e-3, 100.0, 10.0, 10.0]),
    "Stuttering": np.array([1.0, 5.0, 1.0e-5, 100.0, 0.1, 100.0]),
}


def create_filenames(GRB, root="."):
    """generate_data.py:33-44: ``data/synthetic_datasets/<GRB>/<GRB>.csv`` (directories created)."""
    dirname = os.path.join(root, "data", "synthetic_datasets", GRB)
    os.makedirs(dirname, exist_ok=True)
    return os.path.join(dirname, "{0}.csv".format(GRB))


def generate(grb, seed=None, n_points=50, rng=None):
    """(x, y, yerr) of one synthetic dataset; same draw order as generate_data.py:61-67."""
    rng = rng or (np.random.RandomState(seed) if seed is not None else np.random)
    model = model_lum(GRBs[grb])
    if isinstance(model, str):
        raise RuntimeError("model_lum flagged the truth parameters")
    inx = np.sort(rng.randint(low=0, high=model.shape[1], size=n_points))
    x = model[0, inx]
    y = model[1, inx].copy()
    yerr = 0.25 * y
    y += rng.normal(loc=0.0, scale=yerr, size=len(yerr))
    return x, y, yerr


def write_csv(path, x, y, yerr):
    """generate_data.py:70-71: columns x,y,yerr (``DataFrame.to_csv(index=False)`` formatting)."""
    import pandas as pd
    pd.DataFrame({"x": x, "y": y, "yerr": yerr}).to_csv(path, index=False)
