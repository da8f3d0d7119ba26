import ctypes as C
import json
import os
import warnings
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def built():
    import __graft_entry__ as g
    g.build()
    return True


def relerr(a, b):
    a = np.asarray(a, float); b = np.asarray(b, float)
    den = np.maximum(np.abs(a), np.abs(b))
    with np.errstate(all="ignore"):
        r = np.where(den > 0, np.abs(a - b) / np.where(den > 0, den, 1.0), 0.0)
    r = np.where(np.isinf(a) & np.isinf(b) & (a == b), 0.0, r)
    return r


@pytest.fixture(scope="session")
def golden():
    return {n: np.load(os.path.join(GOLDEN, f"{n}.npz")) for n in
            ("reference_fixtures", "curves_script", "curves_packaged", "lnprob_script", "lnprob_packaged")}


@pytest.fixture(scope="session")
def hostsim(built):
    """The per-walker core compiled for the host (test aid; not the product)."""
    from magprop_b200 import _capi as A
    hs = C.CDLL(os.path.join(ROOT, "tests", "hostsim", "_hostsim.so"))
    vp = C.c_void_p
    hs.hs_curves.argtypes = [C.POINTER(A.ModelSpec), vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp]
    hs.hs_lnprob.argtypes = [C.POINTER(A.ModelSpec), C.POINTER(A.PriorSpec), vp, C.c_int, vp, vp, vp, C.c_int, vp,
                             C.c_int, C.c_int, vp, vp, vp, vp]
    hs.hs_model_at.argtypes = [C.POINTER(A.ModelSpec), vp, C.c_int, vp, C.c_int, vp, C.c_int, C.c_int, vp, vp]
    hs.hs_node_program.argtypes = [vp, C.c_int, vp, vp, vp, C.c_int, vp, vp, vp, vp, vp, vp]
    hs.hs_disc_S.restype = C.c_double
    hs.hs_disc_S.argtypes = [C.c_double]
    return hs


def parity_stats(got, ref, tight):
    e_ref, e_tight = relerr(got, ref), relerr(got, tight)
    q = lambda e: {"max": float(e.max()), "p99": float(np.percentile(e, 99)), "median": float(np.median(e)),
                   "n_above_1e-6": int((e > 1e-6).sum())}
    return {"n": int(got.size), "vs_default_oracle": q(e_ref), "vs_converged_oracle": q(e_tight),
            "default_vs_converged": q(relerr(ref, tight))}


def report(name, stats):
    """Error statistics where a reader of the records finds them."""
    warnings.warn(f"parity[{name}] " + json.dumps(stats), UserWarning)
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        path = os.path.join(out, "parity_report.json")
        data = json.load(open(path)) if os.path.exists(path) else {}
        data[name] = stats
        json.dump(data, open(path, "w"), indent=1)
