"""The C-ABI library: builds, loads, exports every symbol the header declares,
struct layouts agree, and -- with no GPU -- compute calls fail loudly instead
of falling back to a CPU path.  CPU only, no compute."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, has_gpu


def header_functions():
    src = open(os.path.join(ROOT, "include", "magprop_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built):
    from magprop_b200 import _capi as A
    lib = A.load()
    names = header_functions()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(A.EXPORTS) == names
    assert lib.mp_abi_version() == A.MP_ABI_VERSION


def test_struct_layouts(built):
    from magprop_b200 import _capi as A
    # 15 doubles, 2 int32, double, 2 int32
    assert C.sizeof(A.ModelSpec) == 15 * 8 + 8 + 8 + 8
    assert C.sizeof(A.PriorSpec) == 8 + 2 * 9 * 8
    s = A.script_model_spec()
    assert (s.inertia_factor, s.mdot_factor, s.rhs_n, s.lum_n, s.breakup_lum, s.unlog_mask) == (0.35, 3.0, 10.0, 10.0, 0.27, 0b111100)
    p = A.packaged_model_spec(n=7.0)
    assert (p.inertia_factor, p.mdot_factor, p.rhs_n, p.lum_n, p.breakup_lum, p.lprop_binding_term) == (0.8, 1.0, 1.0, 7.0, 0.0, 0)


def test_ensemble_struct_matches_the_header(built, tmp_path):
    """mp_ensemble as ctypes lays it out == as a C compiler lays out the header's definition."""
    import subprocess
    from magprop_b200 import _capi as A
    fields = [f[0] for f in A.Ensemble._fields_]
    prog = ['#include <stdio.h>', '#include <stddef.h>', '#include "magprop_b200.h"', 'int main(void) {',
            '  printf("%zu %d\\n", sizeof(mp_ensemble), MP_MAX_PEERS);']
    prog += [f'  printf("%zu\\n", offsetof(mp_ensemble, {f}));' for f in fields]
    prog += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(prog))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    assert int(out[0]) == C.sizeof(A.Ensemble) and int(out[1]) == A.MP_MAX_PEERS
    assert [int(v) for v in out[2:]] == [getattr(A.Ensemble, f).offset for f in fields]


def test_no_cpu_fallback(built):
    """Without a CUDA device every compute entry point must raise."""
    if has_gpu():
        pytest.skip("GPU present")
    from magprop_b200 import _capi as A
    from magprop_b200.engine import Likelihood, time_grid, rhs_batch
    from magprop_b200.synthetic import funcs, mcmc_eqns
    from magprop_b200 import magnetar
    assert A.load().mp_device_count() == 0
    with pytest.raises(A.MagpropCudaError):
        Likelihood(A.script_model_spec(), time_grid(None))
    with pytest.raises(A.MagpropCudaError):
        funcs.model_lum([1, 5, 1e-3, 100, 0.1, 1])
    with pytest.raises(A.MagpropCudaError):
        magnetar.model_lc([1, 5, 1e-3, 100, 0.1, 1])
    with pytest.raises(A.MagpropCudaError):
        mcmc_eqns.lnprob([1, 5, -3, 2, -1, 0], [10.0], [1.0], [0.1], None)
    with pytest.raises(A.MagpropCudaError):
        rhs_batch(A.script_model_spec(), [[1e30, 1e3]], [1.0], [[1, 1e-3, 100, 1, 1]], [10, 0.1, 1, 0.9])


def test_product_never_imports_oracle():
    """The package must not import, call or link anything under oracle/ or tests/."""
    pkg = os.path.join(ROOT, "magprop_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "hostsim" not in text or f == "magprop_core.cuh", f   # only a comment there
                assert "scipy" not in text, f


def test_arg_validation_without_device(built):
    from magprop_b200 import _capi as A
    lib = A.load()
    h = C.c_void_p()
    grid = np.array([1.0, 0.5, 2.0])
    spec, prior = A.script_model_spec(), A.prior_spec()
    rc = lib.mp_create(C.byref(spec), C.byref(prior), A.ptr(grid), 3, None, None, None, 0, 0, C.byref(h))
    assert rc == A.MP_ERR_BAD_GRID
    grid = np.logspace(0, 6, 10001)
    x = np.array([0.5]); one = np.array([1.0])
    rc = lib.mp_create(C.byref(spec), C.byref(prior), A.ptr(grid), grid.size, A.ptr(x), A.ptr(one), A.ptr(one), 1, 0, C.byref(h))
    assert rc == A.MP_ERR_DATA_RANGE          # interp1d bounds_error (funcs.py:233-234)
    with pytest.raises(ValueError):
        A.check(rc)
    rc = lib.mp_create(None, C.byref(prior), A.ptr(grid), grid.size, None, None, None, 0, 0, C.byref(h))
    assert rc == A.MP_ERR_BAD_ARG
