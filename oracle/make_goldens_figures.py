"""TEST INFRASTRUCTURE (build container only).  Pins the oracle's ``figure_spec`` against the copy of the
model that the paper-figure scripts inline (row f3).  The scripts plot at import time (matplotlib is not
installed), so only their module-level constants and the ``init_conds`` / ``odes`` definitions are
executed, unmodified, out of the parsed source; the RHS is then evaluated at random states and stored
with the oracle's values in tests/golden/figure_rhs.npz (asserted equal here).

    python oracle/make_goldens_figures.py
"""
import ast
import os
import sys

import numpy as np

REF = "/root/reference/code"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import magprop_oracle as O     # noqa: E402


def load_defs(path):
    """Execute the constant assignments and function definitions of a figure script (nothing else)."""
    tree = ast.parse(open(path).read())
    keep = []
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("init_conds", "odes", "piroott", "bucciantini"):
            keep.append(node)
        elif isinstance(node, ast.Assign) and all(isinstance(t, ast.Name) for t in node.targets) and \
                node.targets[0].id in ("G", "c", "R", "Msol", "M", "I", "GM", "alpha", "cs7", "k", "n"):
            keep.append(node)
    ns = {"np": np}
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    return ns


rng = np.random.RandomState(4)
out = {}
for script in ("figure_1.py", "figure_4.py"):
    ns = load_defs(os.path.join(REF, script))
    rows = []
    for _ in range(200):
        B, P = rng.uniform(0.5, 10), rng.uniform(0.7, 10)
        MdiscI, RdiscI = 10 ** rng.uniform(-5, -2), 10 ** rng.uniform(1.7, 3.3)
        eps, delta = 10 ** rng.uniform(-1, 2), 10 ** rng.uniform(-1, 3)
        n = rng.choice([1.0, 10.0, 50.0])
        t = 10 ** rng.uniform(0, 6)
        y0 = ns["init_conds"](MdiscI, P)
        y = np.array([y0[0] * 10 ** rng.uniform(-6, 0.5), y0[1] * 10 ** rng.uniform(-1.5, 0.2)])
        ref = ns["odes"](y, t, B, MdiscI, RdiscI, eps, delta, n, 0.1, 1.0, 0.9)
        spec = O.figure_spec(n=n)
        mine = O._rhs_for(spec)(y, t, B, MdiscI, RdiscI, eps, delta)
        assert ref[0] == mine[0] and ref[1] == mine[1], (script, ref, mine)
        rows.append([y[0], y[1], t, B, MdiscI, RdiscI, eps, delta, n, ref[0], ref[1]])
    out[script.replace(".py", "")] = np.array(rows)
# figure_3.py: the two right-hand sides it compares (module-level n = 10, k = 0.9)
ns = load_defs(os.path.join(REF, "figure_3.py"))
rows = []
for _ in range(200):
    B, P = rng.uniform(0.5, 10), rng.uniform(0.7, 10)
    MdiscI, RdiscI = 10 ** rng.uniform(-5, -2), 10 ** rng.uniform(1.7, 3.3)
    eps, delta = 10 ** rng.uniform(-1, 2), 10 ** rng.uniform(-1, 3)
    t = 10 ** rng.uniform(0, 6)
    y0 = ns["init_conds"](MdiscI, P)
    y = np.array([y0[0] * 10 ** rng.uniform(-6, 0.5), y0[1] * 10 ** rng.uniform(-1.5, 0.2)])
    for name, bucc in (("piroott", False), ("bucciantini", True)):
        ref = ns[name](y, t, B, MdiscI, RdiscI, eps, delta)
        mine = O._rhs_for(O.figure_spec(n=10.0, bucciantini=bucc))(y, t, B, MdiscI, RdiscI, eps, delta)
        assert ref[0] == mine[0] and ref[1] == mine[1], (name, ref, mine)
        rows.append([y[0], y[1], t, B, MdiscI, RdiscI, eps, delta, 10.0, ref[0], ref[1], float(bucc)])
out["figure_3"] = np.array(rows)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "figure_rhs.npz"), **out)
print("figure_spec RHS is bit-identical to figure_1.py / figure_4.py odes and figure_3.py piroott / bucciantini on 800 states; wrote figure_rhs.npz")
