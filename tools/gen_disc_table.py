"""Generate magprop_b200/csrc/disc_table.inc: piecewise polynomials for the
disc-mass kernel function

    S(u) = -(3/2) u^(-2/3) 1F1(1; 1/3; -u),        S'(u) = -S(u) + u^(-5/3)

which is the particular solution used to write the reference's (linear,
decoupled) disc-mass equation  dM/dt = Mdot_fb(t) - M/tvisc
(code/synthetic_datasets/funcs.py:126-129, magnetar/funcs.py:85-88) in closed
form:   M(t) = K S(u) + (M_i - K S(u0)) exp(-(u-u0)),   u = (t+tfb)/tvisc,
K = delta*M_i*eps^(2/3).   See DESIGN.md section 3.

Layout: for binade e in [EMIN, EMAX] and sub-interval j in [0, NSUB) the row
(e-EMIN)*NSUB + j holds the monomial coefficients c_0..c_DEG of S on
u = 2^e * (1 + (j + (s+1)/2)/NSUB),  s in [-1, 1), evaluated as sum c_k s^k.
Coefficients come from a 40-digit Chebyshev interpolant (mpmath) converted to
the monomial basis and rounded to double.  The script re-evaluates the rounded
table in double precision at random points and prints the worst error
relative to the interval's max |S| (must be < 2e-15).

    python tools/gen_disc_table.py
"""
import os
import sys

import mpmath as mp
import numpy as np

mp.mp.dps = 45
EMIN, EMAX = -10, 21
NSUB = 8
DEG = 10
ROW = 12  # padded row length (doubles)
# The explicit integrator's own, denser table of S alone: degree 5 on 64 sub-intervals per binade -- 48-byte rows
# (three 16-byte loads, five FMAs per lookup instead of six loads and ten FMAs); 5.9e-14 of the interval's max |S|,
# four orders below the step tolerance it feeds.  (Degree 6 on 32 sub-intervals, 64-byte rows, 3.4e-14: 1 % faster
# on an ensemble whose lanes all read the same row, 3-6 % slower on spread-out ones, where every lane's row is
# another L1 wavefront -- measured.)
FAST_NSUB = 64
FAST_DEG = 5
FAST_ROW = 6


def S(u):
    u = mp.mpf(u)
    return -mp.mpf(3) / 2 * u ** (-mp.mpf(2) / 3) * mp.hyp1f1(1, mp.mpf(1) / 3, -u)


def cheb_to_mono(c):
    """Chebyshev coefficients -> monomial coefficients (exact mp arithmetic)."""
    n = len(c)
    T = [[mp.mpf(0)] * n for _ in range(n)]
    T[0][0] = mp.mpf(1)
    if n > 1:
        T[1][1] = mp.mpf(1)
    for k in range(2, n):
        for i in range(n):
            T[k][i] = (2 * T[k - 1][i - 1] if i > 0 else 0) - T[k - 2][i]
    return [sum(c[k] * T[k][i] for k in range(n)) for i in range(n)]


def Q(u):
    """S(u)^(-1/7): late-time shortcut for Mdisc^(-1/7) (defined for u >= 1, where S > 0)."""
    return S(u) ** (-mp.mpf(1) / 7)


def fit(a, b, N, f=None):
    f = f or S
    nodes = [mp.cos(mp.pi * (2 * j + 1) / (2 * (N + 1))) for j in range(N + 1)]
    fx = [f((a + b) / 2 + (b - a) / 2 * x) for x in nodes]
    c = []
    for n in range(N + 1):
        s = sum(fx[j] * mp.cos(n * mp.pi * (2 * j + 1) / (2 * (N + 1))) for j in range(N + 1))
        c.append(s * 2 / (N + 1))
    c[0] /= 2
    return cheb_to_mono(c)


def main():
    rows = []
    worst = 0.0
    rng = np.random.RandomState(1)
    for e in range(EMIN, EMAX + 1):
        for j in range(NSUB):
            a = mp.mpf(2) ** e * (1 + mp.mpf(j) / NSUB)
            b = mp.mpf(2) ** e * (1 + mp.mpf(j + 1) / NSUB)
            mono = [float(v) for v in fit(a, b, DEG)]
            monoq = [float(v) for v in fit(a, b, DEG, Q)] if e >= 0 else [0.0] * (DEG + 1)
            rows.append(mono + [0.0] * (ROW - len(mono)) + monoq)
            scale = max(abs(S(a)), abs(S(b)))
            for s in rng.uniform(-1, 1, size=6):
                acc = accq = 0.0
                for ck in reversed(mono):
                    acc = acc * s + ck
                for ck in reversed(monoq):
                    accq = accq * s + ck
                u = (a + b) / 2 + (b - a) / 2 * mp.mpf(float(s))
                worst = max(worst, float(abs(acc - S(u)) / scale))
                if e >= 0:
                    worst = max(worst, float(abs(accq - Q(u)) / Q(u)))
        print(f"binade {e:3d} done, worst so far {worst:.2e}", file=sys.stderr)
    print(f"worst error / interval max|S| = {worst:.3e}")
    assert worst < 2e-15
    fast = []
    worst_fast = 0.0
    for e in range(EMIN, EMAX + 1):
        for j in range(FAST_NSUB):
            a = mp.mpf(2) ** e * (1 + mp.mpf(j) / FAST_NSUB)
            b = mp.mpf(2) ** e * (1 + mp.mpf(j + 1) / FAST_NSUB)
            mono = [float(v) for v in fit(a, b, FAST_DEG)]
            fast.append(mono + [0.0] * (FAST_ROW - len(mono)))
            scale = max(abs(S(a)), abs(S(b)))
            for s in list(rng.uniform(-1, 1, size=3)) + [-1.0, 1.0]:
                acc = 0.0
                for ck in reversed(mono):
                    acc = acc * s + ck
                u = (a + b) / 2 + (b - a) / 2 * mp.mpf(float(s))
                worst_fast = max(worst_fast, float(abs(acc - S(u)) / scale))
        print(f"fast table: binade {e:3d} done, worst so far {worst_fast:.2e}", file=sys.stderr)
    print(f"fast table: worst error / interval max|S| = {worst_fast:.3e}")
    assert worst_fast < 1e-13
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "magprop_b200", "csrc", "disc_table.inc")
    with open(out, "w") as fh:
        fh.write("// GENERATED by tools/gen_disc_table.py -- do not edit.\n")
        fh.write(f"// S(u) = -(3/2) u^(-2/3) 1F1(1;1/3;-u); worst error / interval max|S| = {worst:.3e}\n")
        fh.write(f"#define MP_DISC_EMIN ({EMIN})\n#define MP_DISC_EMAX ({EMAX})\n")
        fh.write(f"#define MP_DISC_NSUB_LOG2 ({int(np.log2(NSUB))})\n#define MP_DISC_DEG ({DEG})\n#define MP_DISC_ROW ({ROW})\n")
        fh.write("// row = [S coefficients c0..c10, pad | Q = S^(-1/7) coefficients c0..c10 (binades >= 0), pad]\n")
        fh.write(f"MP_TABLE_QUALIFIER double mp_disc_table[{len(rows)}][{2 * ROW}] = {{\n")
        for r in rows:
            r = r + [0.0] * (2 * ROW - len(r))
            fh.write("  {" + ", ".join(float(v).hex() for v in r) + "},\n")
        fh.write("};\n")
        fh.write(f"// S alone, degree {FAST_DEG} on {FAST_NSUB} sub-intervals per binade (the explicit integrator's stages); worst error / interval max|S| = {worst_fast:.3e}\n")
        fh.write(f"#define MP_DISC_FAST_NSUB_LOG2 ({int(np.log2(FAST_NSUB))})\n#define MP_DISC_FAST_DEG ({FAST_DEG})\n#define MP_DISC_FAST_ROW ({FAST_ROW})\n")
        fh.write(f"MP_TABLE_QUALIFIER double mp_disc_fast[{len(fast)}][{FAST_ROW}] = {{\n")
        for r in fast:
            fh.write("  {" + ", ".join(float(v).hex() for v in r) + "},\n")
        fh.write("};\n")
    print("wrote", os.path.normpath(out), len(rows), "+", len(fast), "rows")


if __name__ == "__main__":
    main()
