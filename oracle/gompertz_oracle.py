"""TEST INFRASTRUCTURE.  CPU restatement of the comparison model of the reference's ``code/figure_5.py:222-363``
("Ben's model" in the script: an accreting-mass magnetar with an exponentially draining disc, advanced by explicit
Euler steps of 1 s in a Python loop), operation for operation.  Pinned by ``oracle/make_goldens_gompertz.py``, which
executes the reference's own loop (the lines of figure_5.py, unmodified, read from /root/reference at generation
time) and stores inputs + outputs in ``tests/golden/gompertz.npz``.  Only tests may import this module."""
import numpy as np

G, c, R, Msol = 6.674e-8, 3.0e10, 1.0e6, 1.99e33            # figure_5.py:8-13


def curves(pars, n_steps, alpha=0.1, cs7=1.0, k=0.9, omass=1.4, dipeff=1.0, propeff=1.0):
    """(t, Ltot, Lprop, Ldip) with the luminosities in erg/s (the script divides by 1e50 when plotting, :357-359)."""
    B, P, MdiscI, RdiscI, epsilon, delta = (float(v) for v in pars)
    n = int(n_steps)
    spin = P * 1.0e-3                                           # :224
    Rdisc = RdiscI * 1.0e5
    visc = alpha * cs7 * 1.0e7 * Rdisc
    mu = 1.0e15 * B * (R ** 3.0)
    omega = (2.0 * np.pi) / spin                                # :229
    t = np.empty(n); Lprop = np.empty(n); Ldip = np.empty(n)
    Mdisc = MdiscI * Msol
    M_bg = omass * Msol
    Mdot0 = (3.0 * Mdisc * visc) / (Rdisc ** 2.0)               # :258
    Mdot, Msum, tt = Mdot0, 0.0, 1.0
    omegadot = 0.0
    with np.errstate(all="ignore"):
        for i in range(n):
            if i > 0:                                           # :302-309
                tt = tt + 1.0
                omega = omega + omegadot
                M_bg = M_bg + Msum
                Mdot = Mdot0 * np.exp((-3.0 * visc * tt) / (Rdisc ** 2.0))
            Rm = (mu ** (4.0 / 7.0)) * ((G * M_bg) ** (-1.0 / 7.0)) * (Mdot ** (-2.0 / 7.0))
            Rc = ((G * M_bg) / (omega ** 2.0)) ** (1.0 / 3.0)
            light = c / omega
            if Rm >= (k * light):
                Rm = k * light
            Ndip = (-2.0 / 3.0) * (((mu ** 2.0) * (omega ** 3.0)) / (c ** 3.0)) * ((light / Rm) ** 3.0)
            w = (Rm / Rc) ** (3.0 / 2.0)
            nn = 1.0 - w
            inertia = 0.35 * M_bg * (R ** 2.0)
            bigT = 0.5 * inertia * (omega ** 2.0)
            modW = 0.6 * M_bg * (c ** 2.0) * (((G * M_bg) / (R * (c ** 2.0))) / (1.0 - 0.5 * ((G * M_bg) / (R * (c ** 2.0)))))
            beta = bigT / modW
            if beta > 0.27:
                Nacc = 0.0
            elif Rm >= R:
                Nacc = nn * ((G * M_bg * Rm) ** 0.5) * Mdot
                if not np.isfinite(Nacc):
                    Nacc = 0.0
            else:
                if i == 0:                                      # :283-285 divides, the loop body (:335-337) multiplies
                    Nacc = (1.0 - (omega / (((G * M_bg) / (R ** 3.0)) ** 0.5))) / ((G * M_bg * R) ** 0.5) * Mdot
                else:
                    Nacc = (1.0 - (omega / (((G * M_bg) / (R ** 3.0)) ** 0.5))) * ((G * M_bg * R) ** 0.5) * Mdot
                if not np.isfinite(Nacc):
                    Nacc = 0.0
            if Rc >= Rm:                                        # :293-296, :341-344
                Msum = Mdot
            elif i == 0:
                Msum = 0.0
            omegadot = (Ndip + Nacc) / inertia
            t[i] = tt
            Lprop[i] = (-1.0 * Nacc * omega) - ((G * M_bg * Mdot) / Rm)
            Ldip[i] = ((mu ** 2.0) * (omega ** 4.0)) / (6.0 * (c ** 3.0))
    Lp = np.where(np.isfinite(Lprop), Lprop, 0.0)               # :354-357
    Lp = np.where(Lp <= 0.0, 0.0, Lprop)
    Ld = np.where(np.isfinite(Ldip), Ldip, 0.0)
    Ld = np.where(Ld <= 0.0, 0.0, Ldip)
    return t, (propeff * Lp) + (dipeff * Ld), Lp, Ld
