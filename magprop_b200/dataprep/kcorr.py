"""``code/kcorr.py`` as functions: the Bloom, Frail & Sari (2001) k-correction of an XRT light curve
to a 1-10000 keV rest-frame luminosity in units of 1e50 erg/s, the luminosity distance it needs, and
the geometric-mean error.

The reference takes the luminosity distance from ``astropy.cosmology.WMAP9`` (``kcorr.py:9,76-77``).
astropy is not in this image, so the same cosmology is restated here: flat LambdaCDM with
H0 = 69.32 km/s/Mpc, Om0 = 0.2865, Tcmb0 = 2.725 K, Neff = 3.04 and massless neutrinos (Hinshaw et al.
2013, table 4, "WMAP9 + eCMB + BAO + H0" -- the parameters astropy ships as WMAP9), photons and
neutrinos included in E(z) as astropy does.  PARITY UNPINNED against astropy itself (absent); pinned
against the comoving distances astropy's documentation quotes for WMAP9 (tests/test_dataprep.py).
"""
import numpy as np


class WMAP9:
    H0 = 69.32            # km / s / Mpc
    Om0 = 0.2865
    Tcmb0 = 2.725         # K
    Neff = 3.04
    # physical constants (CODATA 2018, SI) for the radiation density
    _c = 299792458.0
    _G = 6.67430e-11
    _sigma_sb = 5.670374419e-8
    _Mpc_m = 3.085677581491367e22

    @classmethod
    def radiation(cls):
        """(Ogamma0, Onu0): photon and massless-neutrino density parameters today."""
        H0_s = cls.H0 * 1.0e3 / cls._Mpc_m
        rho_crit = 3.0 * H0_s ** 2 / (8.0 * np.pi * cls._G)                       # kg / m^3
        Ogamma0 = 4.0 * cls._sigma_sb / cls._c ** 3 * cls.Tcmb0 ** 4 / rho_crit
        Onu0 = 0.22710731766 * cls.Neff * Ogamma0                                   # 7/8 (4/11)^(4/3) Neff
        return Ogamma0, Onu0

    @classmethod
    def inv_efunc(cls, z):
        Og, On = cls.radiation()
        Ode0 = 1.0 - cls.Om0 - Og - On                                              # flat
        zp1 = 1.0 + np.asarray(z, dtype=np.float64)
        return 1.0 / np.sqrt(zp1 ** 3 * (cls.Om0 + (Og + On) * zp1) + Ode0)

    @classmethod
    def comoving_distance_mpc(cls, z):
        """(c/H0) * integral_0^z dz'/E(z')  [Mpc], 96-point Gauss-Legendre (the integrand is smooth)."""
        z = np.atleast_1d(np.asarray(z, dtype=np.float64))
        xg, wg = np.polynomial.legendre.leggauss(96)
        zz = 0.5 * z[:, None] * (xg[None, :] + 1.0)
        integral = 0.5 * z * np.sum(wg[None, :] * cls.inv_efunc(zz), axis=1)
        return (cls._c / 1.0e3 / cls.H0) * integral

    @classmethod
    def luminosity_distance_mpc(cls, z):
        z = np.atleast_1d(np.asarray(z, dtype=np.float64))
        return (1.0 + z) * cls.comoving_distance_mpc(z)


def luminosity_distance_cm(z):
    """kcorr.py:76-77: ``cosmo.luminosity_distance(z).value * 3.08568e24`` (the reference's Mpc -> cm factor)."""
    d = WMAP9.luminosity_distance_mpc(z) * 3.08568e24
    return d if np.ndim(z) else float(d[0])


def k_correction(df, gamma, sigma, z, dl_cm):
    """kcorr.py:12-55.  ``df``: mapping with ``t, tpos, tneg, flux, fluxpos, fluxneg`` arrays.  Returns a dict
    with ``t, tpos, tneg`` (rest frame) and ``Lum50, Lum50pos, Lum50neg`` (1e50 erg/s)."""
    e1, e2 = 0.3, 10.0          # XRT band, keV
    eb, et = 1.0, 10000.0       # bolometric band, keV
    a = ((et / (1.0 + z)) ** (2.0 - gamma)) / (2.0 - gamma)
    b = ((eb / (1.0 + z)) ** (2.0 - gamma)) / (2.0 - gamma)
    c = (e2 ** (2.0 - gamma)) / (2.0 - gamma)
    d = (e1 ** (2.0 - gamma)) / (2.0 - gamma)
    k = (a - b) / (c - d)
    factor = 4.0 * np.pi * (dl_cm ** 2.0) * k
    col = lambda name: np.asarray(df[name], dtype=np.float64)
    return {
        "t": col("t") / (1.0 + z), "tpos": col("tpos") / (1.0 + z), "tneg": col("tneg") / (1.0 + z),
        "Lum50": sigma * factor * col("flux") / 1.0e50,
        "Lum50pos": sigma * factor * col("fluxpos") / 1.0e50,
        "Lum50neg": sigma * factor * col("fluxneg") / 1.0e50,
    }


def k_correct_grb(cols, gamma, sigma, z):
    """kcorr.py:101-108 for one burst: k-correction plus ``Lum50err`` = geometric mean of the asymmetric errors
    (the reference's ``gmean`` of the two rows, kcorr.py:8,107-108, = exp(mean(log)))."""
    k = k_correction(cols, gamma, sigma, z, luminosity_distance_cm(z))
    with np.errstate(divide="ignore"):
        k["Lum50err"] = np.exp(np.mean(np.log([k["Lum50pos"], np.abs(k["Lum50neg"])]), axis=0))
    return k
