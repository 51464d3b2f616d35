"""Flat LambdaCDM with the Planck18 parameters NMMA uses by default.

The reference takes its cosmology from astropy (``nmma/core/constants.py:43``:
``cosmology.Planck18``) and only needs d_L(z) and its inverse on the hot path
(``nmma/core/conversion.py:36-55``).  astropy is not part of this stack; the
distance integral is done here with a fixed Gauss-Legendre rule, vectorised over z.
Radiation: photons (Tcmb0 = 2.7255 K) + Neff = 3.046 neutrinos, one of them massive
(0.06 eV) with the Komatsu et al. fitting function astropy uses.
"""
from __future__ import annotations

import numpy as np

_C_KM_S = 299792.458
_GL_X, _GL_W = np.polynomial.legendre.leggauss(96)


class FlatLambdaCDM:
    def __init__(self, H0=67.66, Om0=0.30966, Tcmb0=2.7255, Neff=3.046, m_nu=(0.0, 0.0, 0.06), name="Planck18"):
        self.name = name
        self.H0, self.Om0, self.Tcmb0, self.Neff = float(H0), float(Om0), float(Tcmb0), float(Neff)
        h = self.H0 / 100.0
        # Omega_gamma h^2 = 2.4728e-5 (T/2.7255)^4 follows from a_rad T^4 / rho_crit c^2 (CODATA 2018)
        sigma_sb, G, c, mpc = 5.6703744191844314e-08, 6.6743e-11, 299792458.0, 3.085677581491367e22
        rho_crit = 3.0 * (self.H0 * 1e3 / mpc) ** 2 / (8.0 * np.pi * G)
        self.Ogamma0 = 4.0 * sigma_sb / c ** 3 * self.Tcmb0 ** 4 / rho_crit
        m_nu = np.asarray(m_nu, float)
        self._n_massless = int((m_nu == 0).sum())
        self._nu_y = m_nu[m_nu > 0] / (8.617333262145179e-05 * 0.7137658555036082 * self.Tcmb0)
        self._n_nu = len(m_nu)
        self.Onu0 = self.Ogamma0 * self._nu_rel(np.zeros(1))[0]
        self.Ode0 = 1.0 - self.Om0 - self.Ogamma0 - self.Onu0
        self.h = h

    def _nu_rel(self, z):
        prefac = 0.22710731766
        if self._nu_y.size == 0:
            return np.full_like(z, prefac * self.Neff)
        p, invp, k = 1.83, 0.54644808743, 0.3173
        y = self._nu_y[None, :] / (1.0 + z[:, None])
        rel = ((1.0 + (k * y) ** p) ** invp).sum(axis=1) + self._n_massless
        return prefac * (self.Neff / self._n_nu) * rel

    def inv_efunc(self, z):
        z = np.asarray(z, float)
        zf = z.ravel()
        zp1 = 1.0 + zf
        orel = self.Ogamma0 * (1.0 + self._nu_rel(zf))
        return (1.0 / np.sqrt(zp1 ** 3 * (orel * zp1 + self.Om0) + self.Ode0)).reshape(z.shape)

    def luminosity_distance(self, z):
        """d_L in Mpc (array-friendly)."""
        z = np.asarray(z, float)
        zf = np.atleast_1d(z).ravel()
        # integral_0^z dz'/E(z') on [0, z] mapped from [-1, 1]
        nodes = 0.5 * zf[:, None] * (_GL_X[None, :] + 1.0)
        integ = (self.inv_efunc(nodes) * _GL_W[None, :]).sum(axis=1) * 0.5 * zf
        dl = (1.0 + zf) * (_C_KM_S / self.H0) * integ
        return dl.reshape(z.shape) if z.ndim else float(dl[0])

    def z_at_luminosity_distance(self, d_mpc):
        """Inverse of :meth:`luminosity_distance` (replaces ``astropy.cosmology.z_at_value``)."""
        d = np.atleast_1d(np.asarray(d_mpc, float))
        if np.any(d < 0):
            raise ValueError("luminosity distance must be non-negative")
        lo = np.zeros_like(d)
        hi = np.full_like(d, 1.0)
        while np.any(self.luminosity_distance(hi) < d):
            hi = np.where(self.luminosity_distance(hi) < d, hi * 4.0, hi)
            if np.any(hi > 1e6):
                raise ValueError("luminosity distance out of range")
        for _ in range(200):
            mid = 0.5 * (lo + hi)
            below = self.luminosity_distance(mid) < d
            lo = np.where(below, mid, lo)
            hi = np.where(below, hi, mid)
            if np.all(hi - lo <= 4e-16 * np.maximum(hi, 1e-300)):
                break
        z = 0.5 * (lo + hi)
        return z if np.ndim(d_mpc) else float(z[0])


Planck18 = FlatLambdaCDM()
_COSMOLOGY = Planck18


def get_cosmology():
    return _COSMOLOGY


def set_cosmology(cosmology=None):
    global _COSMOLOGY
    _COSMOLOGY = Planck18 if cosmology is None else cosmology
    return _COSMOLOGY
