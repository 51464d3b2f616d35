"""TEST INFRASTRUCTURE ONLY -- never imported by the product (nmma_b200/).

Flat LambdaCDM restatement of ``astropy.cosmology.Planck18`` as NMMA uses it
(``nmma/core/constants.py:43`` default cosmology; consumers
``nmma/core/conversion.py:36-55``).  astropy is not installed here, so its
published algorithm is restated: photons + 3 neutrino species (one massive,
0.06 eV) with the Komatsu et al. (2011) fitting form for the massive-neutrino
energy density that astropy implements in ``FLRW.nu_relative_density``.

Parity note: the reference inverts d_L(z) with ``astropy.cosmology.z_at_value``
(a bounded Brent minimiser with ztol=1e-8); this file uses a root finder good to
~1e-14, so redshifts agree with the reference to ~1e-8 relative, not bitwise.
The GPU path never computes cosmology: it is handed the (dist_grid, z_grid)
table, so CUDA-vs-oracle parity is unaffected.
"""
from __future__ import annotations

import numpy as np
from scipy.integrate import quad
from scipy.optimize import brentq

# CODATA 2018 (astropy.constants default)
_C_KM_S = 299792.458
_G = 6.6743e-11            # m^3 kg^-1 s^-2
_SIGMA_SB = 5.6703744191844314e-08  # W m^-2 K^-4
_C = 299792458.0
_MPC_M = 3.085677581491367e22
_KB_EV = 8.617333262145179e-05


class FlatLambdaCDM:
    def __init__(self, H0=67.66, Om0=0.30966, Tcmb0=2.7255, Neff=3.046,
                 m_nu=(0.0, 0.0, 0.06)):
        self.H0 = H0
        self.Om0 = Om0
        self.Tcmb0 = Tcmb0
        self.Neff = Neff
        h0_si = H0 * 1000.0 / _MPC_M
        rho_crit = 3.0 * h0_si ** 2 / (8.0 * np.pi * _G)          # kg m^-3
        a_rad = 4.0 * _SIGMA_SB / _C                               # J m^-3 K^-4
        self.Ogamma0 = a_rad * Tcmb0 ** 4 / (_C ** 2) / rho_crit
        m_nu = np.asarray(m_nu, dtype=float)
        self._massive = m_nu[m_nu > 0]
        self._nmassless = int(np.sum(m_nu == 0))
        self._neff_per_nu = Neff / len(m_nu)
        tnu0 = 0.7137658555036082 * Tcmb0
        self._nu_y = self._massive / (_KB_EV * tnu0)
        self.Onu0 = self.Ogamma0 * self.nu_relative_density(0.0)
        self.Ode0 = 1.0 - self.Om0 - self.Ogamma0 - self.Onu0
        self.hubble_distance = _C_KM_S / H0                        # Mpc

    def nu_relative_density(self, z):
        prefac = 0.22710731766
        if self._massive.size == 0:
            return prefac * self.Neff
        p, invp, k = 1.83, 0.54644808743, 0.3173
        curr = self._nu_y / (1.0 + z)
        rel = (1.0 + (k * curr) ** p) ** invp
        return prefac * self._neff_per_nu * (rel.sum() + self._nmassless)

    def inv_efunc(self, z):
        zp1 = 1.0 + z
        orel = self.Ogamma0 * (1.0 + self.nu_relative_density(z))
        return 1.0 / np.sqrt(zp1 ** 3 * (orel * zp1 + self.Om0) + self.Ode0)

    def comoving_distance(self, z):
        return self.hubble_distance * quad(self.inv_efunc, 0.0, z, epsabs=0, epsrel=1e-13)[0]

    def luminosity_distance(self, z):
        z = np.asarray(z, dtype=float)
        if z.ndim == 0:
            return (1.0 + float(z)) * self.comoving_distance(float(z))
        return np.array([(1.0 + zi) * self.comoving_distance(zi) for zi in z])

    def z_at_luminosity_distance(self, d_mpc):
        """Restates ``z_at_value(cosmology.luminosity_distance, d)``."""
        return brentq(lambda z: self.luminosity_distance(z) - d_mpc, 1e-12, 1000.0,
                      xtol=1e-16, rtol=1e-15)


Planck18 = FlatLambdaCDM()


def get_cosmo_grids(distance_min, distance_max, cosmology=Planck18):
    """``nmma/core/conversion.py:49-55``: 50-point geometric z grid."""
    zmin = cosmology.z_at_luminosity_distance(distance_min)
    zmax = cosmology.z_at_luminosity_distance(distance_max)
    z_grid = np.geomspace(zmin, zmax, 50)
    dist_grid = cosmology.luminosity_distance(z_grid)
    return dist_grid, z_grid
