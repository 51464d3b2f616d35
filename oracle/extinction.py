"""TEST INFRASTRUCTURE ONLY -- never imported by the product (nmma_b200/).

CPU restatement of the reference's extinction step, ``LightCurveModelContainer.get_extinction_mags`` /
``apply_extinction_correction`` (``nmma/em/model.py:323-350``) with ``extinctionFactorP92SMC``
(``nmma/em/utils.py:373-433``).

PARITY UNPINNED against the third-party ``dust_extinction`` package (absent offline, floor only in
``pyproject.toml``): its ``shapes.P92`` model is restated here from the published formula, Pei (1992, ApJ 395, 130)
eq. 20, xi(lambda) = sum_i a_i / [(lambda/lambda_i)^n_i + (lambda_i/lambda)^n_i + b_i], with the six SMC terms of
his Table 4 -- the very numbers the reference passes inline at ``nmma/em/utils.py:398-423`` -- and dust_extinction's
conventions as the reference uses them: amplitudes referred to A(V) through ``P92.AbAv = 1/3.08 + 1`` (:395),
validity range ``P92.x_range = [1e-3, 1e3]`` 1/micron (:379-380), R_V = 2.93 (:428).  What CAN be pinned is pinned in
``tests/test_extinction.py``: xi = A_lambda / A_B is 1 at the B band (Pei's normalisation), A(0.55 um)/A(V) = 1, the
curve is the SMC one (no 2175 A bump, steep far-UV rise), and the effective wavelengths (``tools/make_wave_eff.py``)
reproduce the six PS1 values hard-coded in the reference (``nmma/em/utils.py:712-714``).
"""
from __future__ import annotations

import numpy as np

C_CGS = 29979245800.0          # astropy.constants.c.cgs.value
P92_ABAV = 1.0 / 3.08 + 1.0    # dust_extinction.shapes.P92.AbAv
P92_X_RANGE = (1.0 / 1e3, 1.0 / 1e-3)
# (amplitude / AbAv, lambda_i [micron], b_i, n_i): BKG, FUV, NUV (2175 A), SIL1 (9.7 um), SIL2 (18 um), FIR -- Pei 1992 Table 4, SMC
P92_SMC_TERMS = ((185.0, 0.042, 90.0, 2.0), (27.0, 0.08, 5.5, 4.0), (0.005, 0.22, -1.95, 2.0),
                 (0.010, 9.7, -1.95, 2.0), (0.012, 18.0, -1.80, 2.0), (0.030, 25.0, 0.0, 2.0))


def p92_smc_axav(lam_micron):
    """A(lambda)/A(V) of the Pei (1992) SMC curve (dust_extinction ``P92.evaluate`` with the reference's parameters)."""
    lam = np.asarray(lam_micron, float)
    x = 1.0 / lam                       # dust_extinction converts to wavenumbers first ...
    lam = 1.0 / x                       # ... and back
    axav = np.zeros_like(lam)
    for amp, cen, b, n in P92_SMC_TERMS:
        l_norm = lam / cen
        axav = axav + (amp * P92_ABAV) / (np.power(l_norm, n) + np.power(l_norm, -1 * n) + b)
    return axav


def extinction_factor_p92_smc(nu, Ebv, z, cutoff_hi=2e16):
    """``nmma/em/utils.py:373-433``: flux factor 10^(-0.4 A) per observer-frame frequency, dust in the host frame."""
    nu = np.asarray(nu, float)
    ext_range_nu_lo = P92_X_RANGE[0] * 1e4 * C_CGS
    ext_range_nu_hi = min(cutoff_hi, P92_X_RANGE[1] * 1e4 * C_CGS)
    nu_host = nu * (1 + z)
    opt = (nu_host >= ext_range_nu_lo) & (nu_host <= ext_range_nu_hi)
    lam_host_micron = (C_CGS / nu_host[opt]) * 1e4
    Ax_o_Av = p92_smc_axav(lam_host_micron)
    Av = 2.93 * Ebv
    ext = np.ones(nu.shape)
    ext[opt] = np.power(10.0, -0.4 * Ax_o_Av * Av)
    return ext


def get_extinction_mags(nu_0s, Ebv, redshift, law="P92_SMC_host", coef=None):
    """``nmma/em/model.py:323-342``.  ``coef`` (A_f / E(B-V) per filter) stands in for the G23 curve of the linear law."""
    nu_0s = np.asarray(nu_0s, float)
    ext_mag = np.zeros_like(nu_0s)
    if Ebv != 0.0:
        if law == "P92_SMC_host":
            ext = extinction_factor_p92_smc(nu_0s, Ebv, redshift)
        elif law == "G23_MW":
            ext = np.power(10.0, -0.4 * np.asarray(coef, float) * Ebv)
        else:
            raise ValueError(f"Unknown extinction_law {law!r}use 'P92_SMC_host' or 'G23_MW'.")
        with np.errstate(divide="ignore"):
            ext_mag = -2.5 * np.log10(ext)
    return ext_mag
