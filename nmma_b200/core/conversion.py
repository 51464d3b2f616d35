"""Host-side parameter conversions of the hot path (``nmma/core/conversion.py:19-126``)."""
from __future__ import annotations

import numpy as np

from .cosmology import get_cosmology


def distance_modulus_nmma(d_lum=1e-5):
    """``nmma/core/conversion.py:30-34``: mag_app - mag_abs for d_lum in Mpc."""
    return 5.0 * (5 + np.log10(d_lum))


def get_cosmo_grids(distance_min, distance_max, cosmology=None):
    """``nmma/core/conversion.py:49-55``: 50-point geometric redshift grid between the prior bounds."""
    cosmology = cosmology or get_cosmology()
    zmin = cosmology.z_at_luminosity_distance(distance_min) if distance_min > 0 else 0.0
    zmax = cosmology.z_at_luminosity_distance(distance_max)
    if not (zmin > 0):
        # the reference calls np.geomspace(0, ...) here (priors with luminosity_distance minimum 0.0)
        raise ValueError("Geometric sequence cannot include zero")
    z_grid = np.geomspace(zmin, zmax, 50)
    dist_grid = cosmology.luminosity_distance(z_grid)
    return dist_grid, z_grid


def luminosity_distance_to_redshift(distance, cosmology=None):
    """``nmma/core/conversion.py:36-47``."""
    cosmology = cosmology or get_cosmology()
    if hasattr(distance, "__len__") and len(distance) > 50:
        distance = np.asarray(distance, float)
        dist_grid, z_grid = get_cosmo_grids(distance.min(), distance.max(), cosmology)
        return np.interp(distance, dist_grid, z_grid)
    return cosmology.z_at_luminosity_distance(distance)


def get_redshift(parameters):
    """``nmma/core/conversion.py:57-64``."""
    if "redshift" in parameters:
        return parameters["redshift"]
    if "luminosity_distance" in parameters:
        return luminosity_distance_to_redshift(parameters["luminosity_distance"])
    return np.zeros_like(next(iter(parameters.values())))


def observation_angle_conversion(parameters):
    """``nmma/core/conversion.py:119-126``: inclination_EM [rad] <-> KNtheta [deg]."""
    theta_jn = parameters.get("theta_jn", np.arccos(parameters.get("cos_theta_jn", 1.0)))
    theta_jn = np.minimum(theta_jn, np.pi - theta_jn)
    if "KNtheta" not in parameters:
        parameters["KNtheta"] = parameters.get("inclination_EM", theta_jn) * 180.0 / np.pi
    if "inclination_EM" not in parameters:
        parameters["inclination_EM"] = parameters["KNtheta"] / 180.0 * np.pi
    return parameters
