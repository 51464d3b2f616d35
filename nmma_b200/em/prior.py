"""Prior construction for an EM analysis: mirror of ``nmma/em/prior.py:172-244``.

``create_prior_from_args(args, systematics_handler)`` = prior file -> (Hubble prior) -> ``Ebv`` prior ->
(conditional inclination prior) -> (inclination prior from a GW sky map) -> ``em_syserr*`` priors from the
systematics YAML.  The bracketed steps belong to the GRB-afterglow / joint-GW configurations (``bilby``
conditional priors, ``ligo.skymap``, ``healpy``): outside the kilonova hot path, they raise
``NotImplementedError`` when requested instead of being ignored.
"""
from __future__ import annotations

import numpy as np

from ..core.priors import DeltaFunction, Interped, PriorDict


def extinction_prior(priors, args):
    """``nmma/em/prior.py:172-216``: ``Ebv`` = dust-map value (needs ``dustmaps``: not available offline), the
    triangular ``Interped([0, Ebv_max], [2 / Ebv_max, 0])`` density with ``--use-Ebv``, else ``DeltaFunction(0)``."""
    name, latex_label = "Ebv", "$E(B-V)$"
    if getattr(args, "fetch_Ebv_from_dustmap", False):
        raise NotImplementedError("--fetch-Ebv-from-dustmap needs the dustmaps package and the SFD maps (no network here); "
                                  "put `Ebv = <value>` into the prior file instead")
    if "Ebv" not in priors:
        ebv_max = float(getattr(args, "Ebv_max", 0.5724))
        if ebv_max > 0.0 and getattr(args, "use_Ebv", False):
            ebv_c = 1.0 / (0.5 * ebv_max)
            priors["Ebv"] = Interped([0, ebv_max], [ebv_c, 0], 0, ebv_max, name, latex_label)
        else:
            priors["Ebv"] = DeltaFunction(0.0, name, latex_label)
    return priors


def adjust_hubble_prior(priors, args):
    """``nmma/core/base.py:233-255``: a tabulated ``Hubble_constant`` prior from ``--Hubble-weight`` (columns
    ``Hubble prior_weight``).  Sampling the Hubble constant changes the dL <-> z map per point
    (``cosmology_to_distance``), which the batched path does not stage."""
    if getattr(args, "Hubble_weight", None) or "Hubble_constant" in priors:
        raise NotImplementedError("sampling over the Hubble constant (per-point cosmology) is outside the nmma_b200 hot path")
    return priors


def create_prior_from_args(args, systematics_handler=None):
    """``nmma/em/prior.py:221-244``."""
    src = getattr(args, "prior_file", None) or args.prior
    priors = src if isinstance(src, PriorDict) else PriorDict(filename=src)
    priors = adjust_hubble_prior(priors, args)
    priors = extinction_prior(priors, args)
    if getattr(args, "conditional_gaussian_prior_thetaObs", False):
        raise NotImplementedError("--conditional-gaussian-prior-thetaObs (GRB afterglow jets) is outside the kilonova hot path")
    if getattr(args, "fits_file", None):
        raise NotImplementedError("--fits-file (inclination prior from a GW sky map) needs ligo.skymap / healpy")
    if systematics_handler is not None:
        priors = systematics_handler.setup_systematics_priors(priors)
    return priors


__all__ = ["create_prior_from_args", "extinction_prior", "adjust_hubble_prior", "np"]
