from . import constants, cosmology, conversion, priors, base  # noqa: F401
