"""nmma_b200: B200-native (sm_100a) implementation of NMMA's inner kilonova likelihood loop.

Host layer mirrors the reference API (``nmma.em.model.SVDLightCurveModel``,
``nmma.em.em_likelihood.EMTransientLikelihood`` / legacy ``OpticalLightCurve``); compute goes
through the C ABI of ``include/nmma_b200.h`` into hand-written CUDA kernels.
"""
__version__ = "0.1.0"

from . import core, em, mlmodel  # noqa: F401,E402
