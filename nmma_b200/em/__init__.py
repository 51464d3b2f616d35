from . import em_likelihood, io, likelihood, model, systematics, utils  # noqa: F401
from .em_likelihood import BatchPool, EMTransientLikelihood, MultiFilterTransient  # noqa: F401
from .likelihood import OpticalLightCurve  # noqa: F401
from .model import CombinedLightCurveModelContainer, SVDLightCurveModel, model_parameters_dict  # noqa: F401
from .systematics import FilterSystematicsHandler, SystematicsHandler  # noqa: F401
