"""Import the reference's OWN source files in this container.

nmma cannot be imported here because its third-party stack (bilby, sncosmo, astropy, keras, h5py,
healpy, dust_extinction, matplotlib ...) is not installed and there is no network.  None of those
libraries does arithmetic on the likelihood path except:

  * Keras (the per-filter MLP forward pass)     -> float32 NumPy stand-in reading the same .h5 weights
  * astropy Planck18 d_L(z) / z_at_value        -> the flat-LCDM restatement of oracle/cosmology.py

Everything else (np.interp, np.dot, scipy.stats.truncnorm / norm, scikit-learn GP predict, and every
line of nmma/em/{model,em_likelihood,systematics,utils,lightcurve_generation}.py and
nmma/core/{base,conversion}.py) is the real thing, executed unmodified from /root/reference.
This module installs inert stand-ins for the missing imports and exposes ``load_reference()``.
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from unittest.mock import MagicMock

import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class AutoModule(types.ModuleType):
    """Module whose unknown attributes are inert mocks (plot / IO / GW helpers never called here)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = MagicMock(name=f"{self.__name__}.{name}")
        setattr(self, name, m)
        return m


def _mod(name, **attrs):
    m = AutoModule(name)
    m.__path__ = []
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], child, m)
    return m


class QF(float):
    """float that answers the astropy Quantity/Constant attribute chain (.value, .si, .cgs, .to())."""
    value = property(lambda s: float(s))
    si = property(lambda s: s)
    cgs = property(lambda s: s)

    def to(self, *_a, **_k):
        return self

    def _w(self, v):
        return QF(v)

    def __mul__(self, o): return self._w(float(self) * float(o))
    __rmul__ = __mul__
    def __truediv__(self, o): return self._w(float(self) / float(o))
    def __rtruediv__(self, o): return self._w(float(o) / float(self))
    def __pow__(self, o): return self._w(float(self) ** o)


class Q(np.ndarray):
    """ndarray with astropy's ``.value``; NumPy functions applied to it return Q again (np.interp(...).value)."""

    def __new__(cls, a):
        return np.asarray(a, dtype=float).view(cls)

    @property
    def value(self):
        a = np.asarray(self)
        return a.item() if a.ndim == 0 else a

    def __array_function__(self, func, types_, args, kwargs):
        conv = lambda x: np.asarray(x) if isinstance(x, Q) else x
        out = func(*[conv(a) for a in args], **{k: conv(v) for k, v in kwargs.items()})
        return Q(out) if isinstance(out, (np.ndarray, np.floating, float)) else out


class _Planck18:
    """astropy.cosmology.Planck18 stand-in backed by oracle/cosmology.py."""

    def __init__(self):
        from oracle.cosmology import Planck18
        self._c = Planck18

    def luminosity_distance(self, z):
        return Q(self._c.luminosity_distance(np.asarray(z, float)))

    def clone(self, **kw):
        raise ValueError("cloning is not supported by the stand-in")


def _z_at_value(func, fval, *a, **k):
    from oracle.cosmology import Planck18
    return Q(Planck18.z_at_luminosity_distance(float(fval)))


class KerasModelStandIn:
    """``keras.saving.load_model(file, compile=False)`` -> callable returning an object with ``.numpy()``."""

    def __init__(self, model_file):
        from nmma_b200.mlmodel import load_keras_mlp
        from oracle.nmma_oracle import KerasStandIn
        self._m = KerasStandIn(*load_keras_mlp(model_file))

    def __call__(self, x):
        return self._m(x)


def install():
    from nmma_b200.core import priors as P
    from nmma_b200.em.utils import SNCOSMO_BANDPASSES

    # ---- astropy -----------------------------------------------------------------------------
    consts = dict(c=QF(299792458.0), h=QF(6.62607015e-34), e=QF(1.602176634e-19), G=QF(6.6743e-11),
                  M_sun=QF(1.988409870698051e30), pc=QF(3.085677581491367e16), k_B=QF(1.380649e-23),
                  sigma_sb=QF(5.6703744191844314e-08), m_p=QF(1.67262192369e-27))
    _mod("astropy")
    _mod("astropy.constants", **consts)
    _mod("astropy.units", Mpc=1.0)
    planck = _Planck18()
    _mod("astropy.cosmology", Planck18=planck, z_at_value=_z_at_value)
    _mod("astropy.time")
    _mod("astropy.table")
    _mod("astropy.io")
    _mod("astropy.io.fits")
    _mod("astropy.coordinates")

    # ---- bilby ----------------------------------------------------------------------------------
    class Likelihood:
        def __init__(self, parameters=None):
            self.parameters = parameters if parameters is not None else {}

    _mod("bilby", run_sampler=MagicMock())
    _mod("bilby.core")
    _mod("bilby.core.likelihood", Likelihood=Likelihood)
    pm = _mod("bilby.core.prior", analytical=P, **{n: getattr(P, n) for n in dir(P) if isinstance(getattr(P, n), type)})
    for extra in ("ConditionalPriorDict", "MultivariateGaussianDist", "MultivariateGaussian"):
        setattr(pm, extra, type(extra, (), {}))
    _mod("bilby.core.result", FileMovedError=type("FileMovedError", (Exception,), {}))
    _mod("bilby.core.utils")
    _mod("bilby.core.sampler")
    _mod("bilby.gw")
    holder = {"c": planck}

    def set_cosmology(c=None):
        holder["c"] = c if c is not None else planck
        sys.modules["bilby.gw.cosmology"].DEFAULT_COSMOLOGY = holder["c"]

    _mod("bilby.gw.cosmology", set_cosmology=set_cosmology, get_cosmology=lambda *a: holder["c"], DEFAULT_COSMOLOGY=planck)
    _mod("bilby.gw.conversion")
    _mod("bilby.gw.prior")
    _mod("bilby.gw.likelihood")
    _mod("bilby_pipe")
    _mod("bilby_pipe.utils", nonestr=str, nonefloat=float, noneint=int)

    # ---- sncosmo: only names and effective wavelengths (extinction is off: Ebv = 0) -------------------
    class _Registry:
        def get_loaders_metadata(self):
            return [{"name": n} for n in SNCOSMO_BANDPASSES]

    class _Band:
        def __init__(self, name):
            self.name, self.wave_eff = name, 5000.0

    _mod("sncosmo", get_bandpass=lambda name, *a: _Band(name))
    _mod("sncosmo.bandpasses", _BANDPASSES=_Registry(), _BANDPASS_INTERPOLATORS=type("R", (), {"get_loaders_metadata": lambda s: []})())
    _mod("sncosmo.models", _SOURCES=MagicMock())

    # ---- keras: float32 NumPy stand-in for the forward pass ----------------------------------------
    _mod("keras")
    _mod("keras.saving", load_model=lambda f, compile=False: KerasModelStandIn(f))

    # ---- inert: plotting, IO, samplers, dust, healpix ----------------------------------------------
    for name in ("h5py", "healpy", "dust_extinction", "dust_extinction.shapes", "dust_extinction.parameter_averages",
                 "matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.gridspec", "matplotlib.ticker",
                 "matplotlib.lines", "matplotlib.patches", "matplotlib.cm",
                 "mpl_toolkits", "mpl_toolkits.axes_grid1", "configargparse", "afterglowpy", "ligo", "ligo.skymap",
                 "ligo.skymap.io", "ligo.skymap.bayestar", "ligo.skymap.distance", "corner", "seaborn", "mpi4py",
                 "schwimmbad", "dynesty", "lal", "lalsimulation", "arviz", "m4opt", "numba"):
        _mod(name)


def load_reference():
    """Returns the reference modules (model, em_likelihood, systematics, utils, lightcurve_generation)."""
    install()
    for pkg, path in (("nmma", f"{REF}/nmma"), ("nmma.core", f"{REF}/nmma/core"), ("nmma.em", f"{REF}/nmma/em")):
        m = types.ModuleType(pkg)
        m.__path__ = [path]          # real source directory, package __init__ not executed
        sys.modules[pkg] = m
    names = ["nmma.core.constants", "nmma.core.conversion", "nmma.core.base", "nmma.em.utils",
             "nmma.em.lightcurve_generation", "nmma.em.systematics", "nmma.em.model", "nmma.em.em_likelihood"]
    mods = {n.rsplit(".", 1)[1]: importlib.import_module(n) for n in names}
    for n in names:
        assert sys.modules[n].__file__.startswith(REF), n
    return mods
