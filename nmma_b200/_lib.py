"""ctypes binding of the C ABI in ``include/nmma_b200.h``.

The shared library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a)
into ``nmma_b200/lib/libnmma_b200.so``.  There is no CPU fallback: if the library
is missing or no B200 is visible, the product path raises.
"""
from __future__ import annotations

import ctypes as C
import os

LIB_PATH = os.environ.get("NMMA_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libnmma_b200.so")

OK, ERR_ARG, ERR_CUDA, ERR_STATE, ERR_UNSUPPORTED = 0, 1, 2, 3, 4
XF_NONE, XF_RAD2DEG, XF_LOG10, XF_POW10, XF_THETAJN_DEG, XF_COSTHETAJN_DEG = range(6)
Z_ZERO, Z_PARAM, Z_TABLE = 0, 1, 2
SYS_BUDGET, SYS_PARAM, SYS_INTERP = 0, 1, 2
(PR_UNIFORM, PR_DELTA, PR_SINE, PR_COSINE, PR_GAUSSIAN, PR_TRUNC_GAUSS, PR_POWERLAW, PR_TRIANGULAR,
 PR_INTERPED) = range(9)
EXT_NONE, EXT_P92_SMC_HOST, EXT_LINEAR = 0, 1, 2
MAX_P = 32
MAX_CONSTRAINTS = 8


class ParamSrc(C.Structure):
    _fields_ = [("col", C.c_int32), ("transform", C.c_int32), ("value", C.c_double)]

    @classmethod
    def column(cls, col, transform=XF_NONE):
        return cls(int(col), int(transform), 0.0)

    @classmethod
    def const(cls, value, transform=XF_NONE):
        return cls(-1, int(transform), float(value))


class NmmaB200Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"nmma_b200 error {code}: {message}")
        self.code = code


_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int32)
_sp = C.POINTER(ParamSrc)
_h = C.c_void_p

# name -> (restype, argtypes); every symbol include/nmma_b200.h declares
SIGNATURES = {
    "nmma_b200_create": (C.c_int, [C.c_int, C.POINTER(_h)]),
    "nmma_b200_destroy": (C.c_int, [_h]),
    "nmma_b200_last_error": (C.c_char_p, [_h]),
    "nmma_b200_version": (C.c_int, []),
    "nmma_b200_set_svd": (C.c_int, [_h, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp]),
    "nmma_b200_set_mlp": (C.c_int, [_h, C.c_int, C.c_int, _fp, _fp, _fp, _fp]),
    "nmma_b200_set_gp": (C.c_int, [_h, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _dp]),
    "nmma_b200_set_sample_grid": (C.c_int, [_h, C.c_int, _dp]),
    "nmma_b200_set_param_layout": (C.c_int, [_h, C.c_int, _sp, _sp, _sp, _sp, C.c_int]),
    "nmma_b200_set_redshift_table": (C.c_int, [_h, C.c_int, _dp, _dp]),
    "nmma_b200_set_observations": (C.c_int, [_h, C.c_int, _ip, _ip, _ip, _dp, _dp, _dp, _dp]),
    "nmma_b200_set_systematics": (C.c_int, [_h, C.c_int, _ip, _dp, _ip, _ip, _sp, _dp]),
    "nmma_b200_set_constraints": (C.c_int, [_h, C.c_int, _sp, _dp, _dp]),
    "nmma_b200_set_extinction": (C.c_int, [_h, C.c_int, _sp, _dp, _dp]),
    "nmma_b200_logl": (C.c_int, [_h, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "nmma_b200_logl_host": (C.c_int, [_h, _dp, C.c_int64, _dp]),
    "nmma_b200_logl_host_to_device": (C.c_int, [_h, _dp, C.c_int64, C.c_void_p]),
    "nmma_b200_mags": (C.c_int, [_h, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nmma_b200_coeffs": (C.c_int, [_h, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "nmma_b200_set_priors": (C.c_int, [_h, C.c_int, _ip, _dp, _ip, _dp, _dp]),
    "nmma_b200_prior_transform": (C.c_int, [_h, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "nmma_b200_prior_sample": (C.c_int, [_h, C.c_uint64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nmma_b200_logl_sweep": (C.c_int, [_h, C.c_uint64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nmma_b200_set_option": (C.c_int, [_h, C.c_char_p, C.c_int64]),
    "nmma_b200_get_info": (C.c_int, [_h, C.c_char_p, C.POINTER(C.c_int64)]),
    "nmma_b200_ffma_peak": (C.c_int, [_h, C.c_int, C.c_int, _dp]),
    "nmma_b200_tf32_peak": (C.c_int, [_h, C.c_int, _dp]),
    "nmma_b200_dfma_peak": (C.c_int, [_h, C.c_int, _dp]),
    "nmma_b200_obs_terms": (C.c_int, [_h, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp]),
}

_lib = None


def load():
    """Load libnmma_b200.so and attach prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()'). nmma_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
