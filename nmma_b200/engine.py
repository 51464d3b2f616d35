"""Thin Python owner of one ``nmma_b200_t`` handle.

PyTorch is plumbing only: it provides device buffers and the current CUDA
stream; every computation happens inside ``libnmma_b200.so`` behind the C ABI
(``include/nmma_b200.h``).  There is no CPU fallback -- constructing an engine
without the built library or without a B200 raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib as L
from ._lib import ParamSrc

SENTINEL = -1.7976931348623157e308  # np.nan_to_num(-np.inf), nmma/core/base.py:82


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _iptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class KilonovaEngine:
    """Device-resident surrogate + observation tables and the kernels that use them."""

    def __init__(self, device: int = 0):
        self._lib = L.load()
        self._h = C.c_void_p()
        rc = self._lib.nmma_b200_create(int(device), C.byref(self._h))
        if rc != L.OK:
            msg = self._lib.nmma_b200_last_error(None).decode()
            self._h = None
            raise L.NmmaB200Error(rc, msg)
        self.device = int(device)
        self.F = self.d = self.K = self.T = self.P = self.S = 0
        self.prior_P = 0          # set by set_priors; 0 = no device priors staged

    # ---- lifetime -------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.nmma_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != L.OK:
            raise L.NmmaB200Error(rc, self._lib.nmma_b200_last_error(self._h).decode())

    # ---- staging ---------------------------------------------------------------------
    def set_surrogate(self, sw):
        """Stage a :class:`nmma_b200.mlmodel.SurrogateWeights` (replaces load_filt_model)."""
        F, d, K, T = sw.F, sw.d, sw.n_coeff, sw.T
        tt, pm, pM = _f64(sw.tt), _f64(sw.param_mins), _f64(sw.param_maxs)
        VA, mins, maxs = _f64(sw.VA), _f64(sw.mins), _f64(sw.maxs)
        assert VA.shape == (F, T, K) and tt.shape == (F, T) and pm.shape == (F, d)
        self._check(self._lib.nmma_b200_set_svd(self._h, F, d, K, T, _dptr(tt), _dptr(pm), _dptr(pM),
                                                _dptr(VA), _dptr(mins), _dptr(maxs)))
        if sw.kind == "mlp":
            W1 = np.ascontiguousarray(sw.W1, np.float32)
            b1 = np.ascontiguousarray(sw.b1, np.float32)
            W2 = np.ascontiguousarray(sw.W2, np.float32)
            b2 = np.ascontiguousarray(sw.b2, np.float32)
            H, Kout = W1.shape[2], W2.shape[2]
            assert W1.shape == (F, d, H) and W2.shape == (F, H, Kout)
            self._check(self._lib.nmma_b200_set_mlp(self._h, H, Kout, _fptr(W1), _fptr(b1), _fptr(W2), _fptr(b2)))
        elif sw.kind == "gp":
            X, al = _f64(sw.X), _f64(sw.alpha)
            Ntr = X.shape[0]
            assert al.shape == (F, K, Ntr)
            arrs = [_f64(getattr(sw, n)) for n in ("c2", "rq_alpha", "rq_len", "ymean", "ystd")]
            self._check(self._lib.nmma_b200_set_gp(self._h, Ntr, _dptr(X), _dptr(al), *[_dptr(a) for a in arrs]))
        else:
            raise ValueError(sw.kind)
        self.F, self.d, self.K, self.T = F, d, K, T
        self.S = T

    def set_sample_grid(self, sample_times: Optional[Sequence[float]]):
        if sample_times is None:
            self._check(self._lib.nmma_b200_set_sample_grid(self._h, 0, None))
            self.S = self.T
        else:
            st = _f64(sample_times)
            self._check(self._lib.nmma_b200_set_sample_grid(self._h, st.size, _dptr(st)))
            self.S = int(st.size)

    def set_param_layout(self, P: int, model_params: Sequence[ParamSrc], luminosity_distance: ParamSrc = None,
                         timeshift: ParamSrc = None, redshift: ParamSrc = None, z_mode: int = L.Z_ZERO):
        arr = (ParamSrc * len(model_params))(*model_params)
        ref = lambda s: C.byref(s) if s is not None else None      # NULL -> the reference's default (10 pc, 0, 0)
        self._check(self._lib.nmma_b200_set_param_layout(self._h, int(P), arr, ref(luminosity_distance), ref(timeshift),
                                                         ref(redshift), int(z_mode)))
        self.P = int(P)

    def set_redshift_table(self, dist_grid, z_grid):
        dg, zg = _f64(dist_grid), _f64(z_grid)
        assert dg.shape == zg.shape
        self._check(self._lib.nmma_b200_set_redshift_table(self._h, dg.size, _dptr(dg), _dptr(zg)))

    def set_observations(self, helper_lists, times, mags, sigmas, det_limits):
        """``helper_lists[g]``: model-filter indices of observed filter g (1..3);
        ``times/mags/sigmas[g]``: arrays of that filter; ``det_limits[g]``: float."""
        G = len(helper_lists)
        nh = np.array([len(h) for h in helper_lists], np.int32)
        hidx = np.zeros((G, 3), np.int32)
        for g, h in enumerate(helper_lists):
            hidx[g, :len(h)] = h
        off = np.zeros(G + 1, np.int32)
        off[1:] = np.cumsum([len(t) for t in times])
        cat = lambda xs: _f64(np.concatenate([np.asarray(x, float).ravel() for x in xs])) if off[-1] else np.zeros(0)
        t, m, s = cat(times), cat(mags), cat(sigmas)
        lim = _f64(det_limits)
        self._check(self._lib.nmma_b200_set_observations(self._h, G, _iptr(nh), _iptr(hidx), _iptr(off),
                                                         _dptr(t), _dptr(m), _dptr(s), _dptr(lim)))
        self.G = G

    def set_systematics(self, modes, budgets, node_srcs, node_times):
        """Per observed filter g: ``modes[g]`` in SYS_*, ``budgets[g]`` float,
        ``node_srcs[g]`` list of ParamSrc, ``node_times[g]`` list of float (SYS_INTERP)."""
        G = len(modes)
        mode = np.asarray(modes, np.int32)
        bud = _f64(budgets)
        nn = np.array([len(s) for s in node_srcs], np.int32)
        off = np.zeros(G, np.int32)
        off[1:] = np.cumsum(nn)[:-1]
        flat = [s for lst in node_srcs for s in lst]
        src = (ParamSrc * max(len(flat), 1))(*flat)
        times = []
        for g in range(G):
            tg = list(node_times[g]) if node_times[g] is not None else []
            tg = tg + [0.0] * (nn[g] - len(tg))
            times.extend(tg)
        nt = _f64(times if times else [0.0])
        self._check(self._lib.nmma_b200_set_systematics(self._h, G, _iptr(mode), _dptr(bud), _iptr(nn), _iptr(off),
                                                        src, _dptr(nt)))

    def set_constraints(self, srcs: Sequence[ParamSrc], minimums, maximums):
        """Constraint priors (``nmma/core/base.py:67-68``): ``minimum < value < maximum`` or the sentinel."""
        n = len(srcs)
        arr = (ParamSrc * max(n, 1))(*srcs)
        lo, hi = _f64(minimums if n else [0.0]), _f64(maximums if n else [0.0])
        self._check(self._lib.nmma_b200_set_constraints(self._h, n, arr, _dptr(lo), _dptr(hi)))

    def set_extinction(self, law: int, ebv: ParamSrc = None, nu0=None, coef=None):
        """Extinction law (``nmma/em/model.py:323-350``): ``nu0[F]`` [Hz] for P92_SMC_host, ``coef[F]`` for the linear law."""
        nu = _f64(nu0) if nu0 is not None else None
        cf = _f64(coef) if coef is not None else None
        for a in (nu, cf):
            assert a is None or a.shape == (self.F,)
        self._check(self._lib.nmma_b200_set_extinction(self._h, int(law), C.byref(ebv) if ebv is not None else None,
                                                       _dptr(nu) if nu is not None else None,
                                                       _dptr(cf) if cf is not None else None))

    # ---- compute ----------------------------------------------------------------------
    @staticmethod
    def _check_out(out, n, device):
        """A caller-supplied result tensor goes straight to the C ABI: it must be exactly what the kernel writes."""
        import torch
        if not (isinstance(out, torch.Tensor) and out.is_cuda and out.dtype == torch.float64 and out.is_contiguous()
                and out.numel() == n and out.device == device):
            raise ValueError(f"out must be a contiguous float64 CUDA tensor with {n} elements on {device}")

    @staticmethod
    def _stream():
        import torch
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _as_device_points(self, points):
        import torch
        if isinstance(points, torch.Tensor):
            if not points.is_cuda:
                raise ValueError("tensor points must live on the GPU; pass a NumPy array for host data")
            pts = points.to(dtype=torch.float64).contiguous()
        else:
            pts = torch.from_numpy(_f64(points)).to(f"cuda:{self.device}")
        if pts.ndim != 2 or pts.shape[1] != self.P:
            raise ValueError(f"points must have shape [N, {self.P}], got {tuple(pts.shape)}")
        return pts

    def logl_device(self, points, out=None):
        """log L for a CUDA tensor ``points[N,P]``; returns a CUDA float64 tensor (async on the current stream)."""
        import torch
        pts = self._as_device_points(points)
        N = pts.shape[0]
        if out is None:
            out = torch.empty(N, dtype=torch.float64, device=pts.device)
        else:
            self._check_out(out, N, pts.device)
        with torch.cuda.device(pts.device):
            self._check(self._lib.nmma_b200_logl(self._h, C.c_void_p(pts.data_ptr()), N,
                                                 C.c_void_p(out.data_ptr()), self._stream()))
        return out

    def logl_host(self, points, out: Optional[np.ndarray] = None) -> np.ndarray:
        """log L for host ``points[N,P]`` through ``nmma_b200_logl_host`` (H2D + kernels + D2H).
        Page-locked arrays (e.g. views of ``torch`` pinned tensors) are copied without staging."""
        pts = _f64(points)
        if pts.ndim == 1:
            pts = pts[None, :]
        if pts.shape[1] != self.P:
            raise ValueError(f"points must have shape [N, {self.P}], got {pts.shape}")
        if out is None:
            out = np.empty(pts.shape[0], np.float64)
        if hasattr(out, "is_cuda"):        # CUDA tensor: the result stays on the device (sharded path, NCCL gather next)
            import torch
            self._check_out(out, pts.shape[0], torch.device(f"cuda:{self.device}"))
            self._check(self._lib.nmma_b200_logl_host_to_device(self._h, _dptr(pts), pts.shape[0],
                                                                C.c_void_p(out.data_ptr())))
            return out
        assert out.dtype == np.float64 and out.flags.c_contiguous and out.size == pts.shape[0]
        self._check(self._lib.nmma_b200_logl_host(self._h, _dptr(pts), pts.shape[0], _dptr(out)))
        return out

    def mags(self, points, apparent: bool = False):
        """(mags[N,F,S], tobs[N,S]) CUDA tensors: generate_lightcurve / gen_detector_lc."""
        import torch
        pts = self._as_device_points(points)
        N = pts.shape[0]
        mags = torch.empty((N, self.F, self.S), dtype=torch.float64, device=pts.device)
        tobs = torch.empty((N, self.S), dtype=torch.float64, device=pts.device)
        with torch.cuda.device(pts.device):
            self._check(self._lib.nmma_b200_mags(self._h, C.c_void_p(pts.data_ptr()), N, int(bool(apparent)),
                                                 C.c_void_p(mags.data_ptr()), C.c_void_p(tobs.data_ptr()),
                                                 self._stream()))
        return mags, tobs

    def coeffs(self, points):
        import torch
        pts = self._as_device_points(points)
        N = pts.shape[0]
        out = torch.empty((N, self.F, self.K), dtype=torch.float64, device=pts.device)
        with torch.cuda.device(pts.device):
            self._check(self._lib.nmma_b200_coeffs(self._h, C.c_void_p(pts.data_ptr()), N,
                                                   C.c_void_p(out.data_ptr()), self._stream()))
        return out

    def obs_terms(self, mag, model_mag, sigma_obs, sigma_sys, det_limit) -> np.ndarray:
        arrs = np.broadcast_arrays(*[np.asarray(a, float) for a in (mag, model_mag, sigma_obs, sigma_sys, det_limit)])
        arrs = [_f64(a).ravel() for a in arrs]
        out = np.empty(arrs[0].size, np.float64)
        self._check(self._lib.nmma_b200_obs_terms(self._h, out.size, *[_dptr(a) for a in arrs], _dptr(out)))
        return out

    # ---- priors on the device -------------------------------------------------------------
    def set_priors(self, kinds, params, tables=None):
        """Stage one analytic prior per column (``PriorDict.device_plan``): ``kinds[P]`` in PR_*,
        ``params[P,4]``, ``tables[j] = (cdf, grid)`` for PR_INTERPED columns."""
        kinds = np.ascontiguousarray(kinds, np.int32)
        par = _f64(params).reshape(len(kinds), 4)
        P = len(kinds)
        off = np.zeros(P + 1, np.int32)
        cdfs, grids = [], []
        for j in range(P):
            n = 0
            if tables and tables.get(j) is not None:
                cdf, grid = (np.asarray(a, float).ravel() for a in tables[j])
                assert cdf.shape == grid.shape
                cdfs.append(cdf); grids.append(grid); n = cdf.size
            off[j + 1] = off[j] + n
        if off[-1]:
            cdf, grid = _f64(np.concatenate(cdfs)), _f64(np.concatenate(grids))
            self._check(self._lib.nmma_b200_set_priors(self._h, P, _iptr(kinds), _dptr(par), _iptr(off),
                                                       _dptr(cdf), _dptr(grid)))
        else:
            self._check(self._lib.nmma_b200_set_priors(self._h, P, _iptr(kinds), _dptr(par), _iptr(off), None, None))
        self.prior_P = P

    def _need_priors(self, what):
        if not self.prior_P:
            raise L.NmmaB200Error(L.ERR_STATE, f"{what}: no device priors staged (set_priors was not called or a "
                                               "column's prior has no device transform)")

    def prior_transform(self, unit, out=None):
        """``PriorDict.rescale`` of a CUDA (or NumPy -> copied) ``unit[N,P]``; returns a CUDA tensor."""
        import torch
        if isinstance(unit, torch.Tensor):
            u = unit.to(dtype=torch.float64).contiguous()
            if not u.is_cuda:
                raise ValueError("tensor input must live on the GPU; pass a NumPy array for host data")
        else:
            u = torch.from_numpy(_f64(unit)).to(f"cuda:{self.device}")
        self._need_priors("prior_transform")
        if u.ndim != 2 or u.shape[1] != self.prior_P:
            raise ValueError(f"unit cube must have shape [N, {self.prior_P}], got {tuple(u.shape)}")
        if out is None:
            out = torch.empty_like(u)
        elif not (isinstance(out, torch.Tensor) and out.is_cuda and out.dtype == torch.float64 and out.is_contiguous()
                  and out.shape == u.shape and out.device == u.device):
            raise ValueError("out must be a contiguous float64 CUDA tensor shaped like the unit cube")
        with torch.cuda.device(u.device):
            self._check(self._lib.nmma_b200_prior_transform(self._h, C.c_void_p(u.data_ptr()), u.shape[0],
                                                            C.c_void_p(out.data_ptr()), self._stream()))
        return out

    def prior_sample(self, n: int, seed: int = 0, first_index: int = 0, return_unit: bool = False):
        """``n`` prior draws on the device (Philox4x32-10 keyed by ``seed``, counter = global point index)."""
        import torch
        self._need_priors("prior_sample")
        dev = torch.device(f"cuda:{self.device}")
        pts = torch.empty((int(n), self.prior_P), dtype=torch.float64, device=dev)
        unit = torch.empty_like(pts) if return_unit else None
        with torch.cuda.device(dev):
            self._check(self._lib.nmma_b200_prior_sample(
                self._h, C.c_uint64(int(seed) & (2 ** 64 - 1)), int(first_index), int(n), C.c_void_p(pts.data_ptr()),
                C.c_void_p(unit.data_ptr()) if unit is not None else None, self._stream()))
        return (pts, unit) if return_unit else pts

    def logl_sweep(self, n: int, seed: int = 0, first_index: int = 0, return_points: bool = False, out=None):
        """log L of ``n`` prior draws without host traffic (``nmma_b200_logl_sweep``)."""
        import torch
        self._need_priors("logl_sweep")
        dev = torch.device(f"cuda:{self.device}")
        if out is None:
            out = torch.empty(int(n), dtype=torch.float64, device=dev)
        else:
            self._check_out(out, int(n), dev)
        pts = torch.empty((int(n), self.prior_P), dtype=torch.float64, device=dev) if return_points else None
        with torch.cuda.device(dev):
            self._check(self._lib.nmma_b200_logl_sweep(
                self._h, C.c_uint64(int(seed) & (2 ** 64 - 1)), int(first_index), int(n), C.c_void_p(out.data_ptr()),
                C.c_void_p(pts.data_ptr()) if pts is not None else None, self._stream()))
        return (out, pts) if return_points else out

    # ---- knobs -------------------------------------------------------------------------
    def set_option(self, key: str, value: int):
        self._check(self._lib.nmma_b200_set_option(self._h, key.encode(), int(value)))

    def get_info(self, key: str) -> int:
        v = C.c_int64()
        self._check(self._lib.nmma_b200_get_info(self._h, key.encode(), C.byref(v)))
        return int(v.value)

    def ffma_peak(self, variant: int = 0, iters: int = 20000) -> float:
        v = C.c_double()
        self._check(self._lib.nmma_b200_ffma_peak(self._h, int(variant), int(iters), C.byref(v)))
        return float(v.value)

    def tf32_peak(self, iters: int = 20000) -> float:
        """Dense tcgen05 kind::tf32 FLOP/s measured on this device (roofline denominator of the tensor-core kernel)."""
        v = C.c_double()
        self._check(self._lib.nmma_b200_tf32_peak(self._h, int(iters), C.byref(v)))
        return float(v.value)

    def dfma_peak(self, iters: int = 20000) -> float:
        """fp64 FMA FLOP/s measured on this device (roofline denominator of the GP front end)."""
        v = C.c_double()
        self._check(self._lib.nmma_b200_dfma_peak(self._h, int(iters), C.byref(v)))
        return float(v.value)
