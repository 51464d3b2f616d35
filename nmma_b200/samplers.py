"""Batched nested sampling for the GPU likelihood (the sampler-side adapter).

The reference hands its likelihood to bilby -> pymultinest / dynesty, one point per call
(``nmma/core/base.py:316-329``, ``nmma/em/analysis.py:183-260``).  None of those samplers is available offline and
none of them can feed a kernel that wants 10^4 points per launch, so this module carries the smallest sampler that
can: single-ellipsoid nested sampling in the unit cube (Mukherjee, Parkinson & Liddle 2006; the bound dynesty calls
``'single'``) with the candidate queue evaluated in batches.  It is deliberately sampler-agnostic about the model:
``loglike`` maps unit-cube points ``u[N, ndim]`` to ``logL[N]`` (prior transform included), which is exactly what
``EMTransientLikelihood.vectorized()`` returns.  Multi-modal posteriors need a multi-ellipsoid bound; the kilonova
posteriors of the graded configurations are unimodal.
"""
from __future__ import annotations

from typing import Callable, Dict

import numpy as np

__all__ = ["nested_sample", "equal_weight"]


def _ellipsoid(u: np.ndarray, enlarge: float):
    """Mean and Cholesky factor of the covariance ellipsoid scaled to contain every live point, times `enlarge`
    in volume."""
    ndim = u.shape[1]
    mean = u.mean(axis=0)
    cov = np.cov(u, rowvar=False).reshape(ndim, ndim)
    cov += 1e-18 * np.eye(ndim) * max(np.trace(cov), 1e-300)
    for _ in range(8):
        try:
            chol = np.linalg.cholesky(cov)
            break
        except np.linalg.LinAlgError:
            cov += 1e-12 * np.eye(ndim) * max(np.trace(cov), 1e-300)
    else:  # degenerate live set: fall back to the axis-aligned box
        chol = np.diag(np.maximum(u.std(axis=0), 1e-12))
    d = np.linalg.solve(chol, (u - mean).T)
    r2 = float((d * d).sum(axis=0).max())
    return mean, chol * np.sqrt(r2) * enlarge ** (1.0 / ndim)


def _draw(rng, mean, axes, n, max_raw=1 << 22):
    """About `n` points uniform in the ellipsoid intersected with the unit cube (the raw draw is enlarged by the
    measured in-cube fraction, so a posterior on a prior edge still fills the likelihood batch)."""
    ndim = mean.size
    out, have, raw = [], 0, n
    for _ in range(8):
        z = rng.standard_normal((raw, ndim))
        z *= (rng.random(raw) ** (1.0 / ndim) / np.linalg.norm(z, axis=1))[:, None]
        x = mean + z @ axes.T
        x = x[np.all((x > 0.0) & (x < 1.0), axis=1)]
        out.append(x); have += x.shape[0]
        if have >= n // 2:
            break
        raw = int(min(max_raw, max(n, (n - have) * raw / max(x.shape[0], 1) * 1.2)))
    return np.concatenate(out)[:n]


def nested_sample(loglike: Callable[[np.ndarray], np.ndarray], ndim: int, nlive: int = 512, batch: int = 8192,
                  dlogz: float = 0.1, seed: int = 0, enlarge: float = 2.0, max_calls: int = 200_000_000,
                  floor: float = -1e300) -> Dict[str, object]:
    """Returns ``log_evidence``, ``log_evidence_err``, the dead + final live points in the unit cube (``samples_u``)
    with ``log_weights`` (normalised posterior weights) and ``log_likelihoods``, ``ncall`` and ``niter``.

    `floor`: log-likelihoods at or below it (the reference's sentinel -1.797e308, ``core/base.py:180-181``) are
    treated as zero likelihood."""
    rng = np.random.default_rng(seed)
    u = rng.random((nlive, ndim))
    logl = np.asarray(loglike(u), dtype=float).copy()
    logl[~(logl > floor)] = -np.inf
    ncall = nlive
    # a likelihood that vanishes on most of the prior (sentinel rows): redraw the dead starts until all are live
    for _ in range(200):
        bad = ~np.isfinite(logl)
        if not bad.any():
            break
        u[bad] = rng.random((int(bad.sum()), ndim))
        lb = np.asarray(loglike(u[bad]), dtype=float)
        lb[~(lb > floor)] = -np.inf
        logl[bad] = lb
        ncall += int(bad.sum())
    frac_live = nlive / ncall          # Monte-Carlo estimate of the prior mass with non-zero likelihood
    logx = np.log(frac_live)
    logz = -np.inf
    shrink = 1.0 / nlive
    logdx_fac = np.log1p(-np.exp(-shrink))
    dead_u, dead_logl, dead_logw = [], [], []
    niter = 0
    while ncall < max_calls:
        mean, axes = _ellipsoid(u, enlarge)
        cand = _draw(rng, mean, axes, batch)
        if cand.shape[0] == 0:
            enlarge = max(1.05, enlarge * 0.9)
            continue
        cl = np.asarray(loglike(cand), dtype=float)
        cl[~(cl > floor)] = -np.inf
        ncall += cand.shape[0]
        worst = int(np.argmin(logl))
        # the threshold only rises: candidates at or below the current minimum can never be accepted
        for c in np.nonzero(cl > logl[worst])[0]:
            lmin = logl[worst]
            if not cl[c] > lmin:
                continue
            logw = lmin + logx + logdx_fac
            dead_u.append(u[worst].copy()); dead_logl.append(lmin); dead_logw.append(logw)
            logz = np.logaddexp(logz, logw)
            logx -= shrink
            u[worst] = cand[c]; logl[worst] = cl[c]
            worst = int(np.argmin(logl))
            niter += 1
        remain = logl.max() + logx
        if np.isfinite(logz) and np.logaddexp(logz, remain) - logz < dlogz:
            break
    # final live points share the remaining volume
    logw_live = logl + logx - np.log(nlive)
    all_u = np.concatenate([np.asarray(dead_u).reshape(-1, ndim), u])
    all_logl = np.concatenate([np.asarray(dead_logl, dtype=float), logl])
    all_logw = np.concatenate([np.asarray(dead_logw, dtype=float), logw_live])
    logz = float(np.logaddexp.reduce(all_logw))
    w = np.exp(all_logw - logz)
    finite = np.isfinite(all_logl) & (w > 0)
    info = float(np.sum(w[finite] * all_logl[finite]) - logz)     # H = <logL>_posterior - logZ
    return {
        "log_evidence": logz,
        "log_evidence_err": float(np.sqrt(max(info, 0.0) / nlive)),
        "information": info,
        "samples_u": all_u,
        "log_likelihoods": all_logl,
        "log_weights": all_logw - logz,
        "ncall": int(ncall),
        "niter": int(niter),
        "prior_mass_with_support": float(frac_live),
    }


def equal_weight(result: Dict[str, object], n: int, seed: int = 0) -> np.ndarray:
    """Indices of `n` equally weighted posterior samples (systematic resampling of the nested-sampling weights)."""
    w = np.exp(np.asarray(result["log_weights"], dtype=float))
    w = w / w.sum()
    rng = np.random.default_rng(seed)
    pos = (rng.random() + np.arange(n)) / n
    return np.minimum(np.searchsorted(np.cumsum(w), pos), w.size - 1)
