"""Surrogate weight loading: on-disk NMMA model files -> packed host arrays.

In the reference the surrogate weights are loaded inside
``SVDLightCurveModel.__init__`` / ``load_filt_model`` (``nmma/em/model.py:568-696``)
into Keras / scikit-learn objects.  Here the same files are unpacked into plain
NumPy arrays (:class:`SurrogateWeights`) that ``nmma_b200.engine`` stages once into
device buffers.  On-disk layout contract (``nmma/core/gitlab.py:214-232``,
``nmma/em/model.py:593-606,681-686``):

    {svd_path}/{model}.joblib                     SVD metadata dict, filter keys with '_' for ':'
    {svd_path}/{model}_tf/{filt}.h5 | .keras      per-filter Keras MLP   (interpolation_type tensorflow/keras)
    {svd_path}/{model}/{filt}.joblib              per-filter list of GPs (interpolation_type sklearn_gp)
"""
from __future__ import annotations

import io
import os
import re
import zipfile
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from .h5mini import H5File, H5FormatError

__all__ = ["SurrogateWeights", "load_svd_core", "load_keras_mlp", "load_sklearn_gps",
           "load_surrogate", "random_surrogate", "H5File", "H5FormatError"]


@dataclass
class SurrogateWeights:
    """Everything the device needs for one SVD surrogate (all filters).

    Shapes (F filters, d inputs, T grid nodes, K coefficients kept):
      tt (F,T)  param_mins/param_maxs (F,d)  VA (F,T,K)  mins/maxs (F,T)   float64
      MLP:  W1 (F,d,H) b1 (F,H) W2 (F,H,K_out) b2 (F,K_out)                float32
      GP:   X (Ntr,d)  alpha (F,K,Ntr)  c2/rq_alpha/rq_len/ymean/ystd (F,K) float64
    """
    filters: List[str]
    kind: str                      # "mlp" | "gp"
    tt: np.ndarray
    param_mins: np.ndarray
    param_maxs: np.ndarray
    VA: np.ndarray
    mins: np.ndarray
    maxs: np.ndarray
    n_coeff: int
    W1: Optional[np.ndarray] = None
    b1: Optional[np.ndarray] = None
    W2: Optional[np.ndarray] = None
    b2: Optional[np.ndarray] = None
    X: Optional[np.ndarray] = None
    alpha: Optional[np.ndarray] = None
    c2: Optional[np.ndarray] = None
    rq_alpha: Optional[np.ndarray] = None
    rq_len: Optional[np.ndarray] = None
    ymean: Optional[np.ndarray] = None
    ystd: Optional[np.ndarray] = None
    meta: dict = field(default_factory=dict)

    @property
    def F(self):
        return len(self.filters)

    @property
    def d(self):
        return self.param_mins.shape[1]

    @property
    def T(self):
        return self.tt.shape[1]


def load_svd_core(modelfile: str) -> Dict[str, dict]:
    """``joblib.load`` of ``{model}.joblib`` with the '_' -> ':' key fix of ``em/model.py:602-606``."""
    import joblib
    if not os.path.isfile(modelfile):
        raise ValueError(f"Model file not found: {modelfile}\n If possible, try removing the --local-only flag and rerun.")
    raw = joblib.load(modelfile)
    return {k.replace("_", ":"): v for k, v in raw.items()}


def _dense_chain(tensors: Dict[str, np.ndarray]):
    """Order (kernel, bias) pairs of a Sequential of Dense layers by layer index."""
    layers = {}
    for path, arr in tensors.items():
        parts = path.strip("/").split("/")
        leaf = parts[-1]
        if "optimizer" in path:
            continue
        if leaf in ("kernel:0", "kernel", "bias:0", "bias"):
            lname = parts[-2]
            layers.setdefault(lname, {})["kernel" if leaf.startswith("kernel") else "bias"] = arr
        elif parts[-2:-1] == ["vars"] and leaf in ("0", "1"):   # Keras-3: layers/<name>/vars/{0,1}
            lname = parts[-3]
            layers.setdefault(lname, {})["kernel" if leaf == "0" else "bias"] = arr

    def idx(name):
        m = re.search(r"(\d+)$", name)
        return int(m.group(1)) if m else -1

    ordered = [layers[k] for k in sorted(layers, key=idx) if "kernel" in layers[k]]
    return ordered


def load_keras_mlp(model_file: str):
    """Read ``Dense(H, relu) -> Dropout -> Dense(K)`` weights from a Keras ``.h5`` / ``.keras`` file.

    Replaces ``keras.saving.load_model(model_file, compile=False)`` (``em/model.py:637-648``).
    Returns (W1 (d,H), b1 (H,), W2 (H,K), b2 (K,)) float32, kernels stored (in, out) as Keras does.
    """
    if zipfile.is_zipfile(model_file):
        with zipfile.ZipFile(model_file) as zf:
            member = next((n for n in zf.namelist() if n.endswith("model.weights.h5")), None)
            if member is None:
                raise ValueError(f"{model_file}: no model.weights.h5 inside the .keras archive")
            h5 = H5File(zf.read(member))
    else:
        h5 = H5File(model_file)
    tensors = h5.datasets()
    chain = _dense_chain(tensors)
    if len(chain) != 2:
        raise ValueError(f"{model_file}: expected 2 Dense layers (d->H->K surrogate), found {len(chain)}")
    (l1, l2) = chain
    W1 = np.ascontiguousarray(l1["kernel"], np.float32)
    W2 = np.ascontiguousarray(l2["kernel"], np.float32)
    if W1.shape[1] != W2.shape[0]:           # sorted the wrong way round
        W1, W2, l1, l2 = W2, W1, l2, l1
    if W1.ndim != 2 or W2.ndim != 2 or W1.shape[1] != W2.shape[0]:
        raise ValueError(f"{model_file}: dense kernel shapes {W1.shape}, {W2.shape} do not chain")
    b1 = np.ascontiguousarray(l1.get("bias", np.zeros(W1.shape[1])), np.float32)
    b2 = np.ascontiguousarray(l2.get("bias", np.zeros(W2.shape[1])), np.float32)
    return W1, b1, W2, b2


def load_sklearn_gps(gp_file_or_list):
    """Unpack a list of fitted ``GaussianProcessRegressor`` (``em/training.py:429-453``).

    Only the kernel the reference trains is accepted: ``ConstantKernel * RationalQuadratic``
    with scalar length scale.  Returns dict of arrays X (Ntr,d), alpha (K,Ntr), c2, rq_alpha,
    rq_len, ymean, ystd (K,).
    """
    if isinstance(gp_file_or_list, (str, os.PathLike)):
        import joblib
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            gps = joblib.load(gp_file_or_list)
    else:
        gps = gp_file_or_list
    X = None
    alpha, c2, ra, rl, ym, ys = [], [], [], [], [], []
    for gp in gps:
        k = gp.kernel_
        k1, k2 = getattr(k, "k1", None), getattr(k, "k2", None)
        if type(k).__name__ != "Product" or type(k1).__name__ != "ConstantKernel" \
                or type(k2).__name__ != "RationalQuadratic" or np.ndim(k2.length_scale) != 0:
            raise ValueError(f"unsupported GP kernel {k!r}; expected C**2 * RationalQuadratic(alpha, length_scale)")
        Xi = np.asarray(gp.X_train_, np.float64)
        if X is None:
            X = Xi
        elif X.shape != Xi.shape or not np.array_equal(X, Xi):
            raise ValueError("GPs of one filter must share X_train_")
        alpha.append(np.asarray(gp.alpha_, np.float64).reshape(-1))
        c2.append(float(k1.constant_value))
        ra.append(float(k2.alpha))
        rl.append(float(k2.length_scale))
        ym.append(float(np.ravel(gp._y_train_mean)[0]))
        ys.append(float(np.ravel(gp._y_train_std)[0]))
    return dict(X=X, alpha=np.stack(alpha), c2=np.array(c2), rq_alpha=np.array(ra),
                rq_len=np.array(rl), ymean=np.array(ym), ystd=np.array(ys))


def _stack_core(core: Dict[str, dict], filters: Sequence[str], ncoeff: Optional[int]):
    for f in filters:
        if f not in core:
            raise KeyError(f)
    n_file = min(int(core[f]["n_coeff"]) for f in filters)
    n = min(int(ncoeff), n_file) if ncoeff else n_file        # lightcurve_generation.py:182-185
    tt = np.stack([np.asarray(core[f]["tt"], np.float64) for f in filters])
    pm = np.stack([np.asarray(core[f]["param_mins"], np.float64) for f in filters])
    pM = np.stack([np.asarray(core[f]["param_maxs"], np.float64) for f in filters])
    VA = np.stack([np.ascontiguousarray(np.asarray(core[f]["VA"], np.float64)[:, :n]) for f in filters])
    mins = np.stack([np.asarray(core[f]["mins"], np.float64) for f in filters])
    maxs = np.stack([np.asarray(core[f]["maxs"], np.float64) for f in filters])
    return n, tt, pm, pM, VA, mins, maxs


def pack_surrogate(core: Dict[str, dict], filters: Sequence[str], kind: str,
                   ncoeff: Optional[int] = None) -> SurrogateWeights:
    """Pack an in-memory ``svd_mag_model`` (reference layout, ``['model']`` = (W1,b1,W2,b2) tuple or
    ``['gps']`` = list of GPs / unpacked dict) into :class:`SurrogateWeights`."""
    filters = list(filters)
    n, tt, pm, pM, VA, mins, maxs = _stack_core(core, filters, ncoeff)
    sw = SurrogateWeights(filters=filters, kind=kind, tt=tt, param_mins=pm, param_maxs=pM,
                          VA=VA, mins=mins, maxs=maxs, n_coeff=n)
    if kind == "mlp":
        ws = [core[f]["model"] for f in filters]
        sw.W1 = np.stack([np.asarray(w[0], np.float32) for w in ws])
        sw.b1 = np.stack([np.asarray(w[1], np.float32) for w in ws])
        sw.W2 = np.stack([np.asarray(w[2], np.float32) for w in ws])
        sw.b2 = np.stack([np.asarray(w[3], np.float32) for w in ws])
        if sw.W2.shape[2] != n:
            # np.dot(VA[:, :n], cAproj) in lightcurve_generation.py:214 needs len(cAproj) == n
            raise ValueError(f"shapes ({sw.T},{n}) and ({sw.W2.shape[2]},) not aligned: "
                             f"the network emits {sw.W2.shape[2]} coefficients but n_coeff={n}")
    elif kind == "gp":
        gps = [core[f]["gps"] if isinstance(core[f]["gps"], dict) else load_sklearn_gps(core[f]["gps"])
               for f in filters]
        X = gps[0]["X"]
        for g in gps:
            if g["X"].shape != X.shape or not np.array_equal(g["X"], X):
                raise ValueError("all filters must share the GP training inputs (em/training.py:216,230)")
            if g["alpha"].shape[0] < n:
                raise IndexError("list index out of range")   # gps[i] for i < n_coeff
        sw.X = np.ascontiguousarray(X)
        sw.alpha = np.stack([g["alpha"][:n] for g in gps])
        for name in ("c2", "rq_alpha", "rq_len", "ymean", "ystd"):
            setattr(sw, name, np.stack([g[name][:n] for g in gps]))
    else:
        raise ValueError(kind)
    return sw


def load_surrogate(model: str, svd_path: str, filters: Optional[Sequence[str]] = None,
                   interpolation_type: str = "keras", ncoeff: Optional[int] = None,
                   verbose: bool = True):
    """File-level loader following ``SVDLightCurveModel.__init__`` (``em/model.py:581-653``).

    Returns (core dict with per-filter ``'model'``/``'gps'`` entries attached, filters, SurrogateWeights).
    """
    comps = model.split("_")
    if "tf" in comps:
        comps.remove("tf")
    core_name = "_".join(comps)
    specifier = "_tf" if interpolation_type == "tensorflow" else ""
    core = load_svd_core(os.path.join(svd_path, f"{core_name}.joblib"))
    if filters is None:
        filters = list(core.keys())
    filters = list(filters)

    if interpolation_type == "sklearn_gp":
        exts, target = ["joblib"], "gps"
        outdir = os.path.join(svd_path, f"{model}{specifier}")
    elif interpolation_type in ("keras", "tensorflow", "torch", "jax"):
        exts, target = ["keras", "h5"], "model"
        outdir = os.path.join(svd_path, f"{core_name}{specifier}")
    elif interpolation_type == "api_gp":
        raise ValueError("--interpolation-type api_gp is not supported by nmma_b200 (see DESIGN.md, out of scope)")
    else:
        raise ValueError("--interpolation-type must be sklearn_gp, api_gp or tensorflow")

    found: List[str] = []
    for ext in exts:                                   # em/model.py:641-648: .keras first, then legacy .h5
        found, not_found = [], []
        for filt in filters:
            fn = os.path.join(outdir, f"{filt.replace(':', '_')}.{ext}")
            if os.path.isfile(fn):
                core[filt][target] = load_keras_mlp(fn) if target == "model" else load_sklearn_gps(fn)
                found.append(filt)
            else:
                not_found.append(filt)
        if found:
            if not_found and verbose:
                print(f"Warning: No {ext}-model files found for filters: {not_found} at {outdir}")
            break
    if not found:
        raise ValueError(f"No {exts[-1]}-model files found for {model} in {outdir}")
    kind = "mlp" if target == "model" else "gp"
    return core, filters, found, kind


def random_surrogate(filters: Sequence[str], d: int, kind: str = "mlp", T: int = 211, K: int = 10,
                     H: int = 2048, Ntr: int = 329, seed: int = 0,
                     param_mins=None, param_maxs=None, tt=None) -> Dict[str, dict]:
    """Random-init surrogate of a given architecture in the reference's in-memory layout.

    Used when the Zenodo/GitLab weights are unavailable offline (BASELINE.json north_star):
    He-normal W1, Glorot-uniform W2, zero biases as Keras initialises them
    (``em/training.py:353-364``); ``VA`` = orthonormal T x T; smooth ``mins``/``maxs`` in the
    fixture's magnitude range.  GP variant: smooth alpha vectors from a kernel solve so the
    cancellation in k.alpha is realistic.
    """
    rng = np.random.default_rng(seed)
    tt = np.arange(0.0, 21.0 + 1e-9, 0.1)[:T] if tt is None else np.asarray(tt, float)
    T = len(tt)
    pm = np.zeros(d) if param_mins is None else np.asarray(param_mins, float)
    pM = np.ones(d) if param_maxs is None else np.asarray(param_maxs, float)
    core = {}
    X = rng.uniform(0, 1, size=(Ntr, d)) if kind == "gp" else None
    for filt in filters:
        q, _ = np.linalg.qr(rng.normal(size=(T, T)))
        s = tt / max(tt[-1], 1e-9)
        mins = -16.0 + 6.0 * s + rng.normal(scale=0.2)
        maxs = mins + 3.0 + 40.0 * s ** 2 + rng.uniform(0, 2)
        entry = dict(param_mins=pm.copy(), param_maxs=pM.copy(), mins=mins, maxs=maxs, tt=tt.copy(),
                     n_coeff=K, VA=q)
        if kind == "mlp":
            W1 = (rng.normal(size=(d, H)) * np.sqrt(2.0 / d)).astype(np.float32)
            lim = np.sqrt(6.0 / (H + K))
            W2 = rng.uniform(-lim, lim, size=(H, K)).astype(np.float32)
            b1 = (0.05 * rng.normal(size=H)).astype(np.float32)
            b2 = (0.05 * rng.normal(size=K)).astype(np.float32)
            entry["model"] = (W1, b1, W2, b2)
        else:
            # real scikit-learn regressors with fixed hyper-parameters drawn in the fixture's
            # ranges (C in [0.1,3], alpha in [1e-3,0.4], length_scale in [2e-2,1.2])
            from sklearn.gaussian_process import GaussianProcessRegressor
            from sklearn.gaussian_process.kernels import ConstantKernel, RationalQuadratic
            gps = []
            for i in range(K):
                c2 = float(np.exp(rng.uniform(np.log(0.1 ** 2), np.log(3.0 ** 2))))
                ra = float(np.exp(rng.uniform(np.log(1e-3), np.log(0.4))))
                rl = float(np.exp(rng.uniform(np.log(2e-2), np.log(1.2))))
                y = np.sin(3.0 * X @ rng.normal(size=d)) * (2.0 / (i + 1)) + 0.1 * rng.normal(size=Ntr)
                kern = ConstantKernel(c2, "fixed") * RationalQuadratic(length_scale=rl, alpha=ra,
                                                                        length_scale_bounds="fixed",
                                                                        alpha_bounds="fixed")
                gps.append(GaussianProcessRegressor(kernel=kern, optimizer=None).fit(X, y))
            entry["gps"] = gps
            entry["param_array_postprocess"] = X
        core[filt] = entry
    return core
