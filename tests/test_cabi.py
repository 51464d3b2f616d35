"""The C-ABI library builds, loads and exports every symbol include/nmma_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "nmma_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nmma_b200_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    import __graft_entry__ as ge
    ge.build()
    from nmma_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes prototype in nmma_b200/_lib.py"
    assert set(_lib.SIGNATURES) == set(declared)
    assert _lib.load().nmma_b200_version() == 100


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the product path must fail loudly, not fall back to the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible here")
    from nmma_b200 import _lib
    from nmma_b200.engine import KilonovaEngine
    with pytest.raises(_lib.NmmaB200Error) as ei:
        KilonovaEngine(0)
    assert "no CPU fallback" in str(ei.value)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under nmma_b200/ may import it."""
    pkg = os.path.join(ROOT, "nmma_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(".py"):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dirpath, fn)


def test_sass_has_tma_and_packed_fma():
    """The fused kernel really uses TMA bulk copies (UBLKCP), mbarriers (SYNCS) and packed FFMA2."""
    import shutil
    import subprocess
    from nmma_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.isfile(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-sass", "-fun",
                          "_ZN4nmma21fused_mlp_logl_kernelILi4ELi10ELi2ELb1EEEvNS_6DevCfgEPKdxPd", _lib.LIB_PATH],
                         capture_output=True, text=True).stdout
    if "Function" not in out:
        out = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UBLKCP" in out and "SYNCS" in out and "FFMA2" in out and "LDS.128" in out


def test_sass_tensor_core_kernel_is_tcgen05():
    """The throughput instantiation of the hot path, fused_tc_logl_kernel<10, FAST, unsplit>, really runs on tcgen05:
    UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit), UBLKCP (cp.async.bulk), SYNCS
    (mbarrier) -- and no legacy HMMA / HGMMA.  profiles/r02_sass_fused_tc.txt is the committed histogram."""
    import shutil
    import subprocess
    from nmma_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.isfile(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    blocks = [b for b in re.split(r"\n\s*Function : ", sass) if b.startswith("_ZN4nmma20fused_tc_logl_kernelILi10ELb1ELb0E")]
    assert len(blocks) == 1, "fused_tc_logl_kernel<10, true, false> not found in the library"
    body = blocks[0]
    counts = {op: len(re.findall(r"\b" + op, body))
              for op in ("UTCHMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "SYNCS", "FHFMA", r"F2FP\.RELU")}
    # 9 MMAs per 64-hidden chunk (+ the prologue), in-place [h_hi | h_lo] stores of 32 columns
    assert counts["UTCHMMA"] >= 18 and counts["LDTM"] >= 4 and counts["STTM"] >= 3, counts
    # the 2-instruction-per-hidden-unit ReLU + fp16 hi/lo split: F2FP.RELU (cvt.{rz,rn}.relu.f16x2.f32) + FHFMA (fma.f32.f16)
    assert counts["FHFMA"] >= 64 and counts[r"F2FP\.RELU"] >= 64, counts
    assert counts["UTCBAR"] >= 4 and counts["UBLKCP"] >= 2 and counts["SYNCS"] >= 20, counts
    assert not re.search(r"\bHMMA|\bHGMMA|\bIMMA", body)


def test_no_struct_passed_by_value():
    """Plain pointers and sizes only.  A {int32, int32, double} struct by value travels in one integer and one SSE register;
    ctypes (CPython 3.12 / libffi) gives every such argument of a call the LAST one's double, so a ctypes binder of a
    by-value prototype silently evaluated at luminosity_distance = the redshift constant (found by the
    OpticalLightCurve dict test of round 2).  The header therefore takes nmma_b200_param_src by pointer everywhere."""
    text = open(os.path.join(ROOT, "include", "nmma_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    for proto in re.findall(r"\bnmma_b200_[a-z_0-9]+\s*\(([^;{]*)\)\s*;", text):
        for arg in proto.split(","):
            assert not re.search(r"\bnmma_b200_param_src\s+\w+\s*$", arg.strip()), f"by-value struct argument: {arg.strip()}"
    from nmma_b200 import _lib
    for name, (_, argtypes) in _lib.SIGNATURES.items():
        assert not any(isinstance(a, type) and issubclass(a, ctypes.Structure) for a in argtypes), name
