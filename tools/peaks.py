#!/usr/bin/env python
"""Measured pipe peaks of the device (roofline denominators): FFMA, DFMA, dense tcgen05 kind::tf32."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nmma_b200.engine import KilonovaEngine
e = KilonovaEngine(0)
print(f"ffma scalar {e.ffma_peak(0, 20000) / 1e12:.2f} TFLOP/s, packed {e.ffma_peak(1, 20000) / 1e12:.2f} TFLOP/s")
print(f"dfma        {e.dfma_peak(20000) / 1e12:.2f} TFLOP/s")
for it in (2000, 20000):
    print(f"tf32 tcgen05 (128x128x8, A in TMEM) iters={it}: {e.tf32_peak(it) / 1e12:.1f} TFLOP/s")
