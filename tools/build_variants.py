#!/usr/bin/env python
"""Build timing-experiment variants of one translation unit: libnmma_b200.so relinked with that unit compiled under extra
-D switches (TCV_* in tc_kernel.cuh, GPV_* in gp_kernel.cuh).
    python tools/build_variants.py [--unit launch_tc|launch_gp_d3|launch_tc+api|...] NAME=FLAG[,FLAG...] ...
(several units joined by '+': switches that change a layout shared with the host staging, e.g. TCV_CHUNK, need api too)
Output: nmma_b200/lib/variants/lib_NAME.so (git-ignored, travels to the GPU box); select with NMMA_B200_LIB."""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
g.build()
UNIT = "launch_tc"
if len(sys.argv) > 2 and sys.argv[1] == "--unit":
    UNIT = sys.argv[2]
    del sys.argv[1:3]
VDIR = os.path.join(g.LIB_DIR, "variants")
os.makedirs(VDIR, exist_ok=True)
def one(spec):
    name, _, flags = spec.partition("=")
    defs = [f"-D{f}" for f in flags.split(",") if f]
    extra = ["-DNMMA_DEV_BUILD"] if g.DEV else []
    vobj, spills = {}, []
    for unit in UNIT.split("+"):
        vobj[unit] = os.path.join(VDIR, f"{unit}_{name}.o")
        src, udefs, _ = g.UNITS[unit]
        cmd = ["/usr/local/cuda/bin/nvcc"] + g.NVCC_FLAGS + extra + udefs + defs + ["-c", "-o", vobj[unit], os.path.join(g.CSRC, src)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode: print(r.stderr[-3000:]); raise SystemExit(1)
        if unit == UNIT.split("+")[0]:
            spills = [l.strip() for l in r.stderr.splitlines() if "spill" in l or "Used" in l]
    objs = [vobj.get(u, g._obj(u)) for u in g.UNITS]
    lib = os.path.join(VDIR, f"lib_{name}.so")
    r = subprocess.run(["/usr/local/cuda/bin/nvcc", "--shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs, capture_output=True, text=True)
    if r.returncode: print(r.stderr[-3000:]); raise SystemExit(1)
    for o in vobj.values(): os.remove(o)   # only the linked library travels to the GPU box
    return name, spills
with ThreadPoolExecutor(8) as ex:
    for name, spills in ex.map(one, sys.argv[1:]):
        print(name, spills[:4])
if g.DEV:   # the dev shortcut relinked nmma_b200/lib/libnmma_b200.so without most instantiations: put the full library back
    env = {k: v for k, v in os.environ.items() if k != "NMMA_DEV_BUILD"}
    subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.build()"], cwd=ROOT, env=env, check=True,
                   stdout=subprocess.DEVNULL)
    print("[build_variants] full library restored")
