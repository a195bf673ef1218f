"""Offline fuzz of advance_coord (march_kernels.cuh, the collapsed quiet runs) against the chain of single steps it replaces, on
the host emulation of the device code: seven step sizes (the reference's 0.001 and 0.1, user-edited ones, a power of two), magnitudes
from 1e-6 to 1e6, up to 20 000 steps per run.  The CPU suite runs a 1.3 M-case version (tests/test_emulation_cpu.py); this is the
larger sweep quoted in DESIGN.md (10.5 M cases, 0 mismatches).      python tools/dev/advance_coord_fuzz.py [cases per step size]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from emu import emu  # noqa: E402
from linon_b200 import _capi as capi  # noqa: E402
from linon_b200.main_view import PredefinedFunction  # noqa: E402

m = int(sys.argv[1]) if len(sys.argv) > 1 else 1_500_000
lib = emu.build(PredefinedFunction.MirageSphericalSigmoid.to_code(), capi.RENDER_ENHANCE, capi.MATH_FAST, quiet_zone=1)
rng = np.random.default_rng(2026)
total = bad = 0
for h in (0.001, 0.1, 0.003, 0.05, 0.00037, 2.0 ** -9, 1e-5):
    p = (rng.uniform(-3, 3, m) * rng.choice([1e-6, 1e-3, 0.1, 1.0, 10.0, 1e3, 1e6], m)).astype(np.float32)
    d = (rng.uniform(-1.5, 1.5, m) * rng.choice([1e-4, 1e-2, 1.0, 1.0, 1.0, 30.0], m)).astype(np.float32)
    n = rng.integers(0, 20000, m).astype(np.int32)
    a, b = np.zeros_like(p), np.zeros_like(p)
    lib.emu_advance_coord(C.c_int(m), C.c_void_p(p.ctypes.data), C.c_void_p(d.ctypes.data), C.c_float(float(np.float32(h))),
                          C.c_void_p(n.ctypes.data), C.c_void_p(a.ctypes.data), C.c_void_p(b.ctypes.data))
    miss = int((~((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b)))).sum())
    print(f"h = {h:g}: {miss} mismatches of {m}")
    total, bad = total + m, bad + miss
print(f"{total} cases, {bad} mismatches")
sys.exit(1 if bad else 0)
