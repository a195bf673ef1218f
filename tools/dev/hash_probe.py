"""Developer probe: the tile kernel's code hash of the C3 variant from linon_compile_check in this process (optionally after
importing torch), with the cubin kept for a SASS diff.   python tools/dev/hash_probe.py OUT_PREFIX [torch]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
prefix = sys.argv[1]
if len(sys.argv) > 2 and sys.argv[2] == "torch":
    import torch  # noqa: F401
os.environ["LINON_CHECK_QUIET_ZONE"] = "2"
os.environ["LINON_DUMP_CUBIN_HASH"] = prefix + ".hash"
os.environ["LINON_DUMP_CUBIN"] = prefix + ".cubin"
from linon_b200 import _capi as capi  # noqa: E402
from linon_b200.main_view import PredefinedFunction, compile_check, nvrtc_info  # noqa: E402

compile_check(PredefinedFunction.MirageSphericalSigmoid.to_code(), flags=capi.RENDER_ENHANCE, math_mode=capi.MATH_FAST)
print(open(prefix + ".hash").read().strip(), nvrtc_info(), "torch" if "torch" in sys.modules else "no torch")
