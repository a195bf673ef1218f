"""One configuration of the path rendered a few times (for ncu captures; never a bench value).

    python tools/render_once.py --size 1920x1080 --fn MirageSphericalSigmoid --flags 0 --frames 3 [--quiet 0|1|2] [--overlay 0|1]
Prints the kernel times, the kernel's own counters and the cubin hash of the march kernel (recorded next to the capture).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linon_b200 import _capi as capi  # noqa: E402
from linon_b200.main_view import MainView, PredefinedFunction, default_scene, make_settings  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", default="3840x2160")
ap.add_argument("--fn", default="MirageSphericalSigmoid")
ap.add_argument("--flags", type=int, default=capi.RENDER_ENHANCE)
ap.add_argument("--frames", type=int, default=3)
ap.add_argument("--quiet", type=int, default=None)
ap.add_argument("--overlay", type=int, default=None)
ap.add_argument("--specialize", type=int, default=None)
ap.add_argument("--kernel", type=int, default=None)
ap.add_argument("--block", type=int, default=None)
ap.add_argument("--overlay-mode", type=int, default=1)
ap.add_argument("--delta", type=int, default=1)
a = ap.parse_args()
W, H = [int(x) for x in a.size.split("x")]
verts, faces, center = default_scene()
with MainView(verts, faces, center, W, H) as mv:
    mv.reload_shader(PredefinedFunction[a.fn].to_code())
    mv.update_settings(make_settings(overlay_mode=a.overlay_mode, central_difference_delta=a.delta))
    for opt, val in ((capi.OPT_QUIET_ZONE, a.quiet), (capi.OPT_OVERLAY_KERNEL, a.overlay), (capi.OPT_SPECIALIZE, a.specialize),
                     (capi.OPT_KERNEL, a.kernel), (capi.OPT_BLOCK, a.block)):
        if val is not None:
            mv.set_option(opt, val)
    for _ in range(a.frames):
        mv.render_flags(a.flags)
        mv.sync()
    t, c = mv.timings(), mv.counters()
    print(json.dumps(dict(size=a.size, fn=a.fn, flags=a.flags, march_ms=t.march_ms, overlay_ms=t.overlay_ms, sample_rays_ms=t.sample_rays_ms,
                          regs=t.march_regs, grid=t.march_grid, block=t.march_block, cubin_hash=mv.march_cubin_hash(),
                          accepted=int(c.accepted_steps), rejected=int(c.rejected_steps), prehit=int(c.prehit_steps),
                          face_tests=int(c.face_tests), culling_checks=int(c.culling_checks), quiet_steps=int(mv.debug_profile()[15]))))
