"""Small frames through every kernel of the path, for compute-sanitizer (race detection):

    compute-sanitizer --tool memcheck  --error-exitcode 7 python tools/sanitize_frames.py
    compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_frames.py --small

Exercises: the four march kernels (tiles, persistent two-phase with the device queue, pairs, refill) in fixed-step,
adaptive and Enhance variants, with and without the quiet zone, the hierarchy (BVH) path, sample rays, click positions,
the outline search + finalize, the three overlay kernels (incl. the TMA tile kernel), and peer stores (three views of one process assembling a frame in
the root view's buffers).  Prints one line per case; exits non-zero if a frame disagrees with the tile kernel's.
"""
import sys
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linon_b200 import _capi as capi  # noqa: E402
from linon_b200.main_view import MainView, PredefinedFunction, default_scene, make_settings  # noqa: E402

small = "--small" in sys.argv
enhance = not small or "--enhance" in sys.argv  # the 5 001-step frame: window resolve and collapsed runs of the quiet zone
W, H = (40, 24) if small else (72, 48)
verts, faces, center = default_scene()
bad = 0


def frame(mv):
    return dict(hit_id=mv.hit_ids(), rgba8=mv.ray_casting_texture(), endpoint=mv.mapping_texture())


with MainView(verts, faces, center, W, H) as mv:
    mv.update_settings(make_settings(overlay_mode=1))
    for fn in (PredefinedFunction.MirageSphericalSigmoid, PredefinedFunction.Rotation):
        mv.reload_shader(fn.to_code())
        for flags, name in ((0, "adaptive"), (capi.RENDER_ADAPTIVE_OFF, "fixed"), (capi.RENDER_ENHANCE, "enhance")):
            if flags == capi.RENDER_ENHANCE and (not enhance or fn != PredefinedFunction.MirageSphericalSigmoid):
                continue
            ref = None
            for kernel in (capi.KERNEL_TILES, capi.KERNEL_PERSISTENT, capi.KERNEL_PAIRS, capi.KERNEL_REFILL):
                for quiet in ((1, 0) if kernel == capi.KERNEL_TILES else (0,)):
                    mv.set_option(capi.OPT_MATH, capi.MATH_STRICT if quiet == 0 else capi.MATH_FAST)
                    mv.set_option(capi.OPT_KERNEL, kernel)
                    mv.set_option(capi.OPT_QUIET_ZONE, 2 if quiet else 0)
                    for ok in (2, 1, 0):  # TMA tile kernel (default), direct taps, hand-staged tile
                        mv.set_option(capi.OPT_OVERLAY_KERNEL, ok)
                        mv.render_flags(flags)
                        mv.sync()
                    got = frame(mv)
                    if quiet == 0:
                        if ref is None:
                            ref = got
                        same = all(np.array_equal(ref[k].view(np.uint8), got[k].view(np.uint8)) for k in ref)
                        bad += not same
                        print(fn.name, name, "kernel", kernel, "strict == tiles:", same, flush=True)
                    else:
                        print(fn.name, name, "kernel", kernel, "quiet zone, fast math: ran", flush=True)
    # outline mode (search over the exponents of the previous overlay) + hierarchy path
    mv.set_option(capi.OPT_KERNEL, capi.KERNEL_TILES)
    mv.set_option(capi.OPT_MATH, capi.MATH_FAST)
    mv.render_flags(capi.RENDER_OUTLINE)
    mv.sync()
    print("outline positions", mv.read(capi.BUF_SAMPLE_POS).round(3).tolist(), flush=True)
    mv.set_option(capi.OPT_FORCE_BVH, 1)
    mv.render_flags(capi.RENDER_ADAPTIVE_OFF)
    mv.sync()
    mv.set_option(capi.OPT_FORCE_BVH, 0)
    print("hierarchy path ran", flush=True)

# peer stores: three shards of one frame store into the root view's buffers
views = [MainView(verts, faces, center, W, H) for _ in range(3)]
try:
    bufs = (capi.BUF_RAYCAST_RGBA8, capi.BUF_ENDPOINT_F32X4, capi.BUF_HIT_ID_I32)
    for rank, v in enumerate(views):
        v.set_shard(rank, 3, 8)
        if rank:
            for b in bufs:
                v.redirect_output(b, views[0].device_ptr(b))
    for v in views:
        v.render_flags(capi.RENDER_ADAPTIVE_OFF | capi.RENDER_NO_OVERLAY | capi.RENDER_NO_SAMPLE_RAYS)
        v.sync()
    with MainView(verts, faces, center, W, H) as whole:
        whole.render_flags(capi.RENDER_ADAPTIVE_OFF | capi.RENDER_NO_OVERLAY | capi.RENDER_NO_SAMPLE_RAYS)
        whole.sync()
        same = all(np.array_equal(a.view(np.uint8), b.view(np.uint8)) for a, b in zip(frame(whole).values(), frame(views[0]).values()))
    bad += not same
    print("peer stores == whole frame:", same, flush=True)
finally:
    for v in views:
        v.close()
print("SANITIZE_FRAMES", "ok" if not bad else f"{bad} mismatches")
sys.exit(1 if bad else 0)
