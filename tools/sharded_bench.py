"""linon_sharded_* timed from ONE host process (developer / evidence tool, run under gpurun --gpus N):

    python tools/sharded_bench.py --devices 0,1 [--size 5431x3055] [--frames 10] [--mode 0|1]

Renders the Enhance frame of bench.py's weak-scaling workload (or any size) sharded over the devices and prints ms/frame
(host wall clock around `frames` back-to-back linon_sharded_render calls + one sync: the calls are asynchronous, so frame k's
overlay on the root overlaps nothing but is ordered by events only - no host synchronisation inside the loop), the per-shard
march times and the comparison with one device doing the whole frame.  mode 0 = peer stores, 1 = NCCL gather."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linon_b200 import _capi as capi  # noqa: E402
from linon_b200.main_view import MainView, PredefinedFunction, ShardedFrame, default_scene, make_settings  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--devices", default="0,1")
ap.add_argument("--size", default=None)
ap.add_argument("--frames", type=int, default=10)
ap.add_argument("--mode", type=int, default=0)
ap.add_argument("--rows-per-block", type=int, default=32)
ap.add_argument("--check", type=int, default=1, help="compare with the one-device frame bit for bit (first frame)")
a = ap.parse_args()
devices = [int(x) for x in a.devices.split(",")]
n = len(devices)
if a.size:
    W, H = [int(x) for x in a.size.split("x")]
else:
    s = n ** 0.5
    W, H = int(round(3840 * s)), int(round(2160 * s))
verts, faces, center = default_scene()
fn = PredefinedFunction.MirageSphericalSigmoid
st = make_settings(overlay_mode=1)
flags = capi.RENDER_ENHANCE

with MainView(verts, faces, center, W, H, device=devices[0]) as mv:
    cam = mv.camera.uniform()
    mv.reload_shader(fn.to_code())
    mv.update_settings(st)
    for _ in range(2):
        mv.render_flags(flags)
        mv.sync()
    t0 = time.perf_counter()
    for _ in range(max(2, a.frames // 2)):
        mv.render_flags(flags)
    mv.sync()
    one_ms = (time.perf_counter() - t0) / max(2, a.frames // 2) * 1e3
    ref = {b: mv.read(b) for b in (capi.BUF_HIT_ID_I32, capi.BUF_ENDPOINT_F32X4, capi.BUF_FINAL_RGBA8)} if a.check else None

with ShardedFrame(devices, verts, faces, W, H, rows_per_block=a.rows_per_block, mode=a.mode) as sf:
    sf.set_camera_uniform(cam)
    sf.reload_shader(fn.to_code())
    sf.update_settings(st)
    for _ in range(3):
        sf.render_flags(flags)
        sf.sync()
    equal = None
    if ref is not None:
        equal = all(np.array_equal(sf.read_root(b).view(np.uint8), ref[b].view(np.uint8)) for b in ref)
    t0 = time.perf_counter()
    for _ in range(a.frames):
        sf.render_flags(flags)
    sf.sync()
    ms = (time.perf_counter() - t0) / a.frames * 1e3
    sf.render_flags(flags)
    sf.sync()
    march = [round(sf.timings(r).march_ms, 3) for r in range(n)]
    overlay = round(sf.timings(0).overlay_ms, 3)
print(json.dumps(dict(devices=devices, mode=["peer stores", "nccl gather"][a.mode], size=f"{W}x{H}", frames=a.frames, ms_per_frame=round(ms, 3),
                      one_device_ms_per_frame=round(one_ms, 3), speedup=round(one_ms / ms, 3), efficiency=round(one_ms / ms / n, 4),
                      mrays_s=round(W * H / ms / 1e3, 1), march_ms_per_shard=march, root_overlay_ms=overlay, equals_one_device_frame=equal)))
