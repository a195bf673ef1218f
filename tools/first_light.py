"""First-light / exploration run on a B200 (developer tool, run under gpurun).

Renders small frames with every predefined function in every math mode and both march kernels, compares with the
CPU oracle, then times larger frames.  Writes gpurun_out/first_light.json.
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linon_b200 import _capi as capi  # noqa: E402
from linon_b200 import parity  # noqa: E402
from linon_b200.main_view import MainView, PredefinedFunction, default_scene, make_settings, measure_fp32_peak  # noqa: E402
from oracle import cpu_oracle as co  # noqa: E402

out = {"parity": [], "timing": []}
os.makedirs("gpurun_out", exist_ok=True)

peak, sms = measure_fp32_peak(0)
out["fp32_peak_tflops"] = peak
out["sm_count"] = sms
print("fp32 peak", peak, "TFLOP/s on", sms, "SMs", flush=True)

verts, faces, center = default_scene()
W, H = 160, 120
mv = MainView(verts, faces, center, W, H)
cam = mv.camera.uniform()
cam_d = dict(origin=list(cam.origin[:3]), view_direction=list(cam.view_direction[:3]), up=list(cam.up[:3]))
st_d = dict(field_weight=1.0, mouse_x=0.5, mouse_y=0.6, overlay_mode=0, central_difference_delta=1, lyapunov_scaling=50.0)

for fn in PredefinedFunction:
    if fn == PredefinedFunction.Custom:
        continue
    mv.reload_shader(fn.to_code())
    for adaptive in (False, True):
        ref = co.march(verts, faces, cam_d, st_d, int(fn), W, H, h_initial=0.1, adaptive=adaptive)
        ref["rgba8"] = ref["rgba8"]
        for math in (0, 1, 2):
            for kernel in (0, 1):
                mv.set_option(capi.OPT_MATH, math)
                mv.set_option(capi.OPT_KERNEL, kernel)
                flags = 0 if adaptive else capi.RENDER_ADAPTIVE_OFF
                mv.set_option(capi.OPT_CULL, 0)
                mv.render_flags(flags)
                mv.sync()
                nocull = dict(hit_id=mv.hit_ids(), rgba8=mv.ray_casting_texture(), endpoint=mv.mapping_texture())
                mv.set_option(capi.OPT_CULL, 1)
                mv.render_flags(flags)
                mv.sync()
                test = dict(hit_id=mv.hit_ids(), rgba8=mv.ray_casting_texture(), endpoint=mv.mapping_texture())
                m = parity.compare_frames(test, ref)
                c = mv.counters()
                m["cull_identical"] = bool(all(np.array_equal(test[k].view(np.uint8), nocull[k].view(np.uint8)) for k in test))
                m["face_tests"] = c.face_tests
                m.update(fn=fn.name, adaptive=adaptive, math=math, kernel=kernel,
                         accepted=c.accepted_steps, prehit=c.prehit_steps, rejected=c.rejected_steps,
                         oracle_accepted=int(ref["accepted"].sum()), oracle_prehit=int(ref["prehit"].sum()),
                         oracle_rejected=int(ref["rejected"].sum()))
                out["parity"].append(m)
                print(json.dumps(m), flush=True)

# timing at larger sizes
for (w, h, flags, name) in ((1920, 1080, 0, "1080p default"), (1920, 1080, capi.RENDER_ENHANCE, "1080p enhance")):
    mv.resize_texture(w, h)
    mv.reload_shader(PredefinedFunction.MirageSphericalSigmoid.to_code())
    for math in (1, 2):
        for kernel in (0, 1):
            for cull in (0, 1):
                for refill in ((1, 8) if kernel == 1 else (1,)):
                    mv.set_option(capi.OPT_MATH, math)
                    mv.set_option(capi.OPT_KERNEL, kernel)
                    mv.set_option(capi.OPT_CULL, cull)
                    mv.set_option(capi.OPT_REFILL_MIN, refill)
                    for rep in range(2):
                        t0 = time.time()
                        mv.render_flags(flags)
                        mv.sync()
                        wall = time.time() - t0
                    t = mv.timings()
                    c = mv.counters()
                    rec = dict(name=name, math=math, kernel=kernel, cull=cull, refill=refill, march_ms=t.march_ms,
                               phase_a_ms=t.march_phase_a_ms, overlay_ms=t.overlay_ms,
                               sample_ms=t.sample_rays_ms, jit_ms=t.jit_ms, regs=t.march_regs, grid=t.march_grid, wall_s=wall,
                               accepted=c.accepted_steps, prehit=c.prehit_steps, rejected=c.rejected_steps,
                               face_tests=c.face_tests, plane_tests=c.plane_tests,
                               gsteps_per_s=(c.accepted_steps + c.rejected_steps) / (t.march_ms * 1e-3) / 1e9,
                               mrays_per_s=c.rays / (t.march_ms * 1e-3) / 1e6)
                    out["timing"].append(rec)
                    print(json.dumps(rec), flush=True)

json.dump(out, open("gpurun_out/first_light.json", "w"), indent=1)
print("done")
