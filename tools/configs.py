"""BASELINE.json configs C1, C2 and C4 on one B200 (developer / evidence tool, run under gpurun).

    python tools/configs.py [--out gpurun_out/configs.json] [--reps 3] [--quiet-zone 0|1|2] [--overlay-kernel 0|1]

  C1  800x600, MirageSpherical (first predefined function), default step size (h = 0.1, adaptive sampling on)
  C2  1920x1080, every predefined function of src/functions.rs, default step size
  C4  1920x1080 'Outline' workload: Lyapunov overlay + sample-ray capture in click mode and in outline mode; the march,
      overlay, outline-search and sample-ray kernels are timed separately (CUDA events of the library)
Every line carries the kernel's own step counters (adaptive step counts are implementation specific).
C3 is bench.py; C5 is tools/orbit.py.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linon_b200 import _capi as capi  # noqa: E402
from linon_b200.main_view import MainView, PredefinedFunction, default_scene, make_settings  # noqa: E402

out_path = "gpurun_out/configs.json"
reps = 3
quiet_zone = 2       # the library's defaults (include/linon_b200.h); the sweep of bench.py's `configs` block supersedes this tool
overlay_kernel = 2
args = sys.argv[1:]
for i, a in enumerate(args):
    if a == "--out":
        out_path = args[i + 1]
    if a == "--reps":
        reps = int(args[i + 1])
    if a == "--quiet-zone":  # LINON_OPT_QUIET_ZONE: 0 off, 1 fixed-step variants, 2 adaptive variants as well (default)
        quiet_zone = int(args[i + 1])
    if a == "--overlay-kernel":  # LINON_OPT_OVERLAY_KERNEL: 0 hand-staged tile, 1 direct taps, 2 TMA tile (default)
        overlay_kernel = int(args[i + 1])

verts, faces, center = default_scene()
PREDEFINED = [f for f in PredefinedFunction if f != PredefinedFunction.Custom]


def frame(mv, flags):
    best = None
    for _ in range(reps):
        mv.render_flags(flags)
        mv.sync()
        t = mv.timings()
        if best is None or t.march_ms + t.overlay_ms < best.march_ms + best.overlay_ms:
            best = t
    c = mv.counters()
    W, H = mv.size
    steps = c.accepted_steps + c.rejected_steps
    return dict(march_ms=round(best.march_ms, 4), overlay_ms=round(best.overlay_ms, 4),
                sample_rays_ms=round(best.sample_rays_ms, 4), outline_search_ms=round(best.outline_search_ms, 4),
                jit_ms=round(best.jit_ms, 1), mrays_s=round(W * H / best.march_ms / 1e3, 2),
                gsteps_s=round(steps / best.march_ms / 1e6, 2), ms_per_frame=round(best.march_ms + best.overlay_ms, 4),
                kernel=best.march_kernel, regs=best.march_regs,
                accepted_per_ray=round(c.accepted_steps / c.rays, 2), rejected_per_ray=round(c.rejected_steps / c.rays, 3),
                prehit_fraction=round(c.prehit_steps / max(c.accepted_steps, 1), 4),
                face_tests_per_prehit_step=round(c.face_tests / max(c.prehit_steps, 1), 4),
                culling_checks_per_prehit_step=round(c.culling_checks / max(c.prehit_steps, 1), 4))


result = dict(configs=[], quiet_zone=quiet_zone, overlay_kernel=overlay_kernel)
# ---- C1 -----------------------------------------------------------------------------------------------------------
with MainView(verts, faces, center, 800, 600) as mv:
    mv.set_option(capi.OPT_QUIET_ZONE, quiet_zone)
    mv.set_option(capi.OPT_OVERLAY_KERNEL, overlay_kernel)
    mv.reload_shader(PredefinedFunction.MirageSpherical.to_code())
    for name, flags in (("adaptive on (as the application renders)", 0), ("adaptive off (51 steps)", capi.RENDER_ADAPTIVE_OFF)):
        r = frame(mv, flags)
        r.update(config="C1", size="800x600", function="MirageSpherical", variant=name)
        result["configs"].append(r)
        print(json.dumps(r), flush=True)

# ---- C2 -----------------------------------------------------------------------------------------------------------
with MainView(verts, faces, center, 1920, 1080) as mv:
    mv.set_option(capi.OPT_QUIET_ZONE, quiet_zone)
    mv.set_option(capi.OPT_OVERLAY_KERNEL, overlay_kernel)
    for fn in PREDEFINED:
        mv.reload_shader(fn.to_code())
        for name, flags in (("adaptive on", 0), ("adaptive off", capi.RENDER_ADAPTIVE_OFF)):
            r = frame(mv, flags)
            r.update(config="C2", size="1920x1080", function=fn.name, variant=name)
            result["configs"].append(r)
            print(json.dumps(r), flush=True)

    # ---- C4 -------------------------------------------------------------------------------------------------------
    fn = PredefinedFunction.MirageSphericalSigmoid
    mv.reload_shader(fn.to_code())
    for mouse, label in (((0.5, 0.6), "click (0.5, 0.6)"), ((0.36, 0.55), "click on the mirage rim")):
        mv.update_settings(make_settings(overlay_mode=1, mouse_pos=mouse))
        r = frame(mv, 0)
        r.update(config="C4", size="1920x1080", function=fn.name, variant=label)
        result["configs"].append(r)
        print(json.dumps(r), flush=True)
    # outline mode searches the exponents of the previous overlay pass (src/main_view.rs:838-846)
    mv.update_settings(make_settings(overlay_mode=1))
    r = frame(mv, capi.RENDER_OUTLINE)
    r.update(config="C4", size="1920x1080", function=fn.name, variant="outline mode (8 searches over exponents >= 0.8)",
             sample_positions=mv.read(capi.BUF_SAMPLE_POS).round(5).tolist())
    result["configs"].append(r)
    print(json.dumps(r), flush=True)

os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
json.dump(result, open(out_path, "w"), indent=1)
