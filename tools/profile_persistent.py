"""Developer tool: section timers of the persistent march kernel vs the tile kernel (run under gpurun)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linon_b200 import _capi as capi  # noqa: E402
from linon_b200.main_view import MainView, PredefinedFunction, default_scene  # noqa: E402

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (960, 540)
mode = sys.argv[3] if len(sys.argv) > 3 else "sweep"
verts, faces, center = default_scene()
mv = MainView(verts, faces, center, W, H)
mv.reload_shader(PredefinedFunction.MirageSphericalSigmoid.to_code())
flags = capi.RENDER_ENHANCE | capi.RENDER_NO_SAMPLE_RAYS


def run(kernel, profile=0, refill=1, ctas=0, math=1, block=256, reps=2):
    mv.set_option(capi.OPT_KERNEL, kernel)
    mv.set_option(capi.OPT_PROFILE, profile)
    mv.set_option(capi.OPT_REFILL_MIN, refill)
    mv.set_option(capi.OPT_CTAS_PER_SM, ctas)
    mv.set_option(capi.OPT_MATH, math)
    mv.set_option(capi.OPT_BLOCK, block)
    for _ in range(reps):
        mv.render_flags(flags)
        mv.sync()
    t = mv.timings()
    p = mv.debug_profile()
    rec = dict(kernel=kernel, profile=profile, refill=refill, ctas=ctas, math=math, block=block, march_ms=round(t.march_ms, 2),
               grid=t.march_grid, regs=t.march_regs, accepted=int(p[0]), prehit=int(p[1]))
    if profile:
        warps = t.march_grid * block // 32
        rec.update(refill_cyc_per_warp=int(p[3]) // warps, step_cyc_per_warp=int(p[4]) // warps,
                   push_cyc_per_warp=int(p[5]) // warps, iters_per_warp=int(p[6]) // warps,
                   lane_util=round(int(p[7]) / max(int(p[6]), 1) / 32, 4), polls_per_warp=int(p[8]) // warps,
                   life_cyc_per_warp=int(p[9]) // warps, cyc_per_iter=round(int(p[4]) / max(int(p[6]), 1), 1))
    print(json.dumps(rec), flush=True)


if mode == "ncu":
    run(0, reps=1)
    run(1, reps=1)
else:
    run(0)
    run(1)
    run(1, profile=1)
    run(1, profile=1, refill=8)
    run(1, profile=1, ctas=2)
    run(1, profile=1, ctas=1)
    run(1, profile=1, block=128)
    run(0, block=128)
    run(0, block=64)
    run(1, math=2, profile=1)
    run(0, math=2)
