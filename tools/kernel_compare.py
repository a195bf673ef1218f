"""Developer tool (run under gpurun): times the march kernels against each other on the C3 frame.

    python tools/kernel_compare.py [W H] [--fn N] [--variants "kernel:block:pair_min_blocks:max_regs,..."]
Writes one JSON line per variant to stdout (and gpurun_out/kernel_compare.jsonl).
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linon_b200 import _capi as capi  # noqa: E402
from linon_b200.main_view import MainView, PredefinedFunction, default_scene, make_settings  # noqa: E402

args = sys.argv[1:]
W, H = 3840, 2160
fn = PredefinedFunction.MirageSphericalSigmoid
variants = "0:256:0:0,2:256:0:0,2:128:0:0,2:256:3:0,2:128:5:0"
flags = capi.RENDER_ENHANCE
reps = 3
body_file = None
i = 0
pos = []
while i < len(args):
    if args[i] == "--fn":
        fn = PredefinedFunction(int(args[i + 1])); i += 2
    elif args[i] == "--variants":
        variants = args[i + 1]; i += 2
    elif args[i] == "--body-file":
        body_file = args[i + 1]; i += 2
    elif args[i] == "--reps":
        reps = int(args[i + 1]); i += 2
    elif args[i] == "--default-step":
        flags = 0; i += 1
    else:
        pos.append(int(args[i])); i += 1
if len(pos) == 2:
    W, H = pos

verts, faces, center = default_scene()
os.makedirs("gpurun_out", exist_ok=True)
out = open("gpurun_out/kernel_compare.jsonl", "a")
from linon_b200 import parity  # noqa: E402

with MainView(verts, faces, center, W, H) as mv:
    ref = None
    if body_file:  # a custom body: report its agreement with the predefined function it restates
        mv.reload_shader(fn.to_code())
        mv.render_flags(flags | capi.RENDER_NO_SAMPLE_RAYS)
        mv.sync()
        ref = dict(hit_id=mv.hit_ids(), rgba8=mv.ray_casting_texture(), endpoint=mv.mapping_texture())
    mv.reload_shader(open(body_file).read() if body_file else fn.to_code())
    mv.update_settings(make_settings(overlay_mode=1))
    for var in variants.split(","):
        f = [int(x) for x in var.split(":")] + [0, 0, 0, 1]
        kernel, block, pmb, regs, spec = f[0], f[1], f[2], f[3], (f[4] if len(var.split(":")) > 4 else 1)
        mv.set_option(capi.OPT_SPECIALIZE, spec)
        refill = f[5] if len(var.split(":")) > 5 else 1
        ctas = f[6] if len(var.split(":")) > 6 else 0
        mv.set_option(capi.OPT_REFILL_MIN, refill)
        mv.set_option(capi.OPT_CTAS_PER_SM, ctas)
        mv.set_option(capi.OPT_KERNEL, kernel)
        mv.set_option(capi.OPT_BLOCK, block)
        mv.set_option(capi.OPT_PAIR_MIN_BLOCKS, pmb)
        mv.set_option(capi.OPT_MAX_REGS, regs)
        best = None
        for rep in range(reps):
            mv.render_flags(flags | capi.RENDER_NO_SAMPLE_RAYS)
            mv.sync()
            t = mv.timings()
            if best is None or t.march_ms < best.march_ms:
                best = t
        c = mv.counters()
        line = dict(fn=fn.name, W=W, H=H, kernel=best.march_kernel, block=block, pair_min_blocks=pmb, max_regs=regs,
                    regs=best.march_regs, grid=best.march_grid, specialize=spec, refill_min=refill, ctas_per_sm=ctas, march_ms=round(best.march_ms, 3),
                    mrays_s=round(W * H / best.march_ms / 1e3, 2),
                    gsteps_s=round((c.accepted_steps + c.rejected_steps) / best.march_ms / 1e6, 2))
        if ref is not None:
            got = dict(hit_id=mv.hit_ids(), rgba8=mv.ray_casting_texture(), endpoint=mv.mapping_texture())
            m = parity.compare_frames(got, ref)
            line.update(body=os.path.basename(body_file), vs_predefined={k: (round(float(v), 6) if isinstance(v, float) else v) for k, v in m.items()})
        print(json.dumps(line), flush=True)
        out.write(json.dumps(line) + "\n")
