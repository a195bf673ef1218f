"""Developer / evidence tool (run under gpurun): the march on OBJ-sized meshes (hierarchy path) vs the Cornell box.

    python tools/big_mesh.py [--lat 192 --lon 384] [W H]
Builds the Cornell box plus a UV sphere of 2 * lat * lon triangles (OBJ text -> linon_host_load_obj), renders the default
and the Enhance frame and prints timings and traversal counters.
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from linon_b200 import _capi as capi  # noqa: E402
from linon_b200.main_view import MainView, PredefinedFunction, default_scene, load_obj, make_settings  # noqa: E402

args = sys.argv[1:]
lat, lon, W, H = 192, 384, 1920, 1080
pos = []
i = 0
while i < len(args):
    if args[i] == "--lat":
        lat = int(args[i + 1]); i += 2
    elif args[i] == "--lon":
        lon = int(args[i + 1]); i += 2
    else:
        pos.append(int(args[i])); i += 1
if len(pos) == 2:
    W, H = pos


def sphere_lines(n_lat, n_lon, first_index, radius=120.0, centre=(-278.0, 274.4, -279.6)):
    out = []
    for a in range(n_lat + 1):
        th = np.pi * a / n_lat
        for b in range(n_lon):
            ph = 2 * np.pi * b / n_lon
            out.append("v %.6f %.6f %.6f" % (centre[0] + radius * np.sin(th) * np.cos(ph), centre[1] + radius * np.cos(th),
                                             centre[2] + radius * np.sin(th) * np.sin(ph)))
    for a in range(n_lat):
        for b in range(n_lon):
            p, q = a * n_lon + b + first_index, a * n_lon + (b + 1) % n_lon + first_index
            out.append(f"f {p} {q} {q + n_lon} {p + n_lon}")
    return out


bverts, bfaces, _ = default_scene()
mn, mx = -559.2, 548.8
box = (bverts.astype(np.float64) + 1.0) / 2.0 * (mx - mn) + mn  # back to box units (x, z negated like Application::new)
lines = ["v %.6f %.6f %.6f" % tuple(p) for p in box] + ["f %d %d %d" % tuple(int(k) + 1 for k in f) for f in bfaces]
lines += sphere_lines(lat, lon, box.shape[0] + 1)
verts, faces, center = load_obj("\n".join(lines) + "\n")
out = []
for name, (v, f, c) in (("Cornell box (36 faces, shared-memory plane groups)", default_scene()),
                        (f"box + sphere ({faces.shape[0]} faces, hierarchy)", (verts, faces, center))):
    with MainView(v, f, c, W, H) as mv:
        mv.reload_shader(PredefinedFunction.MirageSphericalSigmoid.to_code())
        mv.update_settings(make_settings(overlay_mode=1))
        for label, flags in (("default step, adaptive", 0), ("Enhance", capi.RENDER_ENHANCE)):
            best = None
            for _ in range(2):
                mv.render_flags(flags | capi.RENDER_NO_SAMPLE_RAYS)
                mv.sync()
                t = mv.timings()
                best = t if best is None or t.march_ms < best.march_ms else best
            cnt = mv.counters()
            rec = dict(mesh=name, variant=label, size=f"{W}x{H}", march_ms=round(best.march_ms, 3), regs=best.march_regs,
                       mrays_s=round(W * H / best.march_ms / 1e3, 2),
                       gsteps_s=round((cnt.accepted_steps + cnt.rejected_steps) / best.march_ms / 1e6, 2),
                       face_tests_per_prehit_step=round(cnt.face_tests / max(cnt.prehit_steps, 1), 4),
                       checks_per_prehit_step=round(cnt.culling_checks / max(cnt.prehit_steps, 1), 4),
                       hit_fraction=round(float((mv.hit_ids() >= 0).mean()), 4))
            print(json.dumps(rec), flush=True)
            out.append(rec)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/big_mesh.json", "w"), indent=1)
