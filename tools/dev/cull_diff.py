"""Developer probe: frames of the tile kernel with culling on and off - which pixels differ and how.
    python tools/dev/cull_diff.py [PKG_ROOT]     PKG_ROOT: directory holding the linon_b200 package to load (default: the repo)"""
import json
import os
import sys

import numpy as np

ROOT = os.path.abspath(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1] != "." else os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from linon_b200 import _capi as capi  # noqa: E402
from linon_b200.main_view import MainView, PredefinedFunction, default_scene, make_settings  # noqa: E402

verts, faces, center = default_scene()
out = dict(pkg=ROOT)
SIZES = [tuple(int(t) for t in a.split("x")) for a in sys.argv[2:]] or [(160, 120), (320, 240), (640, 480)]
if os.environ.get("CULL_DIFF_DIRTY"):
    import ctypes
    rt = ctypes.CDLL("libcudart.so")
    p = ctypes.c_void_p()
    rt.cudaMalloc(ctypes.byref(p), ctypes.c_size_t(3 << 30))
    rt.cudaMemset(p, 0x5A, ctypes.c_size_t(3 << 30))
    rt.cudaDeviceSynchronize()
    rt.cudaFree(p)  # the next allocations recycle memory that holds garbage
for (W, H) in SIZES:
    with MainView(verts, faces, center, W, H) as mv:
        mv.reload_shader(PredefinedFunction.MirageSphericalSigmoid.to_code())
        mv.update_settings(make_settings())
        frames = {}
        for name, opts in (("cull1", {capi.OPT_CULL: 1}), ("cull0", {capi.OPT_CULL: 0}), ("cull1_quiet0", {capi.OPT_CULL: 1, capi.OPT_QUIET_ZONE: 0}),
                           ("cull0_quiet0", {capi.OPT_CULL: 0, capi.OPT_QUIET_ZONE: 0})):
            mv.set_option(capi.OPT_QUIET_ZONE, 1)
            for k, v in opts.items():
                mv.set_option(k, v)
            mv.render_flags(capi.RENDER_ENHANCE)
            mv.sync()
            frames[name] = (mv.hit_ids().copy(), mv.ray_casting_texture().copy(), mv.mapping_texture().copy(), mv.march_cubin_hash())
        rec = dict(hashes={k: v[3] for k, v in frames.items()})
        for a, b in (("cull1", "cull0"), ("cull1_quiet0", "cull0_quiet0"), ("cull1", "cull1_quiet0")):
            ha, ca, ea, _ = frames[a]
            hb, cb, eb, _ = frames[b]
            d = np.argwhere(ha != hb)
            rec[f"{a}_vs_{b}"] = dict(hit_diff=int(len(d)), colour_diff=int((ca != cb).any(axis=-1).sum()),
                                      endpoint_diff=int((ea.view(np.uint32) != eb.view(np.uint32)).any(axis=-1).sum()),
                                      pixels=[dict(y=int(y), x=int(x), hit_a=int(ha[y, x]), hit_b=int(hb[y, x]), rgba_a=ca[y, x].tolist(),
                                                   rgba_b=cb[y, x].tolist()) for y, x in d[:6]])
        out[f"{W}x{H}"] = rec
print(json.dumps(out))
