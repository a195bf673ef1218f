"""C4 outline mode at 1920x1080: how many exponents reach the threshold, the eight positions of the device search and the
oracle's serial scans over the same exponents (developer probe; prints one JSON line)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from linon_b200 import _capi as capi  # noqa: E402
from linon_b200.main_view import MainView, PredefinedFunction, default_scene, make_settings  # noqa: E402

W, H = 1920, 1080
verts, faces, center = default_scene()
out = {}
with MainView(verts, faces, center, W, H) as mv:
    mv.reload_shader(PredefinedFunction.MirageSphericalSigmoid.to_code())
    st = make_settings(overlay_mode=1)
    mv.update_settings(st)
    for i in range(3):
        mv.render_flags(capi.RENDER_OUTLINE)
        mv.sync()
        ex = mv.exponents()
        pos = np.asarray(mv.read(capi.BUF_SAMPLE_POS))
        out[f"frame{i}"] = dict(n_ge=int((ex >= 0.8).sum()), n_nan=int(np.isnan(ex).sum()), n_inf=int(np.isinf(ex).sum()),
                                 max=float(np.nanmax(ex)), pos=pos.round(4).tolist())
    try:
        from oracle import cpu_oracle as co
        co.build()
        cam = dict(origin=np.array([0, 0, 2], np.float32), view_direction=np.array([0, 0, -1], np.float32), up=np.array([0, 1, 0], np.float32))
        sd = dict(field_weight=1.0, mouse_x=0.5, mouse_y=0.6, overlay_mode=1, central_difference_delta=1, lyapunov_scaling=50.0)
        ref = co.sample_positions(cam, sd, W, H, outline=True, exponents=ex)  # positions depend on the exponents only
        out["oracle_pos_from_last_exponents"] = np.asarray(ref).round(4).tolist()
    except Exception as e:  # noqa: BLE001
        out["oracle_error"] = f"{type(e).__name__}: {e}"[:300]
print(json.dumps(out))
