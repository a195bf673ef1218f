"""Generates tests/golden/edited_mirages.npz: frames of mirage bodies with edited constants (tests/edited_bodies.py) as the
REFERENCE's shader renders them - the body spliced into /root/reference/src/main_view.wgsl like `with_field_function`
(src/main_view.rs:852-861), adaptive sampling switched off by the reference's own string replacement, transpiled by
oracle/wgsl_ref and compiled with gcc.  96x72, default camera and settings.  Stored: rgba8, endpoint, and the
high-Lyapunov mask of the reference's overlay on those endpoints.  Run (needs /root/reference): python tests/golden/make_edited_mirages.py
"""
import ctypes as C
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import edited_bodies  # noqa: E402
from oracle import cpu_oracle as co  # noqa: E402
from oracle import np_oracle as npo  # noqa: E402
from oracle import wgsl_reference as wr  # noqa: E402
from oracle.wgsl_ref import build_ref  # noqa: E402

W, H = 96, 72


def render(lib, verts, faces, cam, st):
    f = wr.RefFrame(W, H, co.make_camera(cam), co.make_settings(st), verts.shape[0], faces.shape[0], verts.ctypes.data,
                    faces.ctypes.data, None, None, 0)
    out = dict(rgba8=np.zeros((H, W, 4), np.uint8), endpoint=np.zeros((H, W, 4), np.float32), ray_samples=np.zeros((800, 8), np.float32))
    o = wr.RefOutputs(None, out["rgba8"].ctypes.data, out["endpoint"].ctypes.data, out["ray_samples"].ctypes.data)
    assert lib.wgslref_custom_main_view(C.byref(f), C.byref(o)) == 0
    return out


def main():
    verts, faces, center = npo.default_scene()
    verts = np.ascontiguousarray(verts, np.float32)
    faces = np.ascontiguousarray(faces, np.uint32)
    cam = npo.camera_uniform(npo.default_camera(center, W, H))
    st = npo.default_settings()
    out = {}
    names = []
    for k, (name, (kind, body)) in enumerate(edited_bodies.bodies().items()):
        with tempfile.TemporaryDirectory() as d:
            lib = C.CDLL(build_ref.build_custom(body, d, variant="adaptive_off"))
            fr = render(lib, verts, faces, cam, st)
        _, exps = co.overlay(fr["rgba8"], fr["endpoint"], dict(st, overlay_mode=1))
        out[f"rgba8_{k}"], out[f"endpoint_{k}"], out[f"high_{k}"] = fr["rgba8"], fr["endpoint"], exps >= 0.8
        names.append(name)
        print(name, "hit fraction", (fr["rgba8"][..., :3].max(axis=-1) > 0).mean(), "high lyapunov", (exps >= 0.8).mean())
    np.savez_compressed(os.path.join(HERE, "edited_mirages.npz"), names=np.array(names), **out)


if __name__ == "__main__":
    main()
