"""Generates tests/golden/march_edits.npz: frames of the REFERENCE's shader with its march-level constants edited, or with an
edited shading block of ray_color (tests/march_edit_cases.py: CASES, SHADE_CASES) - the edit applied to the text of /root/reference/src/main_view.wgsl, the MirageSpherical body
spliced in like `with_field_function` (src/main_view.rs:852-861), transpiled by oracle/wgsl_ref and compiled with gcc.
64x48, default camera and settings, as the application renders (adaptive sampling on unless the edit says otherwise).
Run (needs /root/reference): python tests/golden/make_march_edits.py"""
import ctypes as C
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import march_edit_cases as mc  # noqa: E402
from oracle import cpu_oracle as co  # noqa: E402
from oracle import np_oracle as npo  # noqa: E402
from oracle import wgsl_reference as wr  # noqa: E402
from oracle.wgsl_ref import build_ref  # noqa: E402


def main():
    W, H = mc.W, mc.H
    verts, faces, center = npo.default_scene()
    verts, faces = np.ascontiguousarray(verts, np.float32), np.ascontiguousarray(faces, np.uint32)
    cam = npo.camera_uniform(npo.default_camera(center, W, H))
    st = npo.default_settings()
    body = build_ref.predefined_bodies(open(os.path.join(build_ref.REFERENCE, "src", "functions.rs"), encoding="utf-8").read())["MirageSpherical"]
    out = {}
    os.makedirs(os.path.join(HERE, mc.SHADE_MODULE_DIR), exist_ok=True)
    for name, edits in mc.SHADE_CASES.items():  # scratch copies for a GPU box without the mount (git-ignored: reference source)
        with open(os.path.join(HERE, mc.SHADE_MODULE_DIR, name + ".wgsl"), "w", encoding="utf-8") as f:
            f.write(mc.edited_reference_module(build_ref.REFERENCE, edits))
    for name, edits in list(mc.CASES.items()) + list(mc.SHADE_CASES.items()):
        with tempfile.TemporaryDirectory() as d:
            lib = C.CDLL(build_ref.build_custom(body, d, module_edits=edits))
            f = wr.RefFrame(W, H, co.make_camera(cam), co.make_settings(st), verts.shape[0], faces.shape[0], verts.ctypes.data,
                            faces.ctypes.data, None, None, 0)
            rgba8, endpoint = np.zeros((H, W, 4), np.uint8), np.zeros((H, W, 4), np.float32)
            samples = np.zeros((800, 8), np.float32)
            o = wr.RefOutputs(None, rgba8.ctypes.data, endpoint.ctypes.data, samples.ctypes.data)
            assert lib.wgslref_custom_main_view(C.byref(f), C.byref(o)) == 0
        out[f"rgba8_{name}"], out[f"endpoint_{name}"] = rgba8, endpoint
        print(name, "hit fraction", (rgba8[..., :3].max(axis=-1) > 0).mean(), "alpha values", np.unique(rgba8[..., 3])[:6])
    np.savez_compressed(os.path.join(HERE, "march_edits.npz"), names=np.array(list(mc.CASES) + list(mc.SHADE_CASES)), **out)


if __name__ == "__main__":
    main()
