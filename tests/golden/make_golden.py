"""Generates tests/golden/*.npz: outputs of the CPU oracle (oracle/linon_oracle.cpp) for small frames.

The reference ships no golden vectors and its Rust/wgpu host cannot run here.  What CAN run here is its
shader: oracle/_ref is main_view.wgsl + the functions.rs bodies transpiled mechanically to C++ (oracle/wgsl_ref/).
Every fixture is written only if that library reproduces the oracle's rgba8 and endpoint bytes exactly, so the colour
and endpoint planes are outputs of the reference's own shader source; hit ids, step counters and the exclusion masks
come from the oracle (the shader has no such outputs).  tests/test_golden_cpu.py re-checks the oracle against the
fixtures, tests/test_wgsl_reference_cpu.py re-checks the transpiled reference against them, the GPU tests check the
CUDA path against them without running either.  Run (needs /root/reference): python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import cpu_oracle as co  # noqa: E402
from oracle import np_oracle as npo  # noqa: E402
from oracle import wgsl_reference as wr  # noqa: E402

W, H = 96, 72
NAMES = {1: "MirageSpherical", 2: "MiragePlane", 3: "MirageSphericalSigmoid", 4: "MiragePlaneSigmoid", 5: "TranslationX",
         6: "TranslationZ", 7: "Rotation", 8: "LorenzAttractor", 9: "RoesslerAttractor"}


def main():
    verts, faces, center = npo.default_scene()
    cam = npo.camera_uniform(npo.default_camera(center, W, H))
    st = npo.default_settings()
    frames = []
    for fn, name in NAMES.items():
        for adaptive in ((False, True) if fn <= 4 else (False,)):
            out = co.march(verts, faces, cam, st, fn, W, H, h_initial=0.1, adaptive=adaptive)
            ref = wr.march(verts, faces, cam, st, fn, W, H, h_initial=0.1, adaptive=adaptive)
            assert np.array_equal(ref["rgba8"], out["rgba8"]), name
            assert np.array_equal(ref["endpoint"].view(np.uint32), out["endpoint"].view(np.uint32)), name
            _, exps = co.overlay(out["rgba8"], out["endpoint"], dict(st, overlay_mode=1))
            undecided = np.zeros((H, W), bool)
            if adaptive:  # pixels the oracle itself cannot decide in adaptive mode
                from linon_b200 import parity  # pure numpy bookkeeping
                alt64 = co.march(verts, faces, cam, st, fn, W, H, h_initial=0.1, adaptive=True, use_fp64=True)
                st_ulp = dict(st, field_weight=float(np.nextafter(np.float32(1.0), np.float32(0))))
                alt_ulp = co.march(verts, faces, cam, st_ulp, fn, W, H, h_initial=0.1, adaptive=True)
                undecided = parity.disagreement_mask(alt64, out) | parity.disagreement_mask(alt_ulp, out)
            fname = f"{name}_{'adaptive' if adaptive else 'fixed'}.npz"
            np.savez_compressed(os.path.join(HERE, fname), hit_id=out["hit_id"].astype(np.int8), rgba8=out["rgba8"],
                                endpoint=out["endpoint"], high_lyapunov=exps >= 0.8, undecided=undecided,
                                accepted=out["accepted"].astype(np.int16), rejected=out["rejected"].astype(np.int16))
            frames.append(dict(file=fname, function=name, function_id=fn, adaptive=adaptive))
    json.dump(dict(pinned_by="oracle/_ref (transpiled reference WGSL)", reference_sources=wr.info()["sources"],
                   width=W, height=H, h_initial=0.1, camera={k: [float(x) for x in v] for k, v in cam.items()},
                   settings=st, frames=frames), open(os.path.join(HERE, "index.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
