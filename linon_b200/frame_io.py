"""On-disk format of offline frames (the reference has no export at all, F18).

One frame = a directory entry set written by `write_frame`:
    <stem>.ppm          final image (`texture`, after the overlay), binary PPM (P6), row 0 = top
    <stem>.json         sidecar: size, camera uniform, settings, field-function text + SHA-256, render flags, step
                        counters, kernel timings - everything needed to re-render the frame bit for bit
    <stem>.endpoint.f32 (optional) `mapping_texture`: H x W x 4 float32 little endian
    <stem>.hit.i32      (optional) first-hit face ids: H x W int32
    <stem>.samples.f32  (optional) `ray_samples_buffer`: 800 x {vec4 position, vec4 colour} float32, the vertex buffer
                        the reference view draws (src/ray_samples.rs:8-27)
`ray_sample_indices()` restates the index topology the reference view draws that buffer with
(src/ray_samples.rs:29-49: 4 sides x 99 segments x 4 triangles, u16).
"""
from __future__ import annotations

import hashlib
import json
import os

import numpy as np

SAMPLES_PER_POINT = 100  # src/ray_samples.rs:6


def ray_sample_indices() -> np.ndarray:
    """create_indices(), src/ray_samples.rs:42-49 -> [1584, 3] uint16."""
    out = []

    def side(a, b, c):  # push_indices_for_side, :29-40
        for i in range(SAMPLES_PER_POINT - 1):
            ai, bi, ci = a * SAMPLES_PER_POINT + i, b * SAMPLES_PER_POINT + i, c * SAMPLES_PER_POINT + i
            out.extend([[ai, bi + 1, ai + 1], [ai, bi, bi + 1], [bi, ci, bi + 1], [ci, ci + 1, bi + 1]])

    side(0, 1, 2)
    side(2, 3, 4)
    side(4, 5, 6)
    side(6, 7, 0)
    return np.asarray(out, np.uint16)


def write_ppm(path: str, rgba8: np.ndarray):
    h, w = rgba8.shape[:2]
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (w, h))
        f.write(np.ascontiguousarray(rgba8[..., :3]).tobytes())


def read_ppm(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        assert f.readline().strip() == b"P6"
        w, h = (int(x) for x in f.readline().split())
        assert int(f.readline()) == 255
        return np.frombuffer(f.read(w * h * 3), np.uint8).reshape(h, w, 3)


def camera_dict(cam) -> dict:
    return dict(origin=[float(x) for x in cam.origin], view_direction=[float(x) for x in cam.view_direction],
                up=[float(x) for x in cam.up])


def settings_dict(st) -> dict:
    return dict(field_weight=float(st.field_weight), mouse_pos=[float(st.mouse_pos[0]), float(st.mouse_pos[1])],
                overlay_mode=int(st.overlay_mode), central_difference_delta=int(st.central_difference_delta),
                lyapunov_scaling=float(st.lyapunov_scaling))


def write_frame(stem: str, final_rgba8: np.ndarray, camera, settings, field_function: str, flags: int, counters=None,
                timings=None, endpoint=None, hit_id=None, ray_samples=None, extra=None) -> dict:
    os.makedirs(os.path.dirname(stem) or ".", exist_ok=True)
    write_ppm(stem + ".ppm", final_rgba8)
    h, w = final_rgba8.shape[:2]
    side = dict(format="linon_b200 frame v1", width=w, height=h, camera=camera_dict(camera), settings=settings_dict(settings),
                field_function=field_function, field_function_sha256=hashlib.sha256(field_function.encode()).hexdigest(),
                render_flags=int(flags), files=dict(image=os.path.basename(stem) + ".ppm"))
    if counters is not None:
        side["counters"] = {k: int(getattr(counters, k)) for k, _ in counters._fields_}
    if timings is not None:
        side["timings"] = {k: (float(getattr(timings, k)) if "ms" in k else int(getattr(timings, k))) for k, _ in timings._fields_}
    for name, arr, suffix, dt in (("endpoint", endpoint, ".endpoint.f32", "<f4"), ("hit_id", hit_id, ".hit.i32", "<i4"),
                                  ("ray_samples", ray_samples, ".samples.f32", "<f4")):
        if arr is not None:
            np.ascontiguousarray(arr).astype(dt).tofile(stem + suffix)
            side["files"][name] = os.path.basename(stem) + suffix
    if extra:
        side.update(extra)
    with open(stem + ".json", "w") as f:
        json.dump(side, f, indent=1)
    return side
