"""ctypes binding of oracle/_ref/liblinon_wgsl_ref.so — the reference's own WGSL, transpiled mechanically to C++
(oracle/wgsl_ref/) and compiled with gcc.  TEST INFRASTRUCTURE ONLY (oracle/README.md).

Same call shapes as oracle/cpu_oracle.py so a test can run both on the same inputs.  What this library cannot give
(the shader has no such outputs): hit ids and per-ray step counters.
"""
from __future__ import annotations

import ctypes as C
import json
import os

import numpy as np

from oracle.cpu_oracle import CameraPOD, SettingsPOD, make_camera, make_settings
from oracle.wgsl_ref import build_ref

VARIANT_DEFAULT, VARIANT_ENHANCE, VARIANT_ADAPTIVE_OFF, VARIANT_OUTLINE = 0, 1, 2, 3


class RefFrame(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("camera", CameraPOD), ("settings", SettingsPOD),
                ("n_verts", C.c_int32), ("n_faces", C.c_int32), ("verts", C.c_void_p), ("faces", C.c_void_p),
                ("exponents", C.c_void_p), ("mask", C.c_void_p), ("n_threads", C.c_int32)]


class RefOutputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("color_f32x4", "rgba8", "endpoint_f32x4", "ray_samples")]


def available() -> bool:
    """True when the library can be had: the reference is mounted (build container) or a prebuilt .so travelled."""
    return build_ref.available() or os.path.exists(build_ref.LIB)


def build(force: bool = False):
    return build_ref.build(force)


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = build()
        if path is None:
            raise RuntimeError("oracle/_ref/liblinon_wgsl_ref.so is not built and /root/reference is not mounted")
        _lib = C.CDLL(path)
        _lib.wgslref_info.restype = C.c_char_p
    return _lib


def info() -> dict:
    return json.loads(lib().wgslref_info().decode())


def variant_of(h_initial=0.1, adaptive=True, outline=False) -> int:
    """The reference toggles one variant at a time (src/main_view.rs:752-829)."""
    if outline:
        assert adaptive and abs(h_initial - 0.1) < 1e-9
        return VARIANT_OUTLINE
    if abs(h_initial - 0.001) < 1e-9:
        assert adaptive, "Enhance keeps adaptive_sampling = true in the reference"
        return VARIANT_ENHANCE
    assert abs(h_initial - 0.1) < 1e-9
    return VARIANT_DEFAULT if adaptive else VARIANT_ADAPTIVE_OFF


def main_view(verts, faces, camera, settings, fn_id, W, H, variant=VARIANT_DEFAULT, mask=None, exponents=None,
              n_threads=0, want_color_f32=False):
    """One dispatch of the transpiled main_view entry point over the reference's workgroup grid
    (src/main_view.rs:844).  mask: (H, W) bool in texture space — invocations to run (None = all)."""
    verts = np.ascontiguousarray(verts, dtype=np.float32)
    faces = np.ascontiguousarray(faces, dtype=np.uint32)
    m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
    ex = None if exponents is None else np.ascontiguousarray(exponents, dtype=np.float32)
    f = RefFrame(W, H, make_camera(camera), make_settings(settings), verts.shape[0], faces.shape[0], verts.ctypes.data,
                 faces.ctypes.data, None if ex is None else ex.ctypes.data, None if m is None else m.ctypes.data, n_threads)
    out = dict(rgba8=np.zeros((H, W, 4), np.uint8), endpoint=np.zeros((H, W, 4), np.float32),
               ray_samples=np.zeros((800, 8), np.float32))
    if want_color_f32:
        out["color"] = np.zeros((H, W, 4), np.float32)
    o = RefOutputs(out["color"].ctypes.data if want_color_f32 else None, out["rgba8"].ctypes.data,
                   out["endpoint"].ctypes.data, out["ray_samples"].ctypes.data)
    rc = lib().wgslref_main_view(int(fn_id), int(variant), C.byref(f), C.byref(o))
    if rc != 0:
        raise RuntimeError(f"wgslref_main_view rc={rc}")
    return out


def march(verts, faces, camera, settings, fn_id, W, H, h_initial=0.1, adaptive=True, rect=None, x_step=1, y_step=1,
          n_threads=0, want_color_f32=False):
    """cpu_oracle.march's signature (pixel sample = texture-space rectangle + strides)."""
    x0, x1, y0, y1 = rect if rect is not None else (0, W, 0, H)
    mask = np.zeros((H, W), bool)
    mask[y0:y1:y_step, x0:x1:x_step] = True
    return main_view(verts, faces, camera, settings, fn_id, W, H, variant_of(h_initial, adaptive), mask=mask,
                     n_threads=n_threads, want_color_f32=want_color_f32)


def sample_ray_mask(W, H):
    """The eight invocations that also trace the sample rays: gid.x < 8, gid.y == 0 (src/main_view.wgsl:377),
    i.e. the first eight pixels of the bottom texture row."""
    mask = np.zeros((H, W), bool)
    mask[H - 1, :min(8, W)] = True
    return mask


def overlay(raycast_rgba8, endpoint, settings, exponents=None, n_threads=0):
    H, W = raycast_rgba8.shape[:2]
    raycast_rgba8 = np.ascontiguousarray(raycast_rgba8, dtype=np.uint8)
    endpoint = np.ascontiguousarray(endpoint, dtype=np.float32)
    final = np.zeros_like(raycast_rgba8)
    exps = np.zeros((H, W), np.float32) if exponents is None else np.array(exponents, dtype=np.float32, copy=True)
    st = make_settings(settings)
    rc = lib().wgslref_overlay(W, H, C.byref(st), C.c_void_p(raycast_rgba8.ctypes.data), C.c_void_p(endpoint.ctypes.data),
                               C.c_void_p(final.ctypes.data), C.c_void_p(exps.ctypes.data), n_threads)
    if rc != 0:
        raise RuntimeError(f"wgslref_overlay rc={rc}")
    return final, exps


def field(fn_id, settings, p_prev, p, v0, v, t):
    arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (p_prev, p, v0, v, t)]
    n = arrs[4].shape[0]
    out = np.zeros((n, 3), np.float32)
    st = make_settings(settings)
    rc = lib().wgslref_field(int(fn_id), C.byref(st), n, *[C.c_void_p(a.ctypes.data) for a in arrs], C.c_void_p(out.ctypes.data))
    if rc != 0:
        raise RuntimeError(f"wgslref_field rc={rc}")
    return out


def helper(name, n, v=None, nrm=None, p0=None, a=None, b=None, c=None):
    from oracle.cpu_oracle import HELPERS
    z3 = np.zeros((n, 3), np.float32)
    z1 = np.zeros(n, np.float32)
    arrs = [np.ascontiguousarray(x if x is not None else d, dtype=np.float32)
            for x, d in ((v, z3), (nrm, z3), (p0, z3), (a, z1), (b, z1), (c, z1))]
    out = np.zeros((n, 3), np.float32)
    lib().wgslref_helper(HELPERS[name], n, *[C.c_void_p(x.ctypes.data) for x in arrs], C.c_void_p(out.ctypes.data))
    return out


def hit_triangle(tri, origin, direction):
    tri = np.ascontiguousarray(tri, dtype=np.float32).reshape(-1, 9)
    origin = np.ascontiguousarray(origin, dtype=np.float32)
    direction = np.ascontiguousarray(direction, dtype=np.float32)
    out = np.zeros(tri.shape[0], np.float32)
    lib().wgslref_hit_triangle(tri.shape[0], C.c_void_p(tri.ctypes.data), C.c_void_p(origin.ctypes.data),
                               C.c_void_p(direction.ctypes.data), C.c_void_p(out.ctypes.data))
    return out


def vec2squareroots(a01):
    a01 = np.ascontiguousarray(a01, dtype=np.float32).reshape(-1, 2)
    out = np.zeros_like(a01)
    lib().wgslref_vec2squareroots(a01.shape[0], C.c_void_p(a01.ctypes.data), C.c_void_p(out.ctypes.data))
    return out


def num_threads() -> int:
    return lib().wgslref_num_threads()
