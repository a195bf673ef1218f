"""ctypes binding of oracle/linon_oracle.cpp (TEST INFRASTRUCTURE ONLY — see oracle/README.md).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liblinon_oracle.so")


class CameraPOD(C.Structure):  # src/main_view.rs:12-19
    _fields_ = [("origin", C.c_float * 4), ("view_direction", C.c_float * 4), ("up", C.c_float * 4),
                ("view_matrix", C.c_float * 16)]


class SettingsPOD(C.Structure):  # src/main_view.rs:43-51
    _fields_ = [("field_weight", C.c_float), ("mouse_x", C.c_float), ("mouse_y", C.c_float),
                ("overlay_mode", C.c_int32), ("central_difference_delta", C.c_int32), ("lyapunov_scaling", C.c_float)]


class OracleFrame(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("camera", CameraPOD), ("settings", SettingsPOD),
                ("h_initial", C.c_float), ("adaptive", C.c_int32), ("function_id", C.c_int32),
                ("n_verts", C.c_int32), ("n_faces", C.c_int32), ("verts", C.c_void_p), ("faces", C.c_void_p),
                ("x0", C.c_int32), ("x1", C.c_int32), ("y0", C.c_int32), ("y1", C.c_int32),
                ("x_step", C.c_int32), ("y_step", C.c_int32), ("use_fp64", C.c_int32), ("n_threads", C.c_int32)]


class OracleOutputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("color_f32x4", "rgba8", "endpoint_f32x4", "hit_id", "accepted", "rejected", "prehit")]


def build(force: bool = False) -> str:
    """Compile the oracle with its Makefile (building the checker is not using it)."""
    src = os.path.join(_HERE, "linon_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.oracle_march.restype = C.c_int
        _lib.oracle_num_threads.restype = C.c_int
    return _lib


def make_settings(d) -> SettingsPOD:
    return SettingsPOD(float(d["field_weight"]), float(d["mouse_x"]), float(d["mouse_y"]), int(d["overlay_mode"]),
                       int(d["central_difference_delta"]), float(d["lyapunov_scaling"]))


def make_camera(cam) -> CameraPOD:
    c = CameraPOD()
    for k in range(3):
        c.origin[k] = float(cam["origin"][k])
        c.view_direction[k] = float(cam["view_direction"][k])
        c.up[k] = float(cam["up"][k])
    for k in range(4):
        c.view_matrix[5 * k] = 1.0  # identity, src/main_view.rs:29
    return c


def _frame(verts, faces, camera, settings, fn_id, W, H, h_initial, adaptive, rect=None, x_step=1, y_step=1,
           use_fp64=False, n_threads=0):
    verts = np.ascontiguousarray(verts, dtype=np.float32)
    faces = np.ascontiguousarray(faces, dtype=np.uint32)
    x0, x1, y0, y1 = rect if rect is not None else (0, W, 0, H)
    f = OracleFrame(W, H, make_camera(camera), make_settings(settings), float(h_initial), int(bool(adaptive)),
                    int(fn_id), verts.shape[0], faces.shape[0], verts.ctypes.data, faces.ctypes.data,
                    x0, x1, y0, y1, x_step, y_step, int(use_fp64), n_threads)
    f._keep = (verts, faces)
    return f


def march(verts, faces, camera, settings, fn_id, W, H, h_initial=0.1, adaptive=True, rect=None, x_step=1,
          y_step=1, use_fp64=False, n_threads=0, want_color_f32=False):
    """Full-frame-indexed outputs (only the sampled pixels are written; others stay zero / -2)."""
    f = _frame(verts, faces, camera, settings, fn_id, W, H, h_initial, adaptive, rect, x_step, y_step, use_fp64,
               n_threads)
    out = dict(
        rgba8=np.zeros((H, W, 4), np.uint8),
        endpoint=np.zeros((H, W, 4), np.float32),
        hit_id=np.full((H, W), -2, np.int32),
        accepted=np.zeros((H, W), np.int32),
        rejected=np.zeros((H, W), np.int32),
        prehit=np.zeros((H, W), np.int32),
    )
    if want_color_f32:
        out["color"] = np.zeros((H, W, 4), np.float32)
    o = OracleOutputs(out["color"].ctypes.data if want_color_f32 else None, out["rgba8"].ctypes.data,
                      out["endpoint"].ctypes.data, out["hit_id"].ctypes.data, out["accepted"].ctypes.data,
                      out["rejected"].ctypes.data, out["prehit"].ctypes.data)
    rc = lib().oracle_march(C.byref(f), C.byref(o))
    if rc != 0:
        raise RuntimeError(f"oracle_march rc={rc}")
    return out


def sample_positions(camera, settings, W, H, outline=False, exponents=None):
    f = _frame(np.zeros((1, 3), np.float32), np.zeros((1, 3), np.uint32), camera, settings, 0, W, H, 0.1, True)
    pos = np.zeros((8, 2), np.float32)
    ex = None
    if outline:
        exponents = np.ascontiguousarray(exponents, dtype=np.float32)
        ex = exponents.ctypes.data
    lib().oracle_sample_positions(C.byref(f), int(bool(outline)), C.c_void_p(ex), C.c_void_p(pos.ctypes.data))
    return pos


def sample_rays(camera, settings, fn_id, W, H, positions):
    f = _frame(np.zeros((1, 3), np.float32), np.zeros((1, 3), np.uint32), camera, settings, fn_id, W, H, 0.1, True)
    positions = np.ascontiguousarray(positions, dtype=np.float32)
    out = np.zeros((800, 8), np.float32)
    lib().oracle_sample_rays(C.byref(f), C.c_void_p(positions.ctypes.data), C.c_void_p(out.ctypes.data))
    return out


def overlay(raycast_rgba8, endpoint, settings, exponents=None):
    H, W = raycast_rgba8.shape[:2]
    raycast_rgba8 = np.ascontiguousarray(raycast_rgba8, dtype=np.uint8)
    endpoint = np.ascontiguousarray(endpoint, dtype=np.float32)
    assert endpoint.shape == (H, W, 4)
    final = np.zeros_like(raycast_rgba8)
    exps = np.zeros((H, W), np.float32) if exponents is None else np.array(exponents, dtype=np.float32, copy=True)
    st = make_settings(settings)
    lib().oracle_overlay(W, H, C.byref(st), C.c_void_p(raycast_rgba8.ctypes.data), C.c_void_p(endpoint.ctypes.data),
                         C.c_void_p(final.ctypes.data), C.c_void_p(exps.ctypes.data))
    return final, exps


def field(fn_id, settings, p_prev, p, v0, v, t, use_fp64=False):
    """field_function on n samples; use_fp64 evaluates the same restatement in double (diagnostic, rounded to f32)."""
    arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (p_prev, p, v0, v, t)]
    n = arrs[4].shape[0]
    out = np.zeros((n, 3), np.float32)
    st = make_settings(settings)
    fn = lib().oracle_field_fp64 if use_fp64 else lib().oracle_field
    fn(int(fn_id), C.byref(st), n, *[C.c_void_p(a.ctypes.data) for a in arrs], C.c_void_p(out.ctypes.data))
    return out


HELPERS = dict(rotateX=0, rotateY=1, rotateZ=2, translate=3, refraction=4, point_plane_distance=5, sigmoid=6,
               refraction_index=7, reflect=8, normalize=9, cross=10)


def helper(name, n, v=None, nrm=None, p0=None, a=None, b=None, c=None):
    z3 = np.zeros((n, 3), np.float32)
    z1 = np.zeros(n, np.float32)
    arrs = [np.ascontiguousarray(x if x is not None else d, dtype=np.float32)
            for x, d in ((v, z3), (nrm, z3), (p0, z3), (a, z1), (b, z1), (c, z1))]
    out = np.zeros((n, 3), np.float32)
    lib().oracle_helper(HELPERS[name], n, *[C.c_void_p(x.ctypes.data) for x in arrs], C.c_void_p(out.ctypes.data))
    return out


def hit_triangle(tri, origin, direction):
    tri = np.ascontiguousarray(tri, dtype=np.float32).reshape(-1, 9)
    origin = np.ascontiguousarray(origin, dtype=np.float32)
    direction = np.ascontiguousarray(direction, dtype=np.float32)
    out = np.zeros(tri.shape[0], np.float32)
    lib().oracle_hit_triangle(tri.shape[0], C.c_void_p(tri.ctypes.data), C.c_void_p(origin.ctypes.data),
                              C.c_void_p(direction.ctypes.data), C.c_void_p(out.ctypes.data))
    return out


def vec2squareroots(a01):
    a01 = np.ascontiguousarray(a01, dtype=np.float32).reshape(-1, 2)
    out = np.zeros_like(a01)
    lib().oracle_vec2squareroots(a01.shape[0], C.c_void_p(a01.ctypes.data), C.c_void_p(out.ctypes.data))
    return out


def num_threads() -> int:
    return lib().oracle_num_threads()
