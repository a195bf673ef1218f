"""Host emulation of the march module (TEST INFRASTRUCTURE).

build(): asks the product for the exact NVRTC source text + options of a variant (LINON_DUMP_SOURCE, no GPU needed),
compiles that text with g++ against tests/emu/cuda_shim.hpp (one warp per CTA, lanes as fibers) and returns a ctypes
handle; march(): runs linon_march_tiles over a frame on the host cores.  Strict-math modules are bit-identical with the
device; fast-math modules differ from it by the MUFU approximation error only (libm here).  Only tests import this.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import re
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_CACHE = os.path.join(tempfile.gettempdir(), "linon_emu_cache")


class EmuFrame(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("cam_origin", C.c_float * 4), ("cam_view", C.c_float * 4),
                ("cam_up", C.c_float * 4), ("settings", C.c_byte * 24), ("verts", C.c_void_p), ("faces", C.c_void_p),
                ("n_verts", C.c_int32), ("n_faces", C.c_int32), ("rgba8", C.c_void_p), ("endpoint", C.c_void_p),
                ("hit_id", C.c_void_p), ("counters", C.c_void_p), ("n_threads", C.c_int32), ("tile_row_step", C.c_int32),
                ("kernel", C.c_int32), ("grid", C.c_int32), ("pos16", C.c_void_p), ("ray_samples", C.c_void_p)]


def _cxx():
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def build(body: str, flags: int, math_mode: int, quiet_zone: int = 0, cull: int = 1, opt: str = "-O1", pairs: bool = False,
          quiet_radius: int = 0, bvh: bool = False, n_faces: int = 36, module_src: str | None = None, extra_defines=()):
    """-> ctypes library with emu_march for this variant of the march module (pairs: also linon_march_pairs)."""
    from linon_b200.main_view import compile_check
    os.makedirs(_CACHE, exist_ok=True)
    env = {"LINON_CHECK_BLOCK": "32", "LINON_CHECK_QUIET_ZONE": str(quiet_zone), "LINON_CHECK_CULL": str(cull),
           "LINON_CHECK_KERNEL": "2" if pairs else "0", "LINON_CHECK_QUIET_RADIUS": str(quiet_radius),
           "LINON_CHECK_BVH": "1" if bvh else "0"}
    env["LINON_DUMP_SOURCE_ONLY"] = "1"
    old = {k: os.environ.get(k) for k in list(env) + ["LINON_DUMP_SOURCE"]}
    fd, dump = tempfile.mkstemp(suffix=".cu", dir=_CACHE)
    os.close(fd)
    try:
        os.environ.update(env)
        os.environ["LINON_DUMP_SOURCE"] = dump
        try:
            compile_check(body, flags=flags, math_mode=math_mode, n_faces=n_faces, module_src=module_src)
        except Exception as e:  # LINON_DUMP_SOURCE_ONLY: the module text is written, the NVRTC compile is skipped
            if "LINON_DUMP_SOURCE_ONLY" not in str(e):
                raise
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    text = open(dump, "rb").read()
    os.unlink(dump)
    shim = open(os.path.join(_HERE, "cuda_shim.hpp"), "rb").read() + open(os.path.join(_HERE, "emu_driver.cpp"), "rb").read()
    key = hashlib.sha1(text + shim + opt.encode() + " ".join(extra_defines).encode()).hexdigest()[:20]
    src = os.path.join(_CACHE, f"march_{key}.cu.inc")
    so = os.path.join(_CACHE, f"libemu_{key}.so")
    if not os.path.exists(so):
        header = text.split(b"\n", 1)[0].decode()
        assert header.startswith("// nvrtc:"), header[:80]
        defines = [o for o in header.split()[2:] if o.startswith("-D")]
        with open(src, "wb") as f:
            f.write(text)
        cmd = [_cxx(), opt, "-std=c++17", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-w", "-DLINON_WARP_STATS=1",
               f'-DLINON_EMU_SOURCE="{src}"', "-I", _HERE, "-I", os.path.join(_ROOT, "linon_b200", "csrc", "host"),
               *defines, *extra_defines, os.path.join(_HERE, "emu_driver.cpp"), "-o", so + ".tmp"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("emulation build failed:\n" + r.stderr[-6000:])
        os.replace(so + ".tmp", so)
    lib = C.CDLL(so)
    lib.emu_march.restype = C.c_int
    return lib


def march(lib, verts, faces, camera, settings, W, H, n_threads=0, tile_row_step=1, kernel=0, grid=0, sample_pos=None):
    """camera: dict(origin, view_direction, up); settings: dict like the oracle's. -> dict(rgba8, endpoint, hit_id, counters).
    kernel: LINON_OPT_KERNEL (0 tiles, 1 persistent two phases, 2 pairs, 3 persistent one phase).
    sample_pos: float32[8,2] -> runs linon_sample_rays instead and returns dict(ray_samples float32[800,8])."""
    verts = np.ascontiguousarray(verts, np.float32)
    faces = np.ascontiguousarray(faces, np.uint32)
    out = dict(rgba8=np.zeros((H, W, 4), np.uint8), endpoint=np.zeros((H, W, 4), np.float32),
               hit_id=np.full((H, W), -2, np.int32), counters=np.zeros(16, np.uint64))
    f = EmuFrame()
    f.width, f.height = W, H
    for k in range(3):
        f.cam_origin[k] = float(camera["origin"][k])
        f.cam_view[k] = float(camera["view_direction"][k])
        f.cam_up[k] = float(camera["up"][k])
    st = np.zeros(1, np.dtype([("fw", "f4"), ("mx", "f4"), ("my", "f4"), ("mode", "i4"), ("delta", "i4"), ("scale", "f4")]))
    st[0] = (settings["field_weight"], settings["mouse_x"], settings["mouse_y"], settings["overlay_mode"],
             settings["central_difference_delta"], settings["lyapunov_scaling"])
    C.memmove(f.settings, st.ctypes.data, 24)
    f.verts, f.faces = verts.ctypes.data, faces.ctypes.data
    f.n_verts, f.n_faces = verts.shape[0], faces.shape[0]
    f.rgba8, f.endpoint = out["rgba8"].ctypes.data, out["endpoint"].ctypes.data
    f.hit_id, f.counters = out["hit_id"].ctypes.data, out["counters"].ctypes.data
    f.n_threads = n_threads or (os.cpu_count() or 8)
    f.tile_row_step = tile_row_step
    f.kernel, f.grid = kernel, grid
    if sample_pos is not None:
        pos = np.ascontiguousarray(sample_pos, np.float32).reshape(8, 2)
        out["ray_samples"] = np.zeros((800, 8), np.float32)
        f.pos16, f.ray_samples = pos.ctypes.data, out["ray_samples"].ctypes.data
    rc = lib.emu_march(C.byref(f))
    if rc != 0:
        raise RuntimeError(f"emu_march rc={rc}")
    return out
