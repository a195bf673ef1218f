"""Python face of the host mirror: MainView / ArcballCamera / scene / PredefinedFunction.

Thin wrappers over the C shim of the C++ host mirror (include/linon_b200_host.h), which itself sits on the C ABI
(include/linon_b200.h).  Method names follow the reference (src/main_view.rs) so tests read like calls on the
reference's MainView.  Everything is computed by the CUDA library; numpy only carries the bytes.
"""
from __future__ import annotations

import ctypes as C
import enum

import numpy as np

from . import _capi as capi
from ._capi import Camera, Counters, LinonError, Settings, Timings


class PredefinedFunction(enum.IntEnum):  # src/functions.rs:2-13
    Custom = 0
    MirageSpherical = 1
    MiragePlane = 2
    MirageSphericalSigmoid = 3
    MiragePlaneSigmoid = 4
    TranslationX = 5
    TranslationZ = 6
    Rotation = 7
    LorenzAttractor = 8
    RoesslerAttractor = 9

    def to_code(self) -> str:  # src/functions.rs:33-128
        buf = C.create_string_buffer(4096)
        n = capi.lib().linon_host_function_code(int(self), buf, len(buf))
        if n < 0:
            raise LinonError(capi.ERR_INVALID, "unknown predefined function")
        return buf.raw[:n].decode("utf-8")

    def to_string(self) -> str:  # src/functions.rs:15-30
        buf = C.create_string_buffer(256)
        n = capi.lib().linon_host_function_name(int(self), buf, len(buf))
        return buf.raw[:n].decode("utf-8")


def default_scene(negate_xz: bool = True):
    """Application::new's start-up mesh (src/application.rs:105-123): verts [72,3] f32, faces [36,3] u32, centre."""
    verts = np.zeros((72, 3), np.float32)
    faces = np.zeros((36, 3), np.uint32)
    center = np.zeros(3, np.float32)
    capi.lib().linon_host_default_scene(verts.ctypes.data, faces.ctypes.data, center.ctypes.data, int(negate_xz))
    return verts, faces, center


def load_obj(text: str):
    """The OBJ drop path (src/main.rs:139-157 -> Application::load_model): (verts [n,3] f32 normalised, faces [m,3] u32, centre)."""
    L = capi.lib()
    nv, nf = C.c_uint32(), C.c_uint32()
    err = C.create_string_buffer(1024)
    data = text.encode("utf-8")
    rc = L.linon_host_load_obj(data, None, 0, None, 0, C.byref(nv), C.byref(nf), None, err, len(err))
    if rc != capi.OK:
        raise LinonError(rc, err.value.decode("utf-8", "replace"))
    verts = np.zeros((nv.value, 3), np.float32)
    faces = np.zeros((nf.value, 3), np.uint32)
    center = np.zeros(3, np.float32)
    rc = L.linon_host_load_obj(data, verts.ctypes.data, nv.value, faces.ctypes.data, nf.value, C.byref(nv), C.byref(nf),
                               center.ctypes.data, err, len(err))
    if rc != capi.OK:
        raise LinonError(rc, err.value.decode("utf-8", "replace"))
    return verts, faces, center


def make_settings(field_weight=1.0, mouse_pos=(0.5, 0.6), overlay_mode=0, central_difference_delta=1,
                  lyapunov_scaling=50.0) -> Settings:
    """Defaults are MainView::new's initial uniform (src/main_view.rs:133-139)."""
    s = Settings()
    s.field_weight = field_weight
    s.mouse_pos[0], s.mouse_pos[1] = mouse_pos
    s.overlay_mode = overlay_mode
    s.central_difference_delta = central_difference_delta
    s.lyapunov_scaling = lyapunov_scaling
    return s


class ArcballCamera:
    """src/arcball.rs:18-129."""

    def __init__(self, center, zoom_speed, screen, _borrowed=None):
        self._own = _borrowed is None
        if _borrowed is not None:
            self._h = _borrowed
        else:
            c = np.ascontiguousarray(center, np.float32)
            self._h = capi.lib().linon_host_arcball_new(c.ctypes.data, zoom_speed, screen[0], screen[1])

    def __del__(self):
        if getattr(self, "_own", False) and self._h:
            capi.lib().linon_host_arcball_free(self._h)
            self._h = None

    def rotate(self, mouse_prev, mouse_cur):
        capi.lib().linon_host_arcball_rotate(self._h, mouse_prev[0], mouse_prev[1], mouse_cur[0], mouse_cur[1])

    def zoom(self, amount, elapsed):
        capi.lib().linon_host_arcball_zoom(self._h, amount, elapsed)

    def pan(self, mouse_delta):
        capi.lib().linon_host_arcball_pan(self._h, mouse_delta[0], mouse_delta[1])

    def update_screen(self, width, height):
        capi.lib().linon_host_arcball_update_screen(self._h, width, height)

    def uniform(self) -> Camera:
        """CameraUniform::moving, src/main_view.rs:21-31."""
        u = Camera()
        capi.lib().linon_host_arcball_uniform(self._h, C.byref(u))
        return u

    def eye_pos(self):
        return np.array(self.uniform().origin[:3], np.float32)

    def eye_dir(self):
        return np.array(self.uniform().view_direction[:3], np.float32)

    def up_dir(self):
        return np.array(self.uniform().up[:3], np.float32)


def orbit_camera(center, width, height, frame, n_frames=64) -> ArcballCamera:
    """Camera of frame `frame` of the offline orbit: default camera, then `frame` successive
    horizontal drags from the image centre by sin(pi/n_frames) * W/2 pixels (360/n_frames degrees each)."""
    cam = ArcballCamera(center, 1.0, (width, height))
    cam.zoom(-1.0, 1.0)
    w2, h2 = np.float32(width) / np.float32(2), np.float32(height) / np.float32(2)
    cur_x = np.float32(w2 + np.float32(np.sin(np.pi / n_frames)) * w2)
    for _ in range(frame):
        cam.rotate((float(w2), float(h2)), (float(cur_x), float(h2)))
    return cam


class MainView:
    """src/main_view.rs:53-849 (device-facing half) on the CUDA library."""

    def __init__(self, vertices, faces, center, width, height, discrete_gpu=True, device=0):
        L = capi.lib()
        v = np.ascontiguousarray(vertices, np.float32).reshape(-1, 3)
        f = np.ascontiguousarray(faces, np.uint32).reshape(-1, 3)
        c = np.ascontiguousarray(center, np.float32)
        self._h = L.linon_host_mv_new(device, v.ctypes.data, v.shape[0], f.ctypes.data, f.shape[0], c.ctypes.data,
                                      width, height, int(discrete_gpu))
        if not self._h:
            msg = L.linon_host_mv_error(None).decode()
            raise LinonError(capi.ERR_NO_DRIVER if "no CPU fallback" in msg else capi.ERR_CUDA, msg)
        self._view = L.linon_host_mv_handle(self._h)

    # ---- lifetime -------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            capi.lib().linon_host_mv_free(self._h)
            self._h = None
            self._view = None

    def __del__(self):
        self.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc):
        if rc != capi.OK:
            L = capi.lib()
            msg = L.linon_host_mv_error(self._h).decode() or L.linon_last_error(self._view).decode()
            raise LinonError(rc, msg)

    def _check_view(self, rc):
        if rc != capi.OK:
            raise LinonError(rc, capi.lib().linon_last_error(self._view).decode())

    # ---- the reference's pub methods --------------------------------------------------------------------
    @property
    def needs_redraw(self) -> bool:
        return bool(capi.lib().linon_host_mv_needs_redraw(self._h))

    @property
    def size(self):
        w, h = C.c_uint32(), C.c_uint32()
        capi.lib().linon_host_mv_size(self._h, C.byref(w), C.byref(h))
        return w.value, h.value

    @property
    def camera(self) -> ArcballCamera:
        return ArcballCamera(None, None, None, _borrowed=capi.lib().linon_host_mv_camera(self._h))

    def update_model(self, vertices, faces, center):
        v = np.ascontiguousarray(vertices, np.float32).reshape(-1, 3)
        f = np.ascontiguousarray(faces, np.uint32).reshape(-1, 3)
        c = np.ascontiguousarray(center, np.float32)
        self._check(capi.lib().linon_host_mv_update_model(self._h, v.ctypes.data, v.shape[0], f.ctypes.data, f.shape[0],
                                                          c.ctypes.data))

    def update_camera(self):
        self._check(capi.lib().linon_host_mv_update_camera(self._h))

    def set_camera_uniform(self, camera: Camera):
        """Uploads an explicit 112-byte camera uniform (what update_camera does with the arcball's)."""
        self._check_view(capi.lib().linon_set_camera(self._view, C.byref(camera)))

    def reset_camera(self):
        self._check(capi.lib().linon_host_mv_reset_camera(self._h))

    def on_zoom(self, delta):
        self._check(capi.lib().linon_host_mv_on_zoom(self._h, delta))

    def on_pointer_moved(self, camera_op, prev, pos):
        op = {"none": 0, "rotate": 1, "pan": 2}[camera_op]
        self._check(capi.lib().linon_host_mv_on_pointer_moved(self._h, op, prev[0], prev[1], pos[0], pos[1]))

    def update_settings(self, settings: Settings):
        self._check(capi.lib().linon_host_mv_update_settings(self._h, C.byref(settings)))

    def resize_texture(self, width, height):
        self._check(capi.lib().linon_host_mv_resize_texture(self._h, width, height))

    def reload_shader(self, field_function: str):
        """Raises LinonError(ERR_COMPILE) with the diagnostic; the previous function stays active."""
        err = C.create_string_buffer(16384)
        rc = capi.lib().linon_host_mv_reload_shader(self._h, field_function.encode("utf-8"), err, len(err))
        if rc != capi.OK:
            raise LinonError(rc, err.value.decode("utf-8", "replace"))

    def reload_compute_shader(self, new_src: str | None):
        """Application::reload_compute_shader (src/application.rs:228-232): the text of an edited main_view.wgsl; None
        or "" returns to the stock module.  Raises LinonError(ERR_UNSUPPORTED) when the edit changes the march itself,
        LinonError(ERR_COMPILE) on a syntax / compile error; the previous module stays active either way."""
        err = C.create_string_buffer(16384)
        rc = capi.lib().linon_host_mv_reload_compute_shader(self._h, (new_src or "").encode("utf-8"), err, len(err))
        if rc != capi.OK:
            raise LinonError(rc, err.value.decode("utf-8", "replace"))

    def module_overrides(self) -> list[str]:
        buf = C.create_string_buffer(4096)
        self._check_view(capi.lib().linon_module_overrides(self._view, buf, len(buf)))
        return [n for n in buf.value.decode().split(",") if n]

    def enable_adaptive_sampling(self):
        self._check(capi.lib().linon_host_mv_set_adaptive_sampling(self._h, 1))

    def disable_adaptive_sampling(self):
        self._check(capi.lib().linon_host_mv_set_adaptive_sampling(self._h, 0))

    def render(self):
        self._check(capi.lib().linon_host_mv_render(self._h))

    def render_high_accuracy(self, field_function: str | None = None):
        ff = field_function.encode("utf-8") if field_function is not None else None
        self._check(capi.lib().linon_host_mv_render_high_accuracy(self._h, ff))

    def render_outline_rays(self, field_function: str | None = None):
        ff = field_function.encode("utf-8") if field_function is not None else None
        self._check(capi.lib().linon_host_mv_render_outline_rays(self._h, ff))

    # ---- C-ABI level access (flags, options, shards, buffers) ---------------------------------------------
    def render_flags(self, flags: int):
        self._check_view(capi.lib().linon_render(self._view, flags))

    def sync(self):
        self._check_view(capi.lib().linon_sync(self._view))

    def set_option(self, option: int, value: int):
        self._check_view(capi.lib().linon_set_option(self._view, option, value))

    def set_shard(self, rank: int, world: int, rows_per_block: int = 32):
        self._check_view(capi.lib().linon_set_shard(self._view, rank, world, rows_per_block))

    # ---- peer stores: a shard's march writes straight into the root shard's frame buffers (include/linon_b200.h) --------
    def redirect_output(self, buffer: int, device_ptr: int | None):
        self._check_view(capi.lib().linon_redirect_output(self._view, buffer, C.c_void_p(device_ptr or 0)))

    def ipc_export(self, buffer: int) -> bytes:
        h = C.create_string_buffer(64)
        self._check_view(capi.lib().linon_ipc_export(self._view, buffer, h))
        return h.raw

    def ipc_open(self, handle: bytes) -> int:
        p = C.c_void_p(0)
        self._check_view(capi.lib().linon_ipc_open(self._view, C.c_char_p(handle), C.byref(p)))
        return int(p.value)

    def export_buffer_fd(self, buffer: int):
        """(fd, allocation bytes) of LINON_BUF_FINAL_RGBA8 / LINON_BUF_RAY_SAMPLES in an exportable allocation; the caller closes fd."""
        fd, n = C.c_int(-1), C.c_size_t(0)
        self._check_view(capi.lib().linon_export_buffer_fd(self._view, buffer, C.byref(fd), C.byref(n)))
        return fd.value, n.value

    def counters(self) -> Counters:
        c = Counters()
        self._check_view(capi.lib().linon_get_counters(self._view, C.byref(c)))
        return c

    def timings(self) -> Timings:
        t = Timings()
        self._check_view(capi.lib().linon_get_timings(self._view, C.byref(t)))
        return t

    def march_cubin_hash(self) -> str:
        """FNV-1a of the cubin whose march kernel ran last, 16 hex digits (ties ncu captures to the running kernel)."""
        h = C.c_uint64(0)
        self._check_view(capi.lib().linon_march_cubin_hash(self._view, C.byref(h)))
        return f"{h.value:016x}"

    def debug_profile(self) -> np.ndarray:
        out = np.zeros(16, np.uint64)
        self._check_view(capi.lib().linon_debug_profile(self._view, out.ctypes.data))
        return out

    def device_ptr(self, buffer: int) -> int:
        return capi.lib().linon_device_ptr(self._view, buffer)

    def stream(self) -> int:
        return capi.lib().linon_stream(self._view)

    def buffer_bytes(self, buffer: int) -> int:
        return capi.lib().linon_buffer_bytes(self._view, buffer)

    def read_into(self, buffer: int, dst: np.ndarray):
        self._check_view(capi.lib().linon_read(self._view, buffer, dst.ctypes.data, dst.nbytes))

    def read_into_ptr(self, buffer: int, host_ptr: int, nbytes: int):
        """Reads a buffer into caller-owned host memory (e.g. a pinned torch tensor's data_ptr())."""
        self._check_view(capi.lib().linon_read(self._view, buffer, host_ptr, nbytes))

    def read(self, buffer: int) -> np.ndarray:
        w, h = self.size
        shape, dtype = {
            capi.BUF_FINAL_RGBA8: ((h, w, 4), np.uint8),
            capi.BUF_RAYCAST_RGBA8: ((h, w, 4), np.uint8),
            capi.BUF_ENDPOINT_F32X4: ((h, w, 4), np.float32),
            capi.BUF_EXPONENTS_F32: ((h, w), np.float32),
            capi.BUF_HIT_ID_I32: ((h, w), np.int32),
            capi.BUF_RAY_SAMPLES: ((800, 8), np.float32),
            capi.BUF_SAMPLE_POS: ((8, 2), np.float32),
        }[buffer]
        out = np.empty(shape, dtype)
        self.read_into(buffer, out)
        return out

    # convenience names of the reference's resources
    def texture(self):
        return self.read(capi.BUF_FINAL_RGBA8)

    def ray_casting_texture(self):
        return self.read(capi.BUF_RAYCAST_RGBA8)

    def mapping_texture(self):
        return self.read(capi.BUF_ENDPOINT_F32X4)

    def exponents(self):
        return self.read(capi.BUF_EXPONENTS_F32)

    def hit_ids(self):
        return self.read(capi.BUF_HIT_ID_I32)

    def ray_samples(self):
        return self.read(capi.BUF_RAY_SAMPLES)

    # ---- parity / debug --------------------------------------------------------------------------------------
    def debug_eval_field(self, p_prev, p, v0, v, t):
        a = [np.ascontiguousarray(x, np.float32) for x in (p_prev, p, v0, v, t)]
        n = a[4].shape[0]
        out = np.zeros((n, 3), np.float32)
        self._check_view(capi.lib().linon_debug_eval_field(self._view, n, *[x.ctypes.data for x in a], out.ctypes.data))
        return out

    def debug_eval_rk4(self, p, direction, v0, t, h):
        a = [np.ascontiguousarray(x, np.float32) for x in (p, direction, v0, t, h)]
        n = a[3].shape[0]
        out = np.zeros((n, 3), np.float32)
        self._check_view(capi.lib().linon_debug_eval_rk4(self._view, n, *[x.ctypes.data for x in a], out.ctypes.data))
        return out

    def debug_eval_helper(self, which: int, n: int, v=None, nrm=None, p0=None, a=None, b=None, c=None):
        z3, z1 = np.zeros((n, 3), np.float32), np.zeros(n, np.float32)
        arrs = [np.ascontiguousarray(x if x is not None else d, np.float32)
                for x, d in ((v, z3), (nrm, z3), (p0, z3), (a, z1), (b, z1), (c, z1))]
        out = np.zeros((n, 3), np.float32)
        self._check_view(capi.lib().linon_debug_eval_helper(self._view, which, n, *[x.ctypes.data for x in arrs],
                                                            out.ctypes.data))
        return out

    def debug_hit_triangle(self, tri, origin, direction):
        tri = np.ascontiguousarray(tri, np.float32).reshape(-1, 9)
        o = np.ascontiguousarray(origin, np.float32)
        d = np.ascontiguousarray(direction, np.float32)
        out = np.zeros(tri.shape[0], np.float32)
        self._check_view(capi.lib().linon_debug_hit_triangle(self._view, tri.shape[0], tri.ctypes.data, o.ctypes.data,
                                                             d.ctypes.data, out.ctypes.data))
        return out


def match_mirage(body: str):
    """(kind, dict of constants) of the structural mirage-family match (csrc/host/mirage_match.hpp); kind 0 = no match."""
    kind = C.c_int(0)
    p = np.zeros(8, np.float64)
    rc = capi.lib().linon_debug_match_mirage(body.encode("utf-8"), C.byref(kind), p.ctypes.data)
    if rc != capi.OK:
        raise LinonError(rc, "linon_debug_match_mirage")
    return kind.value, dict(t_env=p[0], t_src=p[1], max_dist=p[2], c=(p[3], p[4], p[5]), sig_k=p[6], sig_c=p[7])


def read_exported_buffer(device: int, fd: int, allocation_bytes: int, nbytes: int) -> np.ndarray:
    """Consumer side of the display hand-off: maps the descriptor in THIS process, copies nbytes to the host, unmaps."""
    L = capi.lib()
    ptr, handle = C.c_void_p(0), C.c_void_p(0)
    rc = L.linon_import_buffer_fd(device, fd, allocation_bytes, C.byref(ptr), C.byref(handle))
    if rc != capi.OK:
        raise LinonError(rc, L.linon_last_error(None).decode())
    try:
        out = np.empty(nbytes, np.uint8)
        rc = L.linon_copy_to_host(device, ptr, out.ctypes.data, nbytes)
        if rc != capi.OK:
            raise LinonError(rc, "linon_copy_to_host")
        return out
    finally:
        L.linon_release_import(handle)


def translate_wgsl(body: str) -> str:
    out = C.create_string_buffer(1 << 16)
    err = C.create_string_buffer(4096)
    rc = capi.lib().linon_translate_wgsl(body.encode("utf-8"), out, len(out), err, len(err))
    if rc != capi.OK:
        raise LinonError(rc, err.value.decode("utf-8", "replace"))
    return out.value.decode("utf-8")


def translate_wgsl_module(module_src: str):
    """Helper layer of an edited main_view.wgsl -> (CUDA member functions, names emitted). No GPU needed."""
    out = C.create_string_buffer(1 << 18)
    names = C.create_string_buffer(4096)
    err = C.create_string_buffer(4096)
    rc = capi.lib().linon_translate_wgsl_module(module_src.encode("utf-8"), out, len(out), names, len(names), err, len(err))
    if rc != capi.OK:
        raise LinonError(rc, err.value.decode("utf-8", "replace"))
    return out.value.decode("utf-8"), [n for n in names.value.decode().split(",") if n]


def scan_wgsl_module(module_src: str, full: bool = False):
    """Top-level declarations of a WGSL module: list of (kind, name, hash, line); full: + (hash with the numeric literals
    masked, the literals in order of appearance)."""
    out = C.create_string_buffer(1 << 18)
    err = C.create_string_buffer(4096)
    rc = capi.lib().linon_scan_wgsl_module(module_src.encode("utf-8"), out, len(out), err, len(err))
    if rc != capi.OK:
        raise LinonError(rc, err.value.decode("utf-8", "replace"))
    items = []
    for line in out.value.decode().splitlines():
        f = line.split()
        kind, name, h, ln = f[:4]
        if full:
            items.append((kind, name, int(h, 16), int(ln), int(f[4], 16), [float(x) for x in f[6:6 + int(f[5])]]))
        else:
            items.append((kind, name, int(h, 16), int(ln)))
    return items


def compile_check(body: str, flags: int = 0, math_mode: int = capi.MATH_FUSED, n_faces: int = 36,
                  module_src: str | None = None) -> int:
    """Translate + NVRTC-compile the march module for sm_100a (no GPU needed). Returns the cubin size."""
    nb = C.c_size_t(0)
    err = C.create_string_buffer(16384)
    if module_src is not None:
        rc = capi.lib().linon_compile_check_module(module_src.encode("utf-8"), body.encode("utf-8"), flags, math_mode,
                                                   n_faces, C.byref(nb), err, len(err))
    else:
        rc = capi.lib().linon_compile_check(body.encode("utf-8"), flags, math_mode, n_faces, C.byref(nb), err, len(err))
    if rc != capi.OK:
        raise LinonError(rc, err.value.decode("utf-8", "replace"))
    return nb.value


def nvrtc_info():
    """(path, (major, minor)) of the NVRTC the library compiles with: the toolkit's copy by absolute path, whatever other
    libnvrtc.so.12 the process holds (include/linon_b200.h: linon_nvrtc_info)."""
    path = C.create_string_buffer(1024)
    major, minor = C.c_int(0), C.c_int(0)
    rc = capi.lib().linon_nvrtc_info(path, len(path), C.byref(major), C.byref(minor))
    if rc != capi.OK:
        raise LinonError(rc, "NVRTC could not be loaded")
    return path.value.decode(), (major.value, minor.value)


def measure_fp32_peak(device: int = 0):
    t, sms = C.c_double(0), C.c_int(0)
    rc = capi.lib().linon_measure_fp32_peak(device, C.byref(t), C.byref(sms))
    if rc != capi.OK:
        raise LinonError(rc, capi.lib().linon_last_error(None).decode())
    return t.value, sms.value


class ShardedFrame:
    """linon_sharded_* (include/linon_b200.h): one frame on several devices of a node, driven by one host thread.

    devices[0] is the root: read_root() reads the assembled frame from it."""

    def __init__(self, devices, vertices, faces, width, height, rows_per_block=32, mode=capi.SHARD_PEER_STORES):
        L = capi.lib()
        devs = (C.c_int * len(devices))(*devices)
        h = C.c_void_p(0)
        rc = L.linon_sharded_create(devs, len(devices), width, height, rows_per_block, mode, C.byref(h))
        if rc != capi.OK:
            raise LinonError(rc, L.linon_sharded_last_error(None).decode())
        self._s = h
        self.width, self.height = width, height
        v = np.ascontiguousarray(vertices, np.float32).reshape(-1, 3)
        f = np.ascontiguousarray(faces, np.uint32).reshape(-1, 3)
        self._check(L.linon_sharded_set_mesh(self._s, v.ctypes.data, v.shape[0], f.ctypes.data, f.shape[0]))

    def _check(self, rc):
        if rc != capi.OK:
            raise LinonError(rc, capi.lib().linon_sharded_last_error(self._s).decode())

    def close(self):
        if getattr(self, "_s", None):
            capi.lib().linon_sharded_destroy(self._s)
            self._s = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        self.close()

    @property
    def world(self):
        return capi.lib().linon_sharded_world(self._s)

    def view(self, rank: int) -> int:
        return capi.lib().linon_sharded_view(self._s, rank)

    def set_camera_uniform(self, camera: Camera):
        self._check(capi.lib().linon_sharded_set_camera(self._s, C.byref(camera)))

    def update_settings(self, settings: Settings):
        self._check(capi.lib().linon_sharded_set_settings(self._s, C.byref(settings)))

    def reload_shader(self, field_function: str):
        err = C.create_string_buffer(16384)
        rc = capi.lib().linon_sharded_set_field_function(self._s, field_function.encode("utf-8"), err, len(err))
        if rc != capi.OK:
            raise LinonError(rc, err.value.decode("utf-8", "replace") or capi.lib().linon_sharded_last_error(self._s).decode())

    def set_option(self, option: int, value: int):
        self._check(capi.lib().linon_sharded_set_option(self._s, option, value))

    def resize_texture(self, width: int, height: int):
        self._check(capi.lib().linon_sharded_resize(self._s, width, height))
        self.width, self.height = width, height

    def render_flags(self, flags: int):
        self._check(capi.lib().linon_sharded_render(self._s, flags))

    def sync(self):
        self._check(capi.lib().linon_sharded_sync(self._s))

    def read_root(self, buffer: int) -> np.ndarray:
        w, h = self.width, self.height
        shape, dtype = {capi.BUF_FINAL_RGBA8: ((h, w, 4), np.uint8), capi.BUF_RAYCAST_RGBA8: ((h, w, 4), np.uint8),
                        capi.BUF_ENDPOINT_F32X4: ((h, w, 4), np.float32), capi.BUF_EXPONENTS_F32: ((h, w), np.float32),
                        capi.BUF_HIT_ID_I32: ((h, w), np.int32), capi.BUF_RAY_SAMPLES: ((800, 8), np.float32),
                        capi.BUF_SAMPLE_POS: ((8, 2), np.float32)}[buffer]
        out = np.empty(shape, dtype)
        rc = capi.lib().linon_read(self.view(0), buffer, out.ctypes.data, out.nbytes)
        if rc != capi.OK:
            raise LinonError(rc, capi.lib().linon_last_error(self.view(0)).decode())
        return out

    def counters(self, rank: int) -> Counters:
        c = Counters()
        rc = capi.lib().linon_get_counters(self.view(rank), C.byref(c))
        if rc != capi.OK:
            raise LinonError(rc, capi.lib().linon_last_error(self.view(rank)).decode())
        return c

    def timings(self, rank: int) -> Timings:
        t = Timings()
        rc = capi.lib().linon_get_timings(self.view(rank), C.byref(t))
        if rc != capi.OK:
            raise LinonError(rc, capi.lib().linon_last_error(self.view(rank)).decode())
        return t
