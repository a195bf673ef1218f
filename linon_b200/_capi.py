"""ctypes binding of liblinon_b200.so (include/linon_b200.h + include/linon_b200_host.h).

The product path is the CUDA library; there is no CPU fallback.  Loading the library works without a GPU (the
driver is dlopen'ed lazily), so the translator and the NVRTC compile check can run anywhere; everything that
touches a device raises LinonError when no B200 is present.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "liblinon_b200.so")

# --- enums of include/linon_b200.h ---------------------------------------------------------------
OK, ERR_INVALID, ERR_CUDA, ERR_COMPILE, ERR_NO_DRIVER, ERR_SIZE, ERR_UNSUPPORTED = range(7)
RENDER_DEFAULT, RENDER_ENHANCE, RENDER_OUTLINE, RENDER_ADAPTIVE_OFF = 0, 1, 2, 4
RENDER_NO_OVERLAY, RENDER_NO_MARCH, RENDER_NO_SAMPLE_RAYS, RENDER_KEEP_TIMINGS = 8, 16, 32, 64
SHARD_PEER_STORES, SHARD_NCCL_GATHER = 0, 1
BUF_FINAL_RGBA8, BUF_RAYCAST_RGBA8, BUF_ENDPOINT_F32X4, BUF_EXPONENTS_F32, BUF_HIT_ID_I32 = range(5)
BUF_RAY_SAMPLES, BUF_SAMPLE_POS = 5, 6
OPT_MATH, OPT_KERNEL, OPT_BLOCK, OPT_CTAS_PER_SM, OPT_REFILL_MIN, OPT_FACE_UNROLL, OPT_MAX_REGS, OPT_PROFILE, OPT_CULL = range(9)
OPT_PAIR_MIN_BLOCKS, OPT_SPECIALIZE, OPT_FORCE_BVH, OPT_QUIET_ZONE, OPT_QUIET_RADIUS = 9, 10, 11, 12, 13
OPT_OVERLAY_KERNEL = 14
OPT_COLLAPSE_RUNS = 15
MATH_STRICT, MATH_FUSED, MATH_FAST = 0, 1, 2
KERNEL_TILES, KERNEL_PERSISTENT, KERNEL_PAIRS, KERNEL_REFILL = 0, 1, 2, 3


class LinonError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"linon_b200 error {code}: {message}")
        self.code = code
        self.message = message


class Camera(C.Structure):  # linon_camera, 112 B
    _fields_ = [("origin", C.c_float * 4), ("view_direction", C.c_float * 4), ("up", C.c_float * 4),
                ("view_matrix", C.c_float * 16)]


class Settings(C.Structure):  # linon_settings, 24 B
    _fields_ = [("field_weight", C.c_float), ("mouse_pos", C.c_float * 2), ("overlay_mode", C.c_int32),
                ("central_difference_delta", C.c_int32), ("lyapunov_scaling", C.c_float)]


class Counters(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("accepted_steps", C.c_uint64), ("prehit_steps", C.c_uint64),
                ("rejected_steps", C.c_uint64), ("face_tests", C.c_uint64), ("plane_tests", C.c_uint64),
                ("culling_checks", C.c_uint64)]


class Timings(C.Structure):
    _fields_ = [("march_ms", C.c_float), ("overlay_ms", C.c_float), ("sample_rays_ms", C.c_float),
                ("outline_search_ms", C.c_float), ("jit_ms", C.c_float), ("march_regs", C.c_int32),
                ("march_grid", C.c_int32), ("march_block", C.c_int32), ("march_phase_a_ms", C.c_float),
                ("march_kernel", C.c_int32), ("overlay_kernel", C.c_int32)]


assert C.sizeof(Camera) == 112 and C.sizeof(Settings) == 24

# every symbol the two headers declare: (name, restype, argtypes)
_P, _I, _U, _F, _S = C.c_void_p, C.c_int, C.c_uint32, C.c_float, C.c_size_t
_CP = C.c_char_p
PROTOTYPES = [
    # include/linon_b200.h
    ("linon_create", _I, [_I, _U, _U, C.POINTER(_P)]),
    ("linon_destroy", None, [_P]),
    ("linon_set_mesh", _I, [_P, _P, _U, _P, _U]),
    ("linon_set_camera", _I, [_P, C.POINTER(Camera)]),
    ("linon_set_settings", _I, [_P, C.POINTER(Settings)]),
    ("linon_resize", _I, [_P, _U, _U]),
    ("linon_set_field_function", _I, [_P, _CP, _P, _S]),
    ("linon_reload_shader_source", _I, [_P, _CP, _P, _S]),
    ("linon_reload_shader", _I, [_P, _CP, _CP, _P, _S]),
    ("linon_module_overrides", _I, [_P, _P, _S]),
    ("linon_render", _I, [_P, _U]),
    ("linon_sync", _I, [_P]),
    ("linon_read", _I, [_P, _I, _P, _S]),
    ("linon_buffer_bytes", _S, [_P, _I]),
    ("linon_device_ptr", _P, [_P, _I]),
    ("linon_stream", _P, [_P]),
    ("linon_set_shard", _I, [_P, _U, _U, _U]),
    ("linon_redirect_output", _I, [_P, _I, _P]),
    ("linon_export_buffer_fd", _I, [_P, _I, C.POINTER(_I), C.POINTER(_S)]),
    ("linon_import_buffer_fd", _I, [_I, _I, _S, C.POINTER(_P), C.POINTER(_P)]),
    ("linon_release_import", _I, [_P]),
    ("linon_copy_to_host", _I, [_I, _P, _P, _S]),
    ("linon_ipc_export", _I, [_P, _I, _P]),
    ("linon_ipc_open", _I, [_P, _P, C.POINTER(_P)]),
    ("linon_sharded_create", _I, [_P, _U, _U, _U, _U, _I, C.POINTER(_P)]),
    ("linon_sharded_destroy", None, [_P]),
    ("linon_sharded_world", _U, [_P]),
    ("linon_sharded_view", _P, [_P, _U]),
    ("linon_sharded_set_mesh", _I, [_P, _P, _U, _P, _U]),
    ("linon_sharded_set_camera", _I, [_P, C.POINTER(Camera)]),
    ("linon_sharded_set_settings", _I, [_P, C.POINTER(Settings)]),
    ("linon_sharded_set_field_function", _I, [_P, _CP, _P, _S]),
    ("linon_sharded_set_option", _I, [_P, _I, _I]),
    ("linon_sharded_resize", _I, [_P, _U, _U]),
    ("linon_sharded_render", _I, [_P, _U]),
    ("linon_sharded_sync", _I, [_P]),
    ("linon_sharded_last_error", _CP, [_P]),
    ("linon_set_option", _I, [_P, _I, _I]),
    ("linon_get_counters", _I, [_P, C.POINTER(Counters)]),
    ("linon_get_timings", _I, [_P, C.POINTER(Timings)]),
    ("linon_march_cubin_hash", _I, [_P, C.POINTER(C.c_uint64)]),
    ("linon_nvrtc_info", _I, [_P, _S, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    ("linon_last_error", _CP, [_P]),
    ("linon_translate_wgsl", _I, [_CP, _P, _S, _P, _S]),
    ("linon_compile_check", _I, [_CP, _U, _I, _U, C.POINTER(_S), _P, _S]),
    ("linon_compile_check_module", _I, [_CP, _CP, _U, _I, _U, C.POINTER(_S), _P, _S]),
    ("linon_translate_wgsl_module", _I, [_CP, _P, _S, _P, _S, _P, _S]),
    ("linon_scan_wgsl_module", _I, [_CP, _P, _S, _P, _S]),
    ("linon_debug_eval_field", _I, [_P, _U, _P, _P, _P, _P, _P, _P]),
    ("linon_debug_eval_rk4", _I, [_P, _U, _P, _P, _P, _P, _P, _P]),
    ("linon_debug_eval_helper", _I, [_P, _I, _U, _P, _P, _P, _P, _P, _P, _P]),
    ("linon_debug_hit_triangle", _I, [_P, _U, _P, _P, _P, _P]),
    ("linon_debug_profile", _I, [_P, _P]),
    ("linon_debug_match_mirage", _I, [_CP, C.POINTER(_I), _P]),
    ("linon_measure_fp32_peak", _I, [_I, C.POINTER(C.c_double), C.POINTER(_I)]),
    # include/linon_b200_host.h
    ("linon_host_default_scene", None, [_P, _P, _P, _I]),
    ("linon_host_normalize_vertices", None, [_P, _U]),
    ("linon_host_load_obj", _I, [_CP, _P, _U, _P, _U, C.POINTER(_U), C.POINTER(_U), _P, _P, _S]),
    ("linon_host_get_center", None, [_P, _U, _P]),
    ("linon_host_function_code", _I, [_I, _P, _S]),
    ("linon_host_function_name", _I, [_I, _P, _S]),
    ("linon_host_arcball_new", _P, [_P, _F, _F, _F]),
    ("linon_host_arcball_free", None, [_P]),
    ("linon_host_arcball_rotate", None, [_P, _F, _F, _F, _F]),
    ("linon_host_arcball_zoom", None, [_P, _F, _F]),
    ("linon_host_arcball_pan", None, [_P, _F, _F]),
    ("linon_host_arcball_update_screen", None, [_P, _F, _F]),
    ("linon_host_arcball_uniform", None, [_P, C.POINTER(Camera)]),
    ("linon_host_mv_new", _P, [_I, _P, _U, _P, _U, _P, _U, _U, _I]),
    ("linon_host_mv_free", None, [_P]),
    ("linon_host_mv_error", _CP, [_P]),
    ("linon_host_mv_update_model", _I, [_P, _P, _U, _P, _U, _P]),
    ("linon_host_mv_update_camera", _I, [_P]),
    ("linon_host_mv_reset_camera", _I, [_P]),
    ("linon_host_mv_on_zoom", _I, [_P, _F]),
    ("linon_host_mv_on_pointer_moved", _I, [_P, _I, _F, _F, _F, _F]),
    ("linon_host_mv_update_settings", _I, [_P, C.POINTER(Settings)]),
    ("linon_host_mv_resize_texture", _I, [_P, _U, _U]),
    ("linon_host_mv_reload_shader", _I, [_P, _CP, _P, _S]),
    ("linon_host_mv_reload_compute_shader", _I, [_P, _CP, _P, _S]),
    ("linon_host_mv_set_adaptive_sampling", _I, [_P, _I]),
    ("linon_host_mv_render", _I, [_P]),
    ("linon_host_mv_render_high_accuracy", _I, [_P, _CP]),
    ("linon_host_mv_render_outline_rays", _I, [_P, _CP]),
    ("linon_host_mv_needs_redraw", _I, [_P]),
    ("linon_host_mv_handle", _P, [_P]),
    ("linon_host_mv_camera", _P, [_P]),
    ("linon_host_mv_size", None, [_P, C.POINTER(_U), C.POINTER(_U)]),
]

_lib = None


def lib():
    """Loads the in-tree library; raises if it has not been built (python -m linon_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LinonError(ERR_NO_DRIVER, f"{LIB_PATH} is missing: run `python -m linon_b200.build` "
                                            "(there is no CPU fallback)")
        handle = C.CDLL(LIB_PATH)
        for name, restype, argtypes in PROTOTYPES:
            fn = getattr(handle, name)  # AttributeError if a declared symbol is not exported
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib
