"""Display hand-off without a copy (include/linon_b200.h: linon_export_buffer_fd): the final image and the
ray samples - what the reference hands to egui (src/main_view.rs:114-115) and to the reference view's vertex buffer
(src/reference_view.rs:381-388) - live in exportable device allocations; a consumer maps the file descriptor and sees every
later frame without a copy.  The consumer here is CUDA itself: a second mapping in this process and a SECOND PROCESS (its own
CUDA context) that inherits the descriptor.  -m gpu."""
import ctypes as C
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

from linon_b200 import _capi as capi
from linon_b200.main_view import MainView, PredefinedFunction, make_settings, read_exported_buffer

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = """
import hashlib, sys
sys.path.insert(0, %r)
from linon_b200.main_view import read_exported_buffer
fd, size, nbytes = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
print(hashlib.sha256(read_exported_buffer(0, fd, size, nbytes).tobytes()).hexdigest())
"""


def render(mv, fn, settings, flags=capi.RENDER_ADAPTIVE_OFF):
    mv.reload_shader(fn.to_code())
    mv.update_settings(settings)
    mv.render_flags(flags)
    mv.sync()


@pytest.mark.parametrize("buffer", [capi.BUF_FINAL_RGBA8, capi.BUF_RAY_SAMPLES])
def test_exported_buffers_are_seen_by_another_mapping_and_another_process(product_lib, scene, buffer):
    verts, faces, center = scene
    W, H = 160, 120
    L = capi.lib()
    with MainView(verts, faces, center, W, H) as mv:
        render(mv, PredefinedFunction.MirageSphericalSigmoid, make_settings(overlay_mode=1))
        before = mv.read(buffer).copy()
        nbytes = mv.buffer_bytes(buffer)
        fd, size = mv.export_buffer_fd(buffer)
        try:
            assert fd >= 0 and size >= nbytes
            assert np.array_equal(mv.read(buffer), before)  # the move into the exportable allocation kept the contents
            # (1) a second mapping in this process, kept open across a NEW frame: zero copy means it shows the new frame
            ptr, handle = C.c_void_p(0), C.c_void_p(0)
            assert L.linon_import_buffer_fd(0, fd, size, C.byref(ptr), C.byref(handle)) == capi.OK
            try:
                seen = np.empty(nbytes, np.uint8)
                assert L.linon_copy_to_host(0, ptr, seen.ctypes.data, nbytes) == capi.OK
                assert np.array_equal(seen, before.view(np.uint8).ravel())
                render(mv, PredefinedFunction.MiragePlane, make_settings(overlay_mode=1, mouse_pos=(0.3, 0.4)))
                after = mv.read(buffer).copy()
                assert not np.array_equal(after, before)
                assert L.linon_copy_to_host(0, ptr, seen.ctypes.data, nbytes) == capi.OK
                assert np.array_equal(seen, after.view(np.uint8).ravel())
            finally:
                assert L.linon_release_import(handle) == capi.OK
            # (2) another process (its own CUDA context) that inherits the descriptor
            r = subprocess.run([sys.executable, "-c", CHILD % ROOT, str(fd), str(size), str(nbytes)], pass_fds=[fd],
                               capture_output=True, text=True, timeout=300)
            assert r.returncode == 0, r.stderr[-800:]
            assert r.stdout.strip().splitlines()[-1] == hashlib.sha256(after.tobytes()).hexdigest()
            # (3) every export call hands out a new descriptor of the same allocation
            fd2, size2 = mv.export_buffer_fd(buffer)
            assert fd2 != fd and size2 == size
            assert np.array_equal(read_exported_buffer(0, fd2, size2, nbytes), after.view(np.uint8).ravel())
            os.close(fd2)
        finally:
            os.close(fd)
        # (4) rendering and the other readers are unaffected
        render(mv, PredefinedFunction.MirageSphericalSigmoid, make_settings(overlay_mode=1))
        assert np.array_equal(mv.read(buffer), before)


def test_resize_drops_the_export_and_other_buffers_are_refused(product_lib, scene):
    verts, faces, center = scene
    with MainView(verts, faces, center, 96, 64) as mv:
        render(mv, PredefinedFunction.TranslationX, make_settings(overlay_mode=1))
        fd, size = mv.export_buffer_fd(capi.BUF_FINAL_RGBA8)
        os.close(fd)
        mv.resize_texture(128, 96)  # like MainView::resize_texture: new textures, the host registers them again
        render(mv, PredefinedFunction.TranslationX, make_settings(overlay_mode=1))
        frame = mv.texture().copy()
        fd, size = mv.export_buffer_fd(capi.BUF_FINAL_RGBA8)
        try:
            assert size >= 128 * 96 * 4
            assert np.array_equal(read_exported_buffer(0, fd, size, frame.nbytes), frame.ravel())
        finally:
            os.close(fd)
        for b in (capi.BUF_RAYCAST_RGBA8, capi.BUF_ENDPOINT_F32X4, capi.BUF_HIT_ID_I32, capi.BUF_EXPONENTS_F32):
            with pytest.raises(Exception):
                mv.export_buffer_fd(b)
