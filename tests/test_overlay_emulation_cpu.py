"""Both overlay kernels on the host - the measured default (overlay_kernel, static_kernels.cu: shared-memory stage, one
barrier) and the opt-in direct variant (LINON_OPT_OVERLAY_KERNEL = 1, csrc/device/overlay_direct.cuh): the device text is
compiled by g++ (tests/emu/overlay_staged_driver.cpp, overlay_driver.cpp) and run over the launcher's grid, then compared BIT
FOR BIT with the oracle's overlay (src/overlay.wgsl:99-146) - final colours, exponents, untouched border, mode 0 copy - on ragged frame
sizes, every delta of the GUI slider, and endpoint fields with smooth, chaotic, constant, NaN and infinite regions."""
import ctypes as C
import hashlib
import os
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import ROOT

_DEVICE = os.path.join(ROOT, "linon_b200", "csrc", "device")
_EMU = os.path.join(ROOT, "tests", "emu")


def _staged_kernel_text():
    """overlay_kernel as static_kernels.cu has it: from its tile constants to the end of the kernel."""
    s = open(os.path.join(_DEVICE, "static_kernels.cu")).read()
    a, b = s.index("constexpr int OV_TX = 32;"), s.index('}  // namespace linon\n#include "overlay_direct.cuh"')
    text = s[a:b]
    assert "__global__" in text and "overlay_kernel(" in text and text.count("__syncthreads()") == 1
    return text


def _build(kind):
    cache = os.path.join(tempfile.gettempdir(), "linon_emu_cache")
    os.makedirs(cache, exist_ok=True)
    if kind == "direct":
        driver, extra = os.path.join(_EMU, "overlay_driver.cpp"), open(os.path.join(_DEVICE, "overlay_direct.cuh"), "rb").read()
        flags = ["-fopenmp", "-I", _DEVICE]
    elif kind in ("tile", "tile_tall"):
        kind = "tile"
        driver = os.path.join(_EMU, "overlay_tile_driver.cpp")
        extra = open(os.path.join(_DEVICE, "overlay_direct.cuh"), "rb").read() + open(os.path.join(_DEVICE, "overlay_tile.cuh"), "rb").read()
        flags = ["-fopenmp", "-I", _DEVICE]
    else:
        driver, extra = os.path.join(_EMU, "overlay_staged_driver.cpp"), _staged_kernel_text().encode()
        flags = []
    key = hashlib.sha1(open(driver, "rb").read() + extra).hexdigest()[:16]
    so = os.path.join(cache, f"libov_{kind}_{key}.so")
    if not os.path.exists(so):
        if kind == "staged":
            inc = os.path.join(cache, f"ov_staged_{key}.inc")
            with open(inc, "wb") as f:
                f.write(extra)
            flags = [f'-DLINON_OV_TEXT="{inc}"']
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-w", *flags, driver, "-o", so + ".tmp"],
                       check=True)
        os.replace(so + ".tmp", so)
    return C.CDLL(so)


@pytest.fixture(scope="module", params=["tile", "tile_tall", "direct", "staged"])
def emu_overlay(request):
    """tile: overlay_tile_kernel's per-thread body (the default) on tiles filled the way its TMA load fills them
    (tests/emu/overlay_tile_driver.cpp).  direct: overlay_direct_kernel's per-thread body.  staged: the measured default overlay_kernel, its text
    cut out of static_kernels.cu and run CTA by CTA in two passes around its single barrier (tests/emu/overlay_staged_driver.cpp)."""
    lib = _build(request.param)
    tall = request.param == "tile_tall"
    if request.param.startswith("tile"):
        def entry(W, H, mode, delta, *rest):
            # 32 x 64 tiles exist for delta <= 4 (static shared memory); the launcher uses 32 x 32 tiles beyond that
            rc = lib.emu_overlay_tile(W, H, mode, delta, *rest, 1 if (tall and delta <= 4) else 0)
            assert rc == 0
    else:
        entry = lib.emu_overlay_direct if request.param == "direct" else lib.emu_overlay_staged

    def run(raycast, endpoint, mode, delta, scaling, exponents=None):
        H, W = raycast.shape[:2]
        raycast = np.ascontiguousarray(raycast, np.uint8)
        endpoint = np.ascontiguousarray(endpoint, np.float32)
        final = np.full_like(raycast, 0xAB)  # every pixel must be written
        ex = np.zeros((H, W), np.float32) if exponents is None else np.array(exponents, np.float32, copy=True)
        entry(W, H, int(mode), int(delta), C.c_float(scaling), C.c_void_p(raycast.ctypes.data),
              C.c_void_p(endpoint.ctypes.data), C.c_void_p(final.ctypes.data), C.c_void_p(ex.ctypes.data))
        return final, ex

    return run


def _field(W, H, seed):
    """Endpoints like a frame's: smooth over most of the image, a chaotic patch, a constant patch (zero gradient: the
    square roots get 0), and a few non-finite texels."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:H, 0:W].astype(np.float32)
    ep = np.stack([np.sin(x * 0.011) + 0.002 * y, np.cos(y * 0.013) * 0.9 + 0.001 * x, 0.3 * np.sin((x + y) * 0.007),
                   np.ones_like(x)], axis=-1).astype(np.float32)
    h0, w0 = H // 3, W // 3
    ep[h0:h0 + H // 4, w0:w0 + W // 4, :3] += rng.normal(0, 0.2, (H // 4, W // 4, 3)).astype(np.float32)
    ep[-H // 5:, :W // 4, :3] = np.float32(0.25)
    if W > 40 and H > 40:
        ep[H // 2, W // 2, 0] = np.nan
        ep[H // 2 + 7, W // 2 + 9, 1] = np.inf
        ep[H // 2 - 9, W // 2 - 5, 2] = -np.inf
    raycast = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
    return raycast, ep


def _same_bits(a, b):
    """Same 32-bit words; two NaNs count as equal whatever their sign / payload (x86 keeps the first operand's, the device
    writes its canonical NaN - neither is part of the reference's definition)."""
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    return bool(np.all((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))))


@pytest.mark.parametrize("delta", [1, 2, 3, 5, 10])
@pytest.mark.parametrize("size", [(200, 150), (97, 33), (64, 48)])
def test_overlay_kernels_are_the_oracles_overlay_bit_for_bit(emu_overlay, oracle, delta, size):
    W, H = size
    raycast, ep = _field(W, H, seed=delta * 1000 + W)
    st = dict(field_weight=1.0, mouse_x=0.5, mouse_y=0.6, overlay_mode=1, central_difference_delta=delta, lyapunov_scaling=50.0)
    ref_final, ref_ex = oracle.overlay(raycast, ep, st)
    final, ex = emu_overlay(raycast, ep, 1, delta, 50.0)
    assert np.array_equal(final, ref_final)
    assert _same_bits(ex, ref_ex)
    if 4 * delta < min(W, H):
        assert (ref_ex >= 0.8).any() and (final != raycast).any()  # the blend branch ran
        inner = ex[2 * delta:H - 2 * delta, 2 * delta:W - 2 * delta]
        assert (inner < 0.8).any()  # and so did the pass-through of an inner pixel


@pytest.mark.parametrize("scaling", [0.0, 7.5, 500.0])
def test_overlay_kernels_scaling_values(emu_overlay, oracle, scaling):
    raycast, ep = _field(120, 90, seed=3)
    st = dict(field_weight=1.0, mouse_x=0.5, mouse_y=0.6, overlay_mode=1, central_difference_delta=1, lyapunov_scaling=scaling)
    ref_final, ref_ex = oracle.overlay(raycast, ep, st)
    final, ex = emu_overlay(raycast, ep, 1, 1, scaling)
    assert np.array_equal(final, ref_final) and _same_bits(ex, ref_ex)


@pytest.mark.parametrize("size", [(1, 1), (3, 2), (4, 4), (5, 5), (33, 17), (40, 41)])
def test_overlay_kernels_frames_smaller_than_the_border(emu_overlay, oracle, size):
    """Frames in which no or few pixels are inner (W, H <= 4 delta): every tap stays inside the image."""
    W, H = size
    rng = np.random.default_rng(W * 100 + H)
    raycast = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
    ep = rng.normal(0, 1, (H, W, 4)).astype(np.float32)
    for delta in (1, 10):
        st = dict(field_weight=1.0, mouse_x=0.5, mouse_y=0.6, overlay_mode=1, central_difference_delta=delta, lyapunov_scaling=50.0)
        ref_final, ref_ex = oracle.overlay(raycast, ep, st)
        final, ex = emu_overlay(raycast, ep, 1, delta, 50.0)
        assert np.array_equal(final, ref_final) and _same_bits(ex, ref_ex)


def test_overlay_kernels_mode_0_copies_and_leaves_the_exponents_alone(emu_overlay, oracle):
    raycast, ep = _field(130, 70, seed=9)
    before = np.random.default_rng(1).random((70, 130)).astype(np.float32)
    final, ex = emu_overlay(raycast, ep, 0, 2, 50.0, exponents=before)
    assert np.array_equal(final, raycast) and _same_bits(ex, before)
    st = dict(field_weight=1.0, mouse_x=0.5, mouse_y=0.6, overlay_mode=0, central_difference_delta=2, lyapunov_scaling=50.0)
    ref_final, ref_ex = oracle.overlay(raycast, ep, st, exponents=before)
    assert np.array_equal(final, ref_final) and _same_bits(ex, ref_ex)


def test_overlay_kernels_at_the_c3_frame_size(emu_overlay, oracle):
    """BASELINE config 3's frame size (3840x2160, delta 1, scaling 50): 16 200 CTAs of the launcher's grid, bit-identical."""
    W, H = 3840, 2160
    raycast, ep = _field(W, H, seed=42)
    st = dict(field_weight=1.0, mouse_x=0.5, mouse_y=0.6, overlay_mode=1, central_difference_delta=1, lyapunov_scaling=50.0)
    ref_final, ref_ex = oracle.overlay(raycast, ep, st)
    final, ex = emu_overlay(raycast, ep, 1, 1, 50.0)
    assert np.array_equal(final, ref_final) and _same_bits(ex, ref_ex)
    assert (ref_ex >= 0.8).mean() > 0.01
