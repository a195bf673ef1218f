"""The outline search's key logic (csrc/device/outline_keys.cuh: what outline_search_kernel / outline_finalize_kernel compute)
on the host against the oracle's eight serial scans (src/main_view.wgsl:381-468 with the intended stride W): random exponent
fields with many ties, single pixels, corners, nothing above the threshold, NaN - the eight positions must be identical."""
import ctypes as C
import hashlib
import math
import os
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import ROOT

_SRC = [os.path.join(ROOT, "tests", "emu", "outline_driver.cpp"),
        os.path.join(ROOT, "linon_b200", "csrc", "device", "outline_keys.cuh")]
CAM = dict(origin=np.array([0, 0, 2], np.float32), view_direction=np.array([0, 0, -1], np.float32), up=np.array([0, 1, 0], np.float32))
ST = dict(field_weight=1.0, mouse_x=0.5, mouse_y=0.6, overlay_mode=1, central_difference_delta=1, lyapunov_scaling=50.0)


@pytest.fixture(scope="module")
def emu_outline():
    key = hashlib.sha1(b"".join(open(p, "rb").read() for p in _SRC)).hexdigest()[:16]
    cache = os.path.join(tempfile.gettempdir(), "linon_emu_cache")
    os.makedirs(cache, exist_ok=True)
    so = os.path.join(cache, f"liboutline_{key}.so")
    if not os.path.exists(so):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-w", "-I", os.path.dirname(_SRC[1]),
                        _SRC[0], "-o", so + ".tmp"], check=True)
        os.replace(so + ".tmp", so)
    lib = C.CDLL(so)

    def run(ex):
        H, W = ex.shape
        ex = np.ascontiguousarray(ex, np.float32)
        n = W * H
        stride = next(s for s in range(max(n // 3, 1) | 1, 2 * n + 3) if math.gcd(s, n) == 1)  # visits every pixel once
        pos = np.zeros((8, 2), np.float32)
        lib.emu_outline_search(W, H, C.c_void_p(ex.ctypes.data), C.c_void_p(pos.ctypes.data), C.c_ulonglong(stride))
        return pos

    return run


def _check(emu_outline, oracle, ex):
    H, W = ex.shape
    ref = oracle.sample_positions(CAM, ST, W, H, outline=True, exponents=ex)
    got = emu_outline(ex)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), (got, ref)


@pytest.mark.parametrize("size", [(64, 48), (97, 33), (33, 97), (240, 135), (2, 2), (5, 3)])
@pytest.mark.parametrize("density", [0.5, 0.05, 0.002])
def test_outline_keys_pick_the_serial_scans_pixels(emu_outline, oracle, size, density):
    W, H = size
    for seed in range(8):
        rng = np.random.default_rng(seed * 7919 + W * 31 + H)
        # two values only above the threshold: every extreme has many ties, so the tie-breaking order decides
        ex = np.where(rng.random((H, W)) < density, rng.choice(np.array([0.8, 3.0], np.float32), (H, W)),
                      rng.random((H, W)).astype(np.float32) * 0.79).astype(np.float32)
        _check(emu_outline, oracle, ex)


def test_outline_keys_edge_cases(emu_outline, oracle):
    W, H = 50, 40
    _check(emu_outline, oracle, np.zeros((H, W), np.float32))  # nothing found: pixel (0, 0) -> (0, 1)
    assert np.array_equal(emu_outline(np.zeros((H, W), np.float32)), np.tile(np.array([0.0, 1.0], np.float32), (8, 1)))
    _check(emu_outline, oracle, np.full((H, W), 1.0, np.float32))  # everything found
    for x, y in [(0, 0), (W - 1, 0), (0, H - 1), (W - 1, H - 1), (W // 2, H // 2), (1, H - 2)]:
        ex = np.zeros((H, W), np.float32)
        ex[y, x] = 0.8  # the threshold itself counts (>=)
        _check(emu_outline, oracle, ex)
    ex = np.full((H, W), np.nan, np.float32)  # NaN is not >= 0.8
    ex[7, 9] = np.inf
    _check(emu_outline, oracle, ex)
    # anti-diagonal and diagonal: all sums equal, only the scan order separates the candidates
    ex = np.zeros((40, 40), np.float32)
    ex[np.arange(40), np.arange(40)[::-1]] = 1.0
    _check(emu_outline, oracle, ex)
    ex = np.zeros((40, 40), np.float32)
    ex[np.arange(40), np.arange(40)] = 1.0
    _check(emu_outline, oracle, ex)
