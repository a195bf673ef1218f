"""The two independent restatements (C++ oracle, numpy oracle) must agree, and the scene tables must equal the
reference's own data whenever /root/reference is mounted.  CPU only."""
import os
import re

import numpy as np
import pytest

from oracle import np_oracle as npo

REFERENCE = "/root/reference"
NO_TRANSCENDENTALS = (1, 2, 5, 6, 8, 9)  # bodies whose arithmetic is + - * / sqrt only -> bit-exact across restatements


@pytest.fixture(scope="module")
def frame():
    verts, faces, center = npo.default_scene()
    W, H = 32, 24
    cam = npo.camera_uniform(npo.default_camera(center, W, H))
    ys, xs = np.mgrid[0:H, 0:W]
    return verts, faces, cam, npo.default_settings(), W, H, xs.ravel(), ys.ravel()


@pytest.mark.parametrize("fn", range(1, 10))
@pytest.mark.parametrize("adaptive", [False, True])
def test_cpp_vs_numpy_march(oracle, frame, fn, adaptive):
    verts, faces, cam, st, W, H, xs, ys = frame
    a = oracle.march(verts, faces, cam, st, fn, W, H, h_initial=0.1, adaptive=adaptive)
    b = npo.march(verts, faces, cam, st, fn, W, H, xs, ys, h_initial=0.1, adaptive=adaptive)
    assert (a["hit_id"].ravel() == b["hit_id"]).all()
    assert np.abs(a["rgba8"].reshape(-1, 4).astype(int) - b["rgba8"].astype(int)).max() <= 1
    if fn in NO_TRANSCENDENTALS:
        assert (a["endpoint"].reshape(-1, 4)[:, :3] == b["endpoint"]).all()
        assert (a["accepted"].ravel() == b["accepted"]).all()
        assert (a["rejected"].ravel() == b["rejected"]).all()
        assert (a["prehit"].ravel() == b["prehit"]).all()
    elif not adaptive:
        # exp/sin/cos differ by an ulp between glibc and numpy; fixed step keeps that from growing
        assert np.allclose(a["endpoint"].reshape(-1, 4)[:, :3], b["endpoint"], rtol=0, atol=1e-4)
        assert (a["accepted"] == 51).all()


def test_step_counts_f2(oracle, frame):
    """h=0.1 -> 51 accepted steps; h=0.001 -> 5001, both from fp32 accumulation of t."""
    verts, faces, cam, st, W, H, xs, ys = frame
    a = oracle.march(verts, faces, cam, st, 5, W, H, h_initial=0.1, adaptive=False, rect=(0, 4, 0, 2))
    assert (a["accepted"][0:2, 0:4] == 51).all()
    b = oracle.march(verts, faces, cam, st, 5, W, H, h_initial=0.001, adaptive=True, rect=(0, 2, 0, 1))
    assert (b["accepted"][0, 0:2] == 5001).all() and (b["rejected"][0, 0:2] == 0).all()


def test_field_weight_zero_is_a_straight_ray(oracle, frame):
    """field_weight = 0 keeps cur_dir = start_dir: endpoint = origin + sum(h) * dir and hit = plain ray cast."""
    verts, faces, cam, st, W, H, xs, ys = frame
    st = dict(st, field_weight=0.0)
    a = oracle.march(verts, faces, cam, st, 1, W, H, h_initial=0.1, adaptive=False)
    origin, dirs = npo.pixel_dirs(cam, W, H, xs, ys)
    expect = origin[None, :] + np.float32(5.1) * dirs
    assert np.allclose(a["endpoint"].reshape(-1, 4)[:, :3], expect, atol=2e-5)
    col, hid = npo.ray_color(verts, faces, np.broadcast_to(origin, dirs.shape).copy(), dirs, np.float32(5.1))
    assert (npo_canon(hid) == npo_canon(a["hit_id"].ravel())).mean() > 0.99


def npo_canon(h):
    g = np.arange(36) // 2
    g[[2, 3, 4, 5]] = 0
    g[[8, 9]] = 3
    out = np.full(h.shape, -1)
    out[h >= 0] = g[h[h >= 0]]
    return out


def test_cpp_vs_numpy_overlay(oracle, frame):
    verts, faces, cam, st, W, H, xs, ys = frame
    a = oracle.march(verts, faces, cam, st, 3, W, H, h_initial=0.1, adaptive=False)
    for delta in (1, 2, 3):
        s = dict(st, overlay_mode=1, central_difference_delta=delta, lyapunov_scaling=50.0)
        f1, e1 = oracle.overlay(a["rgba8"], a["endpoint"], s)
        f2, e2 = npo.overlay(a["rgba8"], a["endpoint"], s)
        assert np.allclose(e1, e2, rtol=1e-5, atol=1e-7)
        assert np.abs(f1.astype(int) - f2.astype(int)).max() <= 1
        pad = 2 * delta  # border keeps the previous exponent (zero) and the unblended colour
        assert (e1[:pad] == 0).all() and (e1[:, :pad] == 0).all() and (e1[-pad:] == 0).all() and (e1[:, -pad:] == 0).all()
        assert (f1[:pad] == a["rgba8"][:pad]).all()
    s0 = dict(st, overlay_mode=0)
    f0, e0 = oracle.overlay(a["rgba8"], a["endpoint"], s0)
    assert (f0 == a["rgba8"]).all() and (e0 == 0).all()


def test_vec2squareroots_branches(oracle):
    # real roots: x^2 - 3x + 2 -> (1, 2); complex: x^2 + 2x + 5 -> -1 +- 2i -> (re, im) = (-1, 2)
    r = oracle.vec2squareroots(np.array([[2.0, -3.0], [5.0, 2.0], [0.0, 0.0]], np.float32))
    assert np.allclose(r, [[1.0, 2.0], [-1.0, 2.0], [0.0, 0.0]])


def test_hit_triangle_edges(oracle):
    tri = np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0]], np.float32)
    cases = [
        ((0.25, 0.25, 1.0), (0, 0, -1), 1.0),      # plain hit from the front
        ((0.25, 0.25, -1.0), (0, 0, 1), 1.0),      # no back-face culling (src/main_view.wgsl:65,156)
        ((0.25, 0.25, 1.0), (0, 0, 1), -1.0),      # behind the origin: t <= eps
        ((0.25, 0.25, 1.0), (1, 0, 0), -1.0),      # parallel: |a| < eps
        ((2.0, 0.25, 1.0), (0, 0, -1), -1.0),      # u > 1
        ((0.75, 0.75, 1.0), (0, 0, -1), -1.0),     # u + v > 1
        ((-0.1, 0.25, 1.0), (0, 0, -1), -1.0),     # u < 0
        ((0.0, 0.0, 1.0), (0, 0, -1), 1.0),        # vertex: u = v = 0 is inside
        ((0.25, 0.25, 0.0), (0, 0, -1), -1.0),     # origin on the plane: t = 0 is not > eps
    ]
    o = np.array([c[0] for c in cases], np.float32)
    d = np.array([c[1] for c in cases], np.float32)
    t = oracle.hit_triangle(np.repeat(tri, len(cases), 0), o, d)
    assert np.allclose(t, [c[2] for c in cases])


def test_helper_quirks(oracle):
    v = np.array([[1.0, 2.0, 3.0]], np.float32)
    # rotateZ drops z (zero third column)
    r = oracle.helper("rotateZ", 1, v=v, a=np.array([0.0], np.float32))
    assert np.allclose(r, [[1.0, 2.0, 0.0]])
    r = oracle.helper("rotateZ", 1, v=v, a=np.array([np.pi / 2], np.float32))
    assert np.allclose(r, [[-2.0, 1.0, 0.0]], atol=1e-6)
    r = oracle.helper("rotateX", 1, v=v, a=np.array([np.pi / 2], np.float32))
    assert np.allclose(r, [[1.0, -3.0, 2.0]], atol=1e-6)
    r = oracle.helper("rotateY", 1, v=v, a=np.array([np.pi / 2], np.float32))
    assert np.allclose(r, [[3.0, 2.0, -1.0]], atol=1e-6)
    # misordered clamp has no lower bound: dot = -4 stays -4 -> cosi = 4 -> k = 1 - eta^2 (1 - 16) > 0
    n = np.array([[0.0, 0.0, 1.0]], np.float32)
    vin = np.array([[0.0, 0.0, -4.0]], np.float32)
    r = oracle.helper("refraction", 1, v=vin, nrm=n, a=np.array([15.0], np.float32), b=np.array([15.0], np.float32))
    # eta = 1: result = v_in + (cosi - sqrt(k)) * n with cosi = 4, k = 16 -> v_in
    assert np.allclose(r, vin)
    # refraction index of air at 15 C is about 1.000277 above 1 -> n - 1 here (the formula returns n - 1 scaled)
    ri = oracle.helper("refraction_index", 1, a=np.array([15.0], np.float32))[0, 0]
    assert 0.9 < ri < 1.1
    assert np.isclose(oracle.helper("sigmoid", 1, a=np.array([0.0], np.float32))[0, 0], 0.5)


def test_outline_positions(oracle):
    W, H = 40, 30
    cam = dict(origin=[0, 0, 1], view_direction=[0, 0, -1], up=[0, 1, 0])
    st = npo.default_settings()
    e = np.zeros((H, W), np.float32)
    pos = oracle.sample_positions(cam, st, W, H, outline=True, exponents=e)
    assert np.allclose(pos, [[0.0, 1.0]] * 8)  # nothing found: pixel (0,0) -> (0, 1)
    e[10:20, 15:25] = 1.0  # a square blob: rows 10..19, cols 15..24
    pos = oracle.sample_positions(cam, st, W, H, outline=True, exponents=e)
    px = np.rint(pos[:, 0] * (W - 1)).astype(int)
    py = np.rint((H - 1) - pos[:, 1] * (H - 1)).astype(int)
    got = list(zip(px, py))
    assert got == [(15, 19), (15, 19), (24, 19), (24, 10), (24, 10), (15, 10), (15, 10), (15, 10)]
    click = oracle.sample_positions(cam, st, W, H, outline=False)
    assert np.allclose(click[0], [0.49, 0.59]) and np.allclose(click[4], [0.51, 0.61])


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "src")), reason="reference mount not present")
def test_scene_tables_equal_reference_sources():
    def table(text, name):
        body = text.split(name)[1].split("[", 2)[2].split("];")[0]
        body = re.sub(r"//.*", "", body)
        return [x for x in body.replace("\n", " ").split(",") if x.strip()]

    src = open(os.path.join(REFERENCE, "src/cornell_box.rs")).read()
    ref_v = np.array([float(x) for x in table(src, "VERTICES")], np.float32).reshape(-1, 3)
    ref_i = np.array([int(x) for x in table(src, "INDICES")], np.uint32).reshape(-1, 3)
    v, f = npo.cornell_box()
    assert v.shape == ref_v.shape and (v == ref_v).all()
    assert f.shape == ref_i.shape and (f == ref_i).all()


def test_default_scene_and_camera_facts():
    """SURVEY F12 / F13."""
    verts, faces, center = npo.default_scene()
    assert np.allclose(verts.min(0), [-0.99422383, 0.0093863, -1.0], atol=1e-7)
    assert np.allclose(verts.max(0), [0.0093863, 1.0, 0.0093863], atol=1e-7)
    assert np.allclose(center, [-0.49241877, 0.50469315, -0.49530685], atol=1e-7)
    cam = npo.camera_uniform(npo.default_camera(center, 800, 600))
    assert np.allclose(cam["origin"], center + np.array([0, 0, 2], np.float32), atol=1e-6)
    assert np.allclose(cam["view_direction"], [0, 0, -1]) and np.allclose(cam["up"], [0, 1, 0])
    # 64 drags of the orbit come back to the start (each is 360/64 degrees about the up axis)
    c64 = npo.camera_uniform(npo.orbit_camera(center, 800, 600, 64))
    assert np.allclose(c64["origin"], cam["origin"], atol=2e-4)
    c16 = npo.camera_uniform(npo.orbit_camera(center, 800, 600, 16))
    assert abs(np.dot(c16["view_direction"], cam["view_direction"])) < 1e-4  # a quarter turn
