"""The oracle is pinned against the reference's OWN shader text, executed here.  CPU only.

niklaskorz/linon cannot be built in this image (Rust + wgpu), but its hot path is WGSL.
oracle/wgsl_ref/ transpiles src/main_view.wgsl (with every predefined body of src/functions.rs spliced in exactly like
`with_field_function`, src/main_view.rs:852-861, and the Enhance / adaptive-off / outline variants made by the same
string replacements as MainView) and src/overlay.wgsl mechanically to C++ and compiles them with gcc
(oracle/_ref/liblinon_wgsl_ref.so).  These tests require the hand-written oracle (oracle/linon_oracle.cpp) — the
checker of every GPU parity test — to agree with that library BIT FOR BIT: colour (before and after the rgba8unorm
store), endpoint, ray samples, exponents, final image, for all ten modules, in fixed-step, adaptive and Enhance mode.
The committed golden fixtures (tests/golden/) are checked against the transpiled reference as well, so the GPU test
that compares the CUDA path with those fixtures compares it with outputs of the reference's shader source.

Skipped when neither /root/reference nor a prebuilt oracle/_ref library exists.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import np_oracle as npo
from oracle import wgsl_reference as wr
from oracle.wgsl_ref import wgsl2cpp

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
needs_reference = pytest.mark.skipif(not wr.available(), reason="/root/reference not mounted and oracle/_ref not built")


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.fixture(scope="module")
def frame96():
    W, H = 96, 72
    verts, faces, center = npo.default_scene()
    return W, H, verts, faces, npo.camera_uniform(npo.default_camera(center, W, H)), npo.default_settings()


@pytest.fixture(scope="module")
def frame64():
    W, H = 64, 64  # square: the reference's stride-H indexing of `exponents` coincides with stride W
    verts, faces, center = npo.default_scene()
    return W, H, verts, faces, npo.camera_uniform(npo.default_camera(center, W, H)), npo.default_settings()


# ---- the transpiler itself ---------------------------------------------------------------------------------------------
def test_abstract_literals_are_evaluated_in_f64_and_concretised_where_they_meet_f32():
    cpp = wgsl2cpp.transpile("""
fn f(h: f32, k: vec3<f32>) -> vec3<f32> {
    let beta = 8.0 / 3.0;
    let n = 100;
    var s: f32 = 1;
    let c = 0.5 * h * k;
    let p = pow(10.0, -10.0) * h;
    let q = max(h, 0.0);
    return c + vec3<f32>(beta, 1.0 / 2.4, f32(n)) * p * q * s;
}""")
    assert "const float beta = float((8.0 / 3.0));" in cpp          # const-evaluated in double, one conversion
    assert "const int32_t n = 100;" in cpp
    assert "float s = float(1);" in cpp
    assert "((float(0.5) * h) * k)" in cpp                          # left-associative, literal takes the f32 type
    assert "float(wgsl::pow(double(10.0), double((-10.0))))" in cpp   # literal-only call: abstract-float evaluation
    assert "wgsl::max(h, float(0.0))" in cpp
    assert "float((1.0 / 2.4))" in cpp


def test_precedence_swizzles_indices_and_control_flow_survive():
    cpp = wgsl2cpp.transpile("""
struct S { a: u32, color: vec4<f32>, };
fn g(v: vec4<f32>, s: S, m: mat2x2<f32>) -> f32 {
    var t: f32 = -1.0;
    for (var i: u32 = 0u; i < s.a; i = i + 1u) {
        if (t > 0.0 && t < 2.0 || v.x == 1.0) { continue; }
        t = t + v.rgb.y * m[1].x - s.color.a;
    }
    return t;
}""")
    assert "for (uint32_t i = 0u; (i < s.a); i = (i + 1u))" in cpp
    assert "(((t > float(0.0)) && (t < float(2.0))) || (v.x == float(1.0)))" in cpp   # && binds tighter than ||
    assert "wgsl::swz<0, 1, 2>(v).y" in cpp and "m[1].x" in cpp and "s.color.a" in cpp
    assert "continue;" in cpp


def test_unsupported_constructs_are_refused_not_guessed():
    for src in ("fn f() { discard; }", "fn f() { loop { continuing { } var x = 1; } }", "alias T = f32;",
                "fn f() -> f32 { return unknown_fn(1.0); }", "fn f() { continuing { } }"):
        with pytest.raises(wgsl2cpp.WgslError):
            wgsl2cpp.transpile(src)


def test_loop_continuing_and_switch_keep_wgsl_semantics():
    """Round 2: the transpiler follows WGSL 9.4 for `loop { ... continuing { ... break if e; } }` (a `continue` reaches the
    continuing block, a `break` of the loop does not) and `switch` (several selectors per clause, no fall-through)."""
    cpp = wgsl2cpp.transpile("""
fn f(n: i32) -> i32 {
    var i = 0; var s = 0;
    loop {
        if (i == 2) { continue; }
        if (i > n) { break; }
        s = s + i;
        continuing { i = i + 1; break if i > 100; }
    }
    switch (s % 3) { case 0, 1: { s = s + 10; } default: { s = s - 1; } }
    return s;
}""")
    assert "do" in cpp and "while (false);" in cpp and "wgsl_loop_break_" in cpp and "if ((i > 100)) break;" in cpp
    assert "case 0:" in cpp and "case 1:" in cpp and "default:" in cpp and cpp.count("break;") >= 4
    assert "wgsl::mod(s, 3)" in cpp


def test_a_transpiled_snippet_computes_what_wgsl_says(tmp_path):
    """End to end through g++: abstract-float folding, u32 wrap-around, vector arithmetic, array indexing."""
    cpp = wgsl2cpp.transpile("""
var<private> table: array<vec2<f32>, 2> = array<vec2<f32>, 2>(vec2<f32>(1.0, 2.0), vec2<f32>(3.0, 4.0));
fn f(x: f32) -> f32 {
    let third = 1.0 / 3.0;
    var u: u32 = 0u;
    u = u - 1u;
    let v = table[1] * x + table[0];
    return third + f32(u >> 31u) + dot(v, vec2<f32>(1.0, 1.0));
}""")
    src = tmp_path / "snippet.cpp"
    here = os.path.dirname(wgsl2cpp.__file__)
    src.write_text('#include "wgsl_runtime.hpp"\n#include <cstdio>\nnamespace m {\n' + cpp +
                   '}\nint main() { std::printf("%.9g\\n", m::f(0.5f)); return 0; }\n')
    exe = tmp_path / "snippet"
    gxx = "/usr/bin/g++" if os.access("/usr/bin/g++", os.X_OK) else "g++"
    subprocess.run([gxx, "-std=c++17", "-O2", "-ffp-contract=off", "-I", here, str(src), "-o", str(exe)], check=True)
    got = np.float32(subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout)
    third = np.float32(1.0 / 3.0)
    v = np.array([3.0, 4.0], np.float32) * np.float32(0.5) + np.array([1.0, 2.0], np.float32)
    want = np.float32(np.float32(third + np.float32(1.0)) + np.float32(v[0] + v[1]))
    assert got == want


# ---- the recipe ----------------------------------------------------------------------------------------------------------
@needs_reference
def test_reference_library_records_its_provenance():
    info = wr.info()
    assert set(info["sources"]) == {"src/main_view.wgsl", "src/overlay.wgsl", "src/functions.rs"}
    assert "-ffp-contract=off" in info["cxxflags"]
    names = [v["name"] for v in info["variants"]]
    assert names == ["default", "enhance", "adaptive_off", "outline"]
    # Enhance is the one replacement of the reference that matches its own shader; the other two are dead
    assert info["variants"][1]["replacements"][0]["spelling"] == "reference"
    assert info["variants"][2]["replacements"][0]["spelling"].startswith("intended")
    assert info["variants"][3]["replacements"][0]["spelling"].startswith("intended")


def test_the_nine_bodies_are_read_from_functions_rs(has_reference):
    if not has_reference:
        pytest.skip("reference not mounted")
    from linon_b200.main_view import PredefinedFunction
    from oracle.wgsl_ref import build_ref
    bodies = build_ref.predefined_bodies(open("/root/reference/src/functions.rs", encoding="utf-8").read())
    for fn in PredefinedFunction:
        assert bodies[fn.name] == fn.to_code(), fn  # the product's table equals the Rust string literals byte for byte


# ---- oracle == transpiled reference, bit for bit -----------------------------------------------------------------------
@needs_reference
@pytest.mark.parametrize("fn", range(10))
@pytest.mark.parametrize("adaptive", [False, True])
def test_oracle_frames_equal_the_transpiled_reference(oracle, frame96, fn, adaptive):
    W, H, verts, faces, cam, st = frame96
    a = oracle.march(verts, faces, cam, st, fn, W, H, h_initial=0.1, adaptive=adaptive, want_color_f32=True)
    b = wr.march(verts, faces, cam, st, fn, W, H, h_initial=0.1, adaptive=adaptive, want_color_f32=True)
    assert np.array_equal(a["rgba8"], b["rgba8"])
    assert np.array_equal(bits(a["endpoint"]), bits(b["endpoint"]))
    assert np.array_equal(bits(a["color"]), bits(b["color"]))


@needs_reference
@pytest.mark.parametrize("fn", [1, 2, 3, 4, 5, 7, 8, 9])
def test_oracle_enhance_rays_equal_the_transpiled_reference(oracle, frame64, fn):
    """h_initial = 0.001 (src/main_view.rs:785-787): 5001 steps per ray, every 4th pixel of every 4th row."""
    W, H, verts, faces, cam, st = frame64
    kw = dict(h_initial=0.001, adaptive=True, x_step=4, y_step=4, want_color_f32=True)
    a = oracle.march(verts, faces, cam, st, fn, W, H, **kw)
    b = wr.march(verts, faces, cam, st, fn, W, H, **kw)
    assert (a["accepted"][::4, ::4] == 5001).all()
    assert np.array_equal(a["rgba8"], b["rgba8"])
    assert np.array_equal(bits(a["endpoint"]), bits(b["endpoint"]))
    assert np.array_equal(bits(a["color"]), bits(b["color"]))


@needs_reference
def test_oracle_frames_equal_the_transpiled_reference_for_other_cameras_and_field_weights(oracle):
    W, H = 80, 48
    verts, faces, center = npo.default_scene()
    for frame, fn, fw in ((5, 3, 1.0), (23, 2, 1.0), (40, 8, 0.05), (0, 9, 0.05), (11, 7, 0.5)):
        cam = npo.camera_uniform(npo.orbit_camera(center, W, H, frame))
        st = dict(npo.default_settings(), field_weight=fw)
        a = oracle.march(verts, faces, cam, st, fn, W, H, h_initial=0.1, adaptive=True)
        b = wr.march(verts, faces, cam, st, fn, W, H, h_initial=0.1, adaptive=True)
        assert np.array_equal(a["rgba8"], b["rgba8"]), (frame, fn)
        assert np.array_equal(bits(a["endpoint"]), bits(b["endpoint"])), (frame, fn)


@needs_reference
@pytest.mark.parametrize("delta", [1, 3, 10])
def test_oracle_overlay_equals_the_transpiled_reference(oracle, frame64, delta):
    W, H, verts, faces, cam, st = frame64
    for fn in (3, 1, 7):
        a = oracle.march(verts, faces, cam, st, fn, W, H, h_initial=0.1, adaptive=False)
        for mode in (0, 1):
            s = dict(st, overlay_mode=mode, central_difference_delta=delta, lyapunov_scaling=50.0)
            fa, ea = oracle.overlay(a["rgba8"], a["endpoint"], s)
            fb, eb = wr.overlay(a["rgba8"], a["endpoint"], s)
            assert np.array_equal(fa, fb), (fn, mode)
            assert np.array_equal(bits(ea), bits(eb)), (fn, mode)


@needs_reference
def test_exponent_stride_of_the_reference_is_the_height(oracle):
    """SURVEY F8, shown on the reference's own code: overlay.wgsl:128 writes exponents[size.y * y + x].  On a frame
    with W != H the transpiled reference therefore fills a stride-H layout; the oracle (and the CUDA path) use the
    intended stride W.  Same values, different index - and an identical final image."""
    W, H = 48, 80  # W < H so that every stride-H index stays inside the W*H buffer's first H*(H-1)+W entries
    verts, faces, center = npo.default_scene()
    cam = npo.camera_uniform(npo.default_camera(center, W, H))
    st = dict(npo.default_settings(), overlay_mode=1)
    a = oracle.march(verts, faces, cam, st, 3, W, H, h_initial=0.1, adaptive=False)
    fa, ea = oracle.overlay(a["rgba8"], a["endpoint"], st)
    big = np.zeros((H * H + W,), np.float32)  # room for the stride-H layout
    # the library is handed exactly W*H entries (the reference's buffer size, src/main_view.rs:146-151): writes beyond
    # it are clamped by the runtime, so compare only the rows whose stride-H index is in range
    fb, eb = wr.overlay(a["rgba8"], a["endpoint"], st)
    assert np.array_equal(fa, fb)
    flat = eb.reshape(-1)
    ys, xs = np.nonzero(ea)
    idx = H * ys + xs
    ok = idx < W * H - 1
    assert ok.sum() > 100
    assert np.array_equal(bits(flat[idx[ok]]), bits(ea[ys[ok], xs[ok]]))
    del big


@needs_reference
@pytest.mark.parametrize("fn", [3, 5, 7, 8])
def test_oracle_sample_rays_equal_the_transpiled_reference(oracle, frame64, fn):
    """Click mode (src/main_view.wgsl:469-479 + 281-314): the eight invocations gid.x < 8, gid.y == 0."""
    W, H, verts, faces, cam, st = frame64
    pos = oracle.sample_positions(cam, st, W, H)
    sa = oracle.sample_rays(cam, st, fn, W, H, pos)
    b = wr.main_view(verts, faces, cam, st, fn, W, H, mask=wr.sample_ray_mask(W, H))
    assert np.array_equal(bits(sa), bits(b["ray_samples"]))


@needs_reference
def test_oracle_outline_search_equals_the_transpiled_reference_on_a_square_frame(oracle, frame64):
    """Outline mode (src/main_view.wgsl:381-468) on W == H, where the reference's stride and start-row slips
    vanish: the oracle's search restated with the intended stride must pick the same eight pixels."""
    W, H, verts, faces, cam, st = frame64
    for fn in (3, 1):
        a = oracle.march(verts, faces, cam, st, fn, W, H, h_initial=0.1, adaptive=False)
        _, exps = oracle.overlay(a["rgba8"], a["endpoint"], dict(st, overlay_mode=1))
        assert (exps >= 0.8).sum() > 50
        pos = oracle.sample_positions(cam, st, W, H, outline=True, exponents=exps)
        sa = oracle.sample_rays(cam, st, fn, W, H, pos)
        b = wr.main_view(verts, faces, cam, st, fn, W, H, variant=wr.VARIANT_OUTLINE, mask=wr.sample_ray_mask(W, H),
                         exponents=exps)
        assert np.array_equal(bits(sa), bits(b["ray_samples"]))


@needs_reference
def test_oracle_helpers_and_field_functions_equal_the_transpiled_reference(oracle):
    rng = np.random.default_rng(7)
    n = 20000
    v = rng.normal(size=(n, 3)).astype(np.float32)
    nrm = rng.normal(size=(n, 3)).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    p0 = rng.normal(size=(n, 3)).astype(np.float32)
    a = rng.uniform(-50, 250, n).astype(np.float32)
    b = rng.uniform(-50, 250, n).astype(np.float32)
    c = rng.normal(size=n).astype(np.float32)
    for name in oracle.HELPERS:
        x = oracle.helper(name, n, v=v, nrm=nrm, p0=p0, a=a, b=b, c=c)
        y = wr.helper(name, n, v=v, nrm=nrm, p0=p0, a=a, b=b, c=c)
        assert np.array_equal(bits(x), bits(y)), name
    tri = rng.normal(size=(n, 9)).astype(np.float32)
    o = rng.normal(size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    ha, hb = oracle.hit_triangle(tri, o, d), wr.hit_triangle(tri, o, d)
    assert (ha > 0).sum() > 100 and np.array_equal(bits(ha), bits(hb))
    a01 = rng.normal(size=(n, 2)).astype(np.float32)
    assert np.array_equal(bits(oracle.vec2squareroots(a01)), bits(wr.vec2squareroots(a01)))
    st = npo.default_settings()
    for fn in range(1, 10):
        p = (rng.normal(size=(n, 3)) * 0.7).astype(np.float32)
        pp = p + (rng.normal(size=(n, 3)) * 0.01).astype(np.float32)
        t = rng.uniform(0, 5, n).astype(np.float32)
        assert np.array_equal(bits(oracle.field(fn, st, pp, p, v, nrm, t)), bits(wr.field(fn, st, pp, p, v, nrm, t))), fn


# ---- the committed fixtures are outputs of the reference's shader source -----------------------------------------------
@needs_reference
def test_golden_fixtures_are_reproduced_by_the_transpiled_reference():
    idx = json.load(open(os.path.join(GOLDEN, "index.json")))
    W, H = idx["width"], idx["height"]
    verts, faces, center = npo.default_scene()
    cam = npo.camera_uniform(npo.default_camera(center, W, H))
    assert idx.get("pinned_by") == "oracle/_ref (transpiled reference WGSL)"
    for entry in idx["frames"]:
        z = np.load(os.path.join(GOLDEN, entry["file"]))
        out = wr.march(verts, faces, cam, idx["settings"], entry["function_id"], W, H, h_initial=idx["h_initial"],
                       adaptive=entry["adaptive"])
        assert np.array_equal(out["rgba8"], z["rgba8"]), entry
        assert np.array_equal(bits(out["endpoint"]), bits(z["endpoint"])), entry
