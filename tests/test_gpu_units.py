"""Unit-level parity on the GPU: helper library, hit_triangle, field functions, overlay stencil, outline search, sample
rays, error behaviour of reload_shader, row-block shards.  Needs a B200: -m gpu."""
import os

import numpy as np
import pytest

from conftest import camera_dict, settings_dict
from linon_b200 import _capi as capi
from linon_b200 import sharding
from linon_b200.main_view import LinonError, MainView, PredefinedFunction, default_scene, make_settings

pytestmark = pytest.mark.gpu

PREDEFINED = [f for f in PredefinedFunction if f != PredefinedFunction.Custom]
HELPERS = dict(rotateX=0, rotateY=1, rotateZ=2, translate=3, refraction=4, point_plane_distance=5, sigmoid=6,
               refraction_index=7, reflect=8, normalize=9, cross=10)
EXACT_HELPERS = ("translate", "refraction", "point_plane_distance", "refraction_index", "reflect", "normalize", "cross")


@pytest.fixture(scope="module")
def view(product_lib, scene):
    verts, faces, center = scene
    mv = MainView(verts, faces, center, 128, 96)
    yield mv
    mv.close()


def rand_inputs(n, seed):
    rng = np.random.default_rng(seed)
    v = rng.normal(0, 1, (n, 3)).astype(np.float32)
    nrm = rng.normal(0, 1, (n, 3)).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    p0 = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    a = rng.uniform(-6, 6, n).astype(np.float32)
    b = rng.uniform(10, 200, n).astype(np.float32)
    c = rng.uniform(-2, 2, n).astype(np.float32)
    return v, nrm, p0, a, b, c


@pytest.mark.parametrize("name", list(HELPERS))
def test_helper_library_strict_mode(view, oracle, name):
    n = 8192
    v, nrm, p0, a, b, c = rand_inputs(n, HELPERS[name])
    if name == "refraction":  # temperatures; includes the total-reflection branch (eta > 1, grazing incidence)
        a = np.random.default_rng(1).uniform(10, 200, n).astype(np.float32)
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        v[: n // 4] = (nrm[: n // 4] * 1e-3 + np.cross(nrm[: n // 4], p0[: n // 4])).astype(np.float32)
    view.set_option(capi.OPT_MATH, capi.MATH_STRICT)
    got = view.debug_eval_helper(HELPERS[name], n, v=v, nrm=nrm, p0=p0, a=a, b=b, c=c)
    ref = oracle.helper(name, n, v=v, nrm=nrm, p0=p0, a=a, b=b, c=c)
    if name in EXACT_HELPERS:  # IEEE + - * / sqrt, no contraction: bit-exact
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), name
    else:  # sin / cos / exp: libdevice vs glibc, both within a few ulp
        assert np.allclose(got, ref, rtol=2e-6, atol=2e-7), name


@pytest.mark.parametrize("math,tol", [(capi.MATH_FUSED, 2e-6), (capi.MATH_FAST, 2e-5)])
def test_helper_library_production_modes(view, oracle, math, tol):
    n = 4096
    v, nrm, p0, a, b, c = rand_inputs(n, 99)
    a_t = np.random.default_rng(2).uniform(10, 200, n).astype(np.float32)
    view.set_option(capi.OPT_MATH, math)
    for name in HELPERS:
        aa = a_t if name == "refraction" else a
        got = view.debug_eval_helper(HELPERS[name], n, v=v, nrm=nrm, p0=p0, a=aa, b=b, c=c)
        ref = oracle.helper(name, n, v=v, nrm=nrm, p0=p0, a=aa, b=b, c=c)
        scale = np.maximum(np.abs(ref).max(axis=1, keepdims=True), 1.0)
        assert (np.abs(got - ref) <= tol * 40 * scale).all(), name


def test_rotateZ_drops_z_and_clamp_has_no_lower_bound(view):
    """SURVEY F4 and F3 on the device."""
    view.set_option(capi.OPT_MATH, capi.MATH_STRICT)
    v = np.array([[1.0, 2.0, 3.0]], np.float32)
    r = view.debug_eval_helper(HELPERS["rotateZ"], 1, v=v, a=np.zeros(1, np.float32))
    assert np.array_equal(r, [[1.0, 2.0, 0.0]])
    n = np.array([[0.0, 0.0, 1.0]], np.float32)
    vin = np.array([[0.0, 0.0, -4.0]], np.float32)  # dot = -4: a real clamp(-1, 1) would change the result
    r = view.debug_eval_helper(HELPERS["refraction"], 1, v=vin, nrm=n, a=np.array([15.0], np.float32), b=np.array([15.0], np.float32))
    assert np.allclose(r, vin)


def test_hit_triangle_matches_the_oracle(view, oracle):
    rng = np.random.default_rng(7)
    n = 20000
    tri = rng.uniform(-1, 1, (n, 9)).astype(np.float32)
    o = rng.uniform(-2, 2, (n, 3)).astype(np.float32)
    target = (tri[:, 0:3] + tri[:, 3:6] + tri[:, 6:9]) / 3 + rng.normal(0, 0.3, (n, 3)).astype(np.float32)
    d = target - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    # edge cases: parallel rays, origin on the plane, degenerate triangle
    tri[0] = [0, 0, 0, 1, 0, 0, 0, 1, 0]; o[0] = [0.2, 0.2, 1]; d[0] = [1, 0, 0]
    tri[1] = [0, 0, 0, 1, 0, 0, 0, 1, 0]; o[1] = [0.2, 0.2, 0]; d[1] = [0, 0, -1]
    tri[2] = [0, 0, 0, 1, 1, 1, 2, 2, 2]; o[2] = [0, 0, 1]; d[2] = [0, 0, -1]
    ref = oracle.hit_triangle(tri, o, d)
    view.set_option(capi.OPT_MATH, capi.MATH_STRICT)
    got = view.debug_hit_triangle(tri, o, d)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    assert 0.2 < (ref > 0).mean() < 0.8 and ref[0] == -1 and ref[1] == -1 and ref[2] == -1
    view.set_option(capi.OPT_MATH, capi.MATH_FAST)
    fast = view.debug_hit_triangle(tri, o, d)
    same = (fast > 0) == (ref > 0)
    assert same.mean() > 0.999
    both = (fast > 0) & (ref > 0)
    assert np.allclose(fast[both], ref[both], rtol=1e-4)


@pytest.mark.parametrize("fn", PREDEFINED)
def test_field_functions_match_the_oracle(view, oracle, fn):
    rng = np.random.default_rng(int(fn))
    n = 8192
    p = rng.uniform(-1.2, 1.2, (n, 3)).astype(np.float32)
    p[: n // 2] = np.array([-0.5, 0.5, -0.5], np.float32) + rng.normal(0, 0.15, (n // 2, 3)).astype(np.float32)
    p[n // 2: 3 * n // 4, 1] = 0.1 + rng.normal(0, 0.006, n // 4).astype(np.float32)
    p_prev = (p + rng.normal(0, 0.01, (n, 3))).astype(np.float32)
    v = rng.normal(0, 1, (n, 3)).astype(np.float32)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    v0 = rng.normal(0, 1, (n, 3)).astype(np.float32)
    t = rng.uniform(0, 5.1, n).astype(np.float32)
    ref = oracle.field(int(fn), settings_dict(make_settings()), p_prev, p, v0, v, t)
    view.reload_shader(fn.to_code())
    view.set_option(capi.OPT_MATH, capi.MATH_STRICT)
    got = view.debug_eval_field(p_prev, p, v0, v, t)
    if fn in (PredefinedFunction.MirageSphericalSigmoid, PredefinedFunction.MiragePlaneSigmoid, PredefinedFunction.Rotation):
        assert np.allclose(got, ref, rtol=1e-5, atol=1e-6)
    else:
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    # fast math: within tolerance of the fp32 oracle, or - where the reference's own formulation is ill-conditioned
    # (refraction at grazing incidence rounds sqrt(1 - (1 - cosi^2)); its fp32 result is then itself 1e-4 away from
    # the function it rounds) - of the same restatement evaluated in double.  Both the generic translation and the
    # hand-fused restatement of the mirage bodies (mirage_fused.cuh) are graded.
    ref64 = oracle.field(int(fn), settings_dict(make_settings()), p_prev, p, v0, v, t, use_fp64=True)
    view.set_option(capi.OPT_MATH, capi.MATH_FAST)
    for specialize in (0, 1):
        view.set_option(capi.OPT_SPECIALIZE, specialize)
        fast = view.debug_eval_field(p_prev, p, v0, v, t)
        ok = np.isclose(fast, ref, rtol=2e-4, atol=2e-5) | np.isclose(fast, ref64, rtol=2e-4, atol=2e-5)
        assert ok.all(), (specialize, np.argwhere(~ok)[:5])
    view.set_option(capi.OPT_SPECIALIZE, 1)


def _oracle_rk4(oracle, fn, p, d, v0, t, h, use_fp64):
    """v = (k1 + 2 k2 + 2 k3 + k4) / 6 of src/main_view.wgsl:246-251 out of four calls of the oracle's field function
    (p_prev = p in every stage); stage inputs in fp32 like the shader."""
    st = settings_dict(make_settings())
    f = lambda q, k, tt: oracle.field(int(fn), st, p, q.astype(np.float32), v0, k, tt.astype(np.float32), use_fp64=use_fp64)
    hh = h[:, None]
    k1 = f(p, d, t)
    k2 = f(p + np.float32(0.5) * hh * k1, k1, t + np.float32(0.5) * h)
    k3 = f(p + np.float32(0.5) * hh * k2, k2, t + np.float32(0.5) * h)
    k4 = f(p + hh * k3, k3, t + h)
    return ((k1 + np.float32(2) * k2 + np.float32(2) * k3 + k4) / np.float32(6)).astype(np.float32), k1, k2


@pytest.mark.parametrize("fn", PREDEFINED)
def test_rk4_velocity_matches_the_oracle(view, oracle, fn):
    """The device function the march kernels call for one Runge-Kutta step - in fast mode the whole-step restatement of
    the mirage bodies (scaled coordinates, slope-normalised eta, deferred total reflection) - against four calls of the
    oracle's field function.  Inputs concentrate on the temperature gradient zones and on grazing incidence, where
    Snell's discriminant changes sign (total reflection: the classic mirage rays above the hot plane)."""
    rng = np.random.default_rng(100 + int(fn))
    n = 16384
    centre = np.array([-0.5, 0.5, -0.5], np.float32)
    p = rng.uniform(-1.2, 1.2, (n, 3)).astype(np.float32)
    p[: n // 4] = centre + rng.normal(0, 0.15, (n // 4, 3)).astype(np.float32)
    p[n // 4: n // 2, 1] = 0.1 + rng.uniform(-0.012, 0.012, n // 4).astype(np.float32)
    d = rng.normal(0, 1, (n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    # grazing incidence: almost tangent to the sphere shells / almost parallel to the plane, heading into the hot side
    q = n // 8
    radial = p[:q] - centre
    radial /= np.linalg.norm(radial, axis=1, keepdims=True)
    tang = np.cross(radial, rng.normal(0, 1, (q, 3)))
    tang /= np.linalg.norm(tang, axis=1, keepdims=True)
    d[:q] = (tang - radial * rng.uniform(0, 0.3, (q, 1))).astype(np.float32)
    d[n // 4: n // 4 + q, 1] = -np.abs(rng.uniform(0, 0.02, q)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    v0 = d.copy()
    t = rng.uniform(0, 5, n).astype(np.float32)
    h = np.where(rng.uniform(size=n) < 0.5, np.float32(0.1), np.float32(0.001)).astype(np.float32)
    ref, k1, k2 = _oracle_rk4(oracle, fn, p, d, v0, t, h, use_fp64=False)
    ref64, _, _ = _oracle_rk4(oracle, fn, p, d, v0, t, h, use_fp64=True)
    # Grading.  Near grazing incidence Snell's discriminant k = 1 - eta^2 (1 - cos^2) is close to 0, sqrt(k) amplifies
    # rounding by 1 / (2 sqrt k) and the refracted / reflected decision (k < 0) can flip: the oracle's own fp32 and fp64
    # evaluations then differ by far more than an ulp.  |ref - ref64| measures that local conditioning directly, so a
    # result passes when it lies within a small tolerance plus three times the oracle's own fp32-vs-fp64 spread of
    # either oracle value; what remains (a reflect / refract decision flipped by fast-math rounding where the two oracle
    # runs happen to agree: about 1 % of the deliberately grazing eighth of the inputs, generic translation and
    # restatement alike) must stay below 2e-3 of the samples.
    spread = np.abs(ref.astype(np.float64) - ref64.astype(np.float64))

    def bad_rows(got, rtol, atol):
        e = np.minimum(np.abs(got - ref), np.abs(got - ref64))
        return (e > atol + rtol * np.abs(ref) + 3.0 * spread).any(axis=1)

    def report(got, bad):
        i = np.argwhere(bad).ravel()[:4]
        return dict(n_bad=int(bad.sum()), idx=i.tolist(), got=got[i].tolist(), ref=ref[i].tolist(), ref64=ref64[i].tolist(),
                    p=p[i].tolist(), d=d[i].tolist(), h=h[i].tolist())

    view.reload_shader(fn.to_code())
    view.set_option(capi.OPT_MATH, capi.MATH_STRICT)
    strict = view.debug_eval_rk4(p, d, v0, t, h)
    if fn in (PredefinedFunction.MirageSphericalSigmoid, PredefinedFunction.MiragePlaneSigmoid, PredefinedFunction.Rotation):
        bad = bad_rows(strict, 1e-5, 1e-6)  # exp / sin / cos: libdevice vs glibc, a few ulp
        assert bad.mean() <= 5e-4, report(strict, bad)
    else:
        assert np.array_equal(strict.view(np.uint32), ref.view(np.uint32))
    view.set_option(capi.OPT_MATH, capi.MATH_FAST)
    for specialize in (0, 1):
        view.set_option(capi.OPT_SPECIALIZE, specialize)
        fast = view.debug_eval_rk4(p, d, v0, t, h)
        bad = bad_rows(fast, 3e-4, 3e-5)
        assert bad.mean() <= 2e-3, (specialize, report(fast, bad))
    view.set_option(capi.OPT_SPECIALIZE, 1)
    assert (spread.max(axis=1) > 1e-3).mean() < 0.02  # the ill-conditioned samples are a small minority of the inputs
    if fn in (PredefinedFunction.MiragePlane, PredefinedFunction.MiragePlaneSigmoid):
        # the inputs do exercise total reflection: stage 2 mirrors the incoming direction at the plane
        mirrored = np.isclose(k2[:, 1], -k1[:, 1], rtol=1e-6) & (np.abs(k1[:, 1]) > 1e-4) & np.isclose(k2[:, 0], k1[:, 0], rtol=1e-6)
        assert mirrored.sum() > 50, int(mirrored.sum())


# ---- whole-shader hot reload: the helper layer of an edited main_view.wgsl ---------------------
HELPER_MODULE = """
const swirl_gain: f32 = 0.25;
const up_axis = vec3<f32>(0.0, 1.0, 0.0);
fn swirl(p: vec3<f32>, v: vec3<f32>, k: f32) -> vec3<f32> {
    let r = cross(up_axis, p);
    if (k > 0.5) {
        return normalize(v + swirl_gain * k * r);
    }
    return v;
}
fn rotateZ(v: vec3<f32>, phi: f32) -> vec3<f32> {
    return vec3<f32>(cos(phi) * v.x - sin(phi) * v.y, sin(phi) * v.x + cos(phi) * v.y, v.z);
}
"""


def test_reloaded_module_helpers_run_on_the_device(view):
    rng = np.random.default_rng(5)
    n = 4096
    p = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    v = rng.normal(0, 1, (n, 3)).astype(np.float32)
    z = np.zeros_like(p)
    t = np.zeros(n, np.float32)
    view.set_option(capi.OPT_MATH, capi.MATH_STRICT)
    view.reload_shader("return rotateZ(v, 0.3);")
    stock = view.debug_eval_field(z, p, z, v, t)
    assert np.all(stock[:, 2] == 0.0)  # the reference's rotateZ drops z
    try:
        view.reload_compute_shader(HELPER_MODULE)
        assert view.module_overrides() == ["swirl_gain", "up_axis", "swirl", "rotateZ"]
        got = view.debug_eval_field(z, p, z, v, t)  # same body, now bound to the module's rotateZ
        c, s_ = np.float32(np.cos(np.float32(0.3))), np.float32(np.sin(np.float32(0.3)))
        want = np.stack([c * v[:, 0] - s_ * v[:, 1], s_ * v[:, 0] + c * v[:, 1], v[:, 2]], axis=1)
        assert np.allclose(got, want, rtol=2e-6, atol=2e-7) and np.array_equal(got[:, 2], v[:, 2])
        # a new helper with module constants and control flow
        view.reload_shader("return swirl(p, v, 0.7);")
        got = view.debug_eval_field(z, p, z, v, t)
        w = v + np.float32(0.25) * np.float32(0.7) * np.cross(np.array([0, 1, 0], np.float32), p)
        want = w / np.linalg.norm(w, axis=1, keepdims=True)
        assert np.allclose(got, want, rtol=1e-5, atol=1e-6)
        view.reload_shader("return swirl(p, v, 0.2);")
        assert np.array_equal(view.debug_eval_field(z, p, z, v, t), v)
        # errors keep the installed module and kernel (src/main_view.rs:677-679)
        with pytest.raises(LinonError) as e:
            view.reload_compute_shader(HELPER_MODULE.replace("cross(up_axis, p)", "crossx(up_axis, p)"))
        assert e.value.code == capi.ERR_COMPILE and "main_view.wgsl(5)" in e.value.message
        with pytest.raises(LinonError) as e:  # an edit to the march itself that the kernel template has no switch for
            view.reload_compute_shader(HELPER_MODULE + "fn hit_triangle(a: f32) -> f32 { return a; }")
        assert e.value.code == capi.ERR_UNSUPPORTED and "'hit_triangle'" in e.value.message
        assert view.module_overrides() == ["swirl_gain", "up_axis", "swirl", "rotateZ"]
        assert np.array_equal(view.debug_eval_field(z, p, z, v, t), v)
        # frames render through every march kernel with the helper layer installed
        view.reload_shader("return normalize(v + 0.02 * (swirl(p, rotateZ(v, 0.01), 0.7) - v));")  # gently bent rays
        frames = []
        for kernel in (capi.KERNEL_TILES, capi.KERNEL_PAIRS, capi.KERNEL_REFILL):
            view.set_option(capi.OPT_KERNEL, kernel)
            view.render_flags(capi.RENDER_ADAPTIVE_OFF)
            view.sync()
            frames.append((view.hit_ids(), view.mapping_texture()))
        for h, e_ in frames[1:]:
            assert np.array_equal(h, frames[0][0]) and np.array_equal(e_.view(np.uint32), frames[0][1].view(np.uint32))
        assert (frames[0][0] >= 0).mean() > 0.05
        # the stock module cannot come back while the field function still calls a helper of the edited one - the same
        # compile error the reference reports for reload_shader(Some(stock), body) - and nothing changes
        with pytest.raises(LinonError) as e:
            view.reload_compute_shader(None)
        assert e.value.code == capi.ERR_COMPILE and "swirl" in e.value.message
        assert view.module_overrides() == ["swirl_gain", "up_axis", "swirl", "rotateZ"]
    finally:
        view.set_option(capi.OPT_KERNEL, capi.KERNEL_TILES)
        view.reload_shader("return v;")
        view.reload_compute_shader(None)  # back to the stock module
    assert view.module_overrides() == []
    view.reload_shader("return rotateZ(v, 0.3);")
    assert np.array_equal(view.debug_eval_field(z, p, z, v, t), stock)


def test_a_helper_layer_sends_the_predefined_mirage_bodies_through_the_generic_translation(view):
    """While a module layer is installed a helper the mirage bodies call may have been redefined, so the hand-fused
    restatement is not spliced: the frame equals the generic fast-math build bit for bit."""
    fn = PredefinedFunction.MirageSphericalSigmoid
    view.set_option(capi.OPT_MATH, capi.MATH_FAST)
    view.reload_shader(fn.to_code())

    def frame():
        view.render_flags(capi.RENDER_ADAPTIVE_OFF)
        view.sync()
        return view.hit_ids(), view.mapping_texture()

    try:
        view.set_option(capi.OPT_SPECIALIZE, 0)
        generic = frame()
        view.set_option(capi.OPT_SPECIALIZE, 1)
        fused = frame()
        view.reload_compute_shader("fn unrelated(a: f32) -> f32 { return a + 1.0; }")
        layered = frame()
        assert np.array_equal(layered[0], generic[0]) and np.array_equal(layered[1].view(np.uint32), generic[1].view(np.uint32))
        assert not np.array_equal(fused[1].view(np.uint32), generic[1].view(np.uint32))  # the splice was really off
    finally:
        view.set_option(capi.OPT_SPECIALIZE, 1)
        view.reload_compute_shader(None)


def test_user_edited_function_runs_and_errors_keep_the_old_kernel(view):
    view.set_option(capi.OPT_MATH, capi.MATH_STRICT)
    body = "var q = p * 2.0;\nif (t > 1.0) { q = q + vec3<f32>(1.0, 0.0, 0.0); }\nreturn q.zyx + v0 * settings.field_weight;"
    view.reload_shader(body)
    p = np.array([[1, 2, 3], [4, 5, 6]], np.float32)
    z = np.zeros_like(p)
    v0 = np.array([[10, 20, 30], [10, 20, 30]], np.float32)
    t = np.array([0.5, 2.0], np.float32)
    got = view.debug_eval_field(z, p, v0, z, t)
    assert np.array_equal(got, [[6 + 10, 4 + 20, 2 + 30], [12 + 10, 10 + 20, 9 + 30]])
    with pytest.raises(LinonError) as e:  # src/main_view.rs:677-679: Err returned, old pipeline kept
        view.reload_shader("let x = ;\nreturn v;")
    assert e.value.code == capi.ERR_COMPILE and "field_function.wgsl(1)" in e.value.message
    with pytest.raises(LinonError):
        view.reload_shader("return undefined_thing;")
    again = view.debug_eval_field(z, p, v0, z, t)
    assert np.array_equal(again, got)
    view.render()
    view.sync()  # the frame still renders with the last good function


def test_user_edited_bodies_match_the_reference_semantics(view):
    """Eleven user-edit style bodies (tests/user_bodies.py), evaluated on the device through the product's translator,
    against tests/golden/user_bodies.npz: the same bodies under the reference's shader semantics (spliced into the
    reference's main_view.wgsl, transpiled by oracle/wgsl_ref, run on the CPU; made by tests/golden/make_user_bodies.py).
    Strict mode: bit-exact where the arithmetic is IEEE + - * / sqrt and integer work; last-ulp agreement where
    transcendentals are involved (libdevice vs glibc)."""
    import os
    import user_bodies
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "user_bodies.npz"))
    names = [str(n) for n in z["names"]]
    assert names == list(user_bodies.BODIES)
    ins = user_bodies.inputs(n=1024)
    view.update_settings(make_settings())
    view.set_option(capi.OPT_MATH, capi.MATH_STRICT)
    for k, name in enumerate(names):
        view.reload_shader(user_bodies.BODIES[name])
        got = view.debug_eval_field(*ins)
        want = z[f"body{k}"]
        finite = np.isfinite(want).all(axis=1)
        assert finite.mean() > 0.9, name
        if name in user_bodies.EXACT:
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), name
        else:
            scale = np.maximum(np.abs(want[finite]).max(axis=1, keepdims=True), 1.0)
            err = (np.abs(got[finite] - want[finite]) / scale).max(axis=1)
            # an FMA-contracted fast-math host build of the same text stays below 3e-6 on every row; the margin is for
            # rows near a branch of refraction() (grazing incidence), where one ulp of the cosine moves the result
            assert (err <= 2e-5).mean() >= 0.995 and np.isfinite(got[finite]).all(), (name, float(err.max()))
    view.reload_shader(PredefinedFunction.MirageSphericalSigmoid.to_code())
    view.set_option(capi.OPT_MATH, capi.MATH_FAST)


# ---- overlay ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("delta", [1, 2, 5, 10])
def test_overlay_matches_the_oracle(product_lib, oracle, scene, delta):
    verts, faces, center = scene
    W, H = 200, 150  # not a multiple of the 32x16 tile
    with MainView(verts, faces, center, W, H) as mv:
        fn = PredefinedFunction.MirageSphericalSigmoid
        mv.reload_shader(fn.to_code())
        st = make_settings(overlay_mode=1, central_difference_delta=delta, lyapunov_scaling=50.0)
        mv.update_settings(st)
        mv.disable_adaptive_sampling()
        mv.render()
        mv.sync()
        raycast, ep = mv.ray_casting_texture(), mv.mapping_texture()
        final, ex = mv.texture(), mv.exponents()
        ref_final, ref_ex = oracle.overlay(raycast, ep, settings_dict(st))  # same inputs: isolates the stencil
        assert np.allclose(ex, ref_ex, rtol=1e-5, atol=1e-7)
        assert np.abs(final.astype(int) - ref_final.astype(int)).max() <= 1
        assert (final == ref_final).mean() > 0.999
        pad = 2 * delta
        assert (ex[:pad] == 0).all() and (ex[:, :pad] == 0).all() and (ex[-pad:] == 0).all() and (ex[:, -pad:] == 0).all()
        assert (final[:pad] == raycast[:pad]).all() and (final[:, -pad:] == raycast[:, -pad:]).all()
        assert (ex >= 0.8).any()  # the mirage rim diverges
        # mode 0 is an exact copy and leaves the exponents alone (src/overlay.wgsl:116-118)
        mv.update_settings(make_settings(overlay_mode=0, central_difference_delta=delta))
        mv.render_flags(capi.RENDER_NO_MARCH | capi.RENDER_NO_SAMPLE_RAYS)
        mv.sync()
        assert np.array_equal(mv.texture(), raycast) and np.array_equal(mv.exponents(), ex)


# ---- sample rays + outline search ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("fn", [PredefinedFunction.MirageSphericalSigmoid, PredefinedFunction.TranslationX, PredefinedFunction.LorenzAttractor])
def test_click_mode_sample_rays(product_lib, oracle, scene, fn):
    verts, faces, center = scene
    W, H = 160, 120
    with MainView(verts, faces, center, W, H) as mv:
        mv.set_option(capi.OPT_MATH, capi.MATH_STRICT)
        mv.reload_shader(fn.to_code())
        st = make_settings(mouse_pos=(0.43, 0.57), field_weight=0.3 if fn == PredefinedFunction.LorenzAttractor else 1.0)
        mv.update_settings(st)
        mv.render()
        mv.sync()
        cam = camera_dict(mv.camera.uniform())
        pos = oracle.sample_positions(cam, settings_dict(st), W, H)
        assert np.array_equal(mv.read(capi.BUF_SAMPLE_POS), pos)
        ref = oracle.sample_rays(cam, settings_dict(st), int(fn), W, H, pos)
        got = mv.ray_samples()
        assert got.shape == (800, 8)
        assert np.array_equal(got[:, 3:], ref[:, 3:])  # w = 1, colours, alpha 0.5 (src/ray_samples.rs layout)
        if fn == PredefinedFunction.MirageSphericalSigmoid:  # 5000 steps through the heat sphere: exp() ulps accumulate
            assert np.allclose(got[:, :3], ref[:, :3], rtol=2e-3, atol=2e-3)
        else:
            assert np.array_equal(got[:, :3].view(np.uint32), ref[:, :3].view(np.uint32))


def test_outline_mode_positions_match_the_serial_search(product_lib, oracle, scene):
    """The parallel arg-reduction must pick the same 8 pixels as the reference's serial scans (stride W)."""
    verts, faces, center = scene
    W, H = 240, 135
    with MainView(verts, faces, center, W, H) as mv:
        fn = PredefinedFunction.MirageSphericalSigmoid
        st = make_settings(overlay_mode=1)
        mv.update_settings(st)
        mv.disable_adaptive_sampling()
        mv.render()  # fills the exponents
        mv.sync()
        ex = mv.exponents()
        assert (ex >= 0.8).sum() > 50
        mv.render_outline_rays(fn.to_code())
        mv.sync()
        cam = camera_dict(mv.camera.uniform())
        ref_pos = oracle.sample_positions(cam, settings_dict(st), W, H, outline=True, exponents=ex)
        assert np.array_equal(mv.read(capi.BUF_SAMPLE_POS), ref_pos)
        ref = oracle.sample_rays(cam, settings_dict(st), int(fn), W, H, ref_pos)
        assert np.allclose(mv.ray_samples()[:, :3], ref[:, :3], rtol=2e-3, atol=2e-3)
        assert mv.timings().outline_search_ms > 0
    # nothing above the threshold: every search falls back to pixel (0, 0) -> (0, 1)
    with MainView(verts, faces, center, 64, 48) as mv:
        mv.render_outline_rays(PredefinedFunction.TranslationX.to_code())
        mv.sync()
        assert np.array_equal(mv.read(capi.BUF_SAMPLE_POS), np.tile(np.array([0.0, 1.0], np.float32), (8, 1)))


def test_edited_sample_colors_hot_reload(product_lib, scene):
    """`var<private> sample_colors` with other values (src/main_view.wgsl:326-335; a partial module holds the declaration alone):
    the sample kernel stores the user's colours next to unchanged positions, and the stock module brings the stock colours back."""
    verts, faces, center = scene
    colours = np.array([[1.0, 0.0, 0.0], [0.5, 0.5, 0.0], [0.0, 1.0, 0.0], [0.5, 1.0, 0.0],
                        [1.0, 1.0, 0.0], [0.25, 0.75, 0.5], [0.0, 0.125, 1.0], [0.5, 0.0, 0.5]], np.float32)
    module = ("var<private> sample_colors: array<vec3<f32>, 8> = array<vec3<f32>, 8>(\n" +
              "".join("    vec3<f32>(%s, %s, %s),\n" % tuple(repr(float(x)) for x in c) for c in colours) + ");\n")
    with MainView(verts, faces, center, 96, 64) as mv:
        mv.reload_shader(PredefinedFunction.MirageSpherical.to_code())
        mv.update_settings(make_settings())
        mv.render_flags(0)
        mv.sync()
        stock = mv.ray_samples().copy()
        mv.reload_compute_shader(module)
        assert mv.module_overrides() == ["sample_colors"]
        mv.render_flags(0)
        mv.sync()
        got = mv.ray_samples()
        assert np.array_equal(got[:, :4].view(np.uint32), stock[:, :4].view(np.uint32))      # the same 800 positions
        assert np.array_equal(got[:, 4:7].reshape(8, 100, 3), np.repeat(colours[:, None, :], 100, axis=1))
        assert not np.array_equal(got[:, 4:7], stock[:, 4:7])
        mv.reload_compute_shader(None)
        mv.render_flags(0)
        mv.sync()
        assert np.array_equal(mv.ray_samples().view(np.uint32), stock.view(np.uint32))


# ---- shards ----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("world,rpb,kernel", [(2, 32, capi.KERNEL_TILES), (3, 8, capi.KERNEL_PERSISTENT), (8, 16, capi.KERNEL_TILES)])
def test_row_block_shards_reassemble_bit_exactly(product_lib, scene, world, rpb, kernel):
    """N-GPU output must equal 1-GPU output bit for bit; here the N shards are rendered one after the
    other on one GPU and merged by row ownership."""
    verts, faces, center = scene
    W, H = 200, 150
    fn = PredefinedFunction.MirageSphericalSigmoid
    with MainView(verts, faces, center, W, H) as mv:
        mv.set_option(capi.OPT_KERNEL, kernel)
        mv.reload_shader(fn.to_code())
        mv.render_flags(capi.RENDER_NO_SAMPLE_RAYS)
        mv.sync()
        full = dict(h=mv.hit_ids(), c=mv.ray_casting_texture(), e=mv.mapping_texture())
        rays_full = mv.counters().rays
        merged = {k: np.zeros_like(v) for k, v in full.items()}
        rays = 0
        for rank in range(world):
            mv.resize_texture(W, H)  # fresh (zeroed) buffers, like another GPU's
            mv.set_shard(rank, world, rpb)
            mv.render_flags(capi.RENDER_NO_SAMPLE_RAYS | capi.RENDER_NO_OVERLAY)
            mv.sync()
            rows = sharding.local_rows(H, rank, world, rpb)
            part = dict(h=mv.hit_ids(), c=mv.ray_casting_texture(), e=mv.mapping_texture())
            other = np.setdiff1d(np.arange(H), rows)
            assert (part["c"][other] == 0).all() and (part["h"][other] == -1).all()  # other ranks' rows untouched
            for k in merged:
                merged[k][rows] = part[k][rows]
            rays += mv.counters().rays
        assert rays == rays_full == W * H
        for k in full:
            assert np.array_equal(full[k].view(np.uint8), merged[k].view(np.uint8)), k


# ---- host mirror behaviour ----------------------------------------------------------------------------------------------------
def test_main_view_mirror_follows_the_reference(product_lib, scene):
    verts, faces, center = scene
    with MainView(verts, faces, center, 200, 100, discrete_gpu=False) as mv:
        assert mv.size == (100, 50)  # downscale_factor 2 on non-discrete adapters (src/main_view.rs:93,104-105)
        assert mv.needs_redraw
        mv.render()
        assert not mv.needs_redraw
        mv.update_settings(make_settings(field_weight=0.5))
        assert mv.needs_redraw
        mv.render()
        mv.on_zoom(1.0)
        assert mv.needs_redraw
        mv.resize_texture(320, 240)
        assert mv.size == (160, 120) and mv.texture().shape == (120, 160, 4)
        mv.on_pointer_moved("rotate", (10, 10), (60, 40))  # dragging disables adaptive sampling (:524)
        mv.render()
        mv.sync()
        assert mv.counters().rejected_steps == 0 and mv.counters().accepted_steps == 160 * 120 * 51
        mv.enable_adaptive_sampling()
        mv.render()
        mv.sync()
        assert mv.counters().rejected_steps > 0
        cam0 = np.array(mv.camera.uniform().origin[:3])
        mv.reset_camera()
        assert np.allclose(mv.camera.uniform().origin[:3], center + np.array([0, 0, 2], np.float32), atol=1e-6)
        assert not np.allclose(cam0, mv.camera.uniform().origin[:3])
        # update_model: the `r` key scene (no axis inversion, src/application.rs:194-197)
        v2, f2, c2 = default_scene(negate_xz=False)
        mv.update_model(v2, f2, c2)
        mv.render()
        mv.sync()
        assert (mv.hit_ids() >= 0).any()


@pytest.mark.parametrize("name", ["tint", "noop", "declined", "depth", "linear_shade", "camera", "linear_reach"])
def test_an_edited_shading_block_hot_reloads_against_the_edited_reference_shader(product_lib, scene, name):
    """(The cases "camera" and "linear_reach" edit literals of main_view instead - focal length, viewport height, the reach of
    linear_mode's ray, :355-356, :369 - which become macros of the kernel template; same procedure.)
    ray_color edited inside its shading block (src/main_view.wgsl:199-218): the block is translated from the user's text and
    called at every geometric hit (march_kernels.cuh: shade_hit); a hit the block declines (alpha <= 0) leaves the ray marching
    like `has_color = result.color.a > 0.0` (:262-263).  Frames equal tests/golden/march_edits.npz - the REFERENCE's shader with
    the same edit, transpiled and run on the CPU - endpoints bit for bit in strict math, colours within 1/255, every march
    kernel.  The module text is the reference's own file with the edit: available where the reference is mounted or where
    tests/golden/make_march_edits.py left its scratch copies; skipped elsewhere (the CPU suite runs the same comparison on
    the emulated device code, tests/test_march_edits_cpu.py)."""
    import march_edit_cases as mc
    module = mc.shade_module_text(name)
    if module is None:
        pytest.skip("needs the text of the reference's ray_color (mount or tests/golden/_shade_modules)")
    verts, faces, center = scene
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "march_edits.npz"))
    fn = PredefinedFunction.MirageSpherical
    with MainView(verts, faces, center, mc.W, mc.H) as mv:
        mv.set_option(capi.OPT_MATH, capi.MATH_STRICT)
        mv.reload_shader(fn.to_code())
        mv.update_settings(make_settings())
        mv.render_flags(0)
        mv.sync()
        stock = mv.mapping_texture().copy(), mv.ray_casting_texture().copy()
        mv.reload_compute_shader(module)
        assert mc.OVERRIDES.get(name, "ray_color") in mv.module_overrides()
        frames = []
        for kernel in (capi.KERNEL_TILES, capi.KERNEL_PERSISTENT, capi.KERNEL_PAIRS, capi.KERNEL_REFILL):
            mv.set_option(capi.OPT_KERNEL, kernel)
            mv.render_flags(0)
            mv.sync()
            frames.append((mv.ray_casting_texture(), mv.mapping_texture(), mv.hit_ids()))
        rgba8, endpoint, hit = frames[0]
        assert np.abs(rgba8.astype(int) - z[f"rgba8_{name}"].astype(int)).max() <= 1
        if name.startswith("linear"):
            assert np.array_equal(endpoint, stock[0])  # `mapping` is not written in linear mode
        else:
            assert np.array_equal(endpoint.view(np.uint32), z[f"endpoint_{name}"].view(np.uint32))
        if name == "noop":
            assert np.array_equal(rgba8, stock[1])     # the reference's own block through the translator: the stock frame
        else:
            assert not np.array_equal(rgba8, stock[1])
        if name == "declined":
            # rays pass through the two side walls: 43 % of the pixels end on a surface instead of 63 %
            assert 0 < (hit >= 0).sum() < 0.8 * (z["rgba8_tint"][..., :3].max(axis=-1) > 0).sum()
        for r, e, h in frames[1:]:
            assert np.array_equal(r, rgba8) and np.array_equal(e.view(np.uint32), endpoint.view(np.uint32)) and np.array_equal(h, hit)
        # fixed-step frames with and without culling agree (a declined hit does not disturb the safe-distance budgets)
        per_cull = []
        for cull in (1, 0):
            mv.set_option(capi.OPT_KERNEL, capi.KERNEL_TILES)
            mv.set_option(capi.OPT_CULL, cull)
            mv.render_flags(capi.RENDER_ADAPTIVE_OFF)
            mv.sync()
            per_cull.append((mv.ray_casting_texture(), mv.mapping_texture()))
        mv.set_option(capi.OPT_CULL, 1)
        if not name.startswith("linear"):
            assert np.array_equal(per_cull[0][0], per_cull[1][0]) and np.array_equal(per_cull[0][1].view(np.uint32), per_cull[1][1].view(np.uint32))
        # production math with the edit: same surfaces; colours within 1/255 of the strict frame - except "depth", whose red
        # channel is t / max_dist: where inside its step a ray meets the wall, which in an adaptive frame follows the step
        # sequence and with it rounding noise (45 % agreement measured)
        mv.set_option(capi.OPT_MATH, capi.MATH_FAST)
        mv.render_flags(0)
        mv.sync()
        fast = mv.ray_casting_texture()
        assert (mv.hit_ids() == hit).mean() >= 0.98
        if name != "depth":
            assert (np.abs(fast.astype(int) - rgba8.astype(int)).max(axis=-1) <= 1).mean() >= 0.99
        mv.reload_compute_shader(None)
        mv.set_option(capi.OPT_MATH, capi.MATH_STRICT)
        mv.render_flags(0)
        mv.sync()
        assert np.array_equal(mv.ray_casting_texture(), stock[1])    # back to the stock module


@pytest.mark.parametrize("name", ["eps", "backface_culling", "use_lighting", "light", "linear_mode", "adaptive_sampling"])
def test_march_level_constants_hot_reload_against_the_edited_reference_shader(product_lib, scene, name):
    """The file watcher path (src/main.rs:206-222) with an edit to a march-level constant of main_view.wgsl
    (src/main_view.wgsl:65-76, :227): the module is accepted, the kernel template is rebuilt with the constant, and the frame
    equals tests/golden/march_edits.npz - the REFERENCE's shader with the same edit in its text (transpiled and run on the CPU
    by tests/golden/make_march_edits.py) - endpoints bit for bit in strict math, colours within 1/255; every march kernel agrees."""
    import march_edit_cases as mc
    verts, faces, center = scene
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "march_edits.npz"))
    fn = PredefinedFunction.MirageSpherical
    with MainView(verts, faces, center, mc.W, mc.H) as mv:
        mv.set_option(capi.OPT_MATH, capi.MATH_STRICT)
        mv.reload_shader(fn.to_code())
        mv.update_settings(make_settings())
        mv.render_flags(0)
        mv.sync()
        stock = mv.mapping_texture().copy(), mv.ray_casting_texture().copy()
        mv.reload_compute_shader(mc.partial_module(name))
        assert set(mv.module_overrides()) == {new.split()[1].rstrip(":") for _, new in mc.CASES[name]}
        frames = []
        for kernel in (capi.KERNEL_TILES, capi.KERNEL_PERSISTENT, capi.KERNEL_PAIRS, capi.KERNEL_REFILL):
            mv.set_option(capi.OPT_KERNEL, kernel)
            mv.render_flags(0)
            mv.sync()
            frames.append((mv.ray_casting_texture(), mv.mapping_texture(), mv.hit_ids()))
        rgba8, endpoint, hit = frames[0]
        assert np.abs(rgba8.astype(int) - z[f"rgba8_{name}"].astype(int)).max() <= 1
        if name == "linear_mode":
            assert np.array_equal(endpoint, stock[0])                 # `mapping` is not written in linear mode
            assert (rgba8[..., 3][hit < 0] == 0).all()                # a miss stores vec4(0): alpha 0
        else:
            assert np.array_equal(endpoint.view(np.uint32), z[f"endpoint_{name}"].view(np.uint32))
        if name != "backface_culling":  # every face the start-up camera sees is front-facing: culling leaves this frame as it is
            assert not (np.array_equal(rgba8, stock[1]) and np.array_equal(endpoint, stock[0]))  # the edit changed the frame
        for r, e, h in frames[1:]:
            assert np.array_equal(r, rgba8) and np.array_equal(e.view(np.uint32), endpoint.view(np.uint32)) and np.array_equal(h, hit)
        if name == "adaptive_sampling":
            c = mv.counters()
            assert c.accepted_steps == mc.W * mc.H * 51 and c.rejected_steps == 0
        # production math with the edit: same hits, colours within 1/255 of the strict frame on >= 99.9 %
        mv.set_option(capi.OPT_KERNEL, capi.KERNEL_TILES)
        mv.set_option(capi.OPT_MATH, capi.MATH_FAST)
        mv.render_flags(0)
        mv.sync()
        fast = mv.ray_casting_texture()
        # (an adaptive 64 x 48 frame: the step sequence follows rounding noise, SURVEY F20, and eps = 1e-3 moves grazing hits)
        assert (np.abs(fast.astype(int) - rgba8.astype(int)).max(axis=-1) <= 1).mean() >= (0.98 if name == "eps" else 0.99)
        mv.reload_compute_shader(None)
        mv.set_option(capi.OPT_MATH, capi.MATH_STRICT)
        mv.render_flags(0)
        mv.sync()
        assert np.array_equal(mv.ray_casting_texture(), stock[1])    # back to the stock module


def test_module_and_body_are_reloaded_as_one_step(product_lib, scene):
    """MainView::reload_shader(device, Some(new_src), field_function) compiles the PAIR (src/main_view.rs:649-690): a helper
    renamed in the module and in the body at once is accepted although neither half compiles against the other's old
    state (round-1 advisor finding); a pair that does not compile changes nothing."""
    verts, faces, center = scene
    L = capi.lib()
    import ctypes as C
    with MainView(verts, faces, center, 64, 48) as mv:
        mod_a = "fn bend(v: vec3<f32>, k: f32) -> vec3<f32> { return normalize(v + vec3<f32>(0.0, k, 0.0)); }"
        mod_b = mod_a.replace("bend", "tilt")
        err = C.create_string_buffer(8192)
        assert L.linon_reload_shader(mv._view, mod_a.encode(), b"return bend(v, 0.1);", err, len(err)) == capi.OK, err.value
        z = np.zeros((4, 3), np.float32)
        v = np.tile(np.array([[0.0, 0.0, -1.0]], np.float32), (4, 1))
        t = np.zeros(4, np.float32)
        a = mv.debug_eval_field(z, z, z, v, t)
        # the two-step route fails either way round ...
        assert L.linon_reload_shader_source(mv._view, mod_b.encode(), err, len(err)) == capi.ERR_COMPILE  # body still calls bend
        assert L.linon_set_field_function(mv._view, b"return tilt(v, 0.1);", err, len(err)) == capi.ERR_COMPILE  # module has no tilt
        # ... the pair goes through, and computes the same
        assert L.linon_reload_shader(mv._view, mod_b.encode(), b"return tilt(v, 0.1);", err, len(err)) == capi.OK, err.value
        assert mv.module_overrides() == ["tilt"]
        assert np.array_equal(mv.debug_eval_field(z, z, z, v, t), a)
        # a bad pair changes nothing
        assert L.linon_reload_shader(mv._view, mod_a.encode(), b"return tilt(v, 0.1);", err, len(err)) == capi.ERR_COMPILE
        assert mv.module_overrides() == ["tilt"] and np.array_equal(mv.debug_eval_field(z, z, z, v, t), a)
        assert L.linon_reload_shader(mv._view, None, b"return v;", err, len(err)) == capi.OK
        assert mv.module_overrides() == []


def test_a_rejected_mesh_leaves_the_current_one_in_place(product_lib, scene):
    verts, faces, center = scene
    with MainView(verts, faces, center, 64, 48) as mv:
        mv.reload_shader(PredefinedFunction.TranslationX.to_code())
        mv.render_flags(capi.RENDER_ADAPTIVE_OFF)
        mv.sync()
        before = mv.hit_ids().copy()
        bad_faces = faces.copy()
        bad_faces[5, 1] = 9999
        with pytest.raises(LinonError):
            mv.update_model(verts, bad_faces, center)
        bad_verts = verts.copy()
        bad_verts[3, 0] = np.nan
        with pytest.raises(LinonError):
            mv.update_model(bad_verts, faces, center)
        mv.set_option(capi.OPT_FORCE_BVH, 1)  # re-stages the mesh from the host copies: must be the good one
        mv.render_flags(capi.RENDER_ADAPTIVE_OFF)
        mv.sync()
        assert np.array_equal(mv.hit_ids(), before)
        mv.set_option(capi.OPT_FORCE_BVH, 0)


def test_an_overflowing_body_gives_the_same_frame_in_fast_and_strict_mode(product_lib, scene):
    """field_weight == 1 with a body whose velocity overflows: the reference computes (1 - 1) * cur_dir + 1 * v = NaN once
    cur_dir is infinite; the production build keeps the blend as written for generic bodies (no cur_dir = v shortcut), so
    non-finite endpoints appear at the same pixels as in strict mode."""
    verts, faces, center = scene
    body = "return v * 1e20 * (1.0 + p.x * p.x);"
    with MainView(verts, faces, center, 48, 32) as mv:
        frames = {}
        for math in (capi.MATH_STRICT, capi.MATH_FAST):
            mv.set_option(capi.OPT_MATH, math)
            mv.reload_shader(body)
            mv.render_flags(capi.RENDER_ADAPTIVE_OFF)
            mv.sync()
            frames[math] = mv.mapping_texture().copy(), mv.hit_ids().copy()
        a, b = frames[capi.MATH_STRICT], frames[capi.MATH_FAST]
        assert np.array_equal(np.isnan(a[0][..., :3]).any(axis=-1), np.isnan(b[0][..., :3]).any(axis=-1))
        assert np.isnan(a[0][..., :3]).any()
        assert (a[1] == b[1]).mean() >= 0.99  # FMA contraction may move a hit across a triangle edge
