"""The structural match that selects the hand-fused mirage restatement (csrc/host/mirage_match.hpp): which bodies belong to
the family, which constants are read from the text, and that every member compiles for sm_100a with those constants.  CPU only."""
import os

import pytest

from linon_b200 import _capi as capi
from linon_b200.main_view import PredefinedFunction as PF, compile_check, match_mirage

MIRAGES = [PF.MirageSpherical, PF.MiragePlane, PF.MirageSphericalSigmoid, PF.MiragePlaneSigmoid]


def test_the_predefined_bodies_match_with_their_own_constants(product_lib):
    for f in PF:
        if f == PF.Custom:
            continue
        kind, p = match_mirage(f.to_code())
        assert kind == (int(f) if f in MIRAGES else 0), f
    _, p = match_mirage(PF.MirageSphericalSigmoid.to_code())
    assert p == dict(t_env=15.0, t_src=200.0, max_dist=0.25, c=(-0.5, 0.5, -0.5), sig_k=12.0, sig_c=6.0)
    _, p = match_mirage(PF.MiragePlane.to_code())
    assert (p["t_src"], p["max_dist"], p["c"]) == (30.0, 0.01, (0.0, 0.1, 0.0))


def test_comments_white_space_and_literal_spelling_do_not_count(product_lib):
    b = PF.MirageSphericalSigmoid.to_code()
    for edit in (b + " ", b + "\n// a remark\n", b.replace("\n", "\n   "), b.replace("15.0;", "15;"), b.replace("0.25", "2.5e-1"),
                 b.replace("let max_dist", "/* nested /* comment */ */ let max_dist"), b.replace("200.0", "200.0f")):
        assert match_mirage(edit)[0] == 3, edit


def test_edited_constants_stay_in_the_family(product_lib):
    b = PF.MirageSphericalSigmoid.to_code()
    e = (b.replace("200.0", "201.5").replace("vec3<f32>(-0.5, 0.5, -0.5)", "vec3<f32>( -0.4,0.55 , -.6)")
         .replace("12.0", "10.0").replace("6.0", "5.0").replace("0.25", "0.3"))
    kind, p = match_mirage(e)
    assert kind == 3 and p == dict(t_env=15.0, t_src=201.5, max_dist=0.3, c=(-0.4, 0.55, -0.6), sig_k=10.0, sig_c=5.0)
    kind, p = match_mirage(PF.MiragePlane.to_code().replace("0.1", "0.15").replace("30.0", "45").replace("0.01;", "0.02;"))
    assert kind == 2 and (p["t_src"], p["max_dist"], p["c"][1]) == (45.0, 0.02, 0.15)
    kind, p = match_mirage(PF.MirageSpherical.to_code().replace("15.0", "-5.0"))
    assert kind == 1 and p["t_env"] == -5.0


@pytest.mark.parametrize("edit", [
    lambda b: b.replace("let t_env = 15.0;", "let t_env = 15.0 + 0.0;"),          # an expression, not a literal
    lambda b: b.replace("12.0", "10.0", 1),                                        # different profiles at p_prev and p
    lambda b: b.replace("200.0", "15.5"),                                          # |t_src - t_env| < 1: outside the validated range
    lambda b: b.replace("200.0", "-300.0"),                                        # below absolute zero: B(t) changes sign
    lambda b: b.replace("0.25", "0.0"),                                            # max_dist = 0
    lambda b: PF.MiragePlane.to_code().replace("0.01;", "0.03;"),                  # a thick plane layer (see mirage_match.hpp: validate)
    lambda b: b.replace("clamp(0.0, 1.0,", "clamp(0.0, 2.0,"),                     # a fixed literal changed
    lambda b: b.replace("normalize(center_dest)", "normalize(center_dest) * 1.0"),
    lambda b: b.replace("p_prev - center", "p - center"),
    lambda b: b.replace("return refraction(t_in, t_out, v, normal);", "return refraction(t_out, t_in, v, normal);"),
    lambda b: b + "\nreturn v;",
    lambda b: b.replace("vec3<f32>(-0.5", "vec3<f32>(1e9"),
    lambda b: "",
    lambda b: "let x = \"",
])
def test_anything_else_runs_the_generic_translation(product_lib, edit):
    assert match_mirage(edit(PF.MirageSphericalSigmoid.to_code()))[0] == 0


def test_a_changed_plane_normal_is_not_in_the_family(product_lib):
    b = PF.MiragePlaneSigmoid.to_code()
    assert match_mirage(b)[0] == 4
    assert match_mirage(b.replace("vec3<f32>(0.0, 1.0, 0.0)", "vec3<f32>(0.0, 1.0, 0.1)"))[0] == 0


@pytest.mark.parametrize("fn", MIRAGES)
def test_family_members_compile_with_their_constants(product_lib, fn, monkeypatch):
    """Fast-math build of the fused restatement (with the quiet zone) for a body with every constant edited."""
    b = fn.to_code().replace("15.0", "12.5").replace("200.0", "180.0").replace("30.0", "42.0").replace("0.25", "0.3")
    b = b.replace("0.01;", "0.015;").replace("-0.5, 0.5, -0.5", "-0.45, 0.52, -0.55").replace("0.0, 0.1, 0.0", "0.0, 0.12, 0.0")
    b = b.replace("12.0", "10.0").replace("6.0", "5.0")
    assert match_mirage(b)[0] == int(fn)
    monkeypatch.setenv("LINON_CHECK_QUIET_ZONE", "2")
    for flags in (0, capi.RENDER_ENHANCE, capi.RENDER_ADAPTIVE_OFF):
        assert compile_check(b, flags=flags, math_mode=capi.MATH_FAST) > 0


@pytest.mark.parametrize("index", [0, 1, 2, 3])
def test_emulated_fused_build_of_edited_bodies_matches_the_reference_shader(product_lib, index):
    """The device code of the fused restatement, parameterised by the constants of an EDITED body and executed on the host
    (tests/emu), against the frame the reference's own transpiled shader renders for that body (tests/golden/edited_mirages.npz,
    adaptive off): colour and endpoint under the north-star tolerances, high-Lyapunov pixels excluded; the quiet zone follows the
    edited constants (straight steps happen, every step is counted)."""
    import numpy as np
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import edited_bodies
    from emu import emu
    from linon_b200 import parity
    from oracle import np_oracle as npo
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "edited_mirages.npz"))
    name, (kind, body) = list(edited_bodies.bodies().items())[index]
    assert str(z["names"][index]) == name and match_mirage(body)[0] == kind
    W, H = 96, 72
    verts, faces, center = npo.default_scene()
    cam = npo.camera_uniform(npo.default_camera(center, W, H))
    st = dict(npo.default_settings())
    lib = emu.build(body, capi.RENDER_ADAPTIVE_OFF, capi.MATH_FAST, quiet_zone=1)
    got = emu.march(lib, verts, faces, cam, st, W, H)
    assert int(got["counters"][0]) == W * H * 51 and int(got["counters"][15]) > 0
    ref = dict(hit_id=got["hit_id"], rgba8=z[f"rgba8_{index}"], endpoint=z[f"endpoint_{index}"])
    m = parity.compare_frames(got, ref, exclude=z[f"high_{index}"])
    assert m["colour_within_1"] >= 0.999 and m["endpoint_within_1e3"] >= 0.999, m
    generic = emu.march(emu.build(body + "\nreturn v;", capi.RENDER_ADAPTIVE_OFF, capi.MATH_FAST), verts, faces, cam, st, W, H)
    assert int(generic["counters"][15]) == 0  # not in the family (a second return statement): no zone, generic translation
    mg = parity.compare_frames(got, generic, exclude=z[f"high_{index}"])
    assert mg["hit_group_agree"] >= 0.999, mg
