"""Hot reload of MARCH-LEVEL edits of main_view.wgsl: the constants of
src/main_view.wgsl:65-76 (eps, backface_culling, use_lighting, linear_mode, the light constants), the module switches
adaptive_sampling / sample_outline_rays (:227, :337) and the step-size / horizon / step-control literals of
nonlinear_ray_color and sample_rays (:241-267, :287-288) are mapped to macros of the kernel template instead of being
refused.  Each edit is rendered twice: by the device code of the march kernel executed on the host (tests/emu, strict math)
and by the REFERENCE's own shader with the same edit applied to its text, transpiled by oracle/wgsl_ref - bit for bit.
CPU only; the reference-side half needs the mount."""
import ctypes as C
import os
import tempfile

import numpy as np
import pytest

from conftest import REFERENCE
from linon_b200 import _capi as capi
from linon_b200.main_view import LinonError, PredefinedFunction as PF, compile_check, translate_wgsl_module

HAS_REF = os.path.isdir(os.path.join(REFERENCE, "src"))
W, H = 40, 28

# (name, [(old text, new text) in main_view.wgsl], what the translator must report)
CONST_EDITS = [
    ("eps", [("const eps: f32 = 0.0000001;", "const eps: f32 = 0.001;")], "#define LINON_EPS (0.00100000005f)"),
    ("backface_culling", [("const backface_culling: bool = false;", "const backface_culling: bool = true;")], "#define LINON_BACKFACE_CULLING 1"),
    ("use_lighting", [("const use_lighting: bool = true;", "const use_lighting: bool = false;")], "#define LINON_USE_LIGHTING 0"),
    ("light", [("const light_color: vec3<f32> = vec3<f32>(1.0, 1.0, 1.0);", "const light_color: vec3<f32> = vec3<f32>(0.9, 0.5, 0.25);"),
               ("const ambient_strength: f32 = 0.01;", "const ambient_strength: f32 = 0.2;"),
               ("const shininess: f32 = 64.0;", "const shininess: f32 = 8.0;")], "#define LINON_SHININESS (8.0f)"),
    ("linear_mode", [("const linear_mode: bool = false;", "const linear_mode: bool = true;")], "#define LINON_LINEAR_MODE 1"),
    ("adaptive_sampling", [("const adaptive_sampling: bool = true;", "const adaptive_sampling: bool = false;")], "// adaptive_sampling = 0"),
]
FN_EDITS = [
    ("h_initial", [("let h_initial = 0.1;", "let h_initial = 0.05;")], "// h_initial = 0.05"),
    ("horizon", [("for (; t <= 5.0;) {", "for (; t <= 3.0;) {")], "#define LINON_T_MAX (3.0f)"),
    ("step control", [("diff > 10.0 * last_diff && h > 0.002", "diff > 4.0 * last_diff && h > 0.004"), ("            h = 0.001;\n", "            h = 0.003;\n"),
                      ("if (2.0 * diff < last_diff", "if (3.0 * diff < last_diff")], "#define LINON_RESTORE_FACTOR (3.0f)"),
    ("sample rays", [("    let h = 0.001;\n    let steps: i32 = 5000;", "    let h = 0.002;\n    let steps: i32 = 2500;")], "#define LINON_SAMPLE_STEPS 2500"),
    # the pinhole camera of main_view (:355-356) and the reach of linear_mode's single ray_color call (:369)
    ("camera", [("let focal_length = 0.035;", "let focal_length = 0.05;"), ("let viewport_height = 0.025;", "let viewport_height = 0.03;")],
     "#define LINON_FOCAL_LENGTH (0.0500000007f)"),
    ("linear reach", [("const linear_mode: bool = false;", "const linear_mode: bool = true;"),
                      ("let color = ray_color(origin, dir, 100.0);", "let color = ray_color(origin, dir, 2.25);")], "#define LINON_LINEAR_MAX_DIST (2.25f)"),
    # the colours the sample kernel stores next to each position (`var<private> sample_colors`, :326-335)
    ("sample colours", [("    vec3<f32>(0.5, 0.5, 0.5),\n    vec3<f32>(0.0, 0.0, 1.0),", "    vec3<f32>(0.25, 0.75, 0.5),\n    vec3<f32>(0.0, 0.125, 1.0),")],
     "#define LINON_SAMPLE_COLORS {(1.0f), (0.0f), (0.0f)}"),
]
_FN_OF = {"sample rays": "sample_rays", "camera": "main_view", "linear reach": "main_view", "sample colours": "sample_colors"}

# edits INSIDE the shading block of ray_color (src/main_view.wgsl:199-218): the block is translated from the user's text
import march_edit_cases as mc  # noqa: E402

_EXPECT = {"depth": "const wg_f32 shininess = (32.0f)"}
SHADE_EDITS = [(name, edits, _EXPECT.get(name, "linon_ray_color_shade")) for name, edits in mc.SHADE_CASES.items() if name not in mc.OVERRIDES]


def partial_module(edits):
    """A module that only holds the edited declarations (modules need not repeat what they leave alone)."""
    return "\n".join(new for _, new in edits)


def reference_module(edits):
    text = open(os.path.join(REFERENCE, "src", "main_view.wgsl"), encoding="utf-8").read()
    for old, new in edits:
        assert old in text, old
        text = text.replace(old, new)
    return text


@pytest.mark.parametrize("name,edits,expect", CONST_EDITS, ids=[e[0] for e in CONST_EDITS])
def test_constant_edits_are_mapped_to_the_kernel_template(product_lib, name, edits, expect):
    text, names = translate_wgsl_module(partial_module(edits))
    assert expect in text, text
    assert all(new.split()[1].rstrip(":") in names for _, new in edits)
    assert compile_check(PF.MirageSpherical.to_code(), flags=capi.RENDER_ADAPTIVE_OFF, math_mode=capi.MATH_FAST,
                         module_src=partial_module(edits)) > 100_000


@pytest.mark.skipif(not HAS_REF, reason="needs the text of the reference's nonlinear_ray_color / sample_rays")
@pytest.mark.parametrize("name,edits,expect", FN_EDITS, ids=[e[0] for e in FN_EDITS])
def test_literal_edits_of_the_march_functions_are_mapped_too(product_lib, name, edits, expect):
    text, names = translate_wgsl_module(reference_module(edits))
    assert expect in text, text
    assert _FN_OF.get(name, "nonlinear_ray_color") in names
    assert compile_check(PF.MirageSpherical.to_code(), flags=0, math_mode=capi.MATH_FAST, module_src=reference_module(edits)) > 100_000


def test_the_unedited_module_installs_no_layer(product_lib):
    if HAS_REF:
        text, names = translate_wgsl_module(reference_module([]))
        assert names == [] and text.strip() == ""


@pytest.mark.parametrize("module,needle", [
    ("const eps: f32 = 0.5;", "outside the range"),
    ("const eps: f32 = 1.0 / 3.0;", "was edited"),                      # an expression, not a literal: the march itself
    ("const shininess: f32 = -1.0;", "outside the range"),
    ("const backface_culling: bool = 1;", "was edited"),
    ("struct Camera { origin: vec4<f32> };", "was edited"),
    ("fn hit_triangle(a: f32) -> f32 { return a; }", "was edited"),
])
def test_other_march_edits_are_still_refused_by_name(product_lib, module, needle):
    with pytest.raises(LinonError) as e:
        translate_wgsl_module(module)
    assert e.value.code == capi.ERR_UNSUPPORTED and needle in e.value.message, e.value.message


@pytest.mark.skipif(not HAS_REF, reason="needs the reference's text")
@pytest.mark.parametrize("edits,needle", [
    ([("let k2 = field_function(cur_point, cur_point + 0.5 * h * k1", "let k2 = field_function(cur_point, cur_point + 0.25 * h * k1")], "nothing else of it"),
    ([("let h_initial = 0.1;", "let h_initial = 7.0;")], "outside the range"),
    ([("    let steps: i32 = 5000;", "    let steps: i32 = 5050;")], "multiple of 100"),
    ([("cur_point = cur_point + step_dir;\n        t = t + h;\n        last_v = v;", "cur_point = cur_point + step_dir * 1.0;\n        t = t + h;\n        last_v = v;")], "was edited"),
])
def test_structural_edits_of_the_march_functions_are_refused(product_lib, edits, needle):
    with pytest.raises(LinonError) as e:
        translate_wgsl_module(reference_module(edits))
    assert e.value.code == capi.ERR_UNSUPPORTED and needle in e.value.message, e.value.message


# ---- the edited kernel against the edited reference shader ---------------------------------------------------------------
def _ref_frame(body, edits, variant, verts, faces, cam, st):
    from oracle import cpu_oracle as co
    from oracle import wgsl_reference as wr
    from oracle.wgsl_ref import build_ref
    with tempfile.TemporaryDirectory() as d:
        lib = C.CDLL(build_ref.build_custom(body, d, variant=variant, module_edits=edits))
        f = wr.RefFrame(W, H, co.make_camera(cam), co.make_settings(st), verts.shape[0], faces.shape[0], verts.ctypes.data,
                        faces.ctypes.data, None, None, 0)
        out = dict(rgba8=np.zeros((H, W, 4), np.uint8), endpoint=np.zeros((H, W, 4), np.float32), ray_samples=np.zeros((800, 8), np.float32))
        o = wr.RefOutputs(None, out["rgba8"].ctypes.data, out["endpoint"].ctypes.data, out["ray_samples"].ctypes.data)
        assert lib.wgslref_custom_main_view(C.byref(f), C.byref(o)) == 0
    return out


@pytest.mark.skipif(not HAS_REF, reason="needs the text of the reference's ray_color")
@pytest.mark.parametrize("name,edits,expect", SHADE_EDITS, ids=[e[0] for e in SHADE_EDITS])
def test_edits_of_the_shading_block_are_translated(product_lib, name, edits, expect):
    text, names = translate_wgsl_module(reference_module(edits))
    assert expect in text and "#define LINON_USER_SHADE 1" in text, text
    assert "ray_color" in names
    assert compile_check(PF.MirageSpherical.to_code(), flags=0, math_mode=capi.MATH_FAST, module_src=reference_module(edits)) > 100_000


@pytest.mark.skipif(not HAS_REF, reason="needs the text of the reference's ray_color")
def test_a_partial_module_may_hold_ray_color_alone(product_lib):
    full = reference_module(SHADE_EDITS[0][1])
    fn = full[full.index("fn ray_color("):full.index("struct NonlinearRayColorResult")]
    text, names = translate_wgsl_module(fn + "\nconst ambient_strength: f32 = 0.25;\n")
    assert "const wg_f32 ambient_strength = (0.25f)" in text and set(names) == {"ray_color", "ambient_strength"}


@pytest.mark.skipif(not HAS_REF, reason="needs the text of the reference's ray_color")
def test_the_blocks_variables_may_shadow_a_module_constant_like_in_wgsl(product_lib):
    module = reference_module([("let diff = intensity;", "let diff = intensity * d1_scale;"),
                               ("const PI: f32 = 3.141592653589793;", "const PI: f32 = 3.141592653589793;\nconst d1_scale: f32 = 0.5;\nconst d1: f32 = 2.0;")])
    assert translate_wgsl_module(module)[1] == ["d1_scale", "d1", "ray_color"]
    assert compile_check(PF.Rotation.to_code(), flags=0, math_mode=capi.MATH_FAST, module_src=module) > 100_000


@pytest.mark.skipif(not HAS_REF, reason="needs the text of the reference's ray_color")
@pytest.mark.parametrize("edits,needle", [
    # the face loop of ray_color is the kernel's own intersection code: still refused
    ([("(t < 0.0 || t_new < t)) {", "(t < 0.0 || t_new <= t)) {")], "was edited"),
    ([("    return vec4<f32>(0.0, 0.0, 0.0, 0.0);\n}\n\nstruct NonlinearRayColorResult", "    return vec4<f32>(0.0, 0.0, 0.0, 1.0);\n}\n\nstruct NonlinearRayColorResult")], "was edited"),
])
def test_edits_of_ray_color_outside_the_shading_block_are_refused(product_lib, edits, needle):
    with pytest.raises(LinonError) as e:
        translate_wgsl_module(reference_module(edits))
    assert e.value.code == capi.ERR_UNSUPPORTED and needle in e.value.message, e.value.message


@pytest.mark.skipif(not HAS_REF, reason="needs the text of the reference's ray_color")
def test_an_error_inside_the_shading_block_cites_the_modules_line(product_lib):
    with pytest.raises(LinonError) as e:
        translate_wgsl_module(reference_module([("let diff = intensity;", "let diff = intensity * faces.data[0].a;")]))
    assert e.value.code == capi.ERR_COMPILE and "main_view.wgsl(212)" in e.value.message and "faces" in e.value.message, e.value.message


@pytest.mark.skipif(not HAS_REF, reason="the reference-side half needs the mount")
@pytest.mark.parametrize("name,edits,expect", CONST_EDITS + FN_EDITS + SHADE_EDITS, ids=[e[0] for e in CONST_EDITS + FN_EDITS + SHADE_EDITS])
def test_edited_kernel_equals_the_edited_reference_shader(product_lib, name, edits, expect):
    """Strict math on the host emulation of the device code: endpoints and hit pattern bit for bit, colours within 1/255
    (pow: libm here on both sides, the rgba8 rounding is the reference's)."""
    from emu import emu
    from oracle import np_oracle as npo
    verts, faces, center = npo.default_scene()
    verts, faces = np.ascontiguousarray(verts, np.float32), np.ascontiguousarray(faces, np.uint32)
    cam = npo.camera_uniform(npo.default_camera(center, W, H))
    st = dict(npo.default_settings())
    fn = PF.MirageSpherical
    adaptive_module = name != "adaptive_sampling"
    # as the application renders (adaptive sampling per the module), except: the step-control literals only act in adaptive frames
    flags = 0
    ref = _ref_frame(fn.to_code(), edits, "default", verts, faces, cam, st)
    lib = emu.build(fn.to_code(), flags, capi.MATH_STRICT, module_src=reference_module(edits))
    if name in ("sample rays", "camera", "sample colours"):  # the sample rays leave through the same pinhole (:475-478)
        from oracle import cpu_oracle as co
        pos = co.sample_positions(cam, st, W, H)
        got = emu.march(lib, verts, faces, cam, st, W, H, sample_pos=pos)
        assert np.array_equal(got["ray_samples"].view(np.uint32), ref["ray_samples"].view(np.uint32))
        if name != "camera":
            return
    got = emu.march(lib, verts, faces, cam, st, W, H)
    assert np.abs(got["rgba8"].astype(int) - ref["rgba8"].astype(int)).max() <= 1, name
    if "linear" not in name:  # linear_mode never writes `mapping`
        assert np.array_equal(got["endpoint"].view(np.uint32), ref["endpoint"].view(np.uint32)), name
        assert (got["rgba8"][..., 3] == 255).all()
    else:
        assert (got["endpoint"] == 0).all() and (got["rgba8"][..., 3] < 255).any()  # alpha = hit distance, 0 on a miss
    if name == "adaptive_sampling":
        assert int(got["counters"][0]) == W * H * 51 and int(got["counters"][2]) == 0 and not adaptive_module
    if name == "horizon":
        assert int(got["counters"][0]) < W * H * 80
