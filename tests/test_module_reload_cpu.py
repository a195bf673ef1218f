"""Whole-shader hot reload: the helper layer of an edited main_view.wgsl, translated and compiled
without a GPU.  Reference path: file watcher src/main.rs:206-222 -> Application::reload_compute_shader
(src/application.rs:228-232) -> MainView::reload_shader(device, Some(new_src), field_function) (src/main_view.rs:649-690);
README.md:46 ("Additional helper functions can be defined by adding them to the compute shader").  CPU only."""
import os
import re

import pytest

from conftest import REFERENCE
from linon_b200 import _capi as capi
from linon_b200.main_view import (LinonError, PredefinedFunction, compile_check, scan_wgsl_module, translate_wgsl_module)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# a self-contained "edited module": two constants, a new helper that uses both, and a corrected rotateZ (the
# reference's drops z).  Modules need not repeat the declarations they leave alone.
USER_MODULE = """
const swirl_gain: f32 = 0.25;   // how strongly the swirl bends the ray
const up_axis = vec3<f32>(0.0, 1.0, 0.0);

fn swirl(p: vec3<f32>, v: vec3<f32>, k: f32) -> vec3<f32> {
    let r = cross(up_axis, p);
    if (k > 0.5) {
        return normalize(v + swirl_gain * k * r);
    }
    return v;
}

/* keeps the third component */
fn rotateZ(v: vec3<f32>, phi: f32) -> vec3<f32> {
    return vec3<f32>(cos(phi) * v.x - sin(phi) * v.y, sin(phi) * v.x + cos(phi) * v.y, v.z);
}
"""
USER_BODY = "return swirl(p, rotateZ(v, PI / 4.0), 0.7);"


def test_new_helpers_constants_and_overrides_are_translated(product_lib):
    cuda, names = translate_wgsl_module(USER_MODULE)
    assert names == ["swirl_gain", "up_axis", "swirl", "rotateZ"]
    # member functions of LinonField<S>: class-scope lookup finds them before the prelude's helper of the same name
    assert "static __device__ __forceinline__ vec3f swirl(const vec3f p , const vec3f v , const wg_f32 k )" in cuda
    assert "static __device__ __forceinline__ vec3f rotateZ(" in cuda
    assert "w_cross(up_axis, p)" in cuda and "const wg_f32 swirl_gain  = 0.25f;" in cuda
    # diagnostics map back to the module's own lines
    assert '#line 5 "main_view.wgsl"' in cuda and '#line 14 "main_view.wgsl"' in cuda


@pytest.mark.parametrize("flags,math", [(0, capi.MATH_FUSED), (capi.RENDER_ENHANCE, capi.MATH_FAST),
                                        (capi.RENDER_ADAPTIVE_OFF, capi.MATH_STRICT)])
def test_module_layer_compiles_for_sm100a(product_lib, flags, math):
    assert compile_check(USER_BODY, flags=flags, math_mode=math, module_src=USER_MODULE) > 100_000
    # a predefined mirage body still builds with a helper layer installed (generic translation, no hand-fused splice)
    assert compile_check(PredefinedFunction.MirageSphericalSigmoid.to_code(), flags=flags, math_mode=math,
                         module_src=USER_MODULE) > 100_000


def test_errors_cite_the_module_and_unused_helpers_are_still_compiled(product_lib):
    with pytest.raises(LinonError) as e:  # the helper is never called; naga would reject the module all the same
        compile_check("return v;", module_src=USER_MODULE.replace("cross(up_axis, p)", "crossx(up_axis, p)"))
    assert e.value.code == capi.ERR_COMPILE and "main_view.wgsl(6)" in e.value.message and "crossx" in e.value.message
    with pytest.raises(LinonError) as e:
        compile_check("return swirl2(p);", module_src=USER_MODULE)
    assert e.value.code == capi.ERR_COMPILE and "field_function.wgsl(1)" in e.value.message
    with pytest.raises(LinonError) as e:
        translate_wgsl_module("fn broken(a: f32 -> f32 { return a; }")
    assert e.value.code == capi.ERR_COMPILE and "broken" in e.value.message
    with pytest.raises(LinonError) as e:
        translate_wgsl_module("fn eval(a: f32) -> f32 { return a; }")
    assert e.value.code == capi.ERR_COMPILE and "reserved" in e.value.message
    with pytest.raises(LinonError) as e:
        translate_wgsl_module("fn twice(a: f32) -> f32 { return a; }\nfn twice(a: f32) -> f32 { return 2.0 * a; }")
    assert "redefinition of 'twice'" in e.value.message


def test_edits_to_the_march_itself_are_refused_by_name(product_lib):
    """The CUDA march is a hand-written kernel template: a module that redefines one of its functions, constants, structs
    or bindings is not silently ignored."""
    for text, name in (("fn ray_color(origin: vec3<f32>, direction: vec3<f32>, max_dist: f32) -> vec4<f32> { return vec4<f32>(0.0); }", "ray_color"),
                       ("const eps: f32 = 1.0 / 1000.0;", "eps"),  # a literal value would be mapped (tests/test_march_edits_cpu.py)
                       ("struct Camera { origin: vec4<f32> }", "Camera"),
                       ("@group(0) @binding(7)\nvar<uniform> extra: vec4<f32>;", "extra")):
        with pytest.raises(LinonError) as e:
            translate_wgsl_module(USER_MODULE + text)
        assert e.value.code == capi.ERR_UNSUPPORTED and f"'{name}'" in e.value.message, e.value.message
    # the placeholder field_function of the module is ignored whatever it holds
    _, names = translate_wgsl_module("fn field_function(p_prev: vec3<f32>, p: vec3<f32>, v0: vec3<f32>, v: vec3<f32>, t: f32) -> vec3<f32> { return -v; }")
    assert names == []


def test_scan_lists_declarations_with_attributes_and_lines(product_lib):
    items = scan_wgsl_module("@group(0) @binding(2)\nvar<uniform> camera: Camera;\n\nconst PI: f32 = 3.0;\n"
                             "@compute @workgroup_size(8, 8)\nfn main_view(@builtin(global_invocation_id) gid: vec3<u32>) {\n}\n")
    assert [(k, n, ln) for k, n, _, ln in items] == [("var", "camera", 1), ("const", "PI", 4), ("fn", "main_view", 5)]
    a = scan_wgsl_module("fn f(a: f32) -> f32 { return a * 2.0; }")[0][2]
    b = scan_wgsl_module("fn f(a: f32) -> f32 {\n    // doubled\n    return a * 2.0;\n}")[0][2]
    c = scan_wgsl_module("fn f(a: f32) -> f32 { return a * 3.0; }")[0][2]
    assert a == b != c  # comments and white space do not count


# ---- pins against the mounted reference (build container only) ---------------------------------------------------
needs_reference = pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "src")), reason="reference not mounted")


def reference_module():
    return open(os.path.join(REFERENCE, "src", "main_view.wgsl"), encoding="utf-8").read()


@needs_reference
def test_the_hash_table_is_the_reference_module(product_lib):
    items = scan_wgsl_module(reference_module(), full=True)
    table = open(os.path.join(ROOT, "linon_b200", "csrc", "host", "module_items.inc")).read()
    rows = re.findall(r'\{"(\w+)", "(\w+)", 0x([0-9a-f]{16})ull, 0x([0-9a-f]{16})ull\}', table)
    assert [(k, n, f"{h:016x}", f"{hm:016x}") for k, n, h, _, hm, _ in items] == rows and len(rows) == 45
    # the literal tables of the two march functions (values only) and the ordinals of the literals that are mapped
    lits = open(os.path.join(ROOT, "linon_b200", "csrc", "host", "module_literals.inc")).read()
    by_name = {n: l for _, n, _, _, _, l in items}
    for fn, holes in (("nonlinear_ray_color", {9: 0.1, 10: 5.0, 19: 10.0, 20: 0.002, 21: 0.001, 24: 2.0}), ("sample_rays", {1: 0.001, 2: 5000.0}),
                      ("main_view", {4: 0.035, 5: 0.025, 10: 100.0})):
        m = re.search(r"kLit_%s\[\] = \{([^}]*)\}" % fn, lits)
        assert [float(x) for x in m.group(1).split(",")] == by_name[fn]
        m = re.search(r"kLitHoles_%s\[\] = \{([^}]*)\}" % fn, lits)
        assert [int(x) for x in m.group(1).split(",")] == list(holes)
        assert all(by_name[fn][k] == v for k, v in holes.items())


@needs_reference
def test_the_rest_hash_of_ray_color_is_the_reference_function_without_its_shading_block(product_lib):
    """module_shade.inc: ray_color of the reference with the statements of `if (t > 0.0 && t < max_dist) { ... }` removed hashes
    to the committed value - what tells an edit inside the block from one to the face loop (wgsl_translate.cpp)."""
    src = reference_module()
    start = src.index("fn ray_color(")
    cond = src.rindex("if (t > 0.0 && t < max_dist) {", start, src.index("struct NonlinearRayColorResult", start))
    depth, k = 0, src.index("{", cond)
    open_brace = k
    while True:
        depth += {"{": 1, "}": -1}.get(src[k], 0)
        if depth == 0:
            break
        k += 1
    emptied = src[:open_brace + 1] + src[k:]
    h = [it[2] for it in scan_wgsl_module(emptied, full=True) if it[1] == "ray_color"]
    committed = open(os.path.join(ROOT, "linon_b200", "csrc", "host", "module_shade.inc")).read()
    assert re.search(r"kRayColorRestHash = 0x([0-9a-f]{16})ull", committed).group(1) == f"{h[0]:016x}"
    # an emptied block is itself an edit inside the block: accepted, and every hit is declined (ray_color's final return)
    text, names = translate_wgsl_module(emptied)
    assert names == ["ray_color"] and "linon_ray_color_shade" in text


@needs_reference
def test_the_unedited_reference_module_installs_nothing(product_lib):
    src = reference_module()
    assert translate_wgsl_module(src) == ("", [])
    assert translate_wgsl_module(src.replace("\n    ", "\n\t"))[1] == []  # re-indented


@needs_reference
def test_an_edited_helper_pulls_in_the_unchanged_helpers_that_call_it(product_lib):
    src = reference_module()
    edited = src.replace("let c2 = 0.00366;", "let c2 = 0.004;")
    assert edited != src
    cuda, names = translate_wgsl_module(edited)
    assert names == ["refraction_index", "refraction"]  # refraction is unchanged but must bind to the new index
    assert "const auto c2 = 0.004f;" in cuda and "w_reflect(v_in, n_ref)" in cuda
    assert compile_check(PredefinedFunction.MiragePlane.to_code(), math_mode=capi.MATH_STRICT, module_src=edited) > 100_000
    # PI may be edited (it is only used by field functions); so may the march's own constants (mapped to macros of the
    # kernel template, tests/test_march_edits_cpu.py) - the structure of the march functions may not
    assert translate_wgsl_module(src.replace("const PI: f32 = 3.141592653589793;", "const PI: f32 = 3.14;"))[1] == ["PI"]
    assert translate_wgsl_module(src.replace("const shininess: f32 = 64.0;", "const shininess: f32 = 8.0;"))[1] == ["shininess"]
    for old, new, name in (("let f = 1.0 / a;", "let f = 2.0 / a;", "hit_triangle"),
                           ("if (t_new > 0.0 && t_new < max_dist", "if (t_new >= 0.0 && t_new < max_dist", "ray_color")):  # its face loop
        assert old in src
        with pytest.raises(LinonError) as e:
            translate_wgsl_module(src.replace(old, new))
        assert e.value.code == capi.ERR_UNSUPPORTED and f"'{name}'" in e.value.message
    # ... except inside the shading block of ray_color, which is translated from the user's text
    assert translate_wgsl_module(src.replace("let spec = pow(intensity, shininess);", "let spec = pow(intensity, 2.0 * shininess);"))[1] == ["ray_color"]
