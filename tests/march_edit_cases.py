"""Edits to the march-level constants of main_view.wgsl (src/main_view.wgsl:65-76, :227) as PARTIAL modules - a module need
not repeat the declarations it leaves alone - shared by the fixture generator (tests/golden/make_march_edits.py, which
applies the same edit to the reference's own text) and the GPU test (tests/test_gpu_units.py)."""

# name -> [(text in the reference's main_view.wgsl, replacement)]
CASES = {
    "eps": [("const eps: f32 = 0.0000001;", "const eps: f32 = 0.001;")],
    "backface_culling": [("const backface_culling: bool = false;", "const backface_culling: bool = true;")],
    "use_lighting": [("const use_lighting: bool = true;", "const use_lighting: bool = false;")],
    "light": [("const light_color: vec3<f32> = vec3<f32>(1.0, 1.0, 1.0);", "const light_color: vec3<f32> = vec3<f32>(0.9, 0.5, 0.25);"),
              ("const ambient_strength: f32 = 0.01;", "const ambient_strength: f32 = 0.2;"),
              ("const shininess: f32 = 64.0;", "const shininess: f32 = 8.0;")],
    "linear_mode": [("const linear_mode: bool = false;", "const linear_mode: bool = true;")],
    "adaptive_sampling": [("const adaptive_sampling: bool = true;", "const adaptive_sampling: bool = false;")],
}
W, H = 64, 48

# Edits INSIDE the shading block of ray_color (src/main_view.wgsl:199-218), which the CUDA path translates from the user's
# text.  Such a module has to carry the whole function (its face loop must be the reference's, token for token), so these
# cases exist as text only where the reference is mounted: tests/golden/make_march_edits.py writes the edited modules to
# tests/golden/_shade_modules/ (git-ignored: they are the reference's source) next to the frames it commits.
_NORMAL = "        let normal = normalize(cross(d1, d2));\n"
SHADE_CASES = {
    "tint": [("let result = (ambient + diffuse + specular) * abs(normal);",
              "let result = (ambient + 0.5 * diffuse + specular) * abs(normal) + vec3<f32>(0.1, 0.0, 0.05);")],
    "noop": [(_NORMAL, _NORMAL + "        let unused = 2.0 * t;\n")],
    # a hit the shader declines (alpha 0): the reference keeps marching through that wall (has_color = color.a > 0.0, :262-263)
    "declined": [(_NORMAL, _NORMAL + "        if (abs(normal.x) > 0.5) {\n            return vec4<f32>(1.0, 0.0, 0.0, 0.0);\n        }\n")],
    "depth": [("        let ambient = ambient_strength * light_color;\n",
               "        return vec4<f32>(vec3<f32>(t / max_dist, shininess / 128.0, ambient_strength), 1.0);\n"
               "        let ambient = ambient_strength * light_color;\n"),
              ("const shininess: f32 = 64.0;", "const shininess: f32 = 32.0;")],
    "linear_shade": [("const linear_mode: bool = false;", "const linear_mode: bool = true;"),
                     ("return vec4<f32>(result, t);", "return vec4<f32>(result.zyx, 0.5 * t);")],
}
# literals of main_view the kernel template takes as macros: the pinhole camera (:355-356), the reach of linear_mode's ray (:369).
# Such a module carries the whole entry point, so these too exist as text only where the reference is mounted.
SHADE_CASES.update({
    "camera": [("let focal_length = 0.035;", "let focal_length = 0.05;"), ("let viewport_height = 0.025;", "let viewport_height = 0.03;")],
    "linear_reach": [("const linear_mode: bool = false;", "const linear_mode: bool = true;"),
                     ("let color = ray_color(origin, dir, 100.0);", "let color = ray_color(origin, dir, 2.25);")],
})
OVERRIDES = {"camera": "main_view", "linear_reach": "main_view"}  # the declaration a case edits (default: ray_color)
SHADE_MODULE_DIR = "_shade_modules"


def edited_reference_module(reference_root, edits):
    import os
    text = open(os.path.join(reference_root, "src", "main_view.wgsl"), encoding="utf-8").read()
    for old, new in edits:
        assert old in text, old
        text = text.replace(old, new)
    return text


def shade_module_text(name, reference_root="/root/reference"):
    """The edited module of a SHADE case: from the mounted reference, else from the scratch copy the generator left; None if neither."""
    import os
    if os.path.isdir(os.path.join(reference_root, "src")):
        return edited_reference_module(reference_root, SHADE_CASES[name])
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", SHADE_MODULE_DIR, name + ".wgsl")
    return open(p, encoding="utf-8").read() if os.path.exists(p) else None


def partial_module(name):
    return "\n".join(new for _, new in CASES[name])
