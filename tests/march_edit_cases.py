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


def partial_module(name):
    return "\n".join(new for _, new in CASES[name])
