"""build_ref.py — recipe that builds oracle/_ref/liblinon_wgsl_ref.so from the reference's sources WHERE THEY LIE.

TEST INFRASTRUCTURE ONLY (oracle/README.md).  Needs /root/reference (present in the build container only; the GPU box
receives the built .so, which is git-ignored but not gpurun-ignored).  Steps:
  1. read /root/reference/src/main_view.wgsl, src/overlay.wgsl, src/functions.rs (nothing is copied into the repo);
  2. take the nine predefined bodies out of functions.rs (Rust string literals of `PredefinedFunction::to_code`,
     src/functions.rs:33-128) and splice each into the module exactly like `with_field_function`
     (src/main_view.rs:852-861);
  3. make the variants the way MainView does, by string replacement on the shader text: Enhance
     (src/main_view.rs:785-787), adaptive sampling off (:764-776), outline rays (:811-814).  Two of the three
     replacement strings of the reference say `let` where the shader says `const` and therefore never match
     (SURVEY F7); the INTENDED replacement (`const ...`) is applied and recorded in info.json;
  4. transpile every module mechanically with wgsl2cpp.py, write the C++ under oracle/_ref/ and compile it with g++
     -O2 -ffp-contract=off -fno-fast-math -fopenmp.
"""
from __future__ import annotations

import concurrent.futures
import hashlib
import json
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE = os.path.dirname(HERE)
OUT = os.path.join(ORACLE, "_ref")
LIB = os.path.join(OUT, "liblinon_wgsl_ref.so")
REFERENCE = os.environ.get("LINON_REFERENCE", "/root/reference")
sys.path.insert(0, HERE)
import wgsl2cpp  # noqa: E402

CXX = "/usr/bin/g++" if os.access("/usr/bin/g++", os.X_OK) else "g++"  # the image's $CXX wrapper lacks libgomp.spec
CXXFLAGS = ["-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC", "-Wall",
            "-Wno-unused-function", "-Wno-unused-variable", "-Wno-unused-but-set-variable"]

FUNCTION_NAMES = ["Custom", "MirageSpherical", "MiragePlane", "MirageSphericalSigmoid", "MiragePlaneSigmoid",
                  "TranslationX", "TranslationZ", "Rotation", "LorenzAttractor", "RoesslerAttractor"]
PLACEHOLDER = ("fn field_function(p_prev: vec3<f32>, p: vec3<f32>, v0: vec3<f32>, v: vec3<f32>, t: f32) -> vec3<f32> "
               "{ return v; }")
# (name, what the reference's code searches for, what it puts there, fallback = the intended spelling)
VARIANTS = [
    ("default", None),
    ("enhance", [("let h_initial = 0.1;", "let h_initial = 0.001;")]),
    ("adaptive_off", [("let adaptive_sampling: bool = true;", "let adaptive_sampling: bool = false;"),
                      ("const adaptive_sampling: bool = true;", "const adaptive_sampling: bool = false;")]),
    ("outline", [("let sample_outline_rays: bool = false;", "let sample_outline_rays: bool = true;"),
                 ("const sample_outline_rays: bool = false;", "const sample_outline_rays: bool = true;")]),
]


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE, "src"))


def rust_string_literal(text: str, start: int):
    """Decodes the Rust string literal whose opening quote is at text[start]; returns (value, index after it)."""
    assert text[start] == '"'
    out, i = [], start + 1
    while text[i] != '"':
        c = text[i]
        if c == "\\":
            n = text[i + 1]
            if n == "\n":  # line continuation: skip the newline and leading whitespace
                i += 2
                while text[i] in " \t\r\n":
                    i += 1
                continue
            out.append({"n": "\n", "t": "\t", "r": "\r", "0": "\0", "\\": "\\", '"': '"', "'": "'"}[n])
            i += 2
        else:
            out.append(c)
            i += 1
    return "".join(out), i + 1


def predefined_bodies(functions_rs: str):
    """{name: WGSL body} from `impl PredefinedFunction { pub fn to_code ... }` (src/functions.rs:33-128)."""
    start = functions_rs.index("pub fn to_code")
    bodies = {}
    for m in re.finditer(r"Self::(\w+)\s*=>\s*\"", functions_rs[start:]):
        value, _ = rust_string_literal(functions_rs, start + m.end() - 1)
        bodies[m.group(1)] = value
    return bodies


def with_field_function(shader_src: str, body: str) -> str:
    """src/main_view.rs:852-861."""
    assert PLACEHOLDER in shader_src, "placeholder of with_field_function not found in main_view.wgsl"
    fn = ("fn field_function(p_prev: vec3<f32>, p: vec3<f32>, v0: vec3<f32>, v: vec3<f32>, t: f32) -> vec3<f32> {\n"
          + body + "\n}")
    return shader_src.replace(PLACEHOLDER, fn)


def make_variant(shader_src: str, replacements, notes):
    if replacements is None:
        return shader_src
    for k, (old, new) in enumerate(replacements):
        if old in shader_src:
            notes.append(dict(search=old, replace=new, spelling="reference" if k == 0 else "intended (SURVEY F7)"))
            return shader_src.replace(old, new)
    raise RuntimeError(f"no variant replacement matched: {replacements}")


def sha256(text: str) -> str:
    return hashlib.sha256(text.encode()).hexdigest()


def generate():
    """Writes the generated C++ under oracle/_ref/ and returns the list of sources to compile."""
    os.makedirs(OUT, exist_ok=True)
    read = lambda n: open(os.path.join(REFERENCE, "src", n), encoding="utf-8").read()  # noqa: E731
    main_view, overlay, functions_rs = read("main_view.wgsl"), read("overlay.wgsl"), read("functions.rs")
    bodies = predefined_bodies(functions_rs)
    assert [n for n in FUNCTION_NAMES if n not in bodies] == [], bodies.keys()
    info = dict(reference=REFERENCE, cxx=CXX, cxxflags=CXXFLAGS,
                sources={"src/main_view.wgsl": sha256(main_view), "src/overlay.wgsl": sha256(overlay),
                         "src/functions.rs": sha256(functions_rs)},
                functions=FUNCTION_NAMES, variants=[], transpiler=sha256(open(wgsl2cpp.__file__).read()))
    sources = []
    prologue = ('#include <cstring>\n#include <vector>\n#ifdef _OPENMP\n#include <omp.h>\n#endif\n'
                '#include "wgsl_runtime.hpp"\n#include "harness_types.hpp"\n')
    for fid, name in enumerate(FUNCTION_NAMES):
        module = main_view if fid == 0 else with_field_function(main_view, bodies[name])  # 0: the placeholder as shipped
        parts = [f"// GENERATED by oracle/wgsl_ref/build_ref.py from {REFERENCE}/src/main_view.wgsl + functions.rs "
                 f"({name}); do not commit.\n", prologue]
        for vid, (vname, repl) in enumerate(VARIANTS):
            notes = []
            text = make_variant(module, repl, notes)
            if fid == 0:
                info["variants"].append(dict(id=vid, name=vname, replacements=notes))
            parts.append(f"namespace fn{fid}_{vname} {{\n{wgsl2cpp.transpile(text)}\n"
                         f'#include "harness_main_view.inc"\n}}\n')
        table = ", ".join(f"&fn{fid}_{v}::module_entry" for v, _ in VARIANTS)
        parts.append(f"extern const RefModule* const wgslref_fn{fid}[4];\n"
                     f"const RefModule* const wgslref_fn{fid}[4] = {{{table}}};\n")
        path = os.path.join(OUT, f"gen_fn{fid}.cpp")
        open(path, "w").write("".join(parts))
        sources.append(path)
    path = os.path.join(OUT, "gen_overlay.cpp")
    open(path, "w").write(
        f"// GENERATED by oracle/wgsl_ref/build_ref.py from {REFERENCE}/src/overlay.wgsl; do not commit.\n" + prologue +
        f"namespace overlay_module {{\n{wgsl2cpp.transpile(overlay)}\n#include \"harness_overlay.inc\"\n}}\n"
        "int wgslref_run_overlay(int W, int H, const RefSettings* st, const uint8_t* a, const float* b, uint8_t* c, float* d, int n);\n"
        "int wgslref_run_overlay(int W, int H, const RefSettings* st, const uint8_t* a, const float* b, uint8_t* c, float* d, int n) "
        "{ return overlay_module::run_overlay(W, H, st, a, b, c, d, n); }\n"
        "void wgslref_run_vec2squareroots(int n, const float* a01, float* out);\n"
        "void wgslref_run_vec2squareroots(int n, const float* a01, float* out) { overlay_module::run_vec2squareroots(n, a01, out); }\n")
    sources.append(path)
    path = os.path.join(OUT, "gen_info.cpp")
    open(path, "w").write("extern const char* const wgslref_build_info;\nconst char* const wgslref_build_info = "
                          + json.dumps(json.dumps(info)) + ";\n")
    sources.append(path)
    json.dump(info, open(os.path.join(OUT, "info.json"), "w"), indent=1)
    return sources + [os.path.join(HERE, "ref_api.cpp")]


def build_custom(body: str, workdir: str, name: str = "custom", variant: str = "default", module_edits=()) -> str:
    """One module with a user-edited field-function body spliced in, compiled into `workdir`; exports
    `wgslref_custom_field(settings, n, p_prev, p, v0, v, t, out)` and `wgslref_custom_main_view(frame, outputs)`.
    variant: one of VARIANTS' names (the reference's own string replacements, src/main_view.rs:752-829);
    module_edits: (old, new) pairs applied to the text of main_view.wgsl first - what a user who edits the shader file
    does (src/main.rs:206-222); every `old` must occur.  For differential tests of the product's translator and hot reload."""
    main_view = open(os.path.join(REFERENCE, "src", "main_view.wgsl"), encoding="utf-8").read()
    for old, new in module_edits:
        assert old in main_view, f"module edit does not apply: {old!r}"
        main_view = main_view.replace(old, new)
    text = make_variant(with_field_function(main_view, body), dict(VARIANTS)[variant], [])
    src = os.path.join(workdir, name + ".cpp")
    lib = os.path.join(workdir, name + ".so")
    open(src, "w").write(
        '#include <cstring>\n#include <vector>\n#ifdef _OPENMP\n#include <omp.h>\n#endif\n'
        '#include "wgsl_runtime.hpp"\n#include "harness_types.hpp"\n'
        f'namespace custom_module {{\n{wgsl2cpp.transpile(text)}\n#include "harness_main_view.inc"\n}}\n'
        'extern "C" void wgslref_custom_field(const RefSettings* st, int n, const float* a, const float* b, const float* c, '
        'const float* d, const float* t, float* out) { custom_module::run_field(st, n, a, b, c, d, t, out); }\n'
        'extern "C" int wgslref_custom_main_view(const RefFrame* f, const RefOutputs* o) { return custom_module::run_main_view(f, o); }\n')
    r = subprocess.run([CXX, *CXXFLAGS, "-I", HERE, "-shared", src, "-o", lib], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"{src}:\n{r.stderr[-4000:]}")
    return lib


def _stamp() -> str:
    h = hashlib.sha256()
    for n in sorted(os.listdir(HERE)):
        if n.endswith((".py", ".hpp", ".inc", ".cpp")):
            h.update(open(os.path.join(HERE, n), "rb").read())
    for n in ("main_view.wgsl", "overlay.wgsl", "functions.rs"):
        h.update(open(os.path.join(REFERENCE, "src", n), "rb").read())
    return h.hexdigest()


def build(force: bool = False):
    """Builds the library if the reference is mounted; returns its path, or None when neither the mount nor a prebuilt
    library exists.  A prebuilt library (the GPU box) is used as it is."""
    if not available():
        return LIB if os.path.exists(LIB) else None
    stamp_path = os.path.join(OUT, "stamp")
    stamp = _stamp()
    if not force and os.path.exists(LIB) and os.path.exists(stamp_path) and open(stamp_path).read() == stamp:
        return LIB
    sources = generate()

    def compile_one(src):
        obj = os.path.join(OUT, os.path.basename(src)[:-4] + ".o")
        r = subprocess.run([CXX, *CXXFLAGS, "-I", HERE, "-c", src, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"{src}:\n{r.stderr[-4000:]}")
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        objs = list(ex.map(compile_one, sources))
    r = subprocess.run([CXX, "-shared", "-fopenmp", "-o", LIB, *objs], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stderr[-4000:])
    for o in objs:
        os.remove(o)
    open(stamp_path, "w").write(stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
