"""The C-ABI library loads without a GPU, exports every symbol include/*.h declares, keeps its POD layouts, and
fails loudly (no CPU fallback) when asked to compute without a device.  CPU only."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from linon_b200 import _capi as capi
from linon_b200.main_view import ArcballCamera, LinonError, MainView, default_scene, orbit_camera
from oracle import np_oracle as npo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in ("linon_b200.h", "linon_b200_host.h"):
        text = open(os.path.join(ROOT, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(linon_\w+)\s*\(", text))
    return names


def test_every_declared_symbol_is_exported_and_bound(product_lib):
    declared = declared_symbols()
    assert len(declared) > 50
    bound = {name for name, _, _ in capi.PROTOTYPES}
    assert declared == bound, (declared - bound, bound - declared)
    for name in declared:
        assert hasattr(product_lib, name), name


def test_pod_layouts_match_the_reference_uniforms():
    assert C.sizeof(capi.Camera) == 112       # CameraUniform, src/main_view.rs:12-19
    assert C.sizeof(capi.Settings) == 24      # Settings, src/main_view.rs:43-51
    assert capi.Settings.overlay_mode.offset == 12 and capi.Settings.lyapunov_scaling.offset == 20
    assert capi.Camera.view_direction.offset == 16 and capi.Camera.view_matrix.offset == 48


def test_no_gpu_means_a_loud_failure_not_a_fallback(product_lib, scene):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    verts, faces, center = scene
    with pytest.raises(LinonError) as e:
        MainView(verts, faces, center, 64, 48)
    assert e.value.code == capi.ERR_NO_DRIVER
    assert "no CPU fallback" in e.value.message
    out = C.c_void_p()
    assert product_lib.linon_create(0, 64, 48, C.byref(out)) == capi.ERR_NO_DRIVER and not out.value
    t = C.c_double()
    assert product_lib.linon_measure_fp32_peak(0, C.byref(t), None) != capi.OK


def test_product_never_touches_the_oracle():
    """The product path may not import, link or execute anything under oracle/."""
    for root, _, files in os.walk(os.path.join(ROOT, "linon_b200")):
        if os.path.basename(root) in ("build", "lib", "__pycache__"):
            continue
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h", ".hpp")):
                text = open(os.path.join(root, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert "liblinon_oracle" not in text and "oracle/" not in text.replace("the CPU oracle", ""), f


def test_host_mirror_scene_equals_the_numpy_restatement(scene):
    verts, faces, center = scene
    v2, f2, c2 = npo.default_scene()
    assert np.array_equal(verts, v2) and np.array_equal(faces, f2) and np.array_equal(center, c2)
    raw, _, _ = default_scene(negate_xz=False)  # the `r` key scene, src/application.rs:194-197
    rv, _ = npo.cornell_box()
    assert np.array_equal(raw, npo.normalize_vertices(rv))


def test_host_mirror_arcball_equals_the_numpy_restatement(scene):
    _, _, center = scene
    for (w, h) in ((800, 600), (1920, 1080)):
        a = ArcballCamera(center, 1.0, (w, h))
        b = npo.Arcball(center, 1.0, [w, h])
        a.zoom(-1.0, 1.0)
        b.zoom(-1.0, 1.0)
        for prev, cur in (((w / 2, h / 2), (w / 2 + 50, h / 2 + 20)), ((10, 10), (w - 5, 30)), ((w, h), (0, 0))):
            a.rotate(prev, cur)
            b.rotate(prev, cur)
            u = npo.camera_uniform(b)
            assert np.allclose(a.eye_pos(), u["origin"], atol=1e-6)
            assert np.allclose(a.eye_dir(), u["view_direction"], atol=1e-6)
            assert np.allclose(a.up_dir(), u["up"], atol=1e-6)
    for k in (0, 1, 7, 32):
        a = orbit_camera(center, 640, 480, k)
        u = npo.camera_uniform(npo.orbit_camera(center, 640, 480, k))
        assert np.allclose(a.eye_pos(), u["origin"], atol=1e-5) and np.allclose(a.eye_dir(), u["view_direction"], atol=1e-5)


def test_default_camera_is_the_reference_start_up_camera(scene):
    """eye = centre + (0,0,2), looking down -z, up +y; view_matrix is identity."""
    _, _, center = scene
    cam = ArcballCamera(center, 1.0, (800, 600))
    cam.zoom(-1.0, 1.0)
    u = cam.uniform()
    assert np.allclose(u.origin[:3], center + np.array([0, 0, 2], np.float32), atol=1e-6)
    assert list(u.view_direction[:3]) == [0.0, 0.0, -1.0] and list(u.up[:3]) == [0.0, 1.0, 0.0]
    assert np.array_equal(np.array(u.view_matrix[:]).reshape(4, 4), np.eye(4))


def test_obj_loader_follows_the_drop_path(product_lib):
    """src/main.rs:139-157 (tobj triangulate, first model) -> Application::load_model (normalise, centre)."""
    from linon_b200.main_view import LinonError, load_obj
    text = ("# a quad, a relative-index triangle, then a second object that must be ignored\n"
            "v 0 0 0\nv 2 0 0\nv 2 2 0\nv 0 2 0\nv 9 9 9\nvt 0 0\nvn 0 0 1\n"
            "o first\nf 1/1/1 2/1/1 3/1/1 4/1/1\nf -4//1 -3//1 -2//1\no second\nf 1 2 5\n")
    verts, faces, center = load_obj(text)
    assert faces.tolist() == [[0, 1, 2], [0, 2, 3], [1, 2, 3]]          # fan (0, i, i+1); -4 counts back from the 5th vertex
    assert verts.shape == (4, 3)                                        # vertex 5 belongs to the second model only
    assert verts.min() == -1.0 and verts.max() == 1.0                   # one global min / max (src/vertices.rs:3-17)
    assert np.allclose(center, [0.0, 0.0, -1.0])                        # bounding-box centre (src/vertices.rs:19-57)
    with pytest.raises(LinonError):
        load_obj("v 0 0 0\nf 1 2 3\n")                                  # index past the vertex list
    with pytest.raises(LinonError):
        load_obj("v 0 0 0\nv 1 0 0\n")                                  # no faces


def test_python_constants_equal_the_headers_enumerators():
    """Every LINON_OPT_* / LINON_RENDER_* / LINON_BUF_* enumerator of include/linon_b200.h that the ctypes binding names has
    the header's value (the binding is written by hand)."""
    _capi = capi
    text = open(os.path.join(ROOT, "include", "linon_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    header = {m.group(1): int(m.group(2), 0) for m in re.finditer(r"\bLINON_([A-Z0-9_]+)\s*=\s*(0x[0-9a-fA-F]+|\d+)", text)}
    checked = 0
    for name, value in vars(_capi).items():
        if name.startswith(("OPT_", "RENDER_", "BUF_")) and isinstance(value, int):
            assert name in header, f"{name} is not an enumerator of the header"
            assert header[name] == value, (name, header[name], value)
            checked += 1
    assert checked >= 25
    assert {k for k in header if k.startswith("OPT_")} <= {k for k in vars(_capi)}, "an option of the header has no binding"


def test_frame_sizes_that_would_overflow_the_kernels_indices_are_refused(product_lib):
    """Pixel indices are 32-bit in the kernels and the outline keys hold 21 bits per coordinate (round-1 advisor finding):
    the size check comes before anything touches a device, so it is visible here."""
    import ctypes as C
    from linon_b200 import _capi as capi
    L = capi.lib()
    for w, h in ((1, 100), (100, 1), (65537, 16), (16, 70000), (65536, 32768), (50000, 50000)):
        out = C.c_void_p(0)
        assert L.linon_create(0, w, h, C.byref(out)) == capi.ERR_INVALID, (w, h)
        assert b"width * height < 2^31" in L.linon_last_error(None)


def test_nvrtc_is_the_toolkits_copy_whatever_else_the_process_loaded(product_lib, tmp_path):
    """torch ships its own libnvrtc.so.12 (12.8 here) and loads it on import; a library linked against the soname would compile
    its kernels with whichever copy came first (round 2: bench.py, which imports torch, timed a different cubin than the tests
    and the ncu captures).  The library opens the toolkit's NVRTC by absolute path: same file, same version and the same machine
    code for the headline variant with and without torch in the process."""
    import subprocess
    import sys
    code = (
        "import os, sys\n"
        "if sys.argv[1] == '1':\n"
        "    import torch\n"
        "from linon_b200 import _capi as capi\n"
        "from linon_b200.main_view import PredefinedFunction, compile_check, nvrtc_info\n"
        "compile_check(PredefinedFunction.MirageSphericalSigmoid.to_code(), flags=capi.RENDER_ENHANCE, math_mode=capi.MATH_FAST)\n"
        "path, ver = nvrtc_info()\n"
        "print(path, ver[0], ver[1], open(os.environ['LINON_DUMP_CUBIN_HASH']).read().strip())\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    seen = []
    for with_torch in ("0", "1"):
        env = dict(os.environ, LINON_DUMP_CUBIN_HASH=str(tmp_path / f"h{with_torch}.txt"), LINON_CHECK_QUIET_ZONE="2", PYTHONPATH=root)
        env.pop("LINON_NVRTC_LIB", None)
        r = subprocess.run([sys.executable, "-c", code, with_torch], capture_output=True, text=True, env=env, cwd=root, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        seen.append(r.stdout.strip().splitlines()[-1].split())
    assert seen[0] == seen[1]
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    assert seen[0][0] == os.path.realpath(os.path.join(cuda_home, "lib64", "libnvrtc.so.12"))
    assert "site-packages" not in seen[0][0]


def test_a_production_kernel_does_not_depend_on_what_was_compiled_before_it(product_lib, tmp_path):
    """NVRTC 12.9 keeps --prec-sqrt / --prec-div / --ftz of an earlier compilation of the same process when a later one says
    --use_fast_math: after a strict or fused compile the folds `1 / sqrtf(x)` -> rsqrt and `a / sqrtf(x)` -> a * rsqrt no longer
    happened, and a long session (the GPU test suite, an editing session that toggles math modes) ran production kernels with
    other machine code and other last bits than a fresh process - the culling-identity test failed only late in the full suite.
    The fast build therefore names rsqrt itself and passes every implied option explicitly; the tile kernel's code hash of the
    headline variant, of an adaptive one and of a generic translation must not depend on the compile order."""
    import subprocess
    import sys
    code = (
        "import os, sys\n"
        "from linon_b200 import _capi as capi\n"
        "from linon_b200.main_view import PredefinedFunction as P, compile_check\n"
        "def h(fn, flags, math, spec='1'):\n"
        "    os.environ.update(LINON_CHECK_QUIET_ZONE='2', LINON_CHECK_SPECIALIZE=spec)\n"
        "    compile_check(fn.to_code(), flags=flags, math_mode=math)\n"
        "    return open(os.environ['LINON_DUMP_CUBIN_HASH']).read().strip()\n"
        "V = [(P.MirageSphericalSigmoid, capi.RENDER_ENHANCE, '1'), (P.MirageSpherical, 0, '1'), (P.MiragePlaneSigmoid, capi.RENDER_ADAPTIVE_OFF, '0')]\n"
        "first = [h(fn, fl, capi.MATH_FAST, sp) for fn, fl, sp in V]\n"
        "h(P.Rotation, 0, capi.MATH_FUSED); h(P.MirageSpherical, capi.RENDER_ENHANCE, capi.MATH_STRICT)\n"
        "again = [h(fn, fl, capi.MATH_FAST, sp) for fn, fl, sp in V]\n"
        "print(' '.join(first)); print(' '.join(again))\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, LINON_DUMP_CUBIN_HASH=str(tmp_path / "h.txt"), PYTHONPATH=root)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=root, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    first, again = r.stdout.strip().splitlines()[-2:]
    assert first == again and len(set(first.split())) == 3
