"""Static properties of the machine code the library ships, checked without a GPU (nvcc / NVRTC cross-compile here; NVRTC's output
is host-independent, tests/test_bench_cpu.py): the headline march variant fits its register budget without local memory, and the
default overlay kernel really is the TMA tile kernel (UTMALDG.2D + mbarrier SYNCS in the library's sm_100a SASS, the mnemonics
/opt/skills/guides/B200_PROFILING.md names)."""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
pytestmark = pytest.mark.skipif(not os.path.exists(CUOBJDUMP), reason="cuobjdump not found")


def _run(*args):
    r = subprocess.run(args, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-1000:]
    return r.stdout


@pytest.fixture(scope="module")
def headline_cubin(tmp_path_factory):
    path = str(tmp_path_factory.mktemp("sass") / "headline.cubin")
    code = ("from linon_b200 import _capi as capi\n"
            "from linon_b200.main_view import PredefinedFunction as P, compile_check\n"
            "compile_check(P.MirageSphericalSigmoid.to_code(), flags=capi.RENDER_ENHANCE, math_mode=capi.MATH_FAST)\n")
    env = dict(os.environ, LINON_DUMP_CUBIN=path, LINON_CHECK_QUIET_ZONE="2", PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return path


def test_headline_march_kernel_resources(headline_cubin):
    usage = _run(CUOBJDUMP, "-res-usage", headline_cubin)
    m = re.search(r"Function linon_march_tiles:\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", usage)
    assert m, usage[-500:]
    regs, stack, shared, local = map(int, m.groups())
    assert regs <= 64        # 4 CTAs of 256 threads per SM
    assert local == 0 and stack <= 64   # no spilled state in the step loop
    assert shared <= 8 * 1024  # the triangle / plane-group tables of the start-up mesh
    assert "sm_100a" in _run(CUOBJDUMP, "-lelf", headline_cubin) or "sm_100" in usage


def test_headline_march_kernel_is_fast_math_fp32(headline_cubin):
    """Approximate reciprocal square roots and exponentials (MUFU.RSQ / MUFU.EX2), flush-to-zero FFMA, tables read from shared
    memory; no FP64 and no integer division helpers in the step loop."""
    sass = _run(CUOBJDUMP, "-sass", "-fun", "linon_march_tiles", headline_cubin)
    assert sass.count("MUFU.RSQ") >= 4 and "MUFU.EX2" in sass
    assert sass.count("FFMA") > 500 and "LDS" in sass
    assert "DFMA" not in sass and "DMUL" not in sass and "DADD" not in sass


def test_the_shipped_overlay_is_the_tma_tile_kernel():
    lib = os.path.join(ROOT, "linon_b200", "lib", "liblinon_b200.so")
    assert os.path.exists(lib), "python -m linon_b200.build first"
    sass = _run(CUOBJDUMP, "-sass", lib)
    assert "sm_100a" in sass
    assert "UTMALDG.2D" in sass                       # cp.async.bulk.tensor.2d: tile + halo by one TMA load per CTA
    assert re.search(r"SYNCS\.(EXCH|ARRIVE|PHASECHK)", sass)  # the mbarrier the load completes on
    tile = [ln for ln in sass.splitlines() if "Function :" in ln and "overlay_tile_kernel" in ln]
    assert len(tile) >= 10  # delta 1..10 for the 32-row tile (+ the tall tile for delta <= 4)
