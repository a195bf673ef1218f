"""Compares the machine code of a set of march-module variants between two trees (no GPU needed: NVRTC cross-compiles).

    git archive <old-commit> | tar -x -C /tmp/old && (cd /tmp/old && python -m linon_b200.build)
    python tools/dev/sass_diff.py /tmp/old /root/repo

For every variant both libraries compile the module (linon_compile_check with LINON_DUMP_CUBIN) and `cuobjdump -sass` of the two
cubins is compared with the encodings stripped - line tables differ whenever a source line moves, instructions must not.  Used in
round 2 to show that the shade_hit refactor and the new macros of the kernel template left every kernel of the module as it was."""
import hashlib
import os
import re
import subprocess
import sys
import tempfile

DUMP = r'''
import os, sys
root, out = sys.argv[1], sys.argv[2]
sys.path.insert(0, root)
from linon_b200 import _capi as capi
from linon_b200.main_view import PredefinedFunction as P, compile_check
import linon_b200
assert os.path.dirname(os.path.abspath(linon_b200.__file__)).startswith(os.path.abspath(root))
V = [("sph_adaptive_fast", P.MirageSpherical, 0, capi.MATH_FAST, {"LINON_CHECK_QUIET_ZONE": "2"}),
     ("sph_fixed_fast", P.MirageSpherical, capi.RENDER_ADAPTIVE_OFF, capi.MATH_FAST, {"LINON_CHECK_QUIET_ZONE": "2"}),
     ("sigmoid_enhance_fast", P.MirageSphericalSigmoid, capi.RENDER_ENHANCE, capi.MATH_FAST, {"LINON_CHECK_QUIET_ZONE": "2"}),
     ("plane_adaptive_fast_q0", P.MiragePlaneSigmoid, 0, capi.MATH_FAST, {"LINON_CHECK_QUIET_ZONE": "0"}),
     ("rot_adaptive_fused", P.Rotation, 0, capi.MATH_FUSED, {}),
     ("sph_adaptive_strict", P.MirageSpherical, 0, capi.MATH_STRICT, {}),
     ("sigmoid_enhance_generic", P.MirageSphericalSigmoid, capi.RENDER_ENHANCE, capi.MATH_FAST, {"LINON_CHECK_SPECIALIZE": "0"}),
     ("lorenz_fixed_fast_nocull", P.LorenzAttractor, capi.RENDER_ADAPTIVE_OFF, capi.MATH_FAST, {"LINON_CHECK_CULL": "0"})]
for name, fn, flags, math, env in V:
    for k in ("LINON_CHECK_QUIET_ZONE", "LINON_CHECK_SPECIALIZE", "LINON_CHECK_CULL"):
        os.environ.pop(k, None)
    os.environ.update(env)
    os.environ["LINON_DUMP_CUBIN"] = os.path.join(out, name + ".cubin")
    compile_check(fn.to_code(), flags=flags, math_mode=math)
'''


def sass_digest(cubin):
    text = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True, check=True).stdout
    return hashlib.md5(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", text).encode()).hexdigest()[:12]


def main(old, new):
    with tempfile.TemporaryDirectory() as d:
        outs = []
        for k, root in enumerate((old, new)):
            out = os.path.join(d, str(k))
            os.makedirs(out)
            subprocess.run([sys.executable, "-c", DUMP, root, out], check=True, cwd=d)
            outs.append(out)
        same = True
        for name in sorted(os.listdir(outs[0])):
            a, b = sass_digest(os.path.join(outs[0], name)), sass_digest(os.path.join(outs[1], name))
            same = same and a == b
            print(f"{name:36s} {a} {b} {'same' if a == b else 'DIFFERENT'}")
    return 0 if same else 1


if __name__ == "__main__":
    sys.exit(main(sys.argv[1], sys.argv[2]))
