"""Do the default kernels still compile to the same SASS as at an earlier commit?  (developer / evidence tool, no GPU needed)

    python tools/sass_identity.py <git-rev>

Builds the library of <git-rev> in a temporary git worktree, JIT-compiles a handful of variants of the march module with
both trees (linon_compile_check -> LINON_DUMP_CUBIN -> cuobjdump -sass, encodings stripped) and compares the listings of
all nine kernels.  Used when a change is made without GPU time: byte-identical SASS means the measured path is untouched.
"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# (predefined function, render flags, math mode): the C3 / bench kernel, adaptive fast, fixed-step strict, adaptive fused, plane fast
VARIANTS = [(3, 1, 2), (3, 0, 2), (1, 4, 0), (7, 0, 1), (2, 4, 2)]

DUMP = r'''
import os, re, subprocess, sys
root, out = sys.argv[1], sys.argv[2]
sys.path.insert(0, root)
from linon_b200 import build
build.build()
from linon_b200.main_view import PredefinedFunction, compile_check
for i, (fn, flags, math) in enumerate(%r):
    os.environ["LINON_DUMP_CUBIN"] = f"{out}_{i}.cubin"
    compile_check(PredefinedFunction(fn).to_code(), flags=flags, math_mode=math)
    s = subprocess.run(["cuobjdump", "-sass", f"{out}_{i}.cubin"], capture_output=True, text=True, check=True).stdout
    s = "\n".join(re.sub(r"/\* 0x[0-9a-f]* \*/", "", l) for l in s.split("\n") if not re.match(r"^\s*/\* 0x", l))
    open(f"{out}_{i}.txt", "w").write(s)
''' % (VARIANTS,)


def main():
    rev = sys.argv[1]
    with tempfile.TemporaryDirectory() as tmp:
        old = os.path.join(tmp, "old")
        subprocess.run(["git", "-C", ROOT, "worktree", "add", "-f", old, rev, "-q"], check=True)
        try:
            script = os.path.join(tmp, "dump.py")
            open(script, "w").write(DUMP)
            for tree, tag in ((old, "old"), (ROOT, "new")):
                subprocess.run([sys.executable, script, tree, os.path.join(tmp, tag)], check=True, cwd=tree,
                               stdout=subprocess.DEVNULL)
            same = True
            for i, v in enumerate(VARIANTS):
                a = open(os.path.join(tmp, f"old_{i}.txt")).read()
                b = open(os.path.join(tmp, f"new_{i}.txt")).read()
                print(f"function {v[0]} flags {v[1]} math {v[2]}: {'byte-identical SASS' if a == b else 'DIFFERENT'}")
                same &= a == b
            return 0 if same else 1
        finally:
            subprocess.run(["git", "-C", ROOT, "worktree", "remove", "--force", old])


if __name__ == "__main__":
    sys.exit(main())
