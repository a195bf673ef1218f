"""Generates tests/golden/user_bodies.npz: results of the user-edited field-function bodies of tests/user_bodies.py,
computed by the REFERENCE's shader semantics - the body spliced into /root/reference/src/main_view.wgsl like
`with_field_function` (src/main_view.rs:852-861), transpiled by oracle/wgsl_ref and compiled with gcc.  The GPU test
(tests/test_gpu_units.py::test_user_edited_bodies_match_the_reference_semantics) compares the device's evaluation of
the same bodies with these arrays on the GPU box, where the reference is not mounted.  Settings: the application's
defaults; camera: zero (what linon_debug_eval_field passes).
Run (needs /root/reference): python tests/golden/make_user_bodies.py
"""
import ctypes as C
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import user_bodies  # noqa: E402
from oracle import cpu_oracle as co  # noqa: E402
from oracle import np_oracle as npo  # noqa: E402
from oracle.wgsl_ref import build_ref  # noqa: E402


def main():
    ins = user_bodies.inputs(n=1024)
    n = ins[4].shape[0]
    st = co.make_settings(npo.default_settings())
    out = {}
    for k, (name, body) in enumerate(user_bodies.BODIES.items()):
        with tempfile.TemporaryDirectory() as d:
            lib = C.CDLL(build_ref.build_custom(body, d))
            want = np.zeros((n, 3), np.float32)
            lib.wgslref_custom_field(C.byref(st), n, *[C.c_void_p(a.ctypes.data) for a in ins], C.c_void_p(want.ctypes.data))
        out[f"body{k}"] = want
    np.savez_compressed(os.path.join(HERE, "user_bodies.npz"), names=np.array(list(user_bodies.BODIES)), **out)


if __name__ == "__main__":
    main()
