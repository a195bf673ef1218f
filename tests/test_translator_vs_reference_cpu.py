"""Differential test of the product's WGSL translator against WGSL's own semantics.  CPU only.

Two independently written front ends read the same user-edited field-function body:
  * the product's token rewriter (linon_b200/csrc/host/wgsl_translate.cpp), whose output is compiled for the host against
    the product prelude (device/wgsl_prelude.cuh) exactly as NVRTC compiles it for the GPU;
  * oracle/wgsl_ref/wgsl2cpp.py, a parser-based transpiler with the WGSL specification's typing rules, applied to the
    reference's whole main_view.wgsl with the body spliced in like `with_field_function` (src/main_view.rs:852-861).
Same compiler, same libm, no contraction: every body must give bit-identical results on random inputs.  The bodies
exercise what a user edit is likely to contain and the places where a token rewriter can go wrong:
abstract-float constant evaluation, operator precedence around folded literals, integer arithmetic, swizzles,
compound assignment, shadowing, control flow, every library helper.
"""
import ctypes as C
import os
import tempfile

import numpy as np
import pytest

from oracle import np_oracle as npo
import user_bodies
from test_translator import build_host_field

REFERENCE = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "src")), reason="reference mount not present")

BODIES = user_bodies.BODIES


@pytest.fixture(scope="module")
def inputs():
    return user_bodies.inputs()


@pytest.mark.parametrize("name", list(BODIES))
def test_product_translation_equals_wgsl_semantics_bit_for_bit(product_lib, oracle, inputs, name):
    from oracle.wgsl_ref import build_ref
    body = BODIES[name]
    st = npo.default_settings()
    n = inputs[4].shape[0]
    with tempfile.TemporaryDirectory() as d:
        lib = build_host_field(body, d, "product")
        oracle.lib().oracle_set_custom_field(C.cast(lib.custom_field, C.c_void_p))
        try:
            got = oracle.field(0, st, *inputs)
        finally:
            oracle.lib().oracle_set_custom_field(None)
        ref = C.CDLL(build_ref.build_custom(body, d))
        want = np.zeros((n, 3), np.float32)
        s = oracle.make_settings(st)
        ref.wgslref_custom_field(C.byref(s), n, *[C.c_void_p(a.ctypes.data) for a in inputs], C.c_void_p(want.ctypes.data))
    assert np.isfinite(want).mean() > 0.9
    bad = (got.view(np.uint32) != want.view(np.uint32)).any(axis=1)
    assert not bad.any(), (name, int(bad.sum()), got[bad][:3], want[bad][:3])
