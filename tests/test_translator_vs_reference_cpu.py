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
from test_translator import USER_EDITS, build_host_field

REFERENCE = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "src")), reason="reference mount not present")

BODIES = dict(USER_EDITS)
BODIES.update({
    "abstract-float constant evaluation": """
        let a = 0.1 * 3.0;
        let b = tan(0.1) + sqrt(2.0) * exp(1.0);
        let c = 1.0 / 3.0 + a;
        let k = 7 / 2;
        let d = p.x * 2.0 / 3.0 - 1.0 + 2.0 + -(1.5 * 2.0) * p.y - 3.0 * 0.7;
        return vec3<f32>(a * p.x + b, c * p.y - -0.3 * 0.7, f32(k) * p.z + pow(2.0, 0.5) + d + max(1.0, 0.1 * 7.0,));""",
    "folding must respect precedence and associativity": """
        let x = p.x;
        let a = x * 0.1 * 0.3 + x / 0.7 / 0.9 - 0.1 - 0.2;
        let b = 0.1 + 0.2 * x - (0.1 + 0.2) * x + 0.1 / 0.3 * x;
        let c = x - 0.1 + 0.2 + (x - 0.1) + 0.2 - -0.1 * -0.3;
        var d = 0.7;
        d += 0.1 * 0.3;
        d *= 1.0 / 0.7 + x;
        if (x < 0.1 + 0.2 && x > 0.1 - 0.2 || d == 0.1 * 3.0) { d = -d; }
        return vec3<f32>(a, b, c + d) * (2.0 / 3.0) + vec3<f32>(sin(0.5), cos(0.5) * x, log(2.0) / log2(10.0));""",
    "integers, casts and wrap-around": """
        var u: u32 = 0u;
        u = u - 1u;
        let top = f32(u >> 31u);
        var n: i32 = 7;
        n = n / 2 * 3 % 5 - 10;
        let m = i32(p.x * 10.0) + n;
        var acc = 0.0;
        for (var i = 0; i < 5; i++) {
            if (i % 2 == 0) { continue; }
            if (i == 4) { break; }
            acc = acc + f32(i) * p.y;
        }
        return vec3<f32>(top + acc, f32(m), f32(abs(n)) + f32(0x10) + f32(3u) + 2f + .5 + 1. + 1e-3);""",
    "vector builtins": """
        let a = vec3(v.xy, 1.0);
        let b = vec4<f32>(p, 2.0).wzyx.xyz;
        var c = cross(a, b) + reflect(v, normalize(b)) * dot(a, v0) + abs(-a) * sign(p) + floor(p * 4.0) + fract(b);
        c = clamp(c, vec3<f32>(-2.0), vec3<f32>(2.0)) + min(a, b) - max(a, b) + step(vec3<f32>(0.1), p);
        c += mix(a, b, vec3<f32>(0.25, 0.5, 0.75)) + mix(a, b, 0.3) * smoothstep(0.0, 1.0, t / 5.0);
        c *= select(vec3<f32>(1.0), vec3<f32>(-1.0), p.x > 0.0);
        c.y -= distance(a, b) + length(c.xz);
        c.z /= 1.0 + saturate(t);
        return c + vec3<f32>(exp2(p.x), inverseSqrt(1.0 + dot(p, p)), atan2(p.y, p.x)) + fma(t, p.x, 1.0);""",
    "matrices and helpers": """
        let m = mat3x3<f32>(vec3<f32>(1.0, 2.0, 3.0), vec3<f32>(0.5, -1.0, 0.25), v0);
        let r = m * v + transpose(m) * p + v * m;
        let q = rotateX(rotateY(rotateZ(r, t), 0.3), PI / 3.0);
        let s = translate(q, sigmoid(p.x), refraction_index(20.0 + t), point_plane_distance(p, normalize(v0), p_prev));
        return refraction(15.0 + t, 40.0, normalize(v), normalize(s)) + q * 1e-3;""",
    "shadowing and nested control flow": """
        let s = 0.5;
        var out = v;
        if (p.x > 0.0) {
            let s = 2.0 * p.y;
            if (s > 0.5) { return out * s; } else if (s < -0.5) { out = -out; } else { out.x = s; }
            var k = 0;
            while (k < 3) {
                let s = f32(k) * 0.25;
                out = out + vec3<f32>(s);
                k += 1;
            }
        } else {
            for (var i: u32 = 0u; i < 3u; i = i + 1u) { out[i] = out[i] * s + f32(i); }
        }
        return out * s;""",
    "settings, camera and the loop variable t": """
        let w = settings.field_weight * 0.5 + f32(settings.overlay_mode) + settings.lyapunov_scaling * 1e-2;
        let o = camera.origin.xyz - p;
        return normalize(o) * w + camera.up.xyz * (t * 0.1) - camera.view_direction.xyz / (1.0 + length(o));""",
})


@pytest.fixture(scope="module")
def inputs():
    rng = np.random.default_rng(11)
    n = 4096
    p = rng.uniform(-1.2, 1.2, (n, 3)).astype(np.float32)
    p_prev = (p + rng.normal(0, 0.01, (n, 3))).astype(np.float32)
    v = rng.normal(0, 1, (n, 3)).astype(np.float32)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    v0 = rng.normal(0, 1, (n, 3)).astype(np.float32)
    t = rng.uniform(0, 5.1, n).astype(np.float32)
    return p_prev, p, v0, v, t


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
