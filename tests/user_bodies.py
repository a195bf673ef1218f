"""User-edited field-function bodies shared by the CPU differential test (tests/test_translator_vs_reference_cpu.py),
the fixture generator (tests/golden/make_user_bodies.py) and the GPU test (tests/test_gpu_units.py): what a user edit is
likely to contain and the places where a token rewriter can go wrong - abstract-float constant
evaluation, operator precedence around folded literals, integer arithmetic, swizzles, compound assignment, shadowing,
control flow, every library helper."""
import numpy as np

from test_translator import USER_EDITS

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
    "loop, continuing and break if": """
        var acc = vec3<f32>(0.0);
        var i = 0;
        var w = 1.0;
        loop {
            if (i >= 6) { break; }
            if (i == 2) { continue; }
            acc = acc + v * w + p * f32(i);
            continuing {
                i = i + 1;
                w = w * 0.5;
                break if w < 0.05;
            }
        }
        var n = 0u;
        loop {
            n = n + 1u;
            if (n > 3u) { break; }
        }
        return acc + vec3<f32>(f32(n), w, f32(i));""",
    "switch with several selectors and default": """
        let k = i32(floor(p.x * 3.0)) + 3;
        var out = v;
        switch k {
            case 0, 1: { out = out * 2.0; }
            case 2 { out = -out; }
            case 3, default: { out = out + p; }
            case 5: {
                if (p.y > 0.0) { break; }
                out = vec3<f32>(p.y);
            }
        }
        switch (k % 2) {
            case 0: { out.x = out.x + 1.0; }
            default { out.y = out.y - 1.0; }
        }
        return out;""",
    "float and integer remainder": """
        let a = p.x * 7.0 % 0.75;
        let b = v % 0.3 + p % vec3<f32>(0.5, 0.25, 0.125) + 2.5 % p.yzx;
        var c = t;
        c %= 0.7;
        var m = i32(p.y * 10.0) * 3 % 5 - 7 % 3;
        m %= 4;
        let d = -p.z % 0.4 * 2.0 + (p.x + 1.0) % (0.1 + 0.2) - length(v) % 0.33;
        return b + vec3<f32>(a, c + f32(m), d + 5.5 % 2.0);""",
})

# bodies whose arithmetic is + - * / sqrt, comparisons and integer work only: bit-exact on the device in strict mode
EXACT = ["for loop, i32, swizzles", "integers, casts and wrap-around", "shadowing and nested control flow",
         "settings, camera and the loop variable t", "loop, continuing and break if", "switch with several selectors and default"]


def inputs(n=4096, seed=11):
    rng = np.random.default_rng(seed)
    p = rng.uniform(-1.2, 1.2, (n, 3)).astype(np.float32)
    p_prev = (p + rng.normal(0, 0.01, (n, 3))).astype(np.float32)
    v = rng.normal(0, 1, (n, 3)).astype(np.float32)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    v0 = rng.normal(0, 1, (n, 3)).astype(np.float32)
    t = rng.uniform(0, 5.1, n).astype(np.float32)
    return p_prev, p, v0, v, t
