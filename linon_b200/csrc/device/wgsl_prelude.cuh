// wgsl_prelude.cuh — the WGSL value types, builtins and linon helper library that a translated
// field-function body may use.  Compiled by NVRTC for sm_100a as the first part of the march
// module; also compiles as plain host C++ (tests build translated bodies for the CPU with it).
//
// Builtins are exposed as w_<name>; the translator (host/wgsl_translate.cpp) rewrites calls.
// Semantics follow the WGSL spec where the reference relies on them:
//   clamp(e,lo,hi) = min(max(e,lo),hi)    (src/main_view.wgsl:122 and src/functions.rs:45-46 pass the
//                                          arguments in the wrong order; this form keeps their meaning)
//   normalize(v) = v / length(v), length(v) = sqrt(dot(v,v)), reflect(a,n) = a - 2*dot(n,a)*n
// Helper library restates src/main_view.wgsl:76-146 (PI, rotateX/Y/Z incl. the zero third column of
// rotateZ, translate, refraction_index, refraction, point_plane_distance, sigmoid).
//
// Arithmetic mode is chosen by the NVRTC flags of the module, not here:
//   strict  --fmad=false --prec-div=true --prec-sqrt=true --ftz=false   (bit-comparable with the CPU oracle)
//   fused   --fmad=true                                                  (FMA contraction only)
//   fast    --use_fast_math                                              (MUFU rcp/sqrt/ex2/sin/cos, ftz)
#pragma once

#if defined(__CUDACC__) || defined(__CUDACC_RTC__)
#define LINON_HD __host__ __device__ __forceinline__
#define LINON_DEVICE_BUILD 1
#else
#include <cmath>
#define LINON_HD inline
#define LINON_DEVICE_BUILD 0
#endif

struct vec2f;
struct vec3f;
struct vec4f;

// ---- swizzle declarations -------------------------------------------------------------------------
#define LINON_SWZ2_DECL(a, b) LINON_HD vec2f a##b() const;
#define LINON_SWZ3_DECL(a, b, c) LINON_HD vec3f a##b##c() const;
#define LINON_FOR2_OF2(M) M(x, x) M(x, y) M(y, x) M(y, y)
#define LINON_FOR2_OF3(M) LINON_FOR2_OF2(M) M(x, z) M(y, z) M(z, x) M(z, y) M(z, z)
#define LINON_FOR2_OF4(M) LINON_FOR2_OF3(M) M(x, w) M(y, w) M(z, w) M(w, x) M(w, y) M(w, z) M(w, w)
#define LINON_FOR3_ROW3(M, a, b) M(a, b, x) M(a, b, y) M(a, b, z)
#define LINON_FOR3_OF3(M)                                                                                   \
    LINON_FOR3_ROW3(M, x, x) LINON_FOR3_ROW3(M, x, y) LINON_FOR3_ROW3(M, x, z) LINON_FOR3_ROW3(M, y, x)     \
    LINON_FOR3_ROW3(M, y, y) LINON_FOR3_ROW3(M, y, z) LINON_FOR3_ROW3(M, z, x) LINON_FOR3_ROW3(M, z, y)     \
    LINON_FOR3_ROW3(M, z, z)

struct vec2f {
    float x, y;
    LINON_HD vec2f() : x(0.0f), y(0.0f) {}
    LINON_HD explicit vec2f(float s) : x(s), y(s) {}
    LINON_HD vec2f(float a, float b) : x(a), y(b) {}
    LINON_HD float& operator[](int i) { return (&x)[i]; }
    LINON_HD float operator[](int i) const { return (&x)[i]; }
    LINON_FOR2_OF2(LINON_SWZ2_DECL)
};

struct vec3f {
    float x, y, z;
    LINON_HD vec3f() : x(0.0f), y(0.0f), z(0.0f) {}
    LINON_HD explicit vec3f(float s) : x(s), y(s), z(s) {}
    LINON_HD vec3f(float a, float b, float c) : x(a), y(b), z(c) {}
    LINON_HD vec3f(vec2f a, float c) : x(a.x), y(a.y), z(c) {}
    LINON_HD vec3f(float a, vec2f b) : x(a), y(b.x), z(b.y) {}
    LINON_HD float& operator[](int i) { return (&x)[i]; }
    LINON_HD float operator[](int i) const { return (&x)[i]; }
    LINON_FOR2_OF3(LINON_SWZ2_DECL)
    LINON_FOR3_OF3(LINON_SWZ3_DECL)
};

struct vec4f {
    float x, y, z, w;
    LINON_HD vec4f() : x(0.0f), y(0.0f), z(0.0f), w(0.0f) {}
    LINON_HD explicit vec4f(float s) : x(s), y(s), z(s), w(s) {}
    LINON_HD vec4f(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    LINON_HD vec4f(vec3f a, float d) : x(a.x), y(a.y), z(a.z), w(d) {}
    LINON_HD vec4f(float a, vec3f b) : x(a), y(b.x), z(b.y), w(b.z) {}
    LINON_HD vec4f(vec2f a, vec2f b) : x(a.x), y(a.y), z(b.x), w(b.y) {}
    LINON_HD vec4f(vec2f a, float c, float d) : x(a.x), y(a.y), z(c), w(d) {}
    LINON_HD float& operator[](int i) { return (&x)[i]; }
    LINON_HD float operator[](int i) const { return (&x)[i]; }
    LINON_FOR2_OF4(LINON_SWZ2_DECL)
    LINON_FOR3_OF3(LINON_SWZ3_DECL)
    LINON_HD vec3f yzw() const;
    LINON_HD vec4f xyzw() const { return *this; }
};

#define LINON_SWZ2_DEF_V2(a, b) LINON_HD vec2f vec2f::a##b() const { return vec2f(a, b); }
#define LINON_SWZ2_DEF_V3(a, b) LINON_HD vec2f vec3f::a##b() const { return vec2f(a, b); }
#define LINON_SWZ2_DEF_V4(a, b) LINON_HD vec2f vec4f::a##b() const { return vec2f(a, b); }
#define LINON_SWZ3_DEF_V3(a, b, c) LINON_HD vec3f vec3f::a##b##c() const { return vec3f(a, b, c); }
#define LINON_SWZ3_DEF_V4(a, b, c) LINON_HD vec3f vec4f::a##b##c() const { return vec3f(a, b, c); }
LINON_FOR2_OF2(LINON_SWZ2_DEF_V2)
LINON_FOR2_OF3(LINON_SWZ2_DEF_V3)
LINON_FOR2_OF4(LINON_SWZ2_DEF_V4)
LINON_FOR3_OF3(LINON_SWZ3_DEF_V3)
LINON_FOR3_OF3(LINON_SWZ3_DEF_V4)
LINON_HD vec3f vec4f::yzw() const { return vec3f(y, z, w); }

// ---- arithmetic operators (component-wise, scalar broadcast) ----------------------------------------
#define LINON_VEC_OPS(V, BODY_VV, BODY_VS, BODY_SV, OP)                 \
    LINON_HD V operator OP(V a, V b) { return BODY_VV; }                \
    LINON_HD V operator OP(V a, float s) { return BODY_VS; }            \
    LINON_HD V operator OP(float s, V b) { return BODY_SV; }            \
    LINON_HD V& operator OP##=(V& a, V b) { a = a OP b; return a; }     \
    LINON_HD V& operator OP##=(V& a, float s) { a = a OP s; return a; }

#define LINON_OPS_FOR(OP)                                                                                          \
    LINON_VEC_OPS(vec2f, vec2f(a.x OP b.x, a.y OP b.y), vec2f(a.x OP s, a.y OP s), vec2f(s OP b.x, s OP b.y), OP)  \
    LINON_VEC_OPS(vec3f, vec3f(a.x OP b.x, a.y OP b.y, a.z OP b.z), vec3f(a.x OP s, a.y OP s, a.z OP s),           \
                  vec3f(s OP b.x, s OP b.y, s OP b.z), OP)                                                         \
    LINON_VEC_OPS(vec4f, vec4f(a.x OP b.x, a.y OP b.y, a.z OP b.z, a.w OP b.w),                                    \
                  vec4f(a.x OP s, a.y OP s, a.z OP s, a.w OP s), vec4f(s OP b.x, s OP b.y, s OP b.z, s OP b.w), OP)
LINON_OPS_FOR(+)
LINON_OPS_FOR(-)
LINON_OPS_FOR(*)
LINON_OPS_FOR(/)
LINON_HD vec2f operator-(vec2f a) { return vec2f(-a.x, -a.y); }
LINON_HD vec3f operator-(vec3f a) { return vec3f(-a.x, -a.y, -a.z); }
LINON_HD vec4f operator-(vec4f a) { return vec4f(-a.x, -a.y, -a.z, -a.w); }

// ---- scalar builtins ------------------------------------------------------------------------------------
LINON_HD float w_abs(float x) { return fabsf(x); }
LINON_HD float w_min(float a, float b) { return fminf(a, b); }
LINON_HD float w_max(float a, float b) { return fmaxf(a, b); }
LINON_HD float w_clamp(float e, float lo, float hi) { return fminf(fmaxf(e, lo), hi); }
LINON_HD float w_saturate(float e) { return fminf(fmaxf(e, 0.0f), 1.0f); }
LINON_HD float w_sqrt(float x) { return sqrtf(x); }
LINON_HD float w_inverseSqrt(float x) { return 1.0f / sqrtf(x); }
LINON_HD float w_exp(float x) { return expf(x); }
LINON_HD float w_exp2(float x) { return exp2f(x); }
LINON_HD float w_log(float x) { return logf(x); }
LINON_HD float w_log2(float x) { return log2f(x); }
LINON_HD float w_pow(float x, float y) { return powf(x, y); }
LINON_HD float w_sin(float x) { return sinf(x); }
LINON_HD float w_cos(float x) { return cosf(x); }
LINON_HD float w_tan(float x) { return tanf(x); }
LINON_HD float w_asin(float x) { return asinf(x); }
LINON_HD float w_acos(float x) { return acosf(x); }
LINON_HD float w_atan(float x) { return atanf(x); }
LINON_HD float w_atan2(float y, float x) { return atan2f(y, x); }
LINON_HD float w_sinh(float x) { return sinhf(x); }
LINON_HD float w_cosh(float x) { return coshf(x); }
LINON_HD float w_tanh(float x) { return tanhf(x); }
LINON_HD float w_floor(float x) { return floorf(x); }
LINON_HD float w_ceil(float x) { return ceilf(x); }
LINON_HD float w_round(float x) { return rintf(x); }  // WGSL round: ties to even
LINON_HD float w_trunc(float x) { return truncf(x); }
LINON_HD float w_fract(float x) { return x - floorf(x); }
LINON_HD float w_sign(float x) { return (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f); }
LINON_HD float w_step(float edge, float x) { return (edge <= x) ? 1.0f : 0.0f; }
LINON_HD float w_mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
LINON_HD float w_smoothstep(float lo, float hi, float x) {
    const float t = w_clamp((x - lo) / (hi - lo), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
LINON_HD float w_fma(float a, float b, float c) { return a * b + c; }
LINON_HD float w_degrees(float x) { return x * 57.295779513082322865f; }
LINON_HD float w_radians(float x) { return x * 0.017453292519943295474f; }
LINON_HD float w_select(float f, float t, bool c) { return c ? t : f; }
LINON_HD int w_select(int f, int t, bool c) { return c ? t : f; }
LINON_HD int w_abs(int x) { return x < 0 ? -x : x; }
LINON_HD int w_min(int a, int b) { return a < b ? a : b; }
LINON_HD int w_max(int a, int b) { return a < b ? b : a; }
LINON_HD int w_clamp(int e, int lo, int hi) { return w_min(w_max(e, lo), hi); }

// ---- component-wise lifts -------------------------------------------------------------------------------
#define LINON_LIFT1(fn)                                                                  \
    LINON_HD vec2f fn(vec2f a) { return vec2f(fn(a.x), fn(a.y)); }                       \
    LINON_HD vec3f fn(vec3f a) { return vec3f(fn(a.x), fn(a.y), fn(a.z)); }              \
    LINON_HD vec4f fn(vec4f a) { return vec4f(fn(a.x), fn(a.y), fn(a.z), fn(a.w)); }
#define LINON_LIFT2(fn)                                                                                  \
    LINON_HD vec2f fn(vec2f a, vec2f b) { return vec2f(fn(a.x, b.x), fn(a.y, b.y)); }                    \
    LINON_HD vec3f fn(vec3f a, vec3f b) { return vec3f(fn(a.x, b.x), fn(a.y, b.y), fn(a.z, b.z)); }      \
    LINON_HD vec4f fn(vec4f a, vec4f b) { return vec4f(fn(a.x, b.x), fn(a.y, b.y), fn(a.z, b.z), fn(a.w, b.w)); }
#define LINON_LIFT3(fn)                                                                                              \
    LINON_HD vec2f fn(vec2f a, vec2f b, vec2f c) { return vec2f(fn(a.x, b.x, c.x), fn(a.y, b.y, c.y)); }             \
    LINON_HD vec3f fn(vec3f a, vec3f b, vec3f c) {                                                                   \
        return vec3f(fn(a.x, b.x, c.x), fn(a.y, b.y, c.y), fn(a.z, b.z, c.z));                                       \
    }                                                                                                                \
    LINON_HD vec4f fn(vec4f a, vec4f b, vec4f c) {                                                                   \
        return vec4f(fn(a.x, b.x, c.x), fn(a.y, b.y, c.y), fn(a.z, b.z, c.z), fn(a.w, b.w, c.w));                    \
    }
LINON_LIFT1(w_abs) LINON_LIFT1(w_saturate) LINON_LIFT1(w_sqrt) LINON_LIFT1(w_inverseSqrt) LINON_LIFT1(w_exp)
LINON_LIFT1(w_exp2) LINON_LIFT1(w_log) LINON_LIFT1(w_log2) LINON_LIFT1(w_sin) LINON_LIFT1(w_cos) LINON_LIFT1(w_tan)
LINON_LIFT1(w_asin) LINON_LIFT1(w_acos) LINON_LIFT1(w_atan) LINON_LIFT1(w_sinh) LINON_LIFT1(w_cosh) LINON_LIFT1(w_tanh)
LINON_LIFT1(w_floor) LINON_LIFT1(w_ceil) LINON_LIFT1(w_round) LINON_LIFT1(w_trunc) LINON_LIFT1(w_fract)
LINON_LIFT1(w_sign) LINON_LIFT1(w_degrees) LINON_LIFT1(w_radians)
LINON_LIFT2(w_min) LINON_LIFT2(w_max) LINON_LIFT2(w_pow) LINON_LIFT2(w_atan2) LINON_LIFT2(w_step)
LINON_LIFT3(w_clamp) LINON_LIFT3(w_mix) LINON_LIFT3(w_smoothstep) LINON_LIFT3(w_fma)
LINON_HD vec2f w_mix(vec2f a, vec2f b, float t) { return a * (1.0f - t) + b * t; }
LINON_HD vec3f w_mix(vec3f a, vec3f b, float t) { return a * (1.0f - t) + b * t; }
LINON_HD vec4f w_mix(vec4f a, vec4f b, float t) { return a * (1.0f - t) + b * t; }
LINON_HD vec2f w_select(vec2f f, vec2f t, bool c) { return c ? t : f; }
LINON_HD vec3f w_select(vec3f f, vec3f t, bool c) { return c ? t : f; }
LINON_HD vec4f w_select(vec4f f, vec4f t, bool c) { return c ? t : f; }

// ---- geometric builtins ------------------------------------------------------------------------------------
LINON_HD float w_dot(vec2f a, vec2f b) { return a.x * b.x + a.y * b.y; }
LINON_HD float w_dot(vec3f a, vec3f b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
LINON_HD float w_dot(vec4f a, vec4f b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
LINON_HD vec3f w_cross(vec3f a, vec3f b) {
    return vec3f(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
LINON_HD float w_length(float a) { return fabsf(a); }
LINON_HD float w_length(vec2f a) { return sqrtf(w_dot(a, a)); }
LINON_HD float w_length(vec3f a) { return sqrtf(w_dot(a, a)); }
LINON_HD float w_length(vec4f a) { return sqrtf(w_dot(a, a)); }
LINON_HD float w_distance(vec2f a, vec2f b) { return w_length(a - b); }
LINON_HD float w_distance(vec3f a, vec3f b) { return w_length(a - b); }
LINON_HD float w_distance(vec4f a, vec4f b) { return w_length(a - b); }
LINON_HD vec2f w_normalize(vec2f a) { return a / w_length(a); }
LINON_HD vec3f w_normalize(vec3f a) { return a / w_length(a); }
LINON_HD vec4f w_normalize(vec4f a) { return a / w_length(a); }
LINON_HD vec2f w_reflect(vec2f e1, vec2f e2) { return e1 - (2.0f * w_dot(e2, e1)) * e2; }
LINON_HD vec3f w_reflect(vec3f e1, vec3f e2) { return e1 - (2.0f * w_dot(e2, e1)) * e2; }
LINON_HD vec3f w_faceForward(vec3f e1, vec3f e2, vec3f e3) { return (w_dot(e2, e3) < 0.0f) ? e1 : -e1; }

// ---- small matrices (column major, WGSL semantics) ------------------------------------------------------------
struct mat2x2f {
    vec2f c0, c1;
    LINON_HD mat2x2f() {}
    LINON_HD mat2x2f(vec2f a, vec2f b) : c0(a), c1(b) {}
    LINON_HD mat2x2f(float a, float b, float c, float d) : c0(a, b), c1(c, d) {}
    LINON_HD vec2f& operator[](int i) { return (&c0)[i]; }
    LINON_HD vec2f operator[](int i) const { return (&c0)[i]; }
};
struct mat3x3f {
    vec3f c0, c1, c2;
    LINON_HD mat3x3f() {}
    LINON_HD mat3x3f(vec3f a, vec3f b, vec3f c) : c0(a), c1(b), c2(c) {}
    LINON_HD mat3x3f(float a, float b, float c, float d, float e, float f, float g, float h, float i)
        : c0(a, b, c), c1(d, e, f), c2(g, h, i) {}
    LINON_HD vec3f& operator[](int i) { return (&c0)[i]; }
    LINON_HD vec3f operator[](int i) const { return (&c0)[i]; }
};
LINON_HD vec2f operator*(mat2x2f m, vec2f v) { return m.c0 * v.x + m.c1 * v.y; }
LINON_HD vec2f operator*(vec2f v, mat2x2f m) { return vec2f(w_dot(v, m.c0), w_dot(v, m.c1)); }
LINON_HD mat2x2f operator*(mat2x2f a, mat2x2f b) { return mat2x2f(a * b.c0, a * b.c1); }
LINON_HD vec3f operator*(mat3x3f m, vec3f v) { return m.c0 * v.x + m.c1 * v.y + m.c2 * v.z; }
LINON_HD vec3f operator*(vec3f v, mat3x3f m) { return vec3f(w_dot(v, m.c0), w_dot(v, m.c1), w_dot(v, m.c2)); }
LINON_HD mat3x3f operator*(mat3x3f a, mat3x3f b) { return mat3x3f(a * b.c0, a * b.c1, a * b.c2); }
LINON_HD mat3x3f operator*(mat3x3f a, float s) { return mat3x3f(a.c0 * s, a.c1 * s, a.c2 * s); }
LINON_HD mat3x3f operator*(float s, mat3x3f a) { return mat3x3f(a.c0 * s, a.c1 * s, a.c2 * s); }
LINON_HD mat2x2f w_transpose(mat2x2f m) { return mat2x2f(m.c0.x, m.c1.x, m.c0.y, m.c1.y); }
LINON_HD mat3x3f w_transpose(mat3x3f m) {
    return mat3x3f(m.c0.x, m.c1.x, m.c2.x, m.c0.y, m.c1.y, m.c2.y, m.c0.z, m.c1.z, m.c2.z);
}
LINON_HD float w_determinant(mat2x2f m) { return m.c0.x * m.c1.y - m.c0.y * m.c1.x; }
LINON_HD float w_determinant(mat3x3f m) { return w_dot(m.c0, w_cross(m.c1, m.c2)); }

// ---- module-scope names a body may reference (src/main_view.wgsl:7-25, :76) ------------------------------------
struct WgslSettings {  // byte-identical with Settings, src/main_view.rs:43-51
    float field_weight;
    float mouse_pos_x;
    float mouse_pos_y;
    int overlay_mode;
    int central_difference_delta;
    float lyapunov_scaling;
};
struct WgslCamera {  // src/main_view.wgsl:7-12 (view_matrix is identity and unused on the path)
    vec4f origin;
    vec4f view_direction;
    vec4f up;
};
constexpr float PI = 3.141592653589793f;  // const PI: f32, src/main_view.wgsl:76

// ---- helper library, src/main_view.wgsl:78-146 ---------------------------------------------------------------
LINON_HD vec3f rotateX(vec3f v, float phi) {
    return mat3x3f(vec3f(1.0f, 0.0f, 0.0f), vec3f(0.0f, w_cos(phi), w_sin(phi)), vec3f(0.0f, -w_sin(phi), w_cos(phi))) * v;
}
LINON_HD vec3f rotateY(vec3f v, float phi) {
    return mat3x3f(vec3f(w_cos(phi), 0.0f, -w_sin(phi)), vec3f(0.0f, 1.0f, 0.0f), vec3f(w_sin(phi), 0.0f, w_cos(phi))) * v;
}
LINON_HD vec3f rotateZ(vec3f v, float phi) {  // third column is zero in the reference: z is dropped
    return mat3x3f(vec3f(w_cos(phi), w_sin(phi), 0.0f), vec3f(-w_sin(phi), w_cos(phi), 0.0f), vec3f(0.0f, 0.0f, 0.0f)) * v;
}
LINON_HD vec3f translate(vec3f v, float dx, float dy, float dz) { return vec3f(v.x + dx, v.y + dy, v.z + dz); }

// pow(10.0, -10.0) is folded by naga in abstract float and concretised to f32 (1e-10f)
LINON_HD float refraction_index(float t) {
    const float air_pressure = 101325.0f;
    const float c1 = 0.0000104f;
    const float c2 = 0.00366f;
    return c1 * air_pressure * (1.0f + air_pressure * (60.1f - 0.972f * t) * 1e-10f) / (1.0f + c2 * t);
}
#ifndef LINON_FAST_MATH
#define LINON_FAST_MATH 0
#endif
// eta = n(t_in) / n(t_out).  Fast mode folds the two refraction indices into one quotient: n(t) = A(t) / B(t) with
// A(t) = c1 p (1 + p 1e-10 (60.1 - 0.972 t)) = a0 - a1 t and B(t) = 1 + c2 t, so eta = A(t_in) B(t_out) / (A(t_out) B(t_in))
// costs one reciprocal instead of three (same fp32 error class: every factor is 1 + O(1e-3)).
LINON_HD float refraction_eta(float t_in, float t_out) {
#if LINON_FAST_MATH
    const float c1p = 0.0000104f * 101325.0f;
    const float a0 = c1p * (1.0f + 101325.0f * 1e-10f * 60.1f);
    const float a1 = c1p * 101325.0f * 1e-10f * 0.972f;
    const float c2 = 0.00366f;
    const float num = (a0 - a1 * t_in) * (1.0f + c2 * t_out);
    const float den = (a0 - a1 * t_out) * (1.0f + c2 * t_in);
    return num / den;
#else
    return refraction_index(t_in) / refraction_index(t_out);
#endif
}
// Total reflection is rare; keeping it out of line keeps its instructions out of the hot loop (the compiler otherwise
// hoists them ahead of the branch as predicated-off work in each of the four RK4 stages).
#if LINON_DEVICE_BUILD
__device__ __noinline__
#else
inline
#endif
vec3f linon_total_reflection(vec3f v_in, vec3f n, float d) { return v_in - (2.0f * d) * n; }

// refraction, src/main_view.wgsl:119-136.  Restated without materialising n_ref = +-n: negation is exact in IEEE
// arithmetic, so with s = -1 when cosi >= 0 (n_ref = -n) and +1 otherwise (n_ref = n)
//     eta * v_in + (eta * |cosi| - sqrt(k)) * n_ref  ==  eta * v_in + (s * (eta * |cosi| - sqrt(k))) * n      bit for bit,
//     reflect(v_in, n_ref) = v_in - 2 dot(n_ref, v_in) n_ref  ==  v_in - 2 dot(n, v_in) n                     bit for bit
// (tests/test_translator.py and tests/test_gpu_units.py compare against the oracle's literal transcription).
// Total reflection (k < 0) is rare, so it is a real branch instead of seven predicated instructions per evaluation.
LINON_HD vec3f refraction(float t_in, float t_out, vec3f v_in, vec3f n) {
    const float d = w_dot(v_in, n);
    const float cosi_raw = w_clamp(-1.0f, 1.0f, d);  // misordered in the reference: min(max(-1, 1), d) = min(1, d)
    const float cosi = w_abs(cosi_raw);              // `if (cosi < 0) cosi = -cosi`
    const float eta = refraction_eta(t_in, t_out);   // eta_in / eta_out, src/main_view.wgsl:120-121,129
    const float k = 1.0f - eta * eta * (1.0f - cosi * cosi);
    if (__builtin_expect(k < 0.0f, 0)) {
        return linon_total_reflection(v_in, n, d);  // reflect(v_in, n_ref); dot(n_ref, v_in) n_ref = dot(n, v_in) n
    }
    const float c = eta * cosi - w_sqrt(k);
    const float cs = (cosi_raw < 0.0f) ? c : -c;     // n_ref = n when cosi < 0, else -n
    return eta * v_in + cs * n;
}
LINON_HD float point_plane_distance(vec3f p, vec3f n, vec3f p0) {
    const float d = w_dot(p0, n);
    return w_abs(w_dot(p, n) - d);
}
LINON_HD float sigmoid(float x) { return 1.0f / (1.0f + w_exp(-x)); }
