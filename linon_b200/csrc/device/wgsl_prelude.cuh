// wgsl_prelude.cuh — the WGSL value types, builtins and linon helper library that a translated
// field-function body may use.  Compiled by NVRTC for sm_100a as the first part of the march
// module; also compiles as plain host C++ (tests build translated bodies for the CPU with it).
//
// Everything is generic over the SCALAR type S the body is instantiated with:
//   S = float   one ray per thread (the reference's own mapping, src/main_view.wgsl:339-375)
//   S = f32x2   two rays per thread.  sm_100 has packed fp32 instructions (FFMA2 / FMUL2 / FADD2, PTX
//               fma/mul/add/sub .f32x2): one issue slot does the arithmetic of both rays.  The march kernel is
//               issue-bound, so the packed instantiation is the fast path; per-ray operations that have no packed
//               form (MUFU: rcp, sqrt, ex2, sin, cos; min/max; compare/select) are done per component.
// vec3f etc. are the S = float instantiations; the march module re-declares them per instantiation inside
// LinonField<S>, so the same translated text serves both.
//
// Builtins are exposed as w_<name>; the translator (host/wgsl_translate.cpp) rewrites calls.
// Semantics follow the WGSL spec where the reference relies on them:
//   clamp(e,lo,hi) = min(max(e,lo),hi)    (src/main_view.wgsl:122 and src/functions.rs:45-46 pass the
//                                          arguments in the wrong order; this form keeps their meaning)
//   normalize(v) = v / length(v), length(v) = sqrt(dot(v,v)), reflect(a,n) = a - 2*dot(n,a)*n
// Helper library restates src/main_view.wgsl:76-146 (PI, rotateX/Y/Z incl. the zero third column of
// rotateZ, translate, refraction_index, refraction, point_plane_distance, sigmoid).
//
// Arithmetic mode is chosen by the NVRTC flags of the module:
//   strict  --fmad=false --prec-div=true --prec-sqrt=true --ftz=false   (bit-comparable with the CPU oracle;
//           -DLINON_PACKED_OPS=0: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 whatever --fmad says,
//           so strict f32x2 arithmetic is done per component and the packed kernel stays bit-identical)
//   fused   --fmad=true                                                  (FMA contraction only)
//   fast    --use_fast_math -DLINON_FAST_MATH=1                          (MUFU rcp/sqrt/ex2/sin/cos, ftz)
#pragma once

#if defined(__CUDACC__) || defined(__CUDACC_RTC__)
#define LINON_HD __host__ __device__ __forceinline__
#define LINON_DEVICE_BUILD 1
#else
#include <cmath>
#define LINON_HD inline
#define LINON_DEVICE_BUILD 0
#endif
#ifndef LINON_FAST_MATH
#define LINON_FAST_MATH 0
#endif
#ifndef LINON_PACKED_OPS  // 1: f32x2 + - * use the packed sm_100 instructions (device, non-strict builds)
#define LINON_PACKED_OPS 0
#endif
#if LINON_PACKED_OPS && !(LINON_DEVICE_BUILD && defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000)
#undef LINON_PACKED_OPS
#define LINON_PACKED_OPS 0
#endif
#if LINON_FAST_MATH
#define LINON_F2_FTZ ".ftz"
#else
#define LINON_F2_FTZ ""
#endif

// =====================================================================================================
// f32x2: the fp32 values of two independent rays (lo = ray 0, hi = ray 1)
// =====================================================================================================
struct bool2 {
    bool lo, hi;
    LINON_HD bool2() : lo(false), hi(false) {}
    LINON_HD bool2(bool b) : lo(b), hi(b) {}
    LINON_HD bool2(bool a, bool b) : lo(a), hi(b) {}
};
LINON_HD bool2 operator&&(bool2 a, bool2 b) { return bool2(a.lo && b.lo, a.hi && b.hi); }
LINON_HD bool2 operator||(bool2 a, bool2 b) { return bool2(a.lo || b.lo, a.hi || b.hi); }
LINON_HD bool2 operator!(bool2 a) { return bool2(!a.lo, !a.hi); }
LINON_HD bool s_any(bool b) { return b; }
LINON_HD bool s_any(bool2 b) { return b.lo || b.hi; }
LINON_HD bool s_all(bool b) { return b; }
LINON_HD bool s_all(bool2 b) { return b.lo && b.hi; }

struct alignas(8) f32x2 {
    float lo, hi;
    LINON_HD f32x2() : lo(0.0f), hi(0.0f) {}
    LINON_HD f32x2(float s) : lo(s), hi(s) {}  // a uniform value broadcasts
    LINON_HD f32x2(float a, float b) : lo(a), hi(b) {}
};

#if LINON_PACKED_OPS
__device__ __forceinline__ unsigned long long f2_pack(f32x2 a) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.lo), "f"(a.hi));
    return r;
}
__device__ __forceinline__ f32x2 f2_unpack(unsigned long long v) {
    f32x2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.lo), "=f"(r.hi) : "l"(v));
    return r;
}
__device__ __forceinline__ f32x2 operator+(f32x2 a, f32x2 b) {
    const float2 r = __fadd2_rn(make_float2(a.lo, a.hi), make_float2(b.lo, b.hi));
    return f32x2(r.x, r.y);
}
__device__ __forceinline__ f32x2 operator*(f32x2 a, f32x2 b) {
    const float2 r = __fmul2_rn(make_float2(a.lo, a.hi), make_float2(b.lo, b.hi));
    return f32x2(r.x, r.y);
}
// sub.f32x2 keeps the negation an operand modifier of FADD2 / FFMA2 (a scalar `-x` costs an FADD per component
// under --ftz=true); ptxas contracts it with a preceding mul like any other add
__device__ __forceinline__ f32x2 operator-(f32x2 a, f32x2 b) {
    unsigned long long r;
    asm("sub.rn" LINON_F2_FTZ ".f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_pack(a)), "l"(f2_pack(b)));
    return f2_unpack(r);
}
__device__ __forceinline__ f32x2 operator-(f32x2 a) {
    unsigned long long r;
    asm("sub.rn" LINON_F2_FTZ ".f32x2 %0, %1, %2;" : "=l"(r) : "l"(0ull), "l"(f2_pack(a)));
    return f2_unpack(r);
}
#else
LINON_HD f32x2 operator+(f32x2 a, f32x2 b) { return f32x2(a.lo + b.lo, a.hi + b.hi); }
LINON_HD f32x2 operator*(f32x2 a, f32x2 b) { return f32x2(a.lo * b.lo, a.hi * b.hi); }
LINON_HD f32x2 operator-(f32x2 a, f32x2 b) { return f32x2(a.lo - b.lo, a.hi - b.hi); }
LINON_HD f32x2 operator-(f32x2 a) { return f32x2(-a.lo, -a.hi); }
#endif
#if LINON_FAST_MATH
// fast division is a * rcp(b) (what --use_fast_math makes of a scalar division): two MUFU.RCP and ONE packed multiply
LINON_HD f32x2 operator/(f32x2 a, f32x2 b) { return a * f32x2(1.0f / b.lo, 1.0f / b.hi); }
#else
LINON_HD f32x2 operator/(f32x2 a, f32x2 b) { return f32x2(a.lo / b.lo, a.hi / b.hi); }
#endif
LINON_HD f32x2 operator+(f32x2 a) { return a; }
#define LINON_F2_MIXED(OP)                                                       \
    LINON_HD f32x2 operator OP(f32x2 a, float b) { return a OP f32x2(b); }       \
    LINON_HD f32x2 operator OP(float a, f32x2 b) { return f32x2(a) OP b; }       \
    LINON_HD f32x2& operator OP##=(f32x2& a, f32x2 b) { a = a OP b; return a; }  \
    LINON_HD f32x2& operator OP##=(f32x2& a, float b) { a = a OP f32x2(b); return a; }
LINON_F2_MIXED(+) LINON_F2_MIXED(-) LINON_F2_MIXED(*) LINON_F2_MIXED(/)
#define LINON_F2_CMP(OP)                                                                          \
    LINON_HD bool2 operator OP(f32x2 a, f32x2 b) { return bool2(a.lo OP b.lo, a.hi OP b.hi); }    \
    LINON_HD bool2 operator OP(f32x2 a, float b) { return bool2(a.lo OP b, a.hi OP b); }          \
    LINON_HD bool2 operator OP(float a, f32x2 b) { return bool2(a OP b.lo, a OP b.hi); }
LINON_F2_CMP(<) LINON_F2_CMP(>) LINON_F2_CMP(<=) LINON_F2_CMP(>=) LINON_F2_CMP(==) LINON_F2_CMP(!=)

// ---- tiny type traits (NVRTC has no <type_traits>) ---------------------------------------------------------
template <class T> struct linon_is_pair { static constexpr bool value = false; };
template <> struct linon_is_pair<f32x2> { static constexpr bool value = true; };
template <class T> struct linon_is_arith { static constexpr bool value = false; };
template <> struct linon_is_arith<float> { static constexpr bool value = true; };
template <> struct linon_is_arith<double> { static constexpr bool value = true; };
template <> struct linon_is_arith<int> { static constexpr bool value = true; };
template <> struct linon_is_arith<unsigned> { static constexpr bool value = true; };
template <> struct linon_is_arith<long> { static constexpr bool value = true; };
template <> struct linon_is_arith<unsigned long> { static constexpr bool value = true; };
template <class T> struct linon_is_scalar { static constexpr bool value = linon_is_arith<T>::value || linon_is_pair<T>::value; };
template <bool B, class T = int> struct linon_enable_if {};
template <class T> struct linon_enable_if<true, T> { typedef T type; };
template <bool B, class A, class C> struct linon_cond { typedef A type; };
template <class A, class C> struct linon_cond<false, A, C> { typedef C type; };
template <class... T> struct linon_any_pair { static constexpr bool value = false; };
template <class A, class... T> struct linon_any_pair<A, T...> {
    static constexpr bool value = linon_is_pair<A>::value || linon_any_pair<T...>::value;
};
template <class... T> struct linon_all_scalar { static constexpr bool value = true; };
template <class A, class... T> struct linon_all_scalar<A, T...> {
    static constexpr bool value = linon_is_scalar<A>::value && linon_all_scalar<T...>::value;
};
// P<...>: f32x2 if any argument is per-ray, float otherwise (ints and doubles concretise to f32 like WGSL abstract numbers)
template <class... T> using P = typename linon_cond<linon_any_pair<T...>::value, f32x2, float>::type;
#define LINON_REQ(...) typename linon_enable_if<linon_all_scalar<__VA_ARGS__>::value, int>::type = 0
template <class S> struct linon_mask { typedef bool type; };
template <> struct linon_mask<f32x2> { typedef bool2 type; };

// ---- scalar primitives s_<name>, defined for float and f32x2 ------------------------------------------------
LINON_HD float s_abs(float x) { return fabsf(x); }
LINON_HD float s_min(float a, float b) { return fminf(a, b); }
LINON_HD float s_max(float a, float b) { return fmaxf(a, b); }
LINON_HD float s_sqrt(float x) { return sqrtf(x); }
// Fast mode names the reciprocal square root itself (MUFU.RSQ).  --use_fast_math folds 1 / sqrtf(x) and a / sqrtf(x) into it
// only as long as no earlier compilation of the same process asked for --prec-sqrt=true: NVRTC 12.9 keeps that switch, and a
// production kernel compiled after a strict or fused one came out with MUFU.SQRT + MUFU.RCP instead (12 more special-function
// evaluations per module, different last bits; found in round 2 - tests/test_capi_cpu.py pins the order independence).
#if LINON_FAST_MATH && LINON_DEVICE_BUILD
LINON_HD float s_inverseSqrt(float x) { return rsqrtf(x); }
#else
LINON_HD float s_inverseSqrt(float x) { return 1.0f / sqrtf(x); }
#endif
LINON_HD float s_exp(float x) { return expf(x); }
LINON_HD float s_exp2(float x) { return exp2f(x); }
LINON_HD float s_log(float x) { return logf(x); }
LINON_HD float s_log2(float x) { return log2f(x); }
LINON_HD float s_pow(float x, float y) { return powf(x, y); }
LINON_HD float s_sin(float x) { return sinf(x); }
LINON_HD float s_cos(float x) { return cosf(x); }
LINON_HD float s_tan(float x) { return tanf(x); }
LINON_HD float s_asin(float x) { return asinf(x); }
LINON_HD float s_acos(float x) { return acosf(x); }
LINON_HD float s_atan(float x) { return atanf(x); }
LINON_HD float s_atan2(float y, float x) { return atan2f(y, x); }
LINON_HD float s_sinh(float x) { return sinhf(x); }
LINON_HD float s_cosh(float x) { return coshf(x); }
LINON_HD float s_tanh(float x) { return tanhf(x); }
LINON_HD float s_floor(float x) { return floorf(x); }
LINON_HD float s_ceil(float x) { return ceilf(x); }
LINON_HD float s_round(float x) { return rintf(x); }  // WGSL round: ties to even
LINON_HD float s_trunc(float x) { return truncf(x); }
LINON_HD float s_fract(float x) { return x - floorf(x); }
LINON_HD float s_sign(float x) { return (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f); }
LINON_HD float s_step(float edge, float x) { return (edge <= x) ? 1.0f : 0.0f; }
LINON_HD float s_sel(bool c, float t, float f) { return c ? t : f; }
LINON_HD f32x2 s_sel(bool2 c, f32x2 t, f32x2 f) { return f32x2(c.lo ? t.lo : f.lo, c.hi ? t.hi : f.hi); }
LINON_HD f32x2 s_sel(bool c, f32x2 t, f32x2 f) { return c ? t : f; }
#define LINON_PAIR1(fn) LINON_HD f32x2 fn(f32x2 a) { return f32x2(fn(a.lo), fn(a.hi)); }
#define LINON_PAIR2(fn) LINON_HD f32x2 fn(f32x2 a, f32x2 b) { return f32x2(fn(a.lo, b.lo), fn(a.hi, b.hi)); }
LINON_PAIR1(s_abs) LINON_PAIR1(s_sqrt) LINON_PAIR1(s_inverseSqrt) LINON_PAIR1(s_exp) LINON_PAIR1(s_exp2) LINON_PAIR1(s_log)
LINON_PAIR1(s_log2) LINON_PAIR1(s_sin) LINON_PAIR1(s_cos) LINON_PAIR1(s_tan) LINON_PAIR1(s_asin) LINON_PAIR1(s_acos)
LINON_PAIR1(s_atan) LINON_PAIR1(s_sinh) LINON_PAIR1(s_cosh) LINON_PAIR1(s_tanh) LINON_PAIR1(s_floor) LINON_PAIR1(s_ceil)
LINON_PAIR1(s_round) LINON_PAIR1(s_trunc) LINON_PAIR1(s_fract) LINON_PAIR1(s_sign)
LINON_PAIR2(s_min) LINON_PAIR2(s_max) LINON_PAIR2(s_pow) LINON_PAIR2(s_atan2) LINON_PAIR2(s_step)

// =====================================================================================================
// vectors
// =====================================================================================================
template <class S> struct vec2t;
template <class S> struct vec3t;
template <class S> struct vec4t;

#define LINON_SWZ2_DECL(a, b) LINON_HD vec2t<S> a##b() const;
#define LINON_SWZ3_DECL(a, b, c) LINON_HD vec3t<S> a##b##c() const;
#define LINON_FOR2_OF2(M) M(x, x) M(x, y) M(y, x) M(y, y)
#define LINON_FOR2_OF3(M) LINON_FOR2_OF2(M) M(x, z) M(y, z) M(z, x) M(z, y) M(z, z)
#define LINON_FOR2_OF4(M) LINON_FOR2_OF3(M) M(x, w) M(y, w) M(z, w) M(w, x) M(w, y) M(w, z) M(w, w)
#define LINON_FOR3_ROW3(M, a, b) M(a, b, x) M(a, b, y) M(a, b, z)
#define LINON_FOR3_OF3(M)                                                                                   \
    LINON_FOR3_ROW3(M, x, x) LINON_FOR3_ROW3(M, x, y) LINON_FOR3_ROW3(M, x, z) LINON_FOR3_ROW3(M, y, x)     \
    LINON_FOR3_ROW3(M, y, y) LINON_FOR3_ROW3(M, y, z) LINON_FOR3_ROW3(M, z, x) LINON_FOR3_ROW3(M, z, y)     \
    LINON_FOR3_ROW3(M, z, z)
// three- and four-component swizzles of a vec4 that involve w
#define LINON_FOR3_ROW4W(M, a, b) M(a, b, w)
#define LINON_FOR3_ROW4(M, a, b) LINON_FOR3_ROW3(M, a, b) M(a, b, w)
#define LINON_FOR3_OF4_W(M)                                                                                 \
    LINON_FOR3_ROW4W(M, x, x) LINON_FOR3_ROW4W(M, x, y) LINON_FOR3_ROW4W(M, x, z) LINON_FOR3_ROW4W(M, y, x) \
    LINON_FOR3_ROW4W(M, y, y) LINON_FOR3_ROW4W(M, y, z) LINON_FOR3_ROW4W(M, z, x) LINON_FOR3_ROW4W(M, z, y) \
    LINON_FOR3_ROW4W(M, z, z) LINON_FOR3_ROW4(M, x, w) LINON_FOR3_ROW4(M, y, w) LINON_FOR3_ROW4(M, z, w)    \
    LINON_FOR3_ROW4(M, w, x) LINON_FOR3_ROW4(M, w, y) LINON_FOR3_ROW4(M, w, z) LINON_FOR3_ROW4(M, w, w)
#define LINON_FOR4_ROW(M, a, b, c) M(a, b, c, x) M(a, b, c, y) M(a, b, c, z) M(a, b, c, w)
#define LINON_FOR4_ROW2(M, a, b) LINON_FOR4_ROW(M, a, b, x) LINON_FOR4_ROW(M, a, b, y) LINON_FOR4_ROW(M, a, b, z) LINON_FOR4_ROW(M, a, b, w)
#define LINON_FOR4_ROW3(M, a) LINON_FOR4_ROW2(M, a, x) LINON_FOR4_ROW2(M, a, y) LINON_FOR4_ROW2(M, a, z) LINON_FOR4_ROW2(M, a, w)
#define LINON_FOR4_OF4(M) LINON_FOR4_ROW3(M, x) LINON_FOR4_ROW3(M, y) LINON_FOR4_ROW3(M, z) LINON_FOR4_ROW3(M, w)
#define LINON_SWZ4_DEF(a, b, c, d) LINON_HD vec4t<S> a##b##c##d() const { return vec4t<S>(a, b, c, d); }

template <class S> struct vec2t {
    S x, y;
    LINON_HD vec2t() : x(0.0f), y(0.0f) {}
    template <class A, LINON_REQ(A)> LINON_HD explicit vec2t(A s) : x(S(s)), y(S(s)) {}
    template <class A, class B, LINON_REQ(A, B)> LINON_HD vec2t(A a, B b) : x(S(a)), y(S(b)) {}
    template <class A> LINON_HD vec2t(vec2t<A> o) : x(S(o.x)), y(S(o.y)) {}  // float -> per-ray promotion
    LINON_HD S& operator[](int i) { return (&x)[i]; }
    LINON_HD S operator[](int i) const { return (&x)[i]; }
    LINON_FOR2_OF2(LINON_SWZ2_DECL)
};

template <class S> struct vec3t {
    S x, y, z;
    LINON_HD vec3t() : x(0.0f), y(0.0f), z(0.0f) {}
    template <class A, LINON_REQ(A)> LINON_HD explicit vec3t(A s) : x(S(s)), y(S(s)), z(S(s)) {}
    template <class A, class B, class C, LINON_REQ(A, B, C)> LINON_HD vec3t(A a, B b, C c) : x(S(a)), y(S(b)), z(S(c)) {}
    template <class A, class C, LINON_REQ(C)> LINON_HD vec3t(vec2t<A> a, C c) : x(S(a.x)), y(S(a.y)), z(S(c)) {}
    template <class A, class B, LINON_REQ(A)> LINON_HD vec3t(A a, vec2t<B> b) : x(S(a)), y(S(b.x)), z(S(b.y)) {}
    template <class A> LINON_HD vec3t(vec3t<A> o) : x(S(o.x)), y(S(o.y)), z(S(o.z)) {}
    LINON_HD S& operator[](int i) { return (&x)[i]; }
    LINON_HD S operator[](int i) const { return (&x)[i]; }
    LINON_FOR2_OF3(LINON_SWZ2_DECL)
    LINON_FOR3_OF3(LINON_SWZ3_DECL)
};

template <class S> struct vec4t {
    S x, y, z, w;
    LINON_HD vec4t() : x(0.0f), y(0.0f), z(0.0f), w(0.0f) {}
    template <class A, LINON_REQ(A)> LINON_HD explicit vec4t(A s) : x(S(s)), y(S(s)), z(S(s)), w(S(s)) {}
    template <class A, class B, class C, class D, LINON_REQ(A, B, C, D)>
    LINON_HD vec4t(A a, B b, C c, D d) : x(S(a)), y(S(b)), z(S(c)), w(S(d)) {}
    template <class A, class D, LINON_REQ(D)> LINON_HD vec4t(vec3t<A> a, D d) : x(S(a.x)), y(S(a.y)), z(S(a.z)), w(S(d)) {}
    template <class A, class B, LINON_REQ(A)> LINON_HD vec4t(A a, vec3t<B> b) : x(S(a)), y(S(b.x)), z(S(b.y)), w(S(b.z)) {}
    template <class A, class B> LINON_HD vec4t(vec2t<A> a, vec2t<B> b) : x(S(a.x)), y(S(a.y)), z(S(b.x)), w(S(b.y)) {}
    template <class A, class C, class D, LINON_REQ(C, D)>
    LINON_HD vec4t(vec2t<A> a, C c, D d) : x(S(a.x)), y(S(a.y)), z(S(c)), w(S(d)) {}
    template <class A> LINON_HD vec4t(vec4t<A> o) : x(S(o.x)), y(S(o.y)), z(S(o.z)), w(S(o.w)) {}
    LINON_HD S& operator[](int i) { return (&x)[i]; }
    LINON_HD S operator[](int i) const { return (&x)[i]; }
    LINON_FOR2_OF4(LINON_SWZ2_DECL)
    LINON_FOR3_OF3(LINON_SWZ3_DECL)
    LINON_FOR3_OF4_W(LINON_SWZ3_DECL)
    LINON_FOR4_OF4(LINON_SWZ4_DEF)
};

#define LINON_SWZ2_DEF_V2(a, b) template <class S> LINON_HD vec2t<S> vec2t<S>::a##b() const { return vec2t<S>(a, b); }
#define LINON_SWZ2_DEF_V3(a, b) template <class S> LINON_HD vec2t<S> vec3t<S>::a##b() const { return vec2t<S>(a, b); }
#define LINON_SWZ2_DEF_V4(a, b) template <class S> LINON_HD vec2t<S> vec4t<S>::a##b() const { return vec2t<S>(a, b); }
#define LINON_SWZ3_DEF_V3(a, b, c) template <class S> LINON_HD vec3t<S> vec3t<S>::a##b##c() const { return vec3t<S>(a, b, c); }
#define LINON_SWZ3_DEF_V4(a, b, c) template <class S> LINON_HD vec3t<S> vec4t<S>::a##b##c() const { return vec3t<S>(a, b, c); }
LINON_FOR2_OF2(LINON_SWZ2_DEF_V2)
LINON_FOR2_OF3(LINON_SWZ2_DEF_V3)
LINON_FOR2_OF4(LINON_SWZ2_DEF_V4)
LINON_FOR3_OF3(LINON_SWZ3_DEF_V3)
LINON_FOR3_OF3(LINON_SWZ3_DEF_V4)
LINON_FOR3_OF4_W(LINON_SWZ3_DEF_V4)

typedef float wg_f32;  // what the translator emits for the WGSL type f32; LinonField<S> re-declares it as S
typedef vec2t<float> vec2f;
typedef vec3t<float> vec3f;
typedef vec4t<float> vec4f;

// ---- arithmetic operators (component-wise, scalar broadcast, float <-> per-ray promotion) ----------------------
#define LINON_VEC_OPS(OP)                                                                                              \
    template <class A, class B> LINON_HD vec2t<P<A, B>> operator OP(vec2t<A> a, vec2t<B> b) {                          \
        typedef P<A, B> S; return vec2t<S>(S(a.x) OP S(b.x), S(a.y) OP S(b.y)); }                                      \
    template <class A, class B, LINON_REQ(B)> LINON_HD vec2t<P<A, B>> operator OP(vec2t<A> a, B s) {                    \
        typedef P<A, B> S; return vec2t<S>(S(a.x) OP S(s), S(a.y) OP S(s)); }                                          \
    template <class A, class B, LINON_REQ(A)> LINON_HD vec2t<P<A, B>> operator OP(A s, vec2t<B> b) {                    \
        typedef P<A, B> S; return vec2t<S>(S(s) OP S(b.x), S(s) OP S(b.y)); }                                          \
    template <class A, class B> LINON_HD vec3t<P<A, B>> operator OP(vec3t<A> a, vec3t<B> b) {                          \
        typedef P<A, B> S; return vec3t<S>(S(a.x) OP S(b.x), S(a.y) OP S(b.y), S(a.z) OP S(b.z)); }                    \
    template <class A, class B, LINON_REQ(B)> LINON_HD vec3t<P<A, B>> operator OP(vec3t<A> a, B s) {                    \
        typedef P<A, B> S; return vec3t<S>(S(a.x) OP S(s), S(a.y) OP S(s), S(a.z) OP S(s)); }                          \
    template <class A, class B, LINON_REQ(A)> LINON_HD vec3t<P<A, B>> operator OP(A s, vec3t<B> b) {                    \
        typedef P<A, B> S; return vec3t<S>(S(s) OP S(b.x), S(s) OP S(b.y), S(s) OP S(b.z)); }                          \
    template <class A, class B> LINON_HD vec4t<P<A, B>> operator OP(vec4t<A> a, vec4t<B> b) {                          \
        typedef P<A, B> S; return vec4t<S>(S(a.x) OP S(b.x), S(a.y) OP S(b.y), S(a.z) OP S(b.z), S(a.w) OP S(b.w)); }  \
    template <class A, class B, LINON_REQ(B)> LINON_HD vec4t<P<A, B>> operator OP(vec4t<A> a, B s) {                    \
        typedef P<A, B> S; return vec4t<S>(S(a.x) OP S(s), S(a.y) OP S(s), S(a.z) OP S(s), S(a.w) OP S(s)); }          \
    template <class A, class B, LINON_REQ(A)> LINON_HD vec4t<P<A, B>> operator OP(A s, vec4t<B> b) {                    \
        typedef P<A, B> S; return vec4t<S>(S(s) OP S(b.x), S(s) OP S(b.y), S(s) OP S(b.z), S(s) OP S(b.w)); }          \
    template <class A, class B> LINON_HD vec2t<A>& operator OP##=(vec2t<A>& a, B b) { a = vec2t<A>(a OP b); return a; } \
    template <class A, class B> LINON_HD vec3t<A>& operator OP##=(vec3t<A>& a, B b) { a = vec3t<A>(a OP b); return a; } \
    template <class A, class B> LINON_HD vec4t<A>& operator OP##=(vec4t<A>& a, B b) { a = vec4t<A>(a OP b); return a; }
LINON_VEC_OPS(+)
LINON_VEC_OPS(-)
LINON_VEC_OPS(*)
LINON_VEC_OPS(/)
template <class S> LINON_HD vec2t<S> operator-(vec2t<S> a) { return vec2t<S>(-a.x, -a.y); }
template <class S> LINON_HD vec3t<S> operator-(vec3t<S> a) { return vec3t<S>(-a.x, -a.y, -a.z); }
template <class S> LINON_HD vec4t<S> operator-(vec4t<S> a) { return vec4t<S>(-a.x, -a.y, -a.z, -a.w); }

// ---- scalar builtins (w_<name>) ------------------------------------------------------------------------------
#define LINON_W1(name)                                                                                              \
    template <class A, LINON_REQ(A)> LINON_HD P<A> w_##name(A x) { return s_##name(P<A>(x)); }                      \
    template <class A> LINON_HD vec2t<A> w_##name(vec2t<A> a) { return vec2t<A>(s_##name(a.x), s_##name(a.y)); }    \
    template <class A> LINON_HD vec3t<A> w_##name(vec3t<A> a) {                                                     \
        return vec3t<A>(s_##name(a.x), s_##name(a.y), s_##name(a.z)); }                                             \
    template <class A> LINON_HD vec4t<A> w_##name(vec4t<A> a) {                                                     \
        return vec4t<A>(s_##name(a.x), s_##name(a.y), s_##name(a.z), s_##name(a.w)); }
#define LINON_W2(name)                                                                                              \
    template <class A, class B, LINON_REQ(A, B)> LINON_HD P<A, B> w_##name(A a, B b) {                              \
        typedef P<A, B> S; return s_##name(S(a), S(b)); }                                                           \
    template <class A, class B> LINON_HD vec2t<P<A, B>> w_##name(vec2t<A> a, vec2t<B> b) {                          \
        typedef P<A, B> S; return vec2t<S>(s_##name(S(a.x), S(b.x)), s_##name(S(a.y), S(b.y))); }                   \
    template <class A, class B> LINON_HD vec3t<P<A, B>> w_##name(vec3t<A> a, vec3t<B> b) {                          \
        typedef P<A, B> S;                                                                                          \
        return vec3t<S>(s_##name(S(a.x), S(b.x)), s_##name(S(a.y), S(b.y)), s_##name(S(a.z), S(b.z))); }            \
    template <class A, class B> LINON_HD vec4t<P<A, B>> w_##name(vec4t<A> a, vec4t<B> b) {                          \
        typedef P<A, B> S;                                                                                          \
        return vec4t<S>(s_##name(S(a.x), S(b.x)), s_##name(S(a.y), S(b.y)), s_##name(S(a.z), S(b.z)),               \
                        s_##name(S(a.w), S(b.w))); }
LINON_W1(abs) LINON_W1(sqrt) LINON_W1(inverseSqrt) LINON_W1(exp) LINON_W1(exp2) LINON_W1(log) LINON_W1(log2) LINON_W1(sin)
LINON_W1(cos) LINON_W1(tan) LINON_W1(asin) LINON_W1(acos) LINON_W1(atan) LINON_W1(sinh) LINON_W1(cosh) LINON_W1(tanh)
LINON_W1(floor) LINON_W1(ceil) LINON_W1(round) LINON_W1(trunc) LINON_W1(fract) LINON_W1(sign)
LINON_W2(min) LINON_W2(max) LINON_W2(pow) LINON_W2(atan2) LINON_W2(step)

// three-operand builtins are written on scalars and lifted
template <class S> LINON_HD S s_clamp(S e, S lo, S hi) { return s_min(s_max(e, lo), hi); }
template <class S> LINON_HD S s_mix(S a, S b, S t) { return a * (S(1.0f) - t) + b * t; }
template <class S> LINON_HD S s_smoothstep(S lo, S hi, S x) {
    const S t = s_clamp((x - lo) / (hi - lo), S(0.0f), S(1.0f));
    return t * t * (S(3.0f) - S(2.0f) * t);
}
template <class S> LINON_HD S s_fma(S a, S b, S c) { return a * b + c; }
#define LINON_W3(name)                                                                                              \
    template <class A, class B, class C, LINON_REQ(A, B, C)> LINON_HD P<A, B, C> w_##name(A a, B b, C c) {          \
        typedef P<A, B, C> S; return s_##name(S(a), S(b), S(c)); }                                                  \
    template <class A, class B, class C> LINON_HD vec2t<P<A, B, C>> w_##name(vec2t<A> a, vec2t<B> b, vec2t<C> c) {  \
        typedef P<A, B, C> S;                                                                                       \
        return vec2t<S>(s_##name(S(a.x), S(b.x), S(c.x)), s_##name(S(a.y), S(b.y), S(c.y))); }                      \
    template <class A, class B, class C> LINON_HD vec3t<P<A, B, C>> w_##name(vec3t<A> a, vec3t<B> b, vec3t<C> c) {  \
        typedef P<A, B, C> S;                                                                                       \
        return vec3t<S>(s_##name(S(a.x), S(b.x), S(c.x)), s_##name(S(a.y), S(b.y), S(c.y)),                         \
                        s_##name(S(a.z), S(b.z), S(c.z))); }                                                        \
    template <class A, class B, class C> LINON_HD vec4t<P<A, B, C>> w_##name(vec4t<A> a, vec4t<B> b, vec4t<C> c) {  \
        typedef P<A, B, C> S;                                                                                       \
        return vec4t<S>(s_##name(S(a.x), S(b.x), S(c.x)), s_##name(S(a.y), S(b.y), S(c.y)),                         \
                        s_##name(S(a.z), S(b.z), S(c.z)), s_##name(S(a.w), S(b.w), S(c.w))); }
LINON_W3(clamp) LINON_W3(mix) LINON_W3(smoothstep) LINON_W3(fma)
template <class A, LINON_REQ(A)> LINON_HD P<A> w_saturate(A e) { typedef P<A> S; return s_clamp(S(e), S(0.0f), S(1.0f)); }
template <class A> LINON_HD vec2t<A> w_saturate(vec2t<A> a) { return vec2t<A>(w_saturate(a.x), w_saturate(a.y)); }
template <class A> LINON_HD vec3t<A> w_saturate(vec3t<A> a) { return vec3t<A>(w_saturate(a.x), w_saturate(a.y), w_saturate(a.z)); }
template <class A> LINON_HD vec4t<A> w_saturate(vec4t<A> a) {
    return vec4t<A>(w_saturate(a.x), w_saturate(a.y), w_saturate(a.z), w_saturate(a.w));
}
template <class A, LINON_REQ(A)> LINON_HD P<A> w_degrees(A x) { return P<A>(x) * 57.295779513082322865f; }
template <class A, LINON_REQ(A)> LINON_HD P<A> w_radians(A x) { return P<A>(x) * 0.017453292519943295474f; }
template <class A> LINON_HD vec3t<A> w_degrees(vec3t<A> a) { return a * 57.295779513082322865f; }
template <class A> LINON_HD vec3t<A> w_radians(vec3t<A> a) { return a * 0.017453292519943295474f; }
// mix with a scalar weight on vectors
template <class A, class B, class C, LINON_REQ(C)> LINON_HD vec2t<P<A, B, C>> w_mix(vec2t<A> a, vec2t<B> b, C t) {
    typedef P<A, B, C> S; return vec2t<S>(a) * (S(1.0f) - S(t)) + vec2t<S>(b) * S(t); }
template <class A, class B, class C, LINON_REQ(C)> LINON_HD vec3t<P<A, B, C>> w_mix(vec3t<A> a, vec3t<B> b, C t) {
    typedef P<A, B, C> S; return vec3t<S>(a) * (S(1.0f) - S(t)) + vec3t<S>(b) * S(t); }
template <class A, class B, class C, LINON_REQ(C)> LINON_HD vec4t<P<A, B, C>> w_mix(vec4t<A> a, vec4t<B> b, C t) {
    typedef P<A, B, C> S; return vec4t<S>(a) * (S(1.0f) - S(t)) + vec4t<S>(b) * S(t); }
// select(f, t, cond): cond is bool (uniform) or bool2 (per ray)
template <class A, class B, LINON_REQ(A, B)> LINON_HD P<A, B> w_select(A f, B t, bool c) { typedef P<A, B> S; return c ? S(t) : S(f); }
template <class A, class B, LINON_REQ(A, B)> LINON_HD f32x2 w_select(A f, B t, bool2 c) { return s_sel(c, f32x2(t), f32x2(f)); }
template <class A, class B> LINON_HD vec2t<P<A, B>> w_select(vec2t<A> f, vec2t<B> t, bool c) { typedef P<A, B> S; return c ? vec2t<S>(t) : vec2t<S>(f); }
template <class A, class B> LINON_HD vec3t<P<A, B>> w_select(vec3t<A> f, vec3t<B> t, bool c) { typedef P<A, B> S; return c ? vec3t<S>(t) : vec3t<S>(f); }
template <class A, class B> LINON_HD vec4t<P<A, B>> w_select(vec4t<A> f, vec4t<B> t, bool c) { typedef P<A, B> S; return c ? vec4t<S>(t) : vec4t<S>(f); }
template <class A, class B> LINON_HD vec2t<f32x2> w_select(vec2t<A> f, vec2t<B> t, bool2 c) {
    return vec2t<f32x2>(w_select(f.x, t.x, c), w_select(f.y, t.y, c)); }
template <class A, class B> LINON_HD vec3t<f32x2> w_select(vec3t<A> f, vec3t<B> t, bool2 c) {
    return vec3t<f32x2>(w_select(f.x, t.x, c), w_select(f.y, t.y, c), w_select(f.z, t.z, c)); }
template <class A, class B> LINON_HD vec4t<f32x2> w_select(vec4t<A> f, vec4t<B> t, bool2 c) {
    return vec4t<f32x2>(w_select(f.x, t.x, c), w_select(f.y, t.y, c), w_select(f.z, t.z, c), w_select(f.w, t.w, c)); }
// WGSL's % on floats (the translator rewrites `a % b` to w_mod(a, b): C++ has no float %): x - y * trunc(x / y), lifted like
// every two-operand builtin; vector % scalar and scalar % vector mix like the arithmetic operators
template <class S> LINON_HD S s_mod(S x, S y) { return x - y * s_trunc(x / y); }
LINON_W2(mod)
template <class A, class B, LINON_REQ(B)> LINON_HD vec2t<P<A, B>> w_mod(vec2t<A> a, B s) { typedef P<A, B> S; return w_mod(vec2t<S>(a), vec2t<S>(S(s))); }
template <class A, class B, LINON_REQ(B)> LINON_HD vec3t<P<A, B>> w_mod(vec3t<A> a, B s) { typedef P<A, B> S; return w_mod(vec3t<S>(a), vec3t<S>(S(s))); }
template <class A, class B, LINON_REQ(B)> LINON_HD vec4t<P<A, B>> w_mod(vec4t<A> a, B s) { typedef P<A, B> S; return w_mod(vec4t<S>(a), vec4t<S>(S(s))); }
template <class A, class B, LINON_REQ(A)> LINON_HD vec2t<P<A, B>> w_mod(A s, vec2t<B> b) { typedef P<A, B> S; return w_mod(vec2t<S>(S(s)), vec2t<S>(b)); }
template <class A, class B, LINON_REQ(A)> LINON_HD vec3t<P<A, B>> w_mod(A s, vec3t<B> b) { typedef P<A, B> S; return w_mod(vec3t<S>(S(s)), vec3t<S>(b)); }
template <class A, class B, LINON_REQ(A)> LINON_HD vec4t<P<A, B>> w_mod(A s, vec4t<B> b) { typedef P<A, B> S; return w_mod(vec4t<S>(S(s)), vec4t<S>(b)); }
// integer builtins keep integer results (non-templates win overload resolution)
LINON_HD int w_mod(int a, int b) { return a % b; }
LINON_HD unsigned w_mod(unsigned a, unsigned b) { return a % b; }
LINON_HD int w_select(int f, int t, bool c) { return c ? t : f; }
LINON_HD int w_abs(int x) { return x < 0 ? -x : x; }
LINON_HD int w_min(int a, int b) { return a < b ? a : b; }
LINON_HD int w_max(int a, int b) { return a < b ? b : a; }
LINON_HD int w_clamp(int e, int lo, int hi) { return w_min(w_max(e, lo), hi); }

// ---- geometric builtins ------------------------------------------------------------------------------------
template <class A, class B> LINON_HD P<A, B> w_dot(vec2t<A> a, vec2t<B> b) { typedef P<A, B> S; return S(a.x) * S(b.x) + S(a.y) * S(b.y); }
template <class A, class B> LINON_HD P<A, B> w_dot(vec3t<A> a, vec3t<B> b) {
    typedef P<A, B> S; return S(a.x) * S(b.x) + S(a.y) * S(b.y) + S(a.z) * S(b.z); }
template <class A, class B> LINON_HD P<A, B> w_dot(vec4t<A> a, vec4t<B> b) {
    typedef P<A, B> S; return S(a.x) * S(b.x) + S(a.y) * S(b.y) + S(a.z) * S(b.z) + S(a.w) * S(b.w); }
template <class A, class B> LINON_HD vec3t<P<A, B>> w_cross(vec3t<A> a_, vec3t<B> b_) {
    typedef P<A, B> S;
    const vec3t<S> a(a_), b(b_);
    return vec3t<S>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
template <class A, LINON_REQ(A)> LINON_HD P<A> w_length(A a) { return s_abs(P<A>(a)); }
template <class A> LINON_HD A w_length(vec2t<A> a) { return s_sqrt(w_dot(a, a)); }
template <class A> LINON_HD A w_length(vec3t<A> a) { return s_sqrt(w_dot(a, a)); }
template <class A> LINON_HD A w_length(vec4t<A> a) { return s_sqrt(w_dot(a, a)); }
template <class A, class B> LINON_HD P<A, B> w_distance(vec2t<A> a, vec2t<B> b) { return w_length(a - b); }
template <class A, class B> LINON_HD P<A, B> w_distance(vec3t<A> a, vec3t<B> b) { return w_length(a - b); }
template <class A, class B> LINON_HD P<A, B> w_distance(vec4t<A> a, vec4t<B> b) { return w_length(a - b); }
#if LINON_FAST_MATH && LINON_DEVICE_BUILD
// a * rsqrt(a.a): what --use_fast_math makes of a / sqrt(a.a) in a fresh process, said explicitly (see s_inverseSqrt)
template <class A> LINON_HD vec2t<A> w_normalize(vec2t<A> a) { return a * w_inverseSqrt(w_dot(a, a)); }
template <class A> LINON_HD vec3t<A> w_normalize(vec3t<A> a) { return a * w_inverseSqrt(w_dot(a, a)); }
template <class A> LINON_HD vec4t<A> w_normalize(vec4t<A> a) { return a * w_inverseSqrt(w_dot(a, a)); }
#else
template <class A> LINON_HD vec2t<A> w_normalize(vec2t<A> a) { return a / w_length(a); }
template <class A> LINON_HD vec3t<A> w_normalize(vec3t<A> a) { return a / w_length(a); }
template <class A> LINON_HD vec4t<A> w_normalize(vec4t<A> a) { return a / w_length(a); }
#endif
template <class A, class B> LINON_HD vec2t<P<A, B>> w_reflect(vec2t<A> e1, vec2t<B> e2) {
    typedef P<A, B> S; return e1 - (S(2.0f) * w_dot(e2, e1)) * e2; }
template <class A, class B> LINON_HD vec3t<P<A, B>> w_reflect(vec3t<A> e1, vec3t<B> e2) {
    typedef P<A, B> S; return e1 - (S(2.0f) * w_dot(e2, e1)) * e2; }
template <class A, class B, class C> LINON_HD vec3t<P<A, B, C>> w_faceForward(vec3t<A> e1, vec3t<B> e2, vec3t<C> e3) {
    typedef P<A, B, C> S;
    const vec3t<S> a(e1);
    return w_select(-a, a, w_dot(e2, e3) < S(0.0f));
}

// ---- small matrices (column major, WGSL semantics) ------------------------------------------------------------
template <class S> struct mat2x2t {
    vec2t<S> c0, c1;
    LINON_HD mat2x2t() {}
    template <class A, class B> LINON_HD mat2x2t(vec2t<A> a, vec2t<B> b) : c0(a), c1(b) {}
    template <class A, class B, class C, class D, LINON_REQ(A, B, C, D)> LINON_HD mat2x2t(A a, B b, C c, D d) : c0(a, b), c1(c, d) {}
    LINON_HD vec2t<S>& operator[](int i) { return (&c0)[i]; }
    LINON_HD vec2t<S> operator[](int i) const { return (&c0)[i]; }
};
template <class S> struct mat3x3t {
    vec3t<S> c0, c1, c2;
    LINON_HD mat3x3t() {}
    template <class A, class B, class C> LINON_HD mat3x3t(vec3t<A> a, vec3t<B> b, vec3t<C> c) : c0(a), c1(b), c2(c) {}
    template <class A, class B, class C, class D, class E, class F, class G, class H, class I,
              LINON_REQ(A, B, C, D, E, F, G, H, I)>
    LINON_HD mat3x3t(A a, B b, C c, D d, E e, F f, G g, H h, I i) : c0(a, b, c), c1(d, e, f), c2(g, h, i) {}
    LINON_HD vec3t<S>& operator[](int i) { return (&c0)[i]; }
    LINON_HD vec3t<S> operator[](int i) const { return (&c0)[i]; }
};
typedef mat2x2t<float> mat2x2f;
typedef mat3x3t<float> mat3x3f;
template <class A, class B> LINON_HD vec2t<P<A, B>> operator*(mat2x2t<A> m, vec2t<B> v) { return m.c0 * v.x + m.c1 * v.y; }
template <class A, class B> LINON_HD vec2t<P<A, B>> operator*(vec2t<A> v, mat2x2t<B> m) {
    return vec2t<P<A, B>>(w_dot(v, m.c0), w_dot(v, m.c1)); }
template <class A, class B> LINON_HD mat2x2t<P<A, B>> operator*(mat2x2t<A> a, mat2x2t<B> b) {
    return mat2x2t<P<A, B>>(a * b.c0, a * b.c1); }
template <class A, class B> LINON_HD vec3t<P<A, B>> operator*(mat3x3t<A> m, vec3t<B> v) { return m.c0 * v.x + m.c1 * v.y + m.c2 * v.z; }
template <class A, class B> LINON_HD vec3t<P<A, B>> operator*(vec3t<A> v, mat3x3t<B> m) {
    return vec3t<P<A, B>>(w_dot(v, m.c0), w_dot(v, m.c1), w_dot(v, m.c2)); }
template <class A, class B> LINON_HD mat3x3t<P<A, B>> operator*(mat3x3t<A> a, mat3x3t<B> b) {
    return mat3x3t<P<A, B>>(a * b.c0, a * b.c1, a * b.c2); }
template <class A, class B, LINON_REQ(B)> LINON_HD mat3x3t<P<A, B>> operator*(mat3x3t<A> a, B s) {
    return mat3x3t<P<A, B>>(a.c0 * s, a.c1 * s, a.c2 * s); }
template <class A, class B, LINON_REQ(A)> LINON_HD mat3x3t<P<A, B>> operator*(A s, mat3x3t<B> a) {
    return mat3x3t<P<A, B>>(a.c0 * s, a.c1 * s, a.c2 * s); }
template <class S> LINON_HD mat2x2t<S> w_transpose(mat2x2t<S> m) { return mat2x2t<S>(m.c0.x, m.c1.x, m.c0.y, m.c1.y); }
template <class S> LINON_HD mat3x3t<S> w_transpose(mat3x3t<S> m) {
    return mat3x3t<S>(m.c0.x, m.c1.x, m.c2.x, m.c0.y, m.c1.y, m.c2.y, m.c0.z, m.c1.z, m.c2.z);
}
template <class S> LINON_HD S w_determinant(mat2x2t<S> m) { return m.c0.x * m.c1.y - m.c0.y * m.c1.x; }
template <class S> LINON_HD S w_determinant(mat3x3t<S> m) { return w_dot(m.c0, w_cross(m.c1, m.c2)); }

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
template <class A, class B, LINON_REQ(B)> LINON_HD vec3t<P<A, B>> rotateX(vec3t<A> v, B phi_) {
    typedef P<A, B> S;
    const S phi(phi_);
    return mat3x3t<S>(vec3t<S>(1.0f, 0.0f, 0.0f), vec3t<S>(0.0f, w_cos(phi), w_sin(phi)), vec3t<S>(0.0f, -w_sin(phi), w_cos(phi))) * v;
}
template <class A, class B, LINON_REQ(B)> LINON_HD vec3t<P<A, B>> rotateY(vec3t<A> v, B phi_) {
    typedef P<A, B> S;
    const S phi(phi_);
    return mat3x3t<S>(vec3t<S>(w_cos(phi), 0.0f, -w_sin(phi)), vec3t<S>(0.0f, 1.0f, 0.0f), vec3t<S>(w_sin(phi), 0.0f, w_cos(phi))) * v;
}
template <class A, class B, LINON_REQ(B)> LINON_HD vec3t<P<A, B>> rotateZ(vec3t<A> v, B phi_) {  // third column is zero in the reference: z is dropped
    typedef P<A, B> S;
    const S phi(phi_);
    return mat3x3t<S>(vec3t<S>(w_cos(phi), w_sin(phi), 0.0f), vec3t<S>(-w_sin(phi), w_cos(phi), 0.0f), vec3t<S>(0.0f, 0.0f, 0.0f)) * v;
}
template <class A, class B, class C, class D, LINON_REQ(B, C, D)>
LINON_HD vec3t<P<A, B, C, D>> translate(vec3t<A> v, B dx, C dy, D dz) {
    typedef P<A, B, C, D> S;
    return vec3t<S>(S(v.x) + S(dx), S(v.y) + S(dy), S(v.z) + S(dz));
}

// pow(10.0, -10.0) is folded by naga in abstract float and concretised to f32 (1e-10f)
template <class A, LINON_REQ(A)> LINON_HD P<A> refraction_index(A t_) {
    typedef P<A> S;
    const S t(t_);
    const float air_pressure = 101325.0f;
    const float c1 = 0.0000104f;
    const float c2 = 0.00366f;
    return c1 * air_pressure * (1.0f + air_pressure * (60.1f - 0.972f * t) * 1e-10f) / (1.0f + c2 * t);
}
// eta = n(t_in) / n(t_out).  Fast mode folds the two refraction indices into one quotient: n(t) = A(t) / B(t) with
// A(t) = c1 p (1 + p 1e-10 (60.1 - 0.972 t)) = a0 - a1 t and B(t) = 1 + c2 t, so eta = A(t_in) B(t_out) / (A(t_out) B(t_in))
// costs one reciprocal instead of three (same fp32 error class: every factor is 1 + O(1e-3)).
template <class S> LINON_HD S refraction_eta(S t_in, S t_out) {
#if LINON_FAST_MATH
    const float c1p = 0.0000104f * 101325.0f;
    const float a0 = c1p * (1.0f + 101325.0f * 1e-10f * 60.1f);
    const float a1 = c1p * 101325.0f * 1e-10f * 0.972f;
    const float c2 = 0.00366f;
    const S num = (a0 - a1 * t_in) * (1.0f + c2 * t_out);
    const S den = (a0 - a1 * t_out) * (1.0f + c2 * t_in);
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

// refraction, src/main_view.wgsl:119-136.  Restated without materialising n_ref = +-n or |cosi|: negation is exact in
// IEEE arithmetic and round-to-nearest is sign-symmetric, so with sk = sqrt(k) and cosi the clamped dot product
//     cosi <  0 (n_ref =  n):   (eta |cosi| - sk) n_ref  ==  (-sk - eta cosi) n      bit for bit
//     cosi >= 0 (n_ref = -n):   (eta |cosi| - sk) n_ref  ==  ( sk - eta cosi) n      bit for bit   (also when contracted to FMA),
//     reflect(v_in, n_ref) = v_in - 2 dot(n_ref, v_in) n_ref  ==  v_in - 2 dot(n, v_in) n            bit for bit
// (tests/test_translator.py and tests/test_gpu_units.py compare against the oracle's literal transcription).
// Total reflection (k < 0) is rare, so it is a real branch instead of seven predicated instructions per evaluation.
LINON_HD vec3f refraction_s(float t_in, float t_out, vec3f v_in, vec3f n) {
    const float d = w_dot(v_in, n);
    const float cosi = w_clamp(-1.0f, 1.0f, d);     // misordered in the reference: min(max(-1, 1), d) = min(1, d)
    const float eta = refraction_eta(t_in, t_out);  // eta_in / eta_out, src/main_view.wgsl:120-121,129
    const float k = 1.0f - eta * eta * (1.0f - cosi * cosi);
    if (__builtin_expect(k < 0.0f, 0)) {
        return linon_total_reflection(v_in, n, d);  // reflect(v_in, n_ref); dot(n_ref, v_in) n_ref = dot(n, v_in) n
    }
    const float sk = w_sqrt(k);
    const float cs = ((cosi < 0.0f) ? -sk : sk) - eta * cosi;
    return eta * v_in + cs * n;
}
// The same arithmetic on two rays at once; the rare total-reflection component is patched by a select.
LINON_HD vec3t<f32x2> refraction_s(f32x2 t_in, f32x2 t_out, vec3t<f32x2> v_in, vec3t<f32x2> n) {
    const f32x2 d = w_dot(v_in, n);
    const f32x2 cosi = w_clamp(-1.0f, 1.0f, d);
    const f32x2 eta = refraction_eta(t_in, t_out);
    const f32x2 k = 1.0f - eta * eta * (1.0f - cosi * cosi);
    const f32x2 sk = w_sqrt(k);  // NaN in a totally reflected component, replaced below
    const f32x2 cs = f32x2((cosi.lo < 0.0f) ? -sk.lo : sk.lo, (cosi.hi < 0.0f) ? -sk.hi : sk.hi) - eta * cosi;
    vec3t<f32x2> r = eta * v_in + cs * n;
    const bool2 total = k < 0.0f;
    if (__builtin_expect(s_any(total), 0)) {
        const vec3t<f32x2> refl = v_in - (2.0f * d) * n;
        r = w_select(r, refl, total);
    }
    return r;
}
template <class A, class B, class C, class D, LINON_REQ(A, B)>
LINON_HD vec3t<P<A, B, C, D>> refraction(A t_in, B t_out, vec3t<C> v_in, vec3t<D> n) {
    typedef P<A, B, C, D> S;
    return refraction_s(S(t_in), S(t_out), vec3t<S>(v_in), vec3t<S>(n));
}
template <class A, class B, class C> LINON_HD P<A, B, C> point_plane_distance(vec3t<A> p, vec3t<B> n, vec3t<C> p0) {
    typedef P<A, B, C> S;
    const S d = w_dot(p0, n);
    return w_abs(S(w_dot(p, n)) - d);
}
template <class A, LINON_REQ(A)> LINON_HD P<A> sigmoid(A x) { typedef P<A> S; return 1.0f / (1.0f + w_exp(-S(x))); }
