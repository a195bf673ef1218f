// wgsl_runtime.hpp — C++17 run-time for C++ that oracle/wgsl_ref/wgsl2cpp.py emits from the reference's WGSL.
//
// TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Nothing under linon_b200/ may include or link this.
//
// The control flow, expression order and literals of the compiled code come MECHANICALLY from the reference's own
// shader text (/root/reference/src/main_view.wgsl, src/overlay.wgsl, bodies of src/functions.rs).  What is defined by
// hand is only this file: the WGSL types and builtin functions, which in the reference live in naga 22.1.0 and the
// Vulkan driver compiler (Cargo.lock:1678-1681; not vendored).  They follow the WGSL specification's definitions:
//   dot(a,b)      = a.x*b.x + a.y*b.y (+ a.z*b.z ...)      left to right
//   length(v)     = sqrt(dot(v,v))            normalize(v) = v / length(v)
//   cross(a,b)    = (a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x)
//   reflect(e1,e2)= e1 - 2*dot(e2,e1)*e2      clamp(e,lo,hi) = min(max(e,lo),hi)
//   mat * vec     = sum over columns of column*component, left to right
//   rgba8unorm store: clamp to [0,1], *255, round to nearest; load: byte / 255
// fp32 throughout, no FMA contraction (build with -ffp-contract=off); exp/sin/cos/pow from glibc libm.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <type_traits>

namespace wgsl {

template <class T> struct identity { using type = T; };
template <class T> using identity_t = typename identity<T>::type;  // keeps a scalar operand out of template deduction

using f32 = float;
using i32 = int32_t;
using u32 = uint32_t;

// ---- vectors -------------------------------------------------------------------------------------------------------
template <class T, int N> struct vec;
template <class T> struct vec<T, 2> {
    union { T x; T r; };
    union { T y; T g; };
    constexpr vec() : x(T(0)), y(T(0)) {}
    constexpr vec(T a, T b) : x(a), y(b) {}
    constexpr explicit vec(T s) : x(s), y(s) {}
    template <class U> constexpr explicit vec(const vec<U, 2>& o) : x(T(o.x)), y(T(o.y)) {}
    T& operator[](int i) { return i == 0 ? x : y; }
    const T& operator[](int i) const { return i == 0 ? x : y; }
};
template <class T> struct vec<T, 3> {
    union { T x; T r; };
    union { T y; T g; };
    union { T z; T b; };
    constexpr vec() : x(T(0)), y(T(0)), z(T(0)) {}
    constexpr vec(T a, T b_, T c) : x(a), y(b_), z(c) {}
    constexpr explicit vec(T s) : x(s), y(s), z(s) {}
    constexpr vec(const vec<T, 2>& a, T c) : x(a.x), y(a.y), z(c) {}
    constexpr vec(T a, const vec<T, 2>& bc) : x(a), y(bc.x), z(bc.y) {}
    template <class U> constexpr explicit vec(const vec<U, 3>& o) : x(T(o.x)), y(T(o.y)), z(T(o.z)) {}
    T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    const T& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
template <class T> struct vec<T, 4> {
    union { T x; T r; };
    union { T y; T g; };
    union { T z; T b; };
    union { T w; T a; };
    constexpr vec() : x(T(0)), y(T(0)), z(T(0)), w(T(0)) {}
    constexpr vec(T a_, T b_, T c, T d) : x(a_), y(b_), z(c), w(d) {}
    constexpr explicit vec(T s) : x(s), y(s), z(s), w(s) {}
    constexpr vec(const vec<T, 3>& v, T d) : x(v.x), y(v.y), z(v.z), w(d) {}
    constexpr vec(T a_, const vec<T, 3>& v) : x(a_), y(v.x), z(v.y), w(v.z) {}
    constexpr vec(const vec<T, 2>& p, const vec<T, 2>& q) : x(p.x), y(p.y), z(q.x), w(q.y) {}
    constexpr vec(const vec<T, 2>& p, T c, T d) : x(p.x), y(p.y), z(c), w(d) {}
    template <class U> constexpr explicit vec(const vec<U, 4>& o) : x(T(o.x)), y(T(o.y)), z(T(o.z)), w(T(o.w)) {}
    T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    const T& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
};

// swizzles with two to four components (value only; the reference never assigns through one)
template <int A, int B, class T, int N> constexpr vec<T, 2> swz(const vec<T, N>& v) { return vec<T, 2>(v[A], v[B]); }
template <int A, int B, int C, class T, int N> constexpr vec<T, 3> swz(const vec<T, N>& v) { return vec<T, 3>(v[A], v[B], v[C]); }
template <int A, int B, int C, int D, class T, int N> constexpr vec<T, 4> swz(const vec<T, N>& v) {
    return vec<T, 4>(v[A], v[B], v[C], v[D]);
}

#define WGSL_VEC_BINOP(OP)                                                                                        \
    template <class T, int N> inline vec<T, N> operator OP(const vec<T, N>& a, const vec<T, N>& b) {             \
        vec<T, N> r;                                                                                              \
        for (int i = 0; i < N; ++i) r[i] = a[i] OP b[i];                                                          \
        return r;                                                                                                 \
    }                                                                                                             \
    template <class T, int N> inline vec<T, N> operator OP(const vec<T, N>& a, identity_t<T> s) {                            \
        vec<T, N> r;                                                                                              \
        for (int i = 0; i < N; ++i) r[i] = a[i] OP s;                                                             \
        return r;                                                                                                 \
    }                                                                                                             \
    template <class T, int N> inline vec<T, N> operator OP(identity_t<T> s, const vec<T, N>& b) {                            \
        vec<T, N> r;                                                                                              \
        for (int i = 0; i < N; ++i) r[i] = s OP b[i];                                                             \
        return r;                                                                                                 \
    }
WGSL_VEC_BINOP(+)
WGSL_VEC_BINOP(-)
WGSL_VEC_BINOP(*)
WGSL_VEC_BINOP(/)
#undef WGSL_VEC_BINOP
template <class T, int N> inline vec<T, N> operator-(const vec<T, N>& a) {
    vec<T, N> r;
    for (int i = 0; i < N; ++i) r[i] = -a[i];
    return r;
}
#define WGSL_VEC_CMP(OP)                                                                                          \
    template <class T, int N> inline vec<bool, N> operator OP(const vec<T, N>& a, const vec<T, N>& b) {          \
        vec<bool, N> r;                                                                                           \
        for (int i = 0; i < N; ++i) r[i] = a[i] OP b[i];                                                          \
        return r;                                                                                                 \
    }
WGSL_VEC_CMP(<)
WGSL_VEC_CMP(>)
WGSL_VEC_CMP(<=)
WGSL_VEC_CMP(>=)
#undef WGSL_VEC_CMP

// ---- scalar builtins (float = the shader's f32; double = abstract-float constant evaluation) ------------------------
inline float abs(float x) { return fabsf(x); }
inline double abs(double x) { return fabs(x); }
inline i32 abs(i32 x) { return x < 0 ? -x : x; }
inline float sqrt(float x) { return sqrtf(x); }
inline double sqrt(double x) { return ::sqrt(x); }
inline float exp(float x) { return expf(x); }
inline double exp(double x) { return ::exp(x); }
inline float exp2(float x) { return exp2f(x); }
inline double exp2(double x) { return ::exp2(x); }
inline float log(float x) { return logf(x); }
inline double log(double x) { return ::log(x); }
inline float log2(float x) { return log2f(x); }
inline double log2(double x) { return ::log2(x); }
inline float sin(float x) { return sinf(x); }
inline double sin(double x) { return ::sin(x); }
inline float cos(float x) { return cosf(x); }
inline double cos(double x) { return ::cos(x); }
inline float tan(float x) { return tanf(x); }
inline double tan(double x) { return ::tan(x); }
inline float asin(float x) { return asinf(x); }
inline float acos(float x) { return acosf(x); }
inline float atan(float x) { return atanf(x); }
inline float atan2(float y, float x) { return atan2f(y, x); }
inline float sinh(float x) { return sinhf(x); }
inline float cosh(float x) { return coshf(x); }
inline float tanh(float x) { return tanhf(x); }
inline float pow(float x, float y) { return powf(x, y); }
inline double pow(double x, double y) { return ::pow(x, y); }
inline float floor(float x) { return floorf(x); }
inline double floor(double x) { return ::floor(x); }
inline float ceil(float x) { return ceilf(x); }
inline float round(float x) { return nearbyintf(x); }  // WGSL: ties to even
inline float trunc(float x) { return truncf(x); }
inline float fract(float x) { return x - floorf(x); }
inline float sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
inline float inverseSqrt(float x) { return 1.0f / sqrtf(x); }
// scalar min/max/clamp accept an abstract-int literal beside a concrete operand (common type = the concrete one)
template <class A, class B, class T = typename std::common_type<A, B>::type,
          class = typename std::enable_if<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>::type>
inline T min(A a_, B b_) { const T a = T(a_), b = T(b_); return (b < a) ? b : a; }
template <class A, class B, class T = typename std::common_type<A, B>::type,
          class = typename std::enable_if<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>::type>
inline T max(A a_, B b_) { const T a = T(a_), b = T(b_); return (a < b) ? b : a; }
template <class A, class B, class D, class T = typename std::common_type<A, B, D>::type,
          class = typename std::enable_if<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value && std::is_arithmetic<D>::value>::type>
inline T clamp(A e, B lo, D hi) { return min(max(T(e), T(lo)), T(hi)); }
inline float saturate(float x) { return clamp(x, 0.0f, 1.0f); }
inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline float step(float edge, float x) { return edge <= x ? 1.0f : 0.0f; }
inline float smoothstep(float lo, float hi, float x) {
    const float t = clamp((x - lo) / (hi - lo), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
inline float fma(float a, float b, float c) { return a * b + c; }  // WGSL allows either; unfused like every other product
template <class A, class B, class T = typename std::common_type<A, B>::type,
          class = typename std::enable_if<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>::type>
inline T select(A f, B t, bool c) { return c ? T(t) : T(f); }
inline float degrees(float x) { return x * 57.295779513082322865f; }
inline float radians(float x) { return x * 0.017453292519943295474f; }

// ---- component-wise builtins on vectors -----------------------------------------------------------------------------
#define WGSL_VEC_MAP1(NAME)                                                                                       \
    template <class T, int N> inline vec<T, N> NAME(const vec<T, N>& a) {                                        \
        vec<T, N> r;                                                                                              \
        for (int i = 0; i < N; ++i) r[i] = NAME(a[i]);                                                            \
        return r;                                                                                                 \
    }
WGSL_VEC_MAP1(abs) WGSL_VEC_MAP1(sqrt) WGSL_VEC_MAP1(exp) WGSL_VEC_MAP1(exp2) WGSL_VEC_MAP1(log) WGSL_VEC_MAP1(log2)
WGSL_VEC_MAP1(sin) WGSL_VEC_MAP1(cos) WGSL_VEC_MAP1(tan) WGSL_VEC_MAP1(floor) WGSL_VEC_MAP1(ceil) WGSL_VEC_MAP1(fract)
WGSL_VEC_MAP1(sign) WGSL_VEC_MAP1(saturate) WGSL_VEC_MAP1(round) WGSL_VEC_MAP1(trunc) WGSL_VEC_MAP1(inverseSqrt)
#undef WGSL_VEC_MAP1
#define WGSL_VEC_MAP2(NAME)                                                                                       \
    template <class T, int N> inline vec<T, N> NAME(const vec<T, N>& a, const vec<T, N>& b) {                    \
        vec<T, N> r;                                                                                              \
        for (int i = 0; i < N; ++i) r[i] = NAME(a[i], b[i]);                                                      \
        return r;                                                                                                 \
    }
WGSL_VEC_MAP2(min) WGSL_VEC_MAP2(max) WGSL_VEC_MAP2(pow) WGSL_VEC_MAP2(step) WGSL_VEC_MAP2(atan2)
#undef WGSL_VEC_MAP2
template <class T, int N> inline vec<T, N> clamp(const vec<T, N>& e, const vec<T, N>& lo, const vec<T, N>& hi) {
    return min(max(e, lo), hi);
}
template <class T, int N> inline vec<T, N> mix(const vec<T, N>& a, const vec<T, N>& b, T t) {
    return a * (T(1) - t) + b * t;
}
template <class T, int N> inline vec<T, N> mix(const vec<T, N>& a, const vec<T, N>& b, const vec<T, N>& t) {
    vec<T, N> r;
    for (int i = 0; i < N; ++i) r[i] = mix(a[i], b[i], t[i]);
    return r;
}
template <class T, int N> inline vec<T, N> select(const vec<T, N>& f, const vec<T, N>& t, const vec<bool, N>& c) {
    vec<T, N> r;
    for (int i = 0; i < N; ++i) r[i] = c[i] ? t[i] : f[i];
    return r;
}
template <class T, int N> inline vec<T, N> select(const vec<T, N>& f, const vec<T, N>& t, bool c) { return c ? t : f; }
template <int N> inline bool any(const vec<bool, N>& a) {
    bool r = false;
    for (int i = 0; i < N; ++i) r = r || a[i];
    return r;
}
template <int N> inline bool all(const vec<bool, N>& a) {
    bool r = true;
    for (int i = 0; i < N; ++i) r = r && a[i];
    return r;
}

// ---- geometric builtins ----------------------------------------------------------------------------------------------
// e1 % e2 (WGSL spec 8.7): integers truncate like C++; floats are e1 - e2 * trunc(e1 / e2); vectors component-wise, a scalar
// operand is splatted.  C++ has no % on floats, so the transpiler emits wgsl::mod for every %.
inline i32 mod(i32 a, i32 b) { return a % b; }
inline u32 mod(u32 a, u32 b) { return a % b; }
inline long long mod(long long a, long long b) { return a % b; }  // AbstractInt constant evaluation
inline float mod(float a, float b) { return a - b * truncf(a / b); }
inline double mod(double a, double b) { return a - b * std::trunc(a / b); }  // AbstractFloat constant evaluation
template <class T, int N> inline vec<T, N> mod(const vec<T, N>& a, const vec<T, N>& b) {
    vec<T, N> r;
    for (int i = 0; i < N; ++i) r[i] = mod(a[i], b[i]);
    return r;
}
template <class T, int N> inline vec<T, N> mod(const vec<T, N>& a, identity_t<T> s) { return mod(a, vec<T, N>(s)); }
template <class T, int N> inline vec<T, N> mod(identity_t<T> s, const vec<T, N>& b) { return mod(vec<T, N>(s), b); }
template <class T, int N> inline T dot(const vec<T, N>& a, const vec<T, N>& b) {
    T s = a[0] * b[0];
    for (int i = 1; i < N; ++i) s = s + a[i] * b[i];
    return s;
}
template <class T> inline vec<T, 3> cross(const vec<T, 3>& a, const vec<T, 3>& b) {
    return vec<T, 3>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
template <class T, int N> inline T length(const vec<T, N>& a) { return sqrt(dot(a, a)); }
inline float length(float a) { return fabsf(a); }
template <class T, int N> inline T distance(const vec<T, N>& a, const vec<T, N>& b) { return length(a - b); }
template <class T, int N> inline vec<T, N> normalize(const vec<T, N>& a) { return a / length(a); }
template <class T, int N> inline vec<T, N> reflect(const vec<T, N>& e1, const vec<T, N>& e2) {
    return e1 - (T(2) * dot(e2, e1)) * e2;
}

// ---- matrices: C columns of R rows (matCxR) ----------------------------------------------------------------------------
template <int C, int R> struct mat {
    vec<float, R> col[C];
    mat() {}
    mat(const vec<float, R>& a, const vec<float, R>& b) : col{a, b} { static_assert(C == 2, "two columns"); }
    mat(const vec<float, R>& a, const vec<float, R>& b, const vec<float, R>& c) : col{a, b, c} {
        static_assert(C == 3, "three columns");
    }
    mat(const vec<float, R>& a, const vec<float, R>& b, const vec<float, R>& c, const vec<float, R>& d) : col{a, b, c, d} {
        static_assert(C == 4, "four columns");
    }
    vec<float, R>& operator[](int i) { return col[i]; }
    const vec<float, R>& operator[](int i) const { return col[i]; }
};
template <int C, int R> inline vec<float, R> operator*(const mat<C, R>& m, const vec<float, C>& v) {
    vec<float, R> r = m[0] * v[0];
    for (int c = 1; c < C; ++c) r = r + m[c] * v[c];
    return r;
}
template <int C, int R> inline vec<float, C> operator*(const vec<float, R>& v, const mat<C, R>& m) {
    vec<float, C> r;
    for (int c = 0; c < C; ++c) r[c] = dot(v, m[c]);
    return r;
}
template <int K, int R, int C> inline mat<C, R> operator*(const mat<K, R>& a, const mat<C, K>& b) {
    mat<C, R> r;
    for (int c = 0; c < C; ++c) r[c] = a * b[c];
    return r;
}
template <int C, int R> inline mat<R, C> transpose(const mat<C, R>& m) {
    mat<R, C> t;
    for (int c = 0; c < C; ++c)
        for (int r = 0; r < R; ++r) t[r][c] = m[c][r];
    return t;
}

// ---- arrays ---------------------------------------------------------------------------------------------------------------
template <class T, int N> struct array {
    T e[N];
    array() : e{} {}
    template <class... A> array(const T& first, const A&... rest) : e{first, T(rest)...} {}
    // out-of-range indices are clamped (one of the behaviours WGSL allows; none occurs on the tested path)
    T& operator[](long i) { return e[i < 0 ? 0 : (i >= N ? N - 1 : i)]; }
    const T& operator[](long i) const { return e[i < 0 ? 0 : (i >= N ? N - 1 : i)]; }
};
template <class T> struct rt_array {  // runtime-sized array of a storage buffer
    T* ptr = nullptr;
    u32 len = 0;
    T& operator[](long i) { return ptr[i < 0 ? 0 : (i >= (long)len ? (long)len - 1 : i)]; }
    const T& operator[](long i) const { return ptr[i < 0 ? 0 : (i >= (long)len ? (long)len - 1 : i)]; }
};
template <class T> inline u32 arrayLength(const rt_array<T>* a) { return a->len; }
template <class T> inline u32 arrayLength(rt_array<T>* a) { return a->len; }

// ---- textures ---------------------------------------------------------------------------------------------------------------
struct rgba8unorm {};
struct rgba32float {};
struct write {};
inline uint8_t to_unorm8(float x) {
    if (!(x > 0.0f)) return 0;  // also NaN
    if (x > 1.0f) x = 1.0f;
    return (uint8_t)lrintf(x * 255.0f);
}
// One texture object serves as the storage view of the writing pass and the sampled view of the reading pass.
struct texture {
    int width = 0, height = 0;
    uint8_t* unorm8 = nullptr;   // rgba8unorm texels (4 bytes)   -- exactly one of the two is set
    float* f32x4 = nullptr;      // rgba32float texels (16 bytes)
    float* shadow_f32x4 = nullptr;  // optional: the unquantised value stored into an rgba8unorm texture (diagnostic)
};
template <class Format, class Access> struct texture_storage_2d { texture* t = nullptr; };
template <class T> struct texture_2d { texture* t = nullptr; };
template <class F, class A> inline vec<u32, 2> textureDimensions(const texture_storage_2d<F, A>& s) {
    return vec<u32, 2>((u32)s.t->width, (u32)s.t->height);
}
template <class T> inline vec<u32, 2> textureDimensions(const texture_2d<T>& s) {
    return vec<u32, 2>((u32)s.t->width, (u32)s.t->height);
}
template <class F, class A> inline void textureStore(const texture_storage_2d<F, A>& s, const vec<i32, 2>& c, const vec<float, 4>& v) {
    texture* t = s.t;
    if (c.x < 0 || c.y < 0 || c.x >= t->width || c.y >= t->height) return;  // out-of-bounds stores are discarded
    const size_t i = ((size_t)c.y * t->width + c.x) * 4;
    if (t->unorm8) {
        for (int k = 0; k < 4; ++k) t->unorm8[i + k] = to_unorm8(v[k]);
        if (t->shadow_f32x4)
            for (int k = 0; k < 4; ++k) t->shadow_f32x4[i + k] = v[k];
    } else {
        for (int k = 0; k < 4; ++k) t->f32x4[i + k] = v[k];
    }
}
template <class F, class A> inline void textureStore(const texture_storage_2d<F, A>& s, const vec<u32, 2>& c, const vec<float, 4>& v) {
    textureStore(s, vec<i32, 2>((i32)c.x, (i32)c.y), v);
}
template <class T, class L> inline vec<float, 4> textureLoad(const texture_2d<T>& s, const vec<i32, 2>& c, L /*level*/) {
    const texture* t = s.t;
    if (c.x < 0 || c.y < 0 || c.x >= t->width || c.y >= t->height) return vec<float, 4>();  // out of bounds: zero
    const size_t i = ((size_t)c.y * t->width + c.x) * 4;
    if (t->unorm8) return vec<float, 4>(t->unorm8[i] / 255.0f, t->unorm8[i + 1] / 255.0f, t->unorm8[i + 2] / 255.0f, t->unorm8[i + 3] / 255.0f);
    return vec<float, 4>(t->f32x4[i], t->f32x4[i + 1], t->f32x4[i + 2], t->f32x4[i + 3]);
}

}  // namespace wgsl
