// linon_oracle.cpp — CPU restatement of linon's nonlinear ray-march hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under linon_b200/ (the product) may link,
// import or execute this file; only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py load it, as the checker.
//
// PARITY: the reference (niklaskorz/linon, Rust + wgpu) ships no tests, golden
// vectors or usable images, and neither rustc/cargo nor a Vulkan/lavapipe stack
// exists in this image, so its host cannot be executed here.  Its shader text can:
// oracle/_ref (main_view.wgsl + functions.rs bodies + overlay.wgsl, transpiled
// mechanically to C++ by oracle/wgsl_ref and compiled with gcc) is what pins this
// file - the two are bit-for-bit identical on every output the shader has
// (tests/test_wgsl_reference_cpu.py).  Still unpinned: the last-ulp behaviour of
// naga's / the Vulkan driver's builtins (see below).
// This file follows the reference WGSL line by line (citations below), fp32,
// no FMA contraction (build with -ffp-contract=off), expression order as
// written (left-associative), WGSL builtin definitions:
//   clamp(e,lo,hi) = min(max(e,lo),hi)      normalize(v) = v / length(v)
//   length(v) = sqrt(dot(v,v))              reflect(a,n) = a - 2*dot(n,a)*n
//   dot = x*x' + y*y' + z*z' (left to right)
// The arithmetic of the builtins really lives in naga 22.1.0 + the Vulkan
// driver compiler (Cargo.lock:1678-1681, 3322-3375), which are not vendored
// under /root/reference; WGSL only bounds them by ULP tolerances.
// A second, independent numpy restatement (oracle/np_oracle.py) cross-checks
// this one in tests/test_oracle_selfcheck.py.
//
// Reference lines followed:
//   src/main_view.wgsl:65-76    constants            -> consts below
//   src/main_view.wgsl:78-146   helper library       -> rotateX..sigmoid
//   src/main_view.wgsl:150-175  hit_triangle         -> hit_triangle
//   src/main_view.wgsl:177-220  ray_color            -> ray_color
//   src/main_view.wgsl:229-279  nonlinear_ray_color  -> nonlinear_ray_color
//   src/main_view.wgsl:281-314  sample_rays          -> sample_rays
//   src/main_view.wgsl:316-335  sample tables        -> SAMPLE_POSITIONS/COLORS
//   src/main_view.wgsl:339-375  main_view entry      -> march_pixel
//   src/main_view.wgsl:381-479  sample-ray origins   -> sample_ray_positions
//       (outline search restated with the INTENDED stride W)
//   src/overlay.wgsl:4-40,99-135 Lyapunov overlay    -> overlay_pixel
//   src/functions.rs:36-126     nine field bodies    -> field_* (hand transcribed)
//   src/cornell_box.rs, src/vertices.rs:3-57, src/application.rs:105-123
//                                                    -> oracle_default_scene
//   src/arcball.rs:33-129, src/main_view.rs:21-31,153-155 -> Arcball
//
// Templated on the scalar type so the same restatement can be run in fp64 to
// separate integration sensitivity from implementation error.

#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ---- scalar math shims -------------------------------------------------------
inline float  m_sqrt(float x)  { return sqrtf(x); }
inline double m_sqrt(double x) { return sqrt(x); }
inline float  m_exp(float x)   { return expf(x); }
inline double m_exp(double x)  { return exp(x); }
inline float  m_sin(float x)   { return sinf(x); }
inline double m_sin(double x)  { return sin(x); }
inline float  m_cos(float x)   { return cosf(x); }
inline double m_cos(double x)  { return cos(x); }
inline float  m_pow(float x, float y)   { return powf(x, y); }
inline double m_pow(double x, double y) { return pow(x, y); }
inline float  m_abs(float x)   { return fabsf(x); }
inline double m_abs(double x)  { return fabs(x); }

// WGSL min/max: "if one operand is NaN the other is returned" is allowed; use
// the fmin/fmax behaviour (matches SPIR-V FMin/FMax as usually lowered).
template <class F> inline F m_min(F a, F b) { return (b < a) ? b : a; }
template <class F> inline F m_max(F a, F b) { return (a < b) ? b : a; }
template <class F> inline F m_clamp(F e, F lo, F hi) { return m_min(m_max(e, lo), hi); }

template <class F> struct V3 {
    F x, y, z;
    V3() : x(0), y(0), z(0) {}
    V3(F a, F b, F c) : x(a), y(b), z(c) {}
};
template <class F> inline V3<F> operator+(V3<F> a, V3<F> b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <class F> inline V3<F> operator-(V3<F> a, V3<F> b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <class F> inline V3<F> operator-(V3<F> a) { return {-a.x, -a.y, -a.z}; }
template <class F> inline V3<F> operator*(F s, V3<F> a) { return {s * a.x, s * a.y, s * a.z}; }
template <class F> inline V3<F> operator*(V3<F> a, F s) { return {a.x * s, a.y * s, a.z * s}; }
template <class F> inline V3<F> operator*(V3<F> a, V3<F> b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
template <class F> inline V3<F> operator/(V3<F> a, F s) { return {a.x / s, a.y / s, a.z / s}; }
template <class F> inline F dot(V3<F> a, V3<F> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class F> inline V3<F> cross(V3<F> a, V3<F> b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
template <class F> inline F length(V3<F> a) { return m_sqrt(dot(a, a)); }
template <class F> inline V3<F> normalize(V3<F> a) { return a / length(a); }
template <class F> inline V3<F> vabs(V3<F> a) { return {m_abs(a.x), m_abs(a.y), m_abs(a.z)}; }
template <class F> inline V3<F> reflect(V3<F> e1, V3<F> e2) { return e1 - (F(2.0) * dot(e2, e1)) * e2; }

// mat3x3(c0,c1,c2) * v = c0*v.x + c1*v.y + c2*v.z   (column major, WGSL spec)
template <class F> inline V3<F> mat3_mul(V3<F> c0, V3<F> c1, V3<F> c2, V3<F> v) {
    return c0 * v.x + c1 * v.y + c2 * v.z;
}

// ---- src/main_view.wgsl:65-76 -------------------------------------------------
constexpr bool backface_culling = false;
constexpr bool use_lighting = true;
template <class F> constexpr F EPS() { return F(0.0000001f); }
template <class F> constexpr F PI_C() { return F(3.141592653589793f); }  // const PI: f32

// ---- src/main_view.wgsl:78-108 ------------------------------------------------
template <class F> V3<F> rotateX(V3<F> v, F phi) {
    return mat3_mul<F>({F(1), F(0), F(0)}, {F(0), m_cos(phi), m_sin(phi)}, {F(0), -m_sin(phi), m_cos(phi)}, v);
}
template <class F> V3<F> rotateY(V3<F> v, F phi) {
    return mat3_mul<F>({m_cos(phi), F(0), -m_sin(phi)}, {F(0), F(1), F(0)}, {m_sin(phi), F(0), m_cos(phi)}, v);
}
// third column is zero in the reference: z is dropped.
template <class F> V3<F> rotateZ(V3<F> v, F phi) {
    return mat3_mul<F>({m_cos(phi), m_sin(phi), F(0)}, {-m_sin(phi), m_cos(phi), F(0)}, {F(0), F(0), F(0)}, v);
}
template <class F> V3<F> translate(V3<F> v, F dx, F dy, F dz) { return {v.x + dx, v.y + dy, v.z + dz}; }

// ---- src/main_view.wgsl:111-117 -------------------------------------------------
// pow(10.0, -10.0) is a literal-only const-expression: naga folds it in abstract
// float (f64) and concretises to f32.
template <class F> F refraction_index(F t) {
    const F air_pressure = F(101325.0f);
    const F c1 = F(0.0000104f);
    const F c2 = F(0.00366f);
    const F p10 = F((float)1e-10);
    return c1 * air_pressure * (F(1.0) + air_pressure * (F(60.1f) - F(0.972f) * t) * p10) / (F(1.0) + c2 * t);
}

// ---- src/main_view.wgsl:119-136 (clamp arguments misordered) ---------
template <class F> V3<F> refraction(F t_in, F t_out, V3<F> v_in, V3<F> n) {
    const F eta_in = refraction_index(t_in);
    const F eta_out = refraction_index(t_out);
    F cosi = m_clamp(F(-1.0), F(1.0), dot(v_in, n));  // = min(max(-1, 1), dot) = min(1, dot)
    V3<F> n_ref = n;
    if (cosi < F(0.0)) {
        cosi = -cosi;
    } else {
        n_ref = -n;
    }
    const F eta = eta_in / eta_out;
    const F k = F(1.0) - eta * eta * (F(1.0) - cosi * cosi);
    if (k < F(0.0)) {
        return reflect(v_in, n_ref);
    }
    return eta * v_in + (eta * cosi - m_sqrt(k)) * n_ref;
}

// ---- src/main_view.wgsl:138-146 --------------------------------------------------
template <class F> F point_plane_distance(V3<F> p, V3<F> n, V3<F> p0) {
    const F d = dot(p0, n);
    return m_abs(dot(p, n) - d);
}
template <class F> F sigmoid(F x) { return F(1.0) / (F(1.0) + m_exp(-x)); }

// ---- src/functions.rs:36-126, hand transcribed -----------------------------------
struct SettingsPOD {  // src/main_view.rs:43-51
    float field_weight;
    float mouse_x, mouse_y;
    int32_t overlay_mode;
    int32_t central_difference_delta;
    float lyapunov_scaling;
};

template <class F> V3<F> field_mirage_spherical(V3<F> p_prev, V3<F> p, V3<F>, V3<F> v, F) {
    const F t_env = F(15.0), t_src = F(200.0), max_dist = F(0.25);
    const V3<F> center(F(-0.5), F(0.5), F(-0.5));
    const V3<F> center_dest = p - center;
    const V3<F> normal = normalize(center_dest);
    const F dist_in = length(p_prev - center);
    const F dist_out = length(center_dest);
    const F part_in = m_clamp(F(0.0), F(1.0), dist_in / max_dist);
    const F part_out = m_clamp(F(0.0), F(1.0), dist_out / max_dist);
    const F t_in = part_in * t_env + (F(1.0) - part_in) * t_src;
    const F t_out = part_out * t_env + (F(1.0) - part_out) * t_src;
    return refraction(t_in, t_out, v, normal);
}
template <class F> V3<F> field_mirage_plane(V3<F> p_prev, V3<F> p, V3<F>, V3<F> v, F) {
    const F t_env = F(15.0), t_src = F(30.0), max_dist = F(0.01f);
    const V3<F> plane_p0(F(0.0), F(0.1f), F(0.0));
    const V3<F> plane_n(F(0.0), F(1.0), F(0.0));
    const F dist_in = point_plane_distance(p_prev, plane_n, plane_p0);
    const F dist_out = point_plane_distance(p, plane_n, plane_p0);
    const F part_in = m_clamp(F(0.0), F(1.0), dist_in / max_dist);
    const F part_out = m_clamp(F(0.0), F(1.0), dist_out / max_dist);
    const F t_in = part_in * t_env + (F(1.0) - part_in) * t_src;
    const F t_out = part_out * t_env + (F(1.0) - part_out) * t_src;
    return refraction(t_in, t_out, v, plane_n);
}
template <class F> V3<F> field_mirage_spherical_sigmoid(V3<F> p_prev, V3<F> p, V3<F>, V3<F> v, F) {
    const F t_env = F(15.0), t_src = F(200.0), max_dist = F(0.25);
    const V3<F> center(F(-0.5), F(0.5), F(-0.5));
    const V3<F> center_dest = p - center;
    const V3<F> normal = normalize(center_dest);
    const F dist_in = length(p_prev - center);
    const F dist_out = length(center_dest);
    const F part_in = m_clamp(F(0.0), F(1.0), sigmoid(dist_in / max_dist * F(12.0) - F(6.0)));
    const F part_out = m_clamp(F(0.0), F(1.0), sigmoid(dist_out / max_dist * F(12.0) - F(6.0)));
    const F t_in = part_in * t_env + (F(1.0) - part_in) * t_src;
    const F t_out = part_out * t_env + (F(1.0) - part_out) * t_src;
    return refraction(t_in, t_out, v, normal);
}
template <class F> V3<F> field_mirage_plane_sigmoid(V3<F> p_prev, V3<F> p, V3<F>, V3<F> v, F) {
    const F t_env = F(15.0), t_src = F(30.0), max_dist = F(0.01f);
    const V3<F> plane_p0(F(0.0), F(0.1f), F(0.0));
    const V3<F> plane_n(F(0.0), F(1.0), F(0.0));
    const F dist_in = point_plane_distance(p_prev, plane_n, plane_p0);
    const F dist_out = point_plane_distance(p, plane_n, plane_p0);
    const F part_in = m_clamp(F(0.0), F(1.0), sigmoid(dist_in / max_dist * F(12.0) - F(6.0)));
    const F part_out = m_clamp(F(0.0), F(1.0), sigmoid(dist_out / max_dist * F(12.0) - F(6.0)));
    const F t_in = part_in * t_env + (F(1.0) - part_in) * t_src;
    const F t_out = part_out * t_env + (F(1.0) - part_out) * t_src;
    return refraction(t_in, t_out, v, plane_n);
}
template <class F> V3<F> field_translation_x(V3<F>, V3<F>, V3<F> v0, V3<F>, F t) {
    return translate(v0, F(0.5) * t, F(0.0), F(0.0));
}
template <class F> V3<F> field_translation_z(V3<F>, V3<F>, V3<F> v0, V3<F>, F t) {
    return translate(v0, F(0.0), F(0.0), F(0.5) * t);
}
template <class F> V3<F> field_rotation(V3<F>, V3<F>, V3<F> v0, V3<F>, F t) {
    return normalize(v0 + rotateZ(v0, PI_C<F>() * t));
}
template <class F> V3<F> field_lorenz(V3<F>, V3<F> p, V3<F>, V3<F>, F) {
    const F rho = F(28.0), sigma = F(10.0);
    const F beta = F((float)(8.0 / 3.0));  // literal-only: folded in abstract float, then f32
    return {sigma * (p.y - p.x), p.x * (rho - p.z) - p.y, p.x * p.y - beta * p.z};
}
template <class F> V3<F> field_roessler(V3<F>, V3<F> p, V3<F>, V3<F>, F) {
    const F a = F(0.1f), b = F(0.1f), c = F(14.0);
    return {-p.y - p.z, p.x + a * p.y, b + p.z * (p.x - c)};
}

// Custom body hook (tests compile the product translator's output for the host and
// plug it in here so user-edited bodies have an oracle too).  fp32 only.
typedef void (*custom_field_fn)(const float* p_prev, const float* p, const float* v0, const float* v, float t,
                                const SettingsPOD* settings, float* out3);
custom_field_fn g_custom = nullptr;

template <class F>
V3<F> field_function(int id, const SettingsPOD& st, V3<F> p_prev, V3<F> p, V3<F> v0, V3<F> v, F t) {
    switch (id) {
        case 1: return field_mirage_spherical(p_prev, p, v0, v, t);
        case 2: return field_mirage_plane(p_prev, p, v0, v, t);
        case 3: return field_mirage_spherical_sigmoid(p_prev, p, v0, v, t);
        case 4: return field_mirage_plane_sigmoid(p_prev, p, v0, v, t);
        case 5: return field_translation_x(p_prev, p, v0, v, t);
        case 6: return field_translation_z(p_prev, p, v0, v, t);
        case 7: return field_rotation(p_prev, p, v0, v, t);
        case 8: return field_lorenz(p_prev, p, v0, v, t);
        case 9: return field_roessler(p_prev, p, v0, v, t);
        case 0:
            if (g_custom) {
                float a[3] = {(float)p_prev.x, (float)p_prev.y, (float)p_prev.z};
                float b[3] = {(float)p.x, (float)p.y, (float)p.z};
                float c[3] = {(float)v0.x, (float)v0.y, (float)v0.z};
                float d[3] = {(float)v.x, (float)v.y, (float)v.z};
                float o[3];
                g_custom(a, b, c, d, (float)t, &st, o);
                return {F(o[0]), F(o[1]), F(o[2])};
            }
            return v;  // placeholder body, src/main_view.wgsl:148
        default: return v;
    }
}

// ---- mesh -------------------------------------------------------------------------
struct Mesh {
    const float* verts;     // 3 floats / vertex, src/main_view.wgsl:33-42
    const uint32_t* faces;  // 3 u32 / face,      src/main_view.wgsl:44-53
    int n_faces;
};

// ---- src/main_view.wgsl:150-175 -----------------------------------------------------
template <class F> F hit_triangle(const V3<F> v_in[3], V3<F> origin, V3<F> direction) {
    const F eps = EPS<F>();
    const V3<F> edge1 = v_in[1] - v_in[0];
    const V3<F> edge2 = v_in[2] - v_in[0];
    const V3<F> h = cross(direction, edge2);
    const F a = dot(edge1, h);
    if ((backface_culling || a > -eps) && a < eps) return F(-1.0);
    const F f = F(1.0) / a;
    const V3<F> s = origin - v_in[0];
    const F u = f * dot(s, h);
    if (u < F(0.0) || u > F(1.0)) return F(-1.0);
    const V3<F> q = cross(s, edge1);
    const F v = f * dot(direction, q);
    if (v < F(0.0) || u + v > F(1.0)) return F(-1.0);
    const F t = f * dot(edge2, q);
    if (t > eps) return t;
    return F(-1.0);
}

template <class F> struct Color4 { F r, g, b, a; };

// ---- src/main_view.wgsl:177-220; hit id per SURVEY F10 --------------------------------
template <class F> Color4<F> ray_color(const Mesh& mesh, V3<F> origin, V3<F> direction, F max_dist, int* hit_id) {
    F t = F(-1.0);
    F t_new;
    V3<F> d1, d2;
    int id = -1;
    for (int i = 0; i < mesh.n_faces; ++i) {
        const uint32_t fa = mesh.faces[3 * i], fb = mesh.faces[3 * i + 1], fc = mesh.faces[3 * i + 2];
        const V3<F> tri[3] = {
            V3<F>(F(mesh.verts[3 * fa]), F(mesh.verts[3 * fa + 1]), F(mesh.verts[3 * fa + 2])),
            V3<F>(F(mesh.verts[3 * fb]), F(mesh.verts[3 * fb + 1]), F(mesh.verts[3 * fb + 2])),
            V3<F>(F(mesh.verts[3 * fc]), F(mesh.verts[3 * fc + 1]), F(mesh.verts[3 * fc + 2])),
        };
        t_new = hit_triangle(tri, origin, direction);
        if (t_new > F(0.0) && t_new < max_dist && (t < F(0.0) || t_new < t)) {
            t = t_new;
            d1 = tri[1] - tri[0];
            d2 = tri[2] - tri[0];
            id = i;
        }
    }
    if (t > F(0.0) && t < max_dist) {
        const V3<F> normal = normalize(cross(d1, d2));
        *hit_id = id;
        if (!use_lighting) {
            const V3<F> c = vabs(normal);
            return {c.x, c.y, c.z, F(1.0)};
        }
        const V3<F> light_color(F(1.0), F(1.0), F(1.0));
        const V3<F> ambient = F(0.01f) * light_color;
        F intensity = m_max(dot(normal, -direction), F(0.0));
        if (intensity == F(0.0)) intensity = m_max(dot(-normal, -direction), F(0.0));
        const F diff = intensity;
        const V3<F> diffuse = diff * light_color;
        const F spec = m_pow(intensity, F(64.0));
        const V3<F> specular = spec * light_color;
        const V3<F> result = (ambient + diffuse + specular) * vabs(normal);
        return {result.x, result.y, result.z, t};
    }
    return {F(0), F(0), F(0), F(0)};
}

struct RayStats { int accepted, rejected, prehit; };

template <class F> struct MarchResult {
    Color4<F> color;
    V3<F> mapping_point;
    int hit_id;
    RayStats stats;
};

// ---- src/main_view.wgsl:229-279 ----------------------------------------------------------
template <class F>
MarchResult<F> nonlinear_ray_color(const Mesh& mesh, const SettingsPOD& st, int fn_id, F h_initial, bool adaptive_sampling,
                                   V3<F> start_point, V3<F> start_dir) {
    MarchResult<F> result;
    result.color = {F(0), F(0), F(0), F(0)};
    result.hit_id = -1;
    result.stats = {0, 0, 0};
    const F field_weight = F(st.field_weight);
    bool has_color = false;
    V3<F> cur_point = start_point;
    V3<F> cur_dir = start_dir;
    F t = F(0.0);
    V3<F> last_v(F(0), F(0), F(0));
    F last_diff = F(-1.0);
    F h = h_initial;

    for (; t <= F(5.0);) {
        const V3<F> k1 = field_function(fn_id, st, cur_point, cur_point, start_dir, cur_dir, t);
        const V3<F> k2 = field_function(fn_id, st, cur_point, cur_point + (F(0.5) * h) * k1, start_dir, k1, t + F(0.5) * h);
        const V3<F> k3 = field_function(fn_id, st, cur_point, cur_point + (F(0.5) * h) * k2, start_dir, k2, t + F(0.5) * h);
        const V3<F> k4 = field_function(fn_id, st, cur_point, cur_point + h * k3, start_dir, k3, t + h);
        const V3<F> v = (k1 + F(2.0) * k2 + F(2.0) * k3 + k4) / F(6.0);
        const F diff = length(v - last_v);
        if (adaptive_sampling && last_diff >= F(0.0) && diff > F(10.0) * last_diff && h > F(0.002f)) {
            h = F(0.001f);
            result.stats.rejected++;
            continue;
        }
        cur_dir = (F(1.0) - field_weight) * cur_dir + field_weight * v;

        const V3<F> step_dir = cur_dir * h;
        if (!has_color) {
            int id = -1;
            result.color = ray_color(mesh, cur_point, normalize(step_dir), length(step_dir), &id);
            has_color = result.color.a > F(0.0);
            if (has_color) result.hit_id = id;
            result.stats.prehit++;
        }

        cur_point = cur_point + step_dir;
        t = t + h;
        last_v = v;
        if (F(2.0) * diff < last_diff && h < h_initial) h = h_initial;
        last_diff = diff;
        result.stats.accepted++;
    }

    result.mapping_point = cur_point;
    result.color.a = F(1.0);
    if (!has_color) result.color = {F(0), F(0), F(0), F(1.0)};
    return result;
}

struct CameraPOD {  // src/main_view.rs:12-19 (112 bytes)
    float origin[4];
    float view_direction[4];
    float up[4];
    float view_matrix[16];
};

template <class F> struct CamBasis {
    V3<F> origin, view_direction, nh, nv;  // nh/nv = normalize(horizontal/vertical)
    F viewport_width, viewport_height, focal_length;
};

// ---- src/main_view.wgsl:347-361 --------------------------------------------------------------
template <class F> CamBasis<F> camera_basis(const CameraPOD& cam, int W, int H) {
    CamBasis<F> b;
    const F width = F((float)W), height = F((float)H);
    const F aspect_ratio = width / height;
    b.origin = {F(cam.origin[0]), F(cam.origin[1]), F(cam.origin[2])};
    b.view_direction = {F(cam.view_direction[0]), F(cam.view_direction[1]), F(cam.view_direction[2])};
    const V3<F> up(F(cam.up[0]), F(cam.up[1]), F(cam.up[2]));
    b.focal_length = F(0.035f);
    b.viewport_height = F(0.025f);
    b.viewport_width = aspect_ratio * b.viewport_height;
    const V3<F> w = -b.view_direction;
    const V3<F> horizontal = cross(up, w);
    const V3<F> vertical = cross(w, horizontal);
    b.nh = normalize(horizontal);
    b.nv = normalize(vertical);
    return b;
}

// ---- src/main_view.wgsl:363-366 (gid.y counts from the bottom row of the texture) ----------------
template <class F> V3<F> pixel_dir(const CamBasis<F>& b, int W, int H, int gid_x, int gid_y) {
    const F width = F((float)W), height = F((float)H);
    const F u = F((float)gid_x) / (width - F(1.0)) * b.viewport_width - F(0.5) * b.viewport_width;
    const F v = F((float)gid_y) / (height - F(1.0)) * b.viewport_height - F(0.5) * b.viewport_height;
    const V3<F> s = u * b.nh + v * b.nv + b.focal_length * b.view_direction;
    return normalize(s);
}

// unorm8 store: clamp to [0,1], scale by 255, round to nearest (NaN -> 0)
inline uint8_t to_unorm8(float x) {
    if (!(x > 0.0f)) return 0;
    if (x > 1.0f) x = 1.0f;
    return (uint8_t)lrintf(x * 255.0f);
}

const float SAMPLE_POSITIONS[8][2] = {  // src/main_view.wgsl:316-325
    {-1.0f, -1.0f}, {0.0f, -1.0f}, {1.0f, -1.0f}, {1.0f, 0.0f}, {1.0f, 1.0f}, {0.0f, 1.0f}, {-1.0f, 1.0f}, {-1.0f, 0.0f}};
const float SAMPLE_COLORS[8][3] = {  // src/main_view.wgsl:326-335
    {1.0f, 0.0f, 0.0f}, {0.5f, 0.5f, 0.0f}, {0.0f, 1.0f, 0.0f}, {0.5f, 1.0f, 0.0f},
    {1.0f, 1.0f, 0.0f}, {0.5f, 0.5f, 0.5f}, {0.0f, 0.0f, 1.0f}, {0.5f, 0.0f, 0.5f}};

// ---- src/main_view.wgsl:281-314 ---------------------------------------------------------------------
template <class F>
void sample_rays(const SettingsPOD& st, int fn_id, V3<F> start_point, V3<F> start_dir, int samples_index,
                 const float* sample_color, float* ray_samples /* 800 x 8 floats */) {
    const F field_weight = F(st.field_weight);
    V3<F> cur_point = start_point;
    V3<F> cur_dir = start_dir;
    F t = F(0.0);
    const F h = F(0.001f);
    const int steps = 5000;
    const int sample_steps = 100;
    const int sample_step_size = steps / sample_steps;
    for (int i = 0; i < steps; ++i) {
        const V3<F> k1 = field_function(fn_id, st, cur_point, cur_point, start_dir, cur_dir, t);
        const V3<F> k2 = field_function(fn_id, st, cur_point, cur_point + (F(0.5) * h) * k1, start_dir, k1, t + F(0.5) * h);
        const V3<F> k3 = field_function(fn_id, st, cur_point, cur_point + (F(0.5) * h) * k2, start_dir, k2, t + F(0.5) * h);
        const V3<F> k4 = field_function(fn_id, st, cur_point, cur_point + h * k3, start_dir, k3, t + h);
        const V3<F> v = (k1 + F(2.0) * k2 + F(2.0) * k3 + k4) / F(6.0);
        cur_dir = (F(1.0) - field_weight) * cur_dir + field_weight * v;
        const V3<F> step_dir = cur_dir * h;
        if (i % sample_step_size == 0) {
            const int index = samples_index * sample_steps + i / sample_step_size;
            float* s = ray_samples + 8 * index;
            s[0] = (float)cur_point.x; s[1] = (float)cur_point.y; s[2] = (float)cur_point.z; s[3] = 1.0f;
            s[4] = sample_color[0]; s[5] = sample_color[1]; s[6] = sample_color[2]; s[7] = 0.5f;
        }
        cur_point = cur_point + step_dir;
        t = t + h;
    }
}

// ---- outline search, src/main_view.wgsl:381-468 with the intended stride W and the
//      intended start row H-1 for the bottom-middle scan (the reference starts at size.x-1). --------
void outline_positions(const float* exponents, int W, int H, float pos[8][2]) {
    const float min_exp = 0.8f;
    for (int k = 0; k < 8; ++k) { pos[k][0] = 0.0f; pos[k][1] = 0.0f; }
    auto E = [&](int x, int y) { return exponents[(size_t)y * W + x]; };
    {  // 0: bottom left
        int sum = 0;
        for (int x = 0; x < W; ++x) for (int y = 0; y < H; ++y)
            if (E(x, y) >= min_exp && (W - x) + y >= sum) { pos[0][0] = (float)x; pos[0][1] = (float)y; sum = (W - x) + y; }
    }
    {  // 1: bottom middle
        bool found = false;
        for (int y = H - 1; y >= 0 && !found; --y) for (int x = 0; x < W && !found; ++x)
            if (E(x, y) >= min_exp) { pos[1][0] = (float)x; pos[1][1] = (float)y; found = true; }
    }
    {  // 2: bottom right
        int sum = 0;
        for (int x = 0; x < W; ++x) for (int y = 0; y < H; ++y)
            if (E(x, y) >= min_exp && x + y >= sum) { pos[2][0] = (float)x; pos[2][1] = (float)y; sum = x + y; }
    }
    {  // 3: middle right
        bool found = false;
        for (int x = W - 1; x >= 0 && !found; --x) for (int y = 0; y < H && !found; ++y)
            if (E(x, y) >= min_exp) { pos[3][0] = (float)x; pos[3][1] = (float)y; found = true; }
    }
    {  // 4: top right
        int sum = 0;
        for (int x = 0; x < W; ++x) for (int y = 0; y < H; ++y)
            if (E(x, y) >= min_exp && x + (H - y) >= sum) { pos[4][0] = (float)x; pos[4][1] = (float)y; sum = x + (H - y); }
    }
    {  // 5: top middle
        bool found = false;
        for (int y = 0; y < H && !found; ++y) for (int x = 0; x < W && !found; ++x)
            if (E(x, y) >= min_exp) { pos[5][0] = (float)x; pos[5][1] = (float)y; found = true; }
    }
    {  // 6: top left
        int sum = 0;
        for (int x = 0; x < W; ++x) for (int y = 0; y < H; ++y)
            if (E(x, y) >= min_exp && (W - x) + (H - y) >= sum) { pos[6][0] = (float)x; pos[6][1] = (float)y; sum = (W - x) + (H - y); }
    }
    {  // 7: middle left
        bool found = false;
        for (int x = 0; x < W && !found; ++x) for (int y = 0; y < H && !found; ++y)
            if (E(x, y) >= min_exp) { pos[7][0] = (float)x; pos[7][1] = (float)y; found = true; }
    }
    const float width = (float)W, height = (float)H;
    for (int k = 0; k < 8; ++k) {  // src/main_view.wgsl:467-468
        pos[k][0] = pos[k][0] / (width - 1.0f);
        pos[k][1] = (height - 1.0f - pos[k][1]) / (height - 1.0f);
    }
}

// ---- src/overlay.wgsl:4-40 ------------------------------------------------------------------------------
struct V2 { float x, y; };
inline V2 vec2squareroots(V2 a) {  // a = (a0, a1): x^2 + a1 x + a0 = 0
    const float discrim = a.y * a.y - 4.0f * a.x;
    if (discrim >= 0.0f) {
        const float root = sqrtf(discrim);
        return {(-a.y - root) / 2.0f, (-a.y + root) / 2.0f};
    }
    const float root = sqrtf(-discrim);
    return {-a.y / 2.0f, root / 2.0f};
}
// m = [c0 c1] column major: c0 = (m00, m01), c1 = (m10, m11)
inline V2 mat2eigenvalues(float c0x, float c0y, float c1x, float c1y) {
    const float det = c0x * c1y - c0y * c1x;
    return vec2squareroots({det, -(c0x + c1y)});
}

}  // namespace

// =================================================================================================
// C entry points (ctypes)
// =================================================================================================
struct OracleFrame {
    int32_t width, height;
    CameraPOD camera;      // 112 B
    SettingsPOD settings;  // 24 B
    float h_initial;       // 0.1 default, 0.001 enhance (src/main_view.rs:785-787)
    int32_t adaptive;      // src/main_view.wgsl:227
    int32_t function_id;   // enum order of src/functions.rs:2-13 (0 = custom hook / placeholder)
    int32_t n_verts, n_faces;
    const float* verts;
    const uint32_t* faces;
    // pixel sample: texture-space rectangle [x0,x1) x [y0,y1) visited with the given strides
    int32_t x0, x1, y0, y1, x_step, y_step;
    int32_t use_fp64;  // run the same restatement in double (diagnostic only)
    int32_t n_threads; // 0 = all
};

struct OracleOutputs {  // every pointer may be null; all indexed y*W + x in texture space (row 0 = top)
    float* color_f32x4;      // pre-quantisation colour
    uint8_t* rgba8;          // ray_casting_texture contents
    float* endpoint_f32x4;   // mapping_texture contents
    int32_t* hit_id;
    int32_t* accepted;       // per-ray accepted RK4 steps
    int32_t* rejected;       // per-ray rejected RK4 attempts
    int32_t* prehit;         // per-ray steps that ran the intersection loop
};

template <class F>
static void march_rows(const OracleFrame& f, const OracleOutputs& o) {
    const int W = f.width, H = f.height;
    Mesh mesh{f.verts, f.faces, f.n_faces};
    const CamBasis<F> basis = camera_basis<F>(f.camera, W, H);
    const int ys = std::max(1, f.y_step), xs = std::max(1, f.x_step);
    const int nrows = (f.y1 - f.y0 + ys - 1) / ys;
    const int ncols = (f.x1 - f.x0 + xs - 1) / xs;
    const long long total = (long long)nrows * ncols;
#ifdef _OPENMP
    if (f.n_threads > 0) omp_set_num_threads(f.n_threads);
#endif
    // one ray per loop iteration, dynamic scheduling in chunks of 16 rays (cost per ray varies by ~5x)
#pragma omp parallel for schedule(dynamic, 16)
    for (long long k = 0; k < total; ++k) {
        {
            const int y = f.y0 + (int)(k / ncols) * ys;  // texture row (coords.y)
            const int x = f.x0 + (int)(k % ncols) * xs;
            const int gid_y = H - 1 - y;                  // src/main_view.wgsl:342
            const V3<F> dir = pixel_dir(basis, W, H, x, gid_y);
            const MarchResult<F> res = nonlinear_ray_color<F>(mesh, f.settings, f.function_id, F(f.h_initial),
                                                              f.adaptive != 0, basis.origin, dir);
            const size_t idx = (size_t)y * W + x;
            if (o.color_f32x4) {
                float* c = o.color_f32x4 + 4 * idx;
                c[0] = (float)res.color.r; c[1] = (float)res.color.g; c[2] = (float)res.color.b; c[3] = (float)res.color.a;
            }
            if (o.rgba8) {
                uint8_t* c = o.rgba8 + 4 * idx;
                c[0] = to_unorm8((float)res.color.r); c[1] = to_unorm8((float)res.color.g);
                c[2] = to_unorm8((float)res.color.b); c[3] = to_unorm8((float)res.color.a);
            }
            if (o.endpoint_f32x4) {
                float* e = o.endpoint_f32x4 + 4 * idx;
                e[0] = (float)res.mapping_point.x; e[1] = (float)res.mapping_point.y; e[2] = (float)res.mapping_point.z; e[3] = 1.0f;
            }
            if (o.hit_id) o.hit_id[idx] = res.hit_id;
            if (o.accepted) o.accepted[idx] = res.stats.accepted;
            if (o.rejected) o.rejected[idx] = res.stats.rejected;
            if (o.prehit) o.prehit[idx] = res.stats.prehit;
        }
    }
}

extern "C" {

int oracle_march(const OracleFrame* f, const OracleOutputs* o) {
    if (!f || !o || f->width < 2 || f->height < 2) return 1;
    if (f->use_fp64) march_rows<double>(*f, *o); else march_rows<float>(*f, *o);
    return 0;
}

void oracle_set_custom_field(custom_field_fn fn) { g_custom = fn; }

// Sample-ray start positions (normalised image coordinates), src/main_view.wgsl:381-472.
// outline != 0: search `exponents` (stride W); else click mode around settings.mouse.
void oracle_sample_positions(const OracleFrame* f, int outline, const float* exponents, float* pos16) {
    float pos[8][2];
    if (outline) {
        outline_positions(exponents, f->width, f->height, pos);
    } else {
        for (int k = 0; k < 8; ++k) {
            pos[k][0] = f->settings.mouse_x + 0.01f * SAMPLE_POSITIONS[k][0];
            pos[k][1] = f->settings.mouse_y + 0.01f * SAMPLE_POSITIONS[k][1];
        }
    }
    memcpy(pos16, pos, sizeof(pos));
}

// The 8 sample rays (src/main_view.wgsl:473-479 + 281-314) -> 800 x {vec4 position, vec4 colour}.
void oracle_sample_rays(const OracleFrame* f, const float* pos16, float* ray_samples) {
    const CamBasis<float> b = camera_basis<float>(f->camera, f->width, f->height);
#pragma omp parallel for schedule(dynamic, 1)
    for (int k = 0; k < 8; ++k) {
        const float px = pos16[2 * k], py = pos16[2 * k + 1];
        const float u2 = px * b.viewport_width - 0.5f * b.viewport_width;
        const float v2 = py * b.viewport_height - 0.5f * b.viewport_height;
        const V3<float> s2 = u2 * b.nh + v2 * b.nv + b.focal_length * b.view_direction;
        const V3<float> dir2 = normalize(s2);
        sample_rays<float>(f->settings, f->function_id, b.origin, dir2, k, SAMPLE_COLORS[k], ray_samples);
    }
}

// Overlay pass (src/overlay.wgsl:99-146).  Inputs are the march outputs of the same frame.
// exponents is read-modify-write (border pixels keep their previous value), stride W.
void oracle_overlay(int W, int H, const SettingsPOD* st, const uint8_t* raycast_rgba8, const float* endpoint_f32x4,
                    uint8_t* final_rgba8, float* exponents) {
    const int delta = st->central_difference_delta;
    const int padding = delta * 2;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            const size_t idx = (size_t)y * W + x;
            const uint8_t* cin = raycast_rgba8 + 4 * idx;
            uint8_t* cout = final_rgba8 + 4 * idx;
            bool passthrough = st->overlay_mode == 0 || x < padding || y < padding || x >= W - padding || y >= H - padding;
            if (!passthrough && st->overlay_mode == 1) {
                auto E = [&](int xx, int yy, int c) { return endpoint_f32x4[4 * ((size_t)yy * W + xx) + c]; };
                float gx[3], gy[3];
                for (int c = 0; c < 3; ++c) {
                    gx[c] = 0.5f * (E(x + delta, y, c) - E(x - delta, y, c));
                    gy[c] = 0.5f * (E(x, y + delta, c) - E(x, y - delta, c));
                }
                // m = transpose(G) * G, G = mat2x3(gx, gy): m[i][j] = dot(G col j, G col i) (symmetric)
                const float m00 = gx[0] * gx[0] + gx[1] * gx[1] + gx[2] * gx[2];
                const float m01 = gx[0] * gy[0] + gx[1] * gy[1] + gx[2] * gy[2];
                const float m11 = gy[0] * gy[0] + gy[1] * gy[1] + gy[2] * gy[2];
                const V2 eigen = mat2eigenvalues(m00, m01, m01, m11);
                const float exponent = sqrtf(m_max(eigen.x, eigen.y));
                const float scaled = expf(st->lyapunov_scaling * exponent - 5.0f);
                exponents[idx] = scaled;
                if (scaled >= 0.8f) {
                    const float alpha = m_min(1.0f, scaled) * 0.5f;
                    for (int c = 0; c < 4; ++c) {
                        const float col = (float)cin[c] / 255.0f;
                        cout[c] = to_unorm8(alpha * 1.0f + (1.0f - alpha) * col);
                    }
                    continue;
                }
            }
            cout[0] = cin[0]; cout[1] = cin[1]; cout[2] = cin[2]; cout[3] = cin[3];
        }
    }
}

// ---- unit-level entry points (helper parity tests) ---------------------------------------------------
void oracle_field(int fn_id, const SettingsPOD* st, int n, const float* p_prev, const float* p, const float* v0,
                  const float* v, const float* t, float* out) {
    for (int i = 0; i < n; ++i) {
        const V3<float> r = field_function<float>(fn_id, *st, {p_prev[3 * i], p_prev[3 * i + 1], p_prev[3 * i + 2]},
                                                  {p[3 * i], p[3 * i + 1], p[3 * i + 2]}, {v0[3 * i], v0[3 * i + 1], v0[3 * i + 2]},
                                                  {v[3 * i], v[3 * i + 1], v[3 * i + 2]}, t[i]);
        out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z;
    }
}

// The same restatement evaluated in double (diagnostic): where the reference's own fp32 formulation is ill-conditioned
// (refraction at grazing incidence computes sqrt(1 - (1 - cosi^2))), this is what it is a rounding of.
void oracle_field_fp64(int fn_id, const SettingsPOD* st, int n, const float* p_prev, const float* p, const float* v0,
                       const float* v, const float* t, float* out) {
    for (int i = 0; i < n; ++i) {
        const V3<double> r = field_function<double>(fn_id, *st, {p_prev[3 * i], p_prev[3 * i + 1], p_prev[3 * i + 2]},
                                                    {p[3 * i], p[3 * i + 1], p[3 * i + 2]}, {v0[3 * i], v0[3 * i + 1], v0[3 * i + 2]},
                                                    {v[3 * i], v[3 * i + 1], v[3 * i + 2]}, (double)t[i]);
        out[3 * i] = (float)r.x; out[3 * i + 1] = (float)r.y; out[3 * i + 2] = (float)r.z;
    }
}

// which: 0 rotateX 1 rotateY 2 rotateZ 3 translate(v, a, b, c) 4 refraction(a, b, v, n) 5 point_plane_distance(v, n, p0)
//        6 sigmoid(a) 7 refraction_index(a) 8 reflect(v, n) 9 normalize(v) 10 cross(v, n)
void oracle_helper(int which, int n, const float* v, const float* nrm, const float* p0, const float* a, const float* b,
                   const float* c, float* out) {
    for (int i = 0; i < n; ++i) {
        const V3<float> V(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
        const V3<float> N(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]);
        const V3<float> P(p0[3 * i], p0[3 * i + 1], p0[3 * i + 2]);
        V3<float> r(0, 0, 0);
        switch (which) {
            case 0: r = rotateX(V, a[i]); break;
            case 1: r = rotateY(V, a[i]); break;
            case 2: r = rotateZ(V, a[i]); break;
            case 3: r = translate(V, a[i], b[i], c[i]); break;
            case 4: r = refraction(a[i], b[i], V, N); break;
            case 5: r.x = point_plane_distance(V, N, P); break;
            case 6: r.x = sigmoid(a[i]); break;
            case 7: r.x = refraction_index(a[i]); break;
            case 8: r = reflect(V, N); break;
            case 9: r = normalize(V); break;
            case 10: r = cross(V, N); break;
        }
        out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z;
    }
}

// hit_triangle over n (triangle, origin, direction) triples; tri = 9 floats each.
void oracle_hit_triangle(int n, const float* tri, const float* origin, const float* dir, float* t_out) {
    for (int i = 0; i < n; ++i) {
        const V3<float> T[3] = {{tri[9 * i], tri[9 * i + 1], tri[9 * i + 2]},
                                {tri[9 * i + 3], tri[9 * i + 4], tri[9 * i + 5]},
                                {tri[9 * i + 6], tri[9 * i + 7], tri[9 * i + 8]}};
        t_out[i] = hit_triangle<float>(T, {origin[3 * i], origin[3 * i + 1], origin[3 * i + 2]},
                                       {dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]});
    }
}

void oracle_vec2squareroots(int n, const float* a01, float* out) {
    for (int i = 0; i < n; ++i) {
        const V2 r = vec2squareroots({a01[2 * i], a01[2 * i + 1]});
        out[2 * i] = r.x; out[2 * i + 1] = r.y;
    }
}

int oracle_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
