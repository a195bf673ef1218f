// mirage_fused.cuh — hand-fused restatements of the four predefined mirage field functions (fast math mode only).
//
// The host splices this text in place of the translated WGSL body when (a) the math mode is `fast`, (b) the body the
// application installed is byte-identical with one of PredefinedFunction::{MirageSpherical, MiragePlane,
// MirageSphericalSigmoid, MiragePlaneSigmoid}.to_code() (src/functions.rs:36-94) and (c) LINON_OPT_SPECIALIZE is on.
// Every other body - and every other math mode - runs the generic translation.  -DLINON_MIRAGE_KIND selects the body:
//   1 MirageSpherical   2 MiragePlane   3 MirageSphericalSigmoid   4 MiragePlaneSigmoid
//
// What is restated (same real-number function, fp32 rounding differs like any fast-math build; graded against the
// oracle by tests/test_gpu_units.py::test_field_functions_match_the_oracle and the frame tests of test_gpu_parity.py):
//
//   * The bodies compute part = clamp(0, 1, f(dist)) (misordered: min(1, f)), t = part t_env + (1 - part) t_src for the
//     previous and the current point and call refraction(t_in, t_out, v, normal), whose only use of the temperatures is
//     eta = n(t_in) / n(t_out) with n(t) = A(t) / B(t), A(t) = a0 - a1 t, B(t) = 1 + c2 t (refraction_index,
//     src/main_view.wgsl:109-117).  t is affine in part, so A and B are affine in part: four FFMAs and ONE reciprocal
//     give eta = A(in) B(out) / (A(out) B(in)).
//   * Sigmoid variants: part = 1 / (1 + E), E = exp(-(12 dist / max_dist - 6)) <= e^6.  Multiplying A and B by
//     (1 + E) > 0 keeps them affine - in E - and the factor cancels in eta, so the sigmoid's reciprocal disappears:
//     A'(E) = (a0 - t_env a1) + (a0 - t_src a1) E,  B'(E) = (1 + t_env c2) + (1 + t_src c2) E.  min(1, sigmoid) never binds.
//   * Spherical variants: normalize(c) and length(c) share one reciprocal square root.
//   * Plane variants: the plane normal is (0, 1, 0) and p0 = (0, 0.1, 0), so point_plane_distance(p) = |p.y - 0.1|
//     (dot(p, n) = 0 p.x + p.y + 0 p.z is exact for finite p) and dot(v, n) = v.y.
//   Special-function evaluations per call: 3 / 2 / 4 / 3 instead of 4 / 3 / 6 / 5.
//
//   * LINON_STAGE1_IDENTITY: the reference's first Runge-Kutta stage evaluates the field at p = p_prev
//     (src/main_view.wgsl:246), where t_in = t_out, eta = 1 and refraction(t, t, v, n) = v for every v and n
//     (k = cosi^2, the normal term cancels).  The reference's own evaluation returns v plus rounding noise of its
//     1 - (1 - cosi^2) formulation; the fast build returns v, i.e. k1 = cur_dir, and evaluates three stages per step.
//     Only the work actually executed is counted in the roofline numerator (linon_b200/flops.py).

#ifndef LINON_MIRAGE_KIND
#error "mirage_fused.cuh needs -DLINON_MIRAGE_KIND=1..4"
#endif
#define LINON_STAGE1_IDENTITY 1

template <class S> struct LinonField {
    // refraction (src/main_view.wgsl:119-136) for a given eta and d = dot(v, n); see refraction_s in wgsl_prelude.cuh
    static __device__ __forceinline__ vec3t<S> refract_eta(S eta, S d, vec3t<S> v, vec3t<S> n) {
        const S cr = w_min(1.0f, d);  // clamp(-1, 1, d) with its misordered arguments
        const S k = 1.0f - eta * eta * (1.0f - cr * cr);
        const S sk = w_sqrt(k);
        const S cs = w_select(sk, -sk, cr < 0.0f) - eta * cr;
        vec3t<S> r = eta * v + cs * n;
        const typename linon_mask<S>::type total = k < 0.0f;
        if (__builtin_expect(s_any(total), 0)) r = w_select(r, v - (2.0f * d) * n, total);
        return r;
    }

    static __device__ __forceinline__ vec3t<S> eval(const vec3t<S> p_prev, const vec3t<S> p, const vec3t<S> v0,
                                                    const vec3t<S> v, const S t, const WgslSettings& settings,
                                                    const WgslCamera& camera) {
        const float t_env = 15.0f;
#if LINON_MIRAGE_KIND == 1 || LINON_MIRAGE_KIND == 3
        const float t_src = 200.0f, max_dist = 0.25f;
#else
        const float t_src = 30.0f, max_dist = 0.01f;
#endif
        // n(t) = A(t) / B(t)
        const float c1p = 0.0000104f * 101325.0f;
        const float a0 = c1p * (1.0f + 101325.0f * 1e-10f * 60.1f);
        const float a1 = c1p * 101325.0f * 1e-10f * 0.972f;
        const float c2 = 0.00366f;

#if LINON_MIRAGE_KIND == 1 || LINON_MIRAGE_KIND == 3
        const vec3f center(-0.5f, 0.5f, -0.5f);
        const vec3t<S> c_out = p - center;
        const S q_out = w_dot(c_out, c_out);
        const S r_out = w_inverseSqrt(q_out);
        const vec3t<S> normal = c_out * r_out;
        const S dist_out = q_out * r_out;
        const vec3t<S> c_in = p_prev - center;
        const S q_in = w_dot(c_in, c_in);
        const S dist_in = q_in * w_inverseSqrt(q_in);
        const S d = w_dot(v, normal);
#else
        const S dist_out = w_abs(p.y - 0.1f);
        const S dist_in = w_abs(p_prev.y - 0.1f);
        const S d = v.y;
#endif

#if LINON_MIRAGE_KIND >= 3
        // E = exp(-(dist / max_dist * 12 - 6)) = exp2(dist * ks + k0)
        const float log2e = 1.4426950408889634f;
        const float ks = -12.0f / max_dist * log2e, k0 = 6.0f * log2e;
        const S e_in = w_exp2(dist_in * ks + k0);
        const S e_out = w_exp2(dist_out * ks + k0);
        const float A1 = a0 - t_env * a1, A2 = a0 - t_src * a1;
        const float B1 = 1.0f + t_env * c2, B2 = 1.0f + t_src * c2;
        const S a_in = A1 + A2 * e_in, b_in = B1 + B2 * e_in;
        const S a_out = A1 + A2 * e_out, b_out = B1 + B2 * e_out;
#else
        // part = min(1, dist / max_dist), t = t_src - (t_src - t_env) part
        const float inv = 1.0f / max_dist, dt = t_src - t_env;
        const S s_in = w_min(1.0f, dist_in * inv);
        const S s_out = w_min(1.0f, dist_out * inv);
        const float A1 = a0 - a1 * t_src, A2 = a1 * dt;
        const float B1 = 1.0f + c2 * t_src, B2 = -c2 * dt;
        const S a_in = A1 + A2 * s_in, b_in = B1 + B2 * s_in;
        const S a_out = A1 + A2 * s_out, b_out = B1 + B2 * s_out;
#endif
        const S eta = (a_in * b_out) / (a_out * b_in);

#if LINON_MIRAGE_KIND == 1 || LINON_MIRAGE_KIND == 3
        return refract_eta(eta, d, v, normal);
#else
        // normal = (0, 1, 0): only the y component receives the normal term
        const S cr = w_min(1.0f, d);
        const S k = 1.0f - eta * eta * (1.0f - cr * cr);
        const S sk = w_sqrt(k);
        const S cs = w_select(sk, -sk, cr < 0.0f) - eta * cr;
        vec3t<S> r(eta * v.x, eta * v.y + cs, eta * v.z);
        const typename linon_mask<S>::type total = k < 0.0f;
        if (__builtin_expect(s_any(total), 0)) r = w_select(r, vec3t<S>(v.x, -v.y, v.z), total);
        return r;
#endif
    }
};
