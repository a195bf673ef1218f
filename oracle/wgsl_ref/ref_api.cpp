// ref_api.cpp — exported C API of oracle/_ref/liblinon_wgsl_ref.so: the reference's own WGSL (main_view.wgsl with each
// predefined field function of functions.rs spliced in, and overlay.wgsl), transpiled mechanically to C++ by
// wgsl2cpp.py and compiled with gcc.  TEST INFRASTRUCTURE ONLY (oracle/README.md).
#include <cstdint>
#include <cstring>
#include "harness_types.hpp"
#ifdef _OPENMP
#include <omp.h>
#endif

#define DECL(N) extern const RefModule* const wgslref_fn##N[4];
DECL(0) DECL(1) DECL(2) DECL(3) DECL(4) DECL(5) DECL(6) DECL(7) DECL(8) DECL(9)
#undef DECL
int wgslref_run_overlay(int W, int H, const RefSettings* st, const uint8_t* raycast_rgba8, const float* mapping_f32x4,
                        uint8_t* final_rgba8, float* exponents_io, int n_threads);
void wgslref_run_vec2squareroots(int n, const float* a01, float* out);
extern const char* const wgslref_build_info;

static const RefModule* module_of(int fn, int variant) {
    static const RefModule* const* const table[10] = {wgslref_fn0, wgslref_fn1, wgslref_fn2, wgslref_fn3, wgslref_fn4,
                                                      wgslref_fn5, wgslref_fn6, wgslref_fn7, wgslref_fn8, wgslref_fn9};
    if (fn < 0 || fn > 9 || variant < 0 || variant > 3) return nullptr;
    return table[fn][variant];
}

extern "C" {

// variant: 0 = as shipped (h_initial 0.1, adaptive sampling), 1 = Enhance (h_initial 0.001, src/main_view.rs:785-787),
// 2 = adaptive sampling off (intent of src/main_view.rs:764-776), 3 = outline sample rays (intent of :811-814)
int wgslref_main_view(int function_id, int variant, const RefFrame* f, const RefOutputs* o) {
    const RefModule* m = module_of(function_id, variant);
    if (!m || !f || !o) return 2;
    return m->main_view(f, o);
}

int wgslref_overlay(int W, int H, const RefSettings* st, const uint8_t* raycast_rgba8, const float* mapping_f32x4,
                    uint8_t* final_rgba8, float* exponents_io, int n_threads) {
    return wgslref_run_overlay(W, H, st, raycast_rgba8, mapping_f32x4, final_rgba8, exponents_io, n_threads);
}

int wgslref_field(int function_id, const RefSettings* st, int n, const float* p_prev, const float* p, const float* v0,
                  const float* v, const float* t, float* out) {
    const RefModule* m = module_of(function_id, 0);
    if (!m) return 2;
    m->field(st, n, p_prev, p, v0, v, t, out);
    return 0;
}

int wgslref_helper(int which, int n, const float* v, const float* nrm, const float* p0, const float* a, const float* b,
                   const float* c, float* out) {
    module_of(0, 0)->helper(which, n, v, nrm, p0, a, b, c, out);
    return 0;
}

int wgslref_hit_triangle(int n, const float* tri, const float* origin, const float* direction, float* out) {
    module_of(0, 0)->hit_triangle(n, tri, origin, direction, out);
    return 0;
}

int wgslref_vec2squareroots(int n, const float* a01, float* out) {
    wgslref_run_vec2squareroots(n, a01, out);
    return 0;
}

const char* wgslref_info(void) { return wgslref_build_info; }

int wgslref_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
