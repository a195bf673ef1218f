// harness_types.hpp — C structs shared by the dispatch harness and the exported C API of liblinon_wgsl_ref.so.
// TEST INFRASTRUCTURE ONLY (oracle/README.md).
#pragma once
#include <cstdint>

struct RefCamera {  // src/main_view.rs:12-19 (112 bytes)
    float origin[4], view_direction[4], up[4], view_matrix[16];
};
struct RefSettings {  // src/main_view.rs:43-51 (24 bytes)
    float field_weight, mouse_pos_x, mouse_pos_y;
    int32_t overlay_mode, central_difference_delta;
    float lyapunov_scaling;
};
struct RefFrame {
    int32_t width, height;
    RefCamera camera;
    RefSettings settings;
    int32_t n_verts, n_faces;
    const float* verts;          // 3 x f32 per vertex  (src/vertices.rs / main_view.wgsl:34-38)
    const uint32_t* faces;       // 3 x u32 per face    (main_view.wgsl:45-49)
    const float* exponents;      // W*H f32 read by the outline search, or null (zeros)
    const uint8_t* mask;         // W*H bytes in texture space (row 0 = top): invocations to run, or null (all)
    int32_t n_threads;           // 0 = all
};
struct RefOutputs {  // texture space, y*W + x; any pointer may be null
    float* color_f32x4;      // value handed to textureStore(ttarget, ...) before the rgba8unorm conversion
    uint8_t* rgba8;          // ttarget (ray_casting_texture)
    float* endpoint_f32x4;   // mapping
    float* ray_samples;      // 800 x {vec4 position, vec4 color}
};
struct RefModule {  // one transpiled main_view.wgsl (one field function, one variant)
    int (*main_view)(const RefFrame*, const RefOutputs*);
    void (*field)(const RefSettings*, int n, const float* p_prev, const float* p, const float* v0, const float* v,
                  const float* t, float* out);
    void (*helper)(int which, int n, const float* v, const float* nrm, const float* p0, const float* a, const float* b,
                   const float* c, float* out);
    void (*hit_triangle)(int n, const float* tri, const float* origin, const float* direction, float* out);
};
