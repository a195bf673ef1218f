/* linon_b200_host.h — C shim over the C++ host mirror (linon_b200/csrc/host/*.hpp) so the Python harness and
 * tests can drive the same MainView / ArcballCamera / scene code a C++ caller would use.
 *
 * NOT part of the FFI boundary a Rust integration binds (that is linon_b200.h): the reference keeps arcball.rs,
 * vertices.rs, cornell_box.rs and functions.rs on its own side.  These are the host inputs of the path, restated
 * so a headless harness can produce exactly the reference's start-up state.
 */
#ifndef LINON_B200_HOST_H
#define LINON_B200_HOST_H

#include "linon_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Application::new start-up scene (src/application.rs:105-123): Cornell box (src/cornell_box.rs), x and z negated,
 * normalised (src/vertices.rs:3-17).  verts: 72*3 f32, faces: 36*3 u32, center: 3 f32 (src/vertices.rs:19-57).
 * negate_xz = 0 gives load_default_model (src/application.rs:194-197). */
void linon_host_default_scene(float* verts216, uint32_t* faces108, float* center3, int negate_xz);
void linon_host_normalize_vertices(float* verts, uint32_t n_floats);
/* The OBJ drag-and-drop path: tobj::load_obj(triangulate) -> models[0] -> Application::load_model (src/main.rs:139-157,
 * src/application.rs:199-226).  Parses `obj_text`, fan-triangulates, keeps the first model, re-indexes its positions in
 * order of first use, normalises them (src/vertices.rs:3-17, no x/z negation on this path) and returns the bounding-box
 * centre.  Call with verts = faces = NULL to query the sizes (vertices / faces, not floats); capacities count vertices
 * and faces.  Returns LINON_ERR_INVALID with a message for malformed text, LINON_ERR_SIZE when a capacity is too small. */
int linon_host_load_obj(const char* obj_text, float* verts, uint32_t verts_capacity, uint32_t* faces, uint32_t faces_capacity,
                        uint32_t* n_verts, uint32_t* n_faces, float* center3, char* err, size_t err_len);
void linon_host_get_center(const float* verts, uint32_t n_floats, float* center3);

/* PredefinedFunction (src/functions.rs): id in enum order 0..9. Returns the byte length (excluding NUL) or -1. */
int linon_host_function_code(int id, char* out, size_t out_len);
int linon_host_function_name(int id, char* out, size_t out_len);

/* ArcballCamera (src/arcball.rs) */
typedef struct linon_host_arcball linon_host_arcball;
linon_host_arcball* linon_host_arcball_new(const float* center3, float zoom_speed, float screen_w, float screen_h);
void linon_host_arcball_free(linon_host_arcball* c);
void linon_host_arcball_rotate(linon_host_arcball* c, float prev_x, float prev_y, float cur_x, float cur_y);
void linon_host_arcball_zoom(linon_host_arcball* c, float amount, float elapsed);
void linon_host_arcball_pan(linon_host_arcball* c, float dx, float dy);
void linon_host_arcball_update_screen(linon_host_arcball* c, float w, float h);
/* CameraUniform::moving (src/main_view.rs:21-31) */
void linon_host_arcball_uniform(const linon_host_arcball* c, linon_camera* out);

/* MainView (src/main_view.rs:53-849).  Errors: functions return 0 / non-zero; message via linon_host_mv_error. */
typedef struct linon_host_mv linon_host_mv;
linon_host_mv* linon_host_mv_new(int device, const float* verts, uint32_t n_verts, const uint32_t* faces,
                                 uint32_t n_faces, const float* center3, uint32_t width, uint32_t height,
                                 int discrete_gpu);
void linon_host_mv_free(linon_host_mv* m);
const char* linon_host_mv_error(const linon_host_mv* m); /* m == NULL: error of the failed _new */
int linon_host_mv_update_model(linon_host_mv* m, const float* verts, uint32_t n_verts, const uint32_t* faces,
                               uint32_t n_faces, const float* center3);
int linon_host_mv_update_camera(linon_host_mv* m);
int linon_host_mv_reset_camera(linon_host_mv* m);
int linon_host_mv_on_zoom(linon_host_mv* m, float delta);
int linon_host_mv_on_pointer_moved(linon_host_mv* m, int camera_op /*0 none 1 rotate 2 pan*/, float prev_x,
                                   float prev_y, float x, float y);
int linon_host_mv_update_settings(linon_host_mv* m, const linon_settings* s);
int linon_host_mv_resize_texture(linon_host_mv* m, uint32_t width, uint32_t height);
int linon_host_mv_reload_shader(linon_host_mv* m, const char* field_function, char* err, size_t err_len);
/* Application::reload_compute_shader(new_src) (src/application.rs:228-232): MainView::reload_shader with
 * new_src = Some(..) and the current field function; see linon_reload_shader_source for what is honoured. */
int linon_host_mv_reload_compute_shader(linon_host_mv* m, const char* new_src, char* err, size_t err_len);
int linon_host_mv_set_adaptive_sampling(linon_host_mv* m, int enabled);
int linon_host_mv_render(linon_host_mv* m);
int linon_host_mv_render_high_accuracy(linon_host_mv* m, const char* field_function);
int linon_host_mv_render_outline_rays(linon_host_mv* m, const char* field_function);
int linon_host_mv_needs_redraw(const linon_host_mv* m);
linon_view* linon_host_mv_handle(linon_host_mv* m);
linon_host_arcball* linon_host_mv_camera(linon_host_mv* m); /* borrowed; owned by the MainView */
void linon_host_mv_size(const linon_host_mv* m, uint32_t* width, uint32_t* height);

#ifdef __cplusplus
}
#endif
#endif /* LINON_B200_HOST_H */
