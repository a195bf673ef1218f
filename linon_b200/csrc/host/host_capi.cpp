// host_capi.cpp — C shim over the C++ host mirror (see include/linon_b200_host.h).
#include "../../../include/linon_b200_host.h"

#include <cstdio>
#include <cstring>
#include <string>

#include "main_view.hpp"
#include "obj_loader.hpp"

struct linon_host_arcball {
    linon::ArcballCamera cam;
};
struct linon_host_mv {
    linon::MainView* view = nullptr;
    std::string error;
};

namespace {
std::string g_mv_new_error;

int copy_out(const std::string& s, char* out, size_t out_len) {
    if (!out || out_len < s.size() + 1) return -1;
    memcpy(out, s.c_str(), s.size() + 1);
    return (int)s.size();
}

template <class F>
int guarded(linon_host_mv* m, F&& f) {
    if (!m || !m->view) return LINON_ERR_INVALID;
    try {
        f(*m->view);
        return LINON_OK;
    } catch (const std::exception& e) {
        m->error = e.what();
        return LINON_ERR_CUDA;
    }
}
}  // namespace

extern "C" {

void linon_host_default_scene(float* verts216, uint32_t* faces108, float* center3, int negate_xz) {
    std::array<float, 3> c{};
    const linon::Mesh m = linon::default_scene(&c, negate_xz != 0);
    if (verts216) memcpy(verts216, m.vertices.data(), m.vertices.size() * sizeof(float));
    if (faces108) memcpy(faces108, m.indices.data(), m.indices.size() * sizeof(uint32_t));
    if (center3) memcpy(center3, c.data(), 12);
}

void linon_host_normalize_vertices(float* verts, uint32_t n_floats) {
    std::vector<float> v(verts, verts + n_floats);
    linon::normalize_vertices(v);
    memcpy(verts, v.data(), n_floats * sizeof(float));
}

void linon_host_get_center(const float* verts, uint32_t n_floats, float* center3) {
    const std::vector<float> v(verts, verts + n_floats);
    const auto c = linon::get_center(v);
    memcpy(center3, c.data(), 12);
}

int linon_host_load_obj(const char* obj_text, float* verts, uint32_t verts_capacity, uint32_t* faces, uint32_t faces_capacity,
                        uint32_t* n_verts, uint32_t* n_faces, float* center3, char* err, size_t err_len) {
    if (!obj_text || !n_verts || !n_faces) return LINON_ERR_INVALID;
    linon::Mesh m;
    std::string error;
    if (!linon::load_obj_text(obj_text, m, error)) {
        if (err && err_len) snprintf(err, err_len, "%s", error.c_str());
        return LINON_ERR_INVALID;
    }
    const std::array<float, 3> c = linon::prepare_loaded_model(m);
    *n_verts = (uint32_t)(m.vertices.size() / 3);
    *n_faces = (uint32_t)(m.indices.size() / 3);
    if (center3) memcpy(center3, c.data(), 12);
    if (!verts && !faces) return LINON_OK;  // size query
    if (*n_verts > verts_capacity || *n_faces > faces_capacity) return LINON_ERR_SIZE;
    if (verts) memcpy(verts, m.vertices.data(), m.vertices.size() * sizeof(float));
    if (faces) memcpy(faces, m.indices.data(), m.indices.size() * sizeof(uint32_t));
    return LINON_OK;
}

int linon_host_function_code(int id, char* out, size_t out_len) {
    if (id < 0 || id >= linon::kNumPredefinedFunctions) return -1;
    return copy_out(linon::to_code((linon::PredefinedFunction)id), out, out_len);
}
int linon_host_function_name(int id, char* out, size_t out_len) {
    if (id < 0 || id >= linon::kNumPredefinedFunctions) return -1;
    return copy_out(linon::to_string((linon::PredefinedFunction)id), out, out_len);
}

linon_host_arcball* linon_host_arcball_new(const float* c, float zoom_speed, float w, float h) {
    return new linon_host_arcball{linon::ArcballCamera({c[0], c[1], c[2]}, zoom_speed, w, h)};
}
void linon_host_arcball_free(linon_host_arcball* c) { delete c; }
void linon_host_arcball_rotate(linon_host_arcball* c, float px, float py, float cx, float cy) { c->cam.rotate(px, py, cx, cy); }
void linon_host_arcball_zoom(linon_host_arcball* c, float amount, float elapsed) { c->cam.zoom(amount, elapsed); }
void linon_host_arcball_pan(linon_host_arcball* c, float dx, float dy) { c->cam.pan(dx, dy); }
void linon_host_arcball_update_screen(linon_host_arcball* c, float w, float h) { c->cam.update_screen(w, h); }
void linon_host_arcball_uniform(const linon_host_arcball* c, linon_camera* out) { *out = linon::camera_uniform_moving(c->cam); }

linon_host_mv* linon_host_mv_new(int device, const float* verts, uint32_t n_verts, const uint32_t* faces, uint32_t n_faces,
                                 const float* center3, uint32_t width, uint32_t height, int discrete_gpu) {
    try {
        auto* m = new linon_host_mv;
        m->view = new linon::MainView(device, std::vector<float>(verts, verts + 3 * (size_t)n_verts),
                                      std::vector<uint32_t>(faces, faces + 3 * (size_t)n_faces),
                                      {center3[0], center3[1], center3[2]}, width, height, discrete_gpu != 0);
        return m;
    } catch (const std::exception& e) {
        g_mv_new_error = e.what();
        return nullptr;
    }
}
void linon_host_mv_free(linon_host_mv* m) {
    if (!m) return;
    delete m->view;
    delete m;
}
const char* linon_host_mv_error(const linon_host_mv* m) { return m ? m->error.c_str() : g_mv_new_error.c_str(); }

int linon_host_mv_update_model(linon_host_mv* m, const float* verts, uint32_t n_verts, const uint32_t* faces, uint32_t n_faces,
                               const float* c) {
    return guarded(m, [&](linon::MainView& v) {
        v.update_model(std::vector<float>(verts, verts + 3 * (size_t)n_verts),
                       std::vector<uint32_t>(faces, faces + 3 * (size_t)n_faces), {c[0], c[1], c[2]});
    });
}
int linon_host_mv_update_camera(linon_host_mv* m) { return guarded(m, [](linon::MainView& v) { v.update_camera(); }); }
int linon_host_mv_reset_camera(linon_host_mv* m) { return guarded(m, [](linon::MainView& v) { v.reset_camera(); }); }
int linon_host_mv_on_zoom(linon_host_mv* m, float delta) { return guarded(m, [&](linon::MainView& v) { v.on_zoom(delta); }); }
int linon_host_mv_on_pointer_moved(linon_host_mv* m, int op, float px, float py, float x, float y) {
    return guarded(m, [&](linon::MainView& v) {
        v.on_pointer_moved(op == 1 ? linon::CameraOperation::Rotate : op == 2 ? linon::CameraOperation::Pan : linon::CameraOperation::None,
                           px, py, x, y);
    });
}
int linon_host_mv_update_settings(linon_host_mv* m, const linon_settings* s) {
    return guarded(m, [&](linon::MainView& v) { v.update_settings(*s); });
}
int linon_host_mv_resize_texture(linon_host_mv* m, uint32_t w, uint32_t h) {
    return guarded(m, [&](linon::MainView& v) { v.resize_texture(w, h); });
}
int linon_host_mv_reload_shader(linon_host_mv* m, const char* field_function, char* err, size_t err_len) {
    if (!m || !m->view || !field_function) return LINON_ERR_INVALID;
    std::string e;
    if (!m->view->reload_shader(field_function, &e)) {
        m->error = e;
        if (err && err_len) snprintf(err, err_len, "%s", e.c_str());
        return LINON_ERR_COMPILE;
    }
    return LINON_OK;
}
int linon_host_mv_reload_compute_shader(linon_host_mv* m, const char* new_src, char* err, size_t err_len) {
    if (!m || !m->view) return LINON_ERR_INVALID;
    std::string e;
    if (!m->view->reload_compute_shader(new_src ? new_src : "", &e)) {
        m->error = e;
        if (err && err_len) snprintf(err, err_len, "%s", e.c_str());
        const char* last = linon_last_error(m->view->handle());
        return (last && e == last && e.find("the CUDA path hot-reloads") != std::string::npos) ? LINON_ERR_UNSUPPORTED : LINON_ERR_COMPILE;
    }
    return LINON_OK;
}
int linon_host_mv_set_adaptive_sampling(linon_host_mv* m, int enabled) {
    return guarded(m, [&](linon::MainView& v) { enabled ? v.enable_adaptive_sampling() : v.disable_adaptive_sampling(); });
}
int linon_host_mv_render(linon_host_mv* m) { return guarded(m, [](linon::MainView& v) { v.render(); }); }
int linon_host_mv_render_high_accuracy(linon_host_mv* m, const char* ff) {
    return guarded(m, [&](linon::MainView& v) { v.render_high_accuracy(ff ? std::string(ff) : v.field_function()); });
}
int linon_host_mv_render_outline_rays(linon_host_mv* m, const char* ff) {
    return guarded(m, [&](linon::MainView& v) { v.render_outline_rays(ff ? std::string(ff) : v.field_function()); });
}
int linon_host_mv_needs_redraw(const linon_host_mv* m) { return (m && m->view && m->view->needs_redraw) ? 1 : 0; }
linon_view* linon_host_mv_handle(linon_host_mv* m) { return (m && m->view) ? m->view->handle() : nullptr; }
linon_host_arcball* linon_host_mv_camera(linon_host_mv* m) {
    // ArcballCamera is the only member of linon_host_arcball, so the MainView's camera can be viewed through it
    return (m && m->view) ? reinterpret_cast<linon_host_arcball*>(&m->view->camera()) : nullptr;
}
void linon_host_mv_size(const linon_host_mv* m, uint32_t* w, uint32_t* h) {
    if (!m || !m->view) return;
    if (w) *w = m->view->width();
    if (h) *h = m->view->height();
}

}  // extern "C"
