// main_view.hpp — host mirror of the reference's MainView (src/main_view.rs:53-849) in C++, written ABOVE the
// C ABI of include/linon_b200.h (the reference is Rust; no Rust toolchain exists in this image, so the host side
// is C++ — INTEGRATION.md shows the equivalent Rust binding).  Same method names, argument meaning and error
// behaviour as the reference for the device-facing half; the egui widget glue (show(), texture registration) is
// out of scope.
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/linon_b200.h"
#include "arcball.hpp"
#include "functions.hpp"
#include "scene.hpp"

namespace linon {

using Settings = linon_settings;  // src/main_view.rs:43-51

// CameraUniform::moving, src/main_view.rs:21-31
inline linon_camera camera_uniform_moving(const ArcballCamera& camera) {
    linon_camera u{};
    const auto eye = camera.eye_pos(), dir = camera.eye_dir(), up = camera.up_dir();
    for (int k = 0; k < 3; ++k) {
        u.origin[k] = eye[k];
        u.view_direction[k] = dir[k];
        u.up[k] = up[k];
    }
    for (int k = 0; k < 4; ++k) u.view_matrix[5 * k] = 1.0f;  // Matrix4::identity()
    return u;
}

class MainView {
public:
    bool needs_redraw = true;  // src/main_view.rs:78

    // MainView::new, src/main_view.rs:82-458.  `device` is the CUDA ordinal (the reference passes a wgpu::Device);
    // vertices/faces are the contents of the storage buffers Application binds (src/application.rs:112-122).
    MainView(int device, const std::vector<float>& vertices, const std::vector<uint32_t>& faces,
             std::array<float, 3> center, uint32_t width, uint32_t height, bool discrete_gpu)
        : downscale_factor_(discrete_gpu ? 1u : 2u),
          width_(width / downscale_factor_),
          height_(height / downscale_factor_),
          field_function_(to_code(PredefinedFunction::MirageSphericalSigmoid)),
          camera_(center, 1.0f, (float)width_, (float)height_) {
        check(linon_create(device, width_, height_, &view_), "linon_create");
        try {
            check(linon_set_mesh(view_, vertices.data(), (uint32_t)(vertices.size() / 3), faces.data(),
                                 (uint32_t)(faces.size() / 3)), "linon_set_mesh");
            camera_.zoom(-1.0f, 1.0f);  // :153-155
            update_camera();
        } catch (...) {  // the destructor does not run for a half-constructed object
            linon_destroy(view_);
            view_ = nullptr;
            throw;
        }
    }
    ~MainView() { linon_destroy(view_); }
    MainView(const MainView&) = delete;
    MainView& operator=(const MainView&) = delete;

    // src/main_view.rs:460-488
    void update_model(const std::vector<float>& vertices, const std::vector<uint32_t>& faces, std::array<float, 3> center) {
        check(linon_set_mesh(view_, vertices.data(), (uint32_t)(vertices.size() / 3), faces.data(),
                             (uint32_t)(faces.size() / 3)), "linon_set_mesh");
        camera_ = ArcballCamera(center, 1.0f, (float)width_, (float)height_);
        camera_.zoom(-1.0f, 1.0f);
        update_camera();
        needs_redraw = true;
    }
    // src/main_view.rs:490-494
    void update_camera() {
        const linon_camera u = camera_uniform_moving(camera_);
        check(linon_set_camera(view_, &u), "linon_set_camera");
        needs_redraw = true;
    }
    // src/main_view.rs:496-502
    void reset_camera() {
        camera_ = ArcballCamera(camera_.center, 1.0f, (float)width_, (float)height_);
        camera_.zoom(-1.0f, 1.0f);
        update_camera();
    }
    // src/main_view.rs:504-510 (desktop branch)
    void on_zoom(float delta) {
        camera_.zoom(delta, 1.0f / 60.0f);
        update_camera();
    }
    // src/main_view.rs:512-545; positions are in panel pixels like the reference's pointer events
    void on_pointer_moved(CameraOperation op, float prev_x, float prev_y, float x, float y) {
        const float d = (float)downscale_factor_;
        if (op == CameraOperation::Rotate) {
            disable_adaptive_sampling();
            camera_.rotate(prev_x / d, prev_y / d, x / d, y / d);
            update_camera();
        } else if (op == CameraOperation::Pan) {
            disable_adaptive_sampling();
            camera_.pan((x - prev_x) / d, (y - prev_y) / d);
            update_camera();
        }
    }
    // src/main_view.rs:547-550
    void update_settings(const Settings& settings) {
        check(linon_set_settings(view_, &settings), "linon_set_settings");
        needs_redraw = true;
    }
    // src/main_view.rs:552-647
    void resize_texture(uint32_t width, uint32_t height) {
        width_ = width / downscale_factor_;
        height_ = height / downscale_factor_;
        check(linon_resize(view_, width_, height_), "linon_resize");
        camera_.update_screen((float)width_, (float)height_);
        update_camera();
        needs_redraw = true;
    }
    // src/main_view.rs:649-690 with new_src = None.  Returns false and fills `error` on a compile error; the old
    // function stays active (:677-679).
    bool reload_shader(const std::string& field_function, std::string* error = nullptr) {
        char buf[8192];
        const int rc = linon_set_field_function(view_, field_function.c_str(), buf, sizeof(buf));
        if (rc != LINON_OK) {
            if (error) *error = buf;
            return false;
        }
        field_function_ = field_function;
        needs_redraw = true;
        return true;
    }
    // src/main_view.rs:649-690 with new_src = Some(..): the whole-shader hot reload (file watcher on src/main_view.wgsl,
    // src/main.rs:206-222 -> Application::reload_compute_shader, src/application.rs:228-232).  The helper layer of the
    // edited module and the field function are validated as a pair and committed together; on an error nothing changes
    // (:677-679 returns before `self.shader_src = new_src`).
    bool reload_shader(const std::string& new_src, const std::string& field_function, std::string* error) {
        char buf[8192];
        const int rc = linon_reload_shader(view_, new_src.c_str(), field_function.c_str(), buf, sizeof(buf));  // both or neither
        if (rc != LINON_OK) {
            if (error) *error = buf;
            return false;
        }
        shader_src_ = new_src;
        field_function_ = field_function;
        needs_redraw = true;
        return true;
    }
    // Application::reload_compute_shader, src/application.rs:228-232
    bool reload_compute_shader(const std::string& new_src, std::string* error = nullptr) {
        return reload_shader(new_src, field_function_, error);
    }
    // src/main_view.rs:752-776.  The reference's string replacement for these two is dead at HEAD; the
    // documented behaviour (fixed step while dragging) is implemented as a render flag.
    void enable_adaptive_sampling() { adaptive_sampling_ = true; }
    void disable_adaptive_sampling() { adaptive_sampling_ = false; }

    // src/main_view.rs:778-802: one frame at h_initial = 0.001, then needs_redraw = false so it stays on screen.
    void render_high_accuracy(const std::string& field_function) {
        set_body_or_throw(field_function);
        check(linon_render(view_, base_flags() | LINON_RENDER_ENHANCE), "linon_render");
        needs_redraw = false;
    }
    // src/main_view.rs:804-829: sample rays start on the outline of the diverging regions.
    void render_outline_rays(const std::string& field_function) {
        set_body_or_throw(field_function);
        check(linon_render(view_, base_flags() | LINON_RENDER_OUTLINE), "linon_render");
        needs_redraw = false;
    }
    // src/main_view.rs:831-849: march dispatch + overlay dispatch.
    void render() {
        check(linon_render(view_, base_flags()), "linon_render");
        needs_redraw = false;
    }

    // ---- outputs (the reference keeps them in textures for egui / ReferenceView) ----
    std::vector<uint8_t> read_bytes(int buffer) {
        std::vector<uint8_t> out(linon_buffer_bytes(view_, buffer));
        check(linon_read(view_, buffer, out.data(), out.size()), "linon_read");
        return out;
    }
    linon_view* handle() { return view_; }
    ArcballCamera& camera() { return camera_; }
    uint32_t width() const { return width_; }
    uint32_t height() const { return height_; }
    const std::string& field_function() const { return field_function_; }
    bool adaptive_sampling() const { return adaptive_sampling_; }

private:
    bool adaptive_sampling_ = true;
    uint32_t downscale_factor_;
    uint32_t width_, height_;
    std::string field_function_;
    std::string shader_src_;  // text of the last accepted main_view.wgsl ("" = stock module), src/main_view.rs:75
    ArcballCamera camera_;
    linon_view* view_ = nullptr;

    uint32_t base_flags() const { return adaptive_sampling_ ? 0u : (uint32_t)LINON_RENDER_ADAPTIVE_OFF; }
    void set_body_or_throw(const std::string& field_function) {
        if (field_function == field_function_) return;
        std::string err;
        if (!reload_shader(field_function, &err)) throw std::runtime_error(err);  // the reference unwrap()s here
    }
    void check(int rc, const char* what) {
        if (rc != LINON_OK)
            throw std::runtime_error(std::string(what) + ": " + linon_last_error(view_));
    }
};

}  // namespace linon
