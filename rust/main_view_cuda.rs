//! Replacement for the device-facing half of `src/main_view.rs` on `liblinon_b200.so` — UNVERIFIED source
//! (no Rust toolchain in the build image).  Method names, argument meaning and error behaviour follow the reference;
//! the egui widget glue (`show`, pointer / zoom handlers, :496-545, :692-750) stays as it is and calls these.
use std::ffi::{CStr, CString};

use cgmath::Vector3;
use linon_cuda_sys as sys;

use crate::arcball::ArcballCamera;
use crate::functions::PredefinedFunction;

pub use sys::linon_settings as Settings; // byte-identical with the reference's Settings (src/main_view.rs:43-51)

fn camera_uniform(camera: &ArcballCamera<f32>) -> sys::linon_camera {
    // CameraUniform::moving, src/main_view.rs:21-31
    let o = camera.eye_pos();
    let d = camera.eye_dir();
    let u = camera.up_dir();
    sys::linon_camera {
        origin: [o.x, o.y, o.z, 0.0],
        view_direction: [d.x, d.y, d.z, 0.0],
        up: [u.x, u.y, u.z, 0.0],
        view_matrix: [[1.0, 0.0, 0.0, 0.0], [0.0, 1.0, 0.0, 0.0], [0.0, 0.0, 1.0, 0.0], [0.0, 0.0, 0.0, 1.0]],
    }
}

pub struct MainView {
    view: *mut sys::linon_view,
    camera: ArcballCamera<f32>,
    pub width: u32,
    pub height: u32,
    adaptive_sampling: bool,
    pub needs_redraw: bool, // src/main_view.rs:78
}

impl MainView {
    fn check(&self, rc: i32) -> Result<(), String> {
        if rc == sys::LINON_OK {
            return Ok(());
        }
        let msg = unsafe { CStr::from_ptr(sys::linon_last_error(self.view)) };
        Err(msg.to_string_lossy().into_owned())
    }

    /// MainView::new, src/main_view.rs:82-458.  `vertices` / `indices` are the CPU-side arrays Application::new already
    /// holds (src/application.rs:105-122); the wgpu storage buffers are no longer needed by this view.
    pub fn new(vertices: &[f32], indices: &[u32], center: Vector3<f32>, width: u32, height: u32, discrete_gpu: bool) -> Result<Self, String> {
        let downscale = if discrete_gpu { 1 } else { 2 }; // :93
        let (w, h) = (width / downscale, height / downscale);
        let mut view = std::ptr::null_mut();
        let rc = unsafe { sys::linon_create(0, w, h, &mut view) };
        if rc != sys::LINON_OK {
            let msg = unsafe { CStr::from_ptr(sys::linon_last_error(std::ptr::null())) };
            return Err(msg.to_string_lossy().into_owned()); // no CUDA device: there is no CPU fallback
        }
        let mut camera = ArcballCamera::new(center, 1.0, [w as f32, h as f32]);
        camera.zoom(-1.0, 1.0); // :153-155
        let mut mv = MainView { view, camera, width: w, height: h, adaptive_sampling: true, needs_redraw: true };
        mv.check(unsafe { sys::linon_set_mesh(view, vertices.as_ptr(), (vertices.len() / 3) as u32, indices.as_ptr(), (indices.len() / 3) as u32) })?;
        mv.update_camera()?;
        mv.reload_shader(PredefinedFunction::MirageSphericalSigmoid.to_code())?; // :100
        Ok(mv)
    }

    /// update_model, :460-488
    pub fn update_model(&mut self, vertices: &[f32], indices: &[u32], center: Vector3<f32>) -> Result<(), String> {
        self.check(unsafe { sys::linon_set_mesh(self.view, vertices.as_ptr(), (vertices.len() / 3) as u32, indices.as_ptr(), (indices.len() / 3) as u32) })?;
        self.camera = ArcballCamera::new(center, 1.0, [self.width as f32, self.height as f32]);
        self.camera.zoom(-1.0, 1.0);
        self.update_camera()
    }

    /// update_camera, :490-494
    pub fn update_camera(&mut self) -> Result<(), String> {
        let u = camera_uniform(&self.camera);
        self.needs_redraw = true;
        self.check(unsafe { sys::linon_set_camera(self.view, &u) })
    }

    /// update_settings, :547-550
    pub fn update_settings(&mut self, settings: Settings) -> Result<(), String> {
        self.needs_redraw = true;
        self.check(unsafe { sys::linon_set_settings(self.view, &settings) })
    }

    /// resize_texture, :552-647 (the caller re-registers the egui texture and calls upload_to_wgpu after the next render)
    pub fn resize_texture(&mut self, width: u32, height: u32) -> Result<(), String> {
        self.width = width;
        self.height = height;
        self.camera.update_screen(width as f32, height as f32);
        self.needs_redraw = true;
        self.check(unsafe { sys::linon_resize(self.view, width, height) })
    }

    /// reload_shader(device, None, body), :649-690: Err carries the diagnostic, the previous function stays active (:677-679)
    pub fn reload_shader(&mut self, field_function: String) -> Result<(), String> {
        let body = CString::new(field_function).map_err(|e| e.to_string())?;
        let mut err = vec![0u8; 16384];
        let rc = unsafe { sys::linon_set_field_function(self.view, body.as_ptr(), err.as_mut_ptr() as *mut _, err.len()) };
        if rc != sys::LINON_OK {
            let end = err.iter().position(|&b| b == 0).unwrap_or(err.len());
            return Err(String::from_utf8_lossy(&err[..end]).into_owned());
        }
        self.needs_redraw = true;
        Ok(())
    }

    /// reload_shader(device, Some(new_src), field_function), :649-690, as Application::reload_compute_shader calls it
    /// (src/application.rs:228-232) when the watcher on src/main_view.wgsl fires (src/main.rs:206-222): the CUDA path
    /// installs the module's helper layer (new / edited helper functions and constants) and its march-level constants (eps,
    /// backface_culling, use_lighting, linear_mode, the light constants, adaptive_sampling, sample_outline_rays, the step /
    /// horizon / step-control literals) and the shading block of ray_color (translated from the user's text); any other edit
    /// to the march is an Err naming the declaration, and nothing changes
    /// on any Err (:677-679).
    pub fn reload_compute_shader(&mut self, new_src: &str, field_function: &str) -> Result<(), String> {
        let src = CString::new(new_src).map_err(|e| e.to_string())?;
        let body = CString::new(field_function).map_err(|e| e.to_string())?;
        let mut err = vec![0u8; 16384];
        // module and body together, like :649-690 compiles them: both are committed or neither is
        let rc = unsafe { sys::linon_reload_shader(self.view, src.as_ptr(), body.as_ptr(), err.as_mut_ptr() as *mut _, err.len()) };
        if rc != sys::LINON_OK {
            let end = err.iter().position(|&b| b == 0).unwrap_or(err.len());
            return Err(String::from_utf8_lossy(&err[..end]).into_owned());
        }
        self.needs_redraw = true;
        Ok(())
    }

    /// :752-776 — a flag instead of the (dead at HEAD) string replacement on the WGSL source
    pub fn enable_adaptive_sampling(&mut self) { self.adaptive_sampling = true; self.needs_redraw = true; }
    pub fn disable_adaptive_sampling(&mut self) { self.adaptive_sampling = false; self.needs_redraw = true; }

    fn flags(&self) -> u32 {
        if self.adaptive_sampling { sys::LINON_RENDER_DEFAULT } else { sys::LINON_RENDER_ADAPTIVE_OFF }
    }

    /// render_high_accuracy, :778-802: the enhanced frame stays until the next change (needs_redraw = false)
    pub fn render_high_accuracy(&mut self) -> Result<(), String> {
        self.check(unsafe { sys::linon_render(self.view, self.flags() | sys::LINON_RENDER_ENHANCE) })?;
        self.needs_redraw = false;
        Ok(())
    }

    /// render_outline_rays, :804-829
    pub fn render_outline_rays(&mut self) -> Result<(), String> {
        self.check(unsafe { sys::linon_render(self.view, self.flags() | sys::LINON_RENDER_OUTLINE) })?;
        self.needs_redraw = false;
        Ok(())
    }

    /// render, :831-849 (the two compute dispatches)
    pub fn render(&mut self) -> Result<(), String> {
        self.check(unsafe { sys::linon_render(self.view, self.flags()) })?;
        self.needs_redraw = false;
        Ok(())
    }

    /// Staged display hand-off: final image into the egui-registered texture (src/main_view.rs:107-115) and the
    /// ray-sample vertices into ReferenceView's vertex buffer (src/application.rs:125-130, src/reference_view.rs:381-388).
    pub fn upload_to_wgpu(&mut self, queue: &wgpu::Queue, texture: &wgpu::Texture, ray_samples_buffer: &wgpu::Buffer) -> Result<(), String> {
        let mut rgba = vec![0u8; (self.width * self.height * 4) as usize];
        self.check(unsafe { sys::linon_read(self.view, sys::LINON_BUF_FINAL_RGBA8, rgba.as_mut_ptr() as *mut _, rgba.len()) })?;
        queue.write_texture(
            wgpu::ImageCopyTexture { texture, mip_level: 0, origin: wgpu::Origin3d::ZERO, aspect: wgpu::TextureAspect::All },
            &rgba,
            wgpu::ImageDataLayout { offset: 0, bytes_per_row: Some(4 * self.width), rows_per_image: Some(self.height) },
            wgpu::Extent3d { width: self.width, height: self.height, depth_or_array_layers: 1 },
        );
        let mut samples = vec![0u8; 800 * 32];
        self.check(unsafe { sys::linon_read(self.view, sys::LINON_BUF_RAY_SAMPLES, samples.as_mut_ptr() as *mut _, samples.len()) })?;
        queue.write_buffer(ray_samples_buffer, 0, &samples);
        Ok(())
    }
}

/// File descriptors of the two buffers the application shows, for a host that imports them into its graphics API instead of
/// staging them through `upload_to_wgpu`: the final image (re-export after every resize_texture,
/// like the reference re-registers its texture, :569-570) and the ray samples (once).  The caller owns the descriptors.
/// With wgpu 22's Vulkan backend: `vkAllocateMemory` with `VkImportMemoryFdInfoKHR { handleType: OPAQUE_FD, fd }` of
/// `allocation_bytes`, `vkBindBufferMemory` on a linear `VkBuffer`, `wgpu_hal::vulkan::Device::buffer_from_raw`, then
/// `wgpu::Device::create_buffer_from_hal`.  UNVERIFIED source like the rest of this file (no Rust toolchain in the image);
/// the C entry point itself is GPU-tested from a second process (tests/test_gpu_handoff.py).
pub struct SharedBuffers {
    pub final_rgba8_fd: i32,
    pub final_rgba8_bytes: usize,
    pub ray_samples_fd: i32,
    pub ray_samples_bytes: usize,
}

impl MainView {
    pub fn import_shared(&mut self) -> Result<SharedBuffers, String> {
        let (mut fd_a, mut n_a, mut fd_b, mut n_b) = (-1i32, 0usize, -1i32, 0usize);
        self.check(unsafe { sys::linon_export_buffer_fd(self.view, sys::LINON_BUF_FINAL_RGBA8, &mut fd_a, &mut n_a) })?;
        self.check(unsafe { sys::linon_export_buffer_fd(self.view, sys::LINON_BUF_RAY_SAMPLES, &mut fd_b, &mut n_b) })?;
        Ok(SharedBuffers { final_rgba8_fd: fd_a, final_rgba8_bytes: n_a, ray_samples_fd: fd_b, ray_samples_bytes: n_b })
    }
}

impl Drop for MainView {
    fn drop(&mut self) {
        unsafe { sys::linon_destroy(self.view) };
    }
}
