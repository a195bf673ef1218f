//! FFI declarations of `include/linon_b200.h` — UNVERIFIED source (no Rust toolchain in the build image).
//! Every item cites the declaration it mirrors; sizes are asserted at compile time.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct linon_view {
    _private: [u8; 0],
}

/// `linon_camera` == `CameraUniform`, src/main_view.rs:12-19 (112 bytes).
#[repr(C)]
#[derive(Copy, Clone, Debug, bytemuck::Pod, bytemuck::Zeroable)]
pub struct linon_camera {
    pub origin: [f32; 4],
    pub view_direction: [f32; 4],
    pub up: [f32; 4],
    pub view_matrix: [[f32; 4]; 4],
}

/// `linon_settings` == `Settings`, src/main_view.rs:43-51 (24 bytes).
#[repr(C)]
#[derive(Copy, Clone, Debug, bytemuck::Pod, bytemuck::Zeroable)]
pub struct linon_settings {
    pub field_weight: f32,
    pub mouse_pos: [f32; 2],
    pub overlay_mode: i32,
    pub central_difference_delta: i32,
    pub lyapunov_scaling: f32,
}

#[repr(C)]
#[derive(Copy, Clone, Debug, Default)]
pub struct linon_counters {
    pub rays: u64,
    pub accepted_steps: u64,
    pub prehit_steps: u64,
    pub rejected_steps: u64,
    pub face_tests: u64,
    pub plane_tests: u64,
    pub culling_checks: u64,
}

#[repr(C)]
#[derive(Copy, Clone, Debug, Default)]
pub struct linon_timings {
    pub march_ms: f32,
    pub overlay_ms: f32,
    pub sample_rays_ms: f32,
    pub outline_search_ms: f32,
    pub jit_ms: f32,
    pub march_regs: i32,
    pub march_grid: i32,
    pub march_block: i32,
    pub march_phase_a_ms: f32,
    pub march_kernel: i32,
    pub overlay_kernel: i32,
}

const _: () = assert!(std::mem::size_of::<linon_camera>() == 112);
const _: () = assert!(std::mem::size_of::<linon_settings>() == 24);

pub const LINON_OK: c_int = 0;
pub const LINON_ERR_INVALID: c_int = 1;
pub const LINON_ERR_CUDA: c_int = 2;
pub const LINON_ERR_COMPILE: c_int = 3;
pub const LINON_ERR_NO_DRIVER: c_int = 4;
pub const LINON_ERR_SIZE: c_int = 5;
pub const LINON_ERR_UNSUPPORTED: c_int = 6;

pub const LINON_RENDER_DEFAULT: u32 = 0;
pub const LINON_RENDER_ENHANCE: u32 = 1; // render_high_accuracy, src/main_view.rs:778-802
pub const LINON_RENDER_OUTLINE: u32 = 2; // render_outline_rays, :804-829
pub const LINON_RENDER_ADAPTIVE_OFF: u32 = 4; // disable_adaptive_sampling, :765-776
pub const LINON_RENDER_NO_OVERLAY: u32 = 8;
pub const LINON_RENDER_NO_MARCH: u32 = 16;
pub const LINON_RENDER_NO_SAMPLE_RAYS: u32 = 32;
pub const LINON_RENDER_KEEP_TIMINGS: u32 = 64;

pub const LINON_SHARD_PEER_STORES: c_int = 0; // every shard's march stores into the root's buffers over NVLink
pub const LINON_SHARD_NCCL_GATHER: c_int = 1; // one grouped ncclSend / ncclRecv to the root after the march

#[repr(C)]
pub struct linon_sharded {
    _private: [u8; 0],
}

pub const LINON_BUF_FINAL_RGBA8: c_int = 0; // `texture`
pub const LINON_BUF_RAYCAST_RGBA8: c_int = 1; // `ray_casting_texture`
pub const LINON_BUF_ENDPOINT_F32X4: c_int = 2; // `mapping_texture`
pub const LINON_BUF_EXPONENTS_F32: c_int = 3; // `exponents_buffer`
pub const LINON_BUF_HIT_ID_I32: c_int = 4;
pub const LINON_BUF_RAY_SAMPLES: c_int = 5; // `ray_samples_buffer`, 800 x 32 B (src/ray_samples.rs:8-27)
pub const LINON_BUF_SAMPLE_POS: c_int = 6;

// linon_set_option keys (include/linon_b200.h, enum linon_option); defaults are the shipped configuration
pub const LINON_OPT_MATH: c_int = 0;
pub const LINON_OPT_KERNEL: c_int = 1;
pub const LINON_OPT_BLOCK: c_int = 2;
pub const LINON_OPT_CTAS_PER_SM: c_int = 3;
pub const LINON_OPT_REFILL_MIN: c_int = 4;
pub const LINON_OPT_FACE_UNROLL: c_int = 5;
pub const LINON_OPT_MAX_REGS: c_int = 6;
pub const LINON_OPT_PROFILE: c_int = 7;
pub const LINON_OPT_CULL: c_int = 8;
pub const LINON_OPT_PAIR_MIN_BLOCKS: c_int = 9;
pub const LINON_OPT_SPECIALIZE: c_int = 10;
pub const LINON_OPT_FORCE_BVH: c_int = 11;
pub const LINON_OPT_QUIET_ZONE: c_int = 12; // 2 (default): fixed-step and adaptive frames
pub const LINON_OPT_QUIET_RADIUS: c_int = 13; // 1500 (default); 0 = the saturation radius
pub const LINON_OPT_OVERLAY_KERNEL: c_int = 14; // 0 staged, 1 direct, 2 TMA tile (default)
pub const LINON_OPT_COLLAPSE_RUNS: c_int = 15; // 1 (default): long straight runs taken binade by binade

extern "C" {
    pub fn linon_create(device: c_int, width: u32, height: u32, out: *mut *mut linon_view) -> c_int;
    pub fn linon_destroy(v: *mut linon_view);
    pub fn linon_set_mesh(v: *mut linon_view, verts: *const f32, n_verts: u32, faces: *const u32, n_faces: u32) -> c_int;
    pub fn linon_set_camera(v: *mut linon_view, camera: *const linon_camera) -> c_int;
    pub fn linon_set_settings(v: *mut linon_view, settings: *const linon_settings) -> c_int;
    pub fn linon_resize(v: *mut linon_view, width: u32, height: u32) -> c_int;
    pub fn linon_set_field_function(v: *mut linon_view, wgsl_body: *const c_char, err: *mut c_char, err_len: usize) -> c_int;
    /// reload_shader(device, Some(new_src), ..): the helper layer of an edited main_view.wgsl; NULL / "" = stock module
    pub fn linon_reload_shader_source(v: *mut linon_view, wgsl_module_source: *const c_char, err: *mut c_char, err_len: usize) -> c_int;
    /// reload_shader(device, Some(new_src), field_function) as one step: module and body validated and committed together
    pub fn linon_reload_shader(v: *mut linon_view, wgsl_module_source: *const c_char, wgsl_body: *const c_char, err: *mut c_char,
                               err_len: usize) -> c_int;
    pub fn linon_module_overrides(v: *mut linon_view, names: *mut c_char, names_len: usize) -> c_int;
    pub fn linon_render(v: *mut linon_view, flags: u32) -> c_int;
    pub fn linon_sync(v: *mut linon_view) -> c_int;
    pub fn linon_read(v: *mut linon_view, buffer: c_int, dst: *mut c_void, dst_bytes: usize) -> c_int;
    pub fn linon_buffer_bytes(v: *const linon_view, buffer: c_int) -> usize;
    pub fn linon_device_ptr(v: *mut linon_view, buffer: c_int) -> *mut c_void;
    pub fn linon_stream(v: *mut linon_view) -> *mut c_void;
    pub fn linon_set_shard(v: *mut linon_view, rank: u32, world: u32, rows_per_block: u32) -> c_int;
    pub fn linon_redirect_output(v: *mut linon_view, buffer: c_int, device_ptr: *mut c_void) -> c_int;
    pub fn linon_ipc_export(v: *mut linon_view, buffer: c_int, handle64: *mut c_void) -> c_int;
    pub fn linon_ipc_open(v: *mut linon_view, handle64: *const c_void, device_ptr: *mut *mut c_void) -> c_int;
    pub fn linon_set_option(v: *mut linon_view, option: c_int, value: c_int) -> c_int;
    pub fn linon_get_counters(v: *mut linon_view, out: *mut linon_counters) -> c_int;
    pub fn linon_get_timings(v: *mut linon_view, out: *mut linon_timings) -> c_int;
    pub fn linon_last_error(v: *const linon_view) -> *const c_char;
    pub fn linon_translate_wgsl(wgsl_body: *const c_char, out: *mut c_char, out_len: usize, err: *mut c_char, err_len: usize) -> c_int;
    pub fn linon_compile_check(wgsl_body: *const c_char, flags: u32, math_mode: c_int, n_faces: u32, cubin_bytes: *mut usize,
                               err: *mut c_char, err_len: usize) -> c_int;
    pub fn linon_measure_fp32_peak(device: c_int, tflops: *mut f64, sm_count: *mut c_int) -> c_int;
    pub fn linon_march_cubin_hash(v: *mut linon_view, out: *mut u64) -> c_int;
    pub fn linon_nvrtc_info(path: *mut c_char, path_len: usize, major: *mut c_int, minor: *mut c_int) -> c_int;

    // one frame on several B200s of a node, from one host thread (devices[0] is the root)
    pub fn linon_sharded_create(devices: *const c_int, n_devices: u32, width: u32, height: u32, rows_per_block: u32, mode: c_int,
                                out: *mut *mut linon_sharded) -> c_int;
    pub fn linon_sharded_destroy(s: *mut linon_sharded);
    pub fn linon_sharded_world(s: *const linon_sharded) -> u32;
    pub fn linon_sharded_view(s: *mut linon_sharded, rank: u32) -> *mut linon_view;
    pub fn linon_sharded_set_mesh(s: *mut linon_sharded, verts: *const f32, n_verts: u32, faces: *const u32, n_faces: u32) -> c_int;
    pub fn linon_sharded_set_camera(s: *mut linon_sharded, camera: *const linon_camera) -> c_int;
    pub fn linon_sharded_set_settings(s: *mut linon_sharded, settings: *const linon_settings) -> c_int;
    pub fn linon_sharded_set_field_function(s: *mut linon_sharded, wgsl_body: *const c_char, err: *mut c_char, err_len: usize) -> c_int;
    pub fn linon_sharded_set_option(s: *mut linon_sharded, option: c_int, value: c_int) -> c_int;
    pub fn linon_sharded_resize(s: *mut linon_sharded, width: u32, height: u32) -> c_int;
    pub fn linon_sharded_render(s: *mut linon_sharded, flags: u32) -> c_int;
    pub fn linon_sharded_sync(s: *mut linon_sharded) -> c_int;
    pub fn linon_sharded_last_error(s: *const linon_sharded) -> *const c_char;

    // display hand-off without a copy: LINON_BUF_FINAL_RGBA8 / LINON_BUF_RAY_SAMPLES as an exportable allocation (POSIX fd,
    // importable as VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT); the caller closes the descriptor
    pub fn linon_export_buffer_fd(v: *mut linon_view, buffer: c_int, fd: *mut c_int, allocation_bytes: *mut usize) -> c_int;
    pub fn linon_import_buffer_fd(device: c_int, fd: c_int, allocation_bytes: usize, device_ptr: *mut *mut c_void,
                                  import_handle: *mut *mut c_void) -> c_int;
    pub fn linon_release_import(import_handle: *mut c_void) -> c_int;
    pub fn linon_copy_to_host(device: c_int, device_ptr: *const c_void, dst: *mut c_void, bytes: usize) -> c_int;
}
