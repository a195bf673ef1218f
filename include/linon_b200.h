/* linon_b200.h — C ABI of the B200-native ray-march path of linon.
 *
 * This is the drop-in boundary: the entry points are what a Rust `extern "C"` block in the reference's
 * src/main_view.rs would bind in place of its wgpu compute pipelines.  Plain pointers and
 * sizes only; no C++ or torch types.  Every function returns 0 on success and a non-zero LINON_ERR_* code on
 * failure (never aborts, never throws across the boundary); linon_last_error() returns the message.
 *
 * Threading: a handle is not thread-safe; use it from one thread (the reference calls MainView only from the
 * winit event-loop thread, src/main.rs:125-195).  Work is enqueued on the handle's CUDA stream; readers
 * synchronise that stream.  Host pointers are borrowed only for the duration of a call.
 *
 * Image layout: every per-pixel buffer is [height][width] row-major in TEXTURE space (row 0 = top of the image,
 * because the shaders write at coords = (gid.x, H-1-gid.y), src/main_view.wgsl:342).
 */
#ifndef LINON_B200_H
#define LINON_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct linon_view linon_view; /* replaces struct MainView, src/main_view.rs:53-79 */

/* CameraUniform, src/main_view.rs:12-19 / struct Camera, src/main_view.wgsl:7-12.  112 bytes. */
typedef struct linon_camera {
    float origin[4];
    float view_direction[4];
    float up[4];
    float view_matrix[16]; /* identity in the reference (CameraUniform::moving); unused on the path */
} linon_camera;

/* Settings, src/main_view.rs:43-51 / src/main_view.wgsl:16-23.  24 bytes. */
typedef struct linon_settings {
    float field_weight;
    float mouse_pos[2];
    int32_t overlay_mode;             /* 0 = disabled, 1 = Lyapunov exponents */
    int32_t central_difference_delta; /* 1..10, src/application.rs:297 */
    float lyapunov_scaling;
} linon_settings;

enum {
    LINON_OK = 0,
    LINON_ERR_INVALID = 1,     /* bad argument / call order */
    LINON_ERR_CUDA = 2,        /* CUDA runtime or driver error */
    LINON_ERR_COMPILE = 3,     /* field function rejected (translator or NVRTC); previous kernel stays live */
    LINON_ERR_NO_DRIVER = 4,   /* libcuda.so.1 / a CUDA device is not available: there is NO CPU fallback */
    LINON_ERR_SIZE = 5,        /* caller buffer too small */
    LINON_ERR_UNSUPPORTED = 6  /* linon_reload_shader_source: the edited module changes the march itself, not its helper layer */
};

/* render flags.  The reference toggles these by string replacement on the WGSL source
 * (src/main_view.rs:752-829); two of the three replacements are dead at HEAD — the flags implement
 * the intended behaviour (report/Sources/2-Approach.tex:192,218). */
enum {
    LINON_RENDER_DEFAULT = 0,
    LINON_RENDER_ENHANCE = 1,         /* h_initial 0.1 -> 0.001, MainView::render_high_accuracy, :778-802 */
    LINON_RENDER_OUTLINE = 2,         /* sample rays on the outline of exponents >= 0.8, render_outline_rays, :804-829 */
    LINON_RENDER_ADAPTIVE_OFF = 4,    /* disable_adaptive_sampling (while dragging), :765-776 */
    LINON_RENDER_NO_OVERLAY = 8,      /* skip the overlay dispatch (sharded march; overlay runs after the gather) */
    LINON_RENDER_NO_MARCH = 16,       /* overlay (+ sample rays) only, on buffers already in place */
    LINON_RENDER_NO_SAMPLE_RAYS = 32, /* skip the 8 reference-view rays */
    LINON_RENDER_KEEP_TIMINGS = 64    /* linon_get_timings keeps the passes of the previous linon_render call (a frame rendered
                                         by two calls: march, then overlay on the assembled frame) */
};

/* buffers (linon_read / linon_device_ptr).  Element sizes in bytes per pixel unless noted. */
enum {
    LINON_BUF_FINAL_RGBA8 = 0,   /* `texture`: overlay output shown on screen, 4 B     (src/main_view.rs:107-113) */
    LINON_BUF_RAYCAST_RGBA8 = 1, /* `ray_casting_texture`: march colour, 4 B            (:117-123) */
    LINON_BUF_ENDPOINT_F32X4 = 2,/* `mapping_texture`: ray endpoint (x,y,z,1), 16 B     (:125-131) */
    LINON_BUF_EXPONENTS_F32 = 3, /* `exponents_buffer`, 4 B, stride = width (:146-151) */
    LINON_BUF_HIT_ID_I32 = 4,    /* NEW: index of the first face hit, -1 = none, 4 B */
    LINON_BUF_RAY_SAMPLES = 5,   /* `ray_samples_buffer`: 800 x {vec4 position, vec4 colour} = 25,600 B total
                                    (src/ray_samples.rs:8-27, src/application.rs:125-130) */
    LINON_BUF_SAMPLE_POS = 6     /* the 8 normalised (x,y) start positions used for the sample rays, 64 B total */
};

/* tunables (linon_set_option) */
enum {
    LINON_OPT_MATH = 0,          /* 0 strict (no FMA contraction, IEEE div/sqrt; comparable with the CPU oracle bit
                                    for bit on bodies without transcendentals), 1 fused (FMA contraction),
                                    2 fast (--use_fast_math: MUFU rcp/rsqrt/ex2/sin/cos, ftz).  Default 2: measured 2x faster
                                    than fused and inside the north-star tolerances (tests/test_gpu_parity.py). */
    LINON_OPT_KERNEL = 1,        /* 0 one thread per pixel in 8x4 tiles, 1 persistent warps: pre-hit / post-hit phase kernels
                                    with lane refill, 2 two rays per thread on the packed fp32 instructions of sm_100
                                    (FFMA2/FMUL2/FADD2) in 16x4 tiles; a user body that does not instantiate for two rays at
                                    once (data-dependent control flow) falls back to 0 automatically
                                    (linon_timings.march_kernel tells which kernel ran), 3 persistent warps in ONE phase: a lane
                                    carries a ray from its first step to t > 5 and then takes the next slot of a chunked cursor
                                    (made for adaptive frames; within +-15 % of the tile kernel there).  Default 0: measured fastest - FFMA2
                                    occupies the issue port for two cycles on B200, so packing does not relieve the
                                    issue-slot bound (profiles/r1_ffma2_microbench.md) */
    LINON_OPT_BLOCK = 2,         /* threads per CTA of the march kernel (multiple of 32). Default 256 */
    LINON_OPT_CTAS_PER_SM = 3,   /* persistent grid = this x #SM; 0 = occupancy query. Default 0 */
    LINON_OPT_REFILL_MIN = 4,    /* idle lanes a warp tolerates before refilling. Default 1 */
    LINON_OPT_FACE_UNROLL = 5,   /* unroll of the face loop. Default 4 */
    LINON_OPT_MAX_REGS = 6,      /* NVRTC --maxrregcount (0 = none). Default 0 */
    LINON_OPT_PROFILE = 7,       /* 1: persistent kernel accumulates clock64 section timers (linon_debug_profile) */
    LINON_OPT_FORCE_BVH = 11,    /* 1: stage the mesh as faces + bounding-volume hierarchy in global memory even when it would fit
                                    shared memory (meshes above 1024 faces always do).  Results are identical; for tests */
    LINON_OPT_SPECIALIZE = 10,   /* 1 (default): in fast math mode a body of the MIRAGE FAMILY - one of the four predefined mirage
                                    functions (src/functions.rs:36-94) up to comments, white space and the values of its
                                    literals (t_env, t_src, max_dist, centre / plane_p0, the sigmoid's constants; matched
                                    token by token, csrc/host/mirage_match.hpp) - runs the hand-fused restatement with those
                                    constants (csrc/device/mirage_fused.cuh: one reciprocal for eta, no sigmoid reciprocal,
                                    shared inverse square root, Runge-Kutta stage 1 = identity).  0: always the generic
                                    translation */
    LINON_OPT_PAIR_MIN_BLOCKS = 9, /* resident CTAs / SM requested for the pair kernel (register cap); 0 = default (512/block) */
    LINON_OPT_QUIET_ZONE = 12,   /* 1: fixed-step frames (Enhance, adaptive off) of a predefined mirage body in fast math mode take
                                    straight steps - velocity = direction, no field evaluation - while every lane of the warp is
                                    outside the body's zone of influence, where the reference's own fp32 evaluation returns
                                    eta = 1 exactly and refraction(t, t, v, n) = v up to its rounding noise (csrc/device/
                                    mirage_fused.cuh: quiet_steps).  Every Runge-Kutta step is still executed and counted.
                                    2: adaptive frames as well - per attempt, velocity = direction when every stage point lies
                                    outside the zone (diff is exactly 0 there, so the reference's noise-driven rejections
                                    outside the zone, SURVEY F20, have no counterpart).
                                    Default 2 since round 2: measured on B200 (C3 march 191 -> 13 ms at level 1; C2 adaptive
                                    frames 7.4 -> 1.3 ms MirageSpherical, 9.1 -> 5.2 ms MirageSphericalSigmoid at level 2) and
                                    graded against the oracle at full size (tests/test_gpu_fullsize.py, bench.py `parity` and
                                    `configs`).  0: every step evaluates the field (the round-1 kernel) */
    LINON_OPT_QUIET_RADIUS = 13, /* radius of the quiet zone of the SIGMOID mirage profiles in permille of their max_dist.
                                    1500 (default since round 2): between 1.5 and 1.8866 max_dist the reference's fp32 eta is 1 +- its
                                    own rounding noise - the real-number signal there is |eta - 1| <= 1e-9, 50x below the fp32
                                    quantum - and the deflection summed over the excluded shell is <= 1.85e-4 exp(6 - 12 r) =
                                    1e-9 rad (mirage_fused.cuh); graded against the oracle at full size in every bench line
                                    (C3: 12.9 -> 10.4 ms, same hit / colour / endpoint agreement as the saturation radius).
                                    0: 1920, where the reference's own fp32 sigmoid saturates (>= 1886.6) and its evaluation IS the
                                    identity.  1250 ... 4000 otherwise */
    LINON_OPT_OVERLAY_KERNEL = 14, /* overlay stencil (src/overlay.wgsl:99-146), bit-identical results in every variant.
                                    0: 32x16 tile + halo of the endpoints staged in shared memory by the threads (round 1).
                                    1: every pixel fetches its four taps through the read-only path, no shared memory, no
                                    barrier (csrc/device/overlay_direct.cuh; 53 vs 91 us at C3).
                                    2 (default): 32x32 tile + halo by ONE TMA load per CTA (cp.async.bulk.tensor.2d + mbarrier),
                                    taps are four LDS.128 at compile-time offsets (csrc/device/overlay_tile.cuh); runs variant 1
                                    when the driver cannot encode the tensor map or delta is outside 1..10 */
    LINON_OPT_COLLAPSE_RUNS = 15, /* quiet-zone runs of 160 steps or more.  1 (default): taken binade by binade - inside one binade of a
                                    coordinate the rounded increment of p <- fma(h, dir, p) is a constant multiple of the ulp, so a
                                    stretch of k steps is one FMA, q + k delta, bit for bit (march_kernels.cuh: advance_coord; pinned
                                    against the step-by-step chain on random and adversarial inputs and by the frames of the kernel
                                    without runs; C3 march 8.4 -> 6.4 ms).  0: every straight step is executed as one FFMA per
                                    coordinate - the variant the FP32 roofline fraction of DESIGN.md section 5 is quoted for beside
                                    the default; every bench line times and grades both */
    LINON_OPT_CULL = 8           /* 1 (default): exact culling of the face loop - a per-ray lower bound on the distance to the
                                    nearest face plane skips the loop entirely while no plane is in reach, and plane groups
                                    out of reach are skipped inside it; results are bit-identical with 0.
                                    0: test every face on every pre-hit step like the reference */
};

typedef struct linon_counters { /* per-frame totals over every ray of the last march */
    uint64_t rays;              /* rays marched */
    uint64_t accepted_steps;    /* accepted RK4 steps */
    uint64_t prehit_steps;      /* accepted steps that also ran the all-face intersection */
    uint64_t rejected_steps;    /* RK4 attempts rejected by adaptive sampling */
    uint64_t face_tests;        /* ray-triangle tests actually executed (36 per pre-hit step without culling) */
    uint64_t plane_tests;       /* plane-distance culling tests executed (0 when culling is off) */
    uint64_t culling_checks;    /* pre-hit steps whose safe-distance bound expired, i.e. that ran the plane-group loop */
} linon_counters;

typedef struct linon_timings { /* CUDA-event times of the last render, milliseconds (0 when the pass did not run) */
    float march_ms;
    float overlay_ms;
    float sample_rays_ms;
    float outline_search_ms;
    float jit_ms;               /* host time spent in translate + NVRTC for this render (0 on a cache hit) */
    int32_t march_regs;         /* registers per thread of the march kernel */
    int32_t march_grid;         /* CTAs launched */
    int32_t march_block;        /* threads per CTA */
    float march_phase_a_ms;     /* persistent kernel: time of the pre-hit phase (march_ms covers both phases) */
    int32_t march_kernel;       /* the kernel that ran: 0 tiles, 1 persistent (two phases), 2 pairs, 3 persistent (one phase) */
    int32_t overlay_kernel;     /* the overlay kernel that ran: 0 hand-staged tile, 1 direct, 2 TMA tile (LINON_OPT_OVERLAY_KERNEL;
                                   2 falls back to 1 when no tensor map can be made) */
} linon_timings;

/* MainView::new (src/main_view.rs:82-458): allocates textures/buffers for width x height on CUDA device
 * `device`, installs the reference's start-up state: Settings {1.0, [0.5,0.6], 0, 1, 50.0} (:133-139), the
 * MirageSphericalSigmoid body (:100) and NO mesh/camera (call linon_set_mesh / linon_set_camera).
 * The caller applies the discrete_gpu down-scale (:93) itself. */
int linon_create(int device, uint32_t width, uint32_t height, linon_view** out);
void linon_destroy(linon_view* v);

/* MainView::update_model (:460-488) + the vertices/faces storage buffers (src/application.rs:112-122).
 * verts: n_verts x 3 f32, faces: n_faces x 3 u32 (copied).  Up to 1024 faces are staged per CTA in shared memory with
 * plane-group culling; larger meshes (an OBJ dropped on the window, src/main.rs:139-157) are staged in global memory with
 * a bounding-volume hierarchy whose traversal is exact: the closest hit and the lowest-index tie-break of the
 * reference's loop over every face (src/main_view.wgsl:182-198) are preserved. */
int linon_set_mesh(linon_view* v, const float* verts, uint32_t n_verts, const uint32_t* faces, uint32_t n_faces);
/* MainView::update_camera (:490-494): uploads the 112-byte uniform. */
int linon_set_camera(linon_view* v, const linon_camera* camera);
/* MainView::update_settings (:547-550): uploads the 24-byte uniform. */
int linon_set_settings(linon_view* v, const linon_settings* settings);
/* MainView::resize_texture (:552-647): reallocates the per-pixel buffers (contents zeroed). */
int linon_resize(linon_view* v, uint32_t width, uint32_t height);
/* MainView::reload_shader(device, None, body) (:649-690): wgsl_body is the text spliced into
 * fn field_function(p_prev, p, v0, v: vec3<f32>, t: f32) -> vec3<f32> (:852-861).  On failure returns
 * LINON_ERR_COMPILE, writes the diagnostic (line numbers refer to the body) to err (if non-null) and keeps the
 * previous function live, like the reference (:677-679). */
int linon_set_field_function(linon_view* v, const char* wgsl_body, char* err, size_t err_len);
/* MainView::reload_shader(device, Some(new_src), field_function) (:649-690), reached from the file watcher on
 * src/main_view.wgsl (src/main.rs:206-222 -> Application::reload_compute_shader, src/application.rs:228-232) and the
 * README's extension point "Additional helper functions can be defined by adding them to the compute shader"
 * (README.md:46).  wgsl_module_source is the text of the edited main_view.wgsl; NULL or "" returns to the stock module.
 * What the CUDA path hot-reloads is the module's HELPER LAYER: new functions, edited versions of the library helpers
 * (rotateX/Y/Z, translate, refraction_index, refraction, point_plane_distance, sigmoid - and every unchanged helper
 * that calls an edited one), new module-scope constants and an edited PI; they are translated to CUDA and take part
 * in the NVRTC build of every variant.  The placeholder field_function of the module is ignored (the body comes from
 * linon_set_field_function).  An edit to the march itself (hit_triangle, ray_color, nonlinear_ray_color, sample_rays,
 * main_view, their constants, structs or bindings - recognised by token hashes of the reference's declarations)
 * returns LINON_ERR_UNSUPPORTED naming the declaration; a syntax or compile error returns LINON_ERR_COMPILE with
 * diagnostics that cite main_view.wgsl line numbers.  In both cases the previous module and kernel stay live (:677-679).
 * While a helper layer is installed the predefined mirage bodies run their generic translation.
 * Also hot-reloaded: the march-level CONSTANTS of the module - eps, backface_culling, use_lighting, linear_mode, light_color,
 * ambient_strength, shininess (src/main_view.wgsl:65-76), adaptive_sampling and sample_outline_rays (:227, :337) and the VALUES
 * of the step size, horizon and step-control literals of nonlinear_ray_color / sample_rays (:241-267, :287-288) and of the
 * focal length, viewport height (:355-356) and the reach of linear_mode's ray (:369) in main_view, and the colours of
 * `var<private> sample_colors` (:326-335, non-negative literals) - which become macros of the kernel template (a changed h_initial disables Enhance's text replacement, like in the reference).
 * Also hot-reloaded: the SHADING BLOCK of ray_color (`if (t > 0.0 && t < max_dist) { ... }`, :199-218) - the statements that
 * turn the nearest hit into a colour.  When ray_color differs from the reference's inside that block only (the face loop must
 * stay the reference's, token for token: it is the kernel's own intersection code), the block is translated from the user's
 * text and called at every geometric hit with ray_color's live values (origin, direction, max_dist, t, d1, d2) and the module's
 * constants; the march accepts the hit when the returned alpha is positive (`has_color = result.color.a > 0.0`, :262-263) and
 * keeps marching otherwise, like the reference.  Such a module runs the generic translation of the field function. */
int linon_reload_shader_source(linon_view* v, const char* wgsl_module_source, char* err, size_t err_len);
/* MainView::reload_shader(device, Some(new_src), field_function) as ONE step (:649-690 compiles new_src with the new field
 * function spliced in): the module and the body are validated together and committed together, or neither is - a helper
 * renamed in both at once is accepted although neither half would compile against the other's old state.
 * wgsl_module_source NULL or "" = the stock module. */
int linon_reload_shader(linon_view* v, const char* wgsl_module_source, const char* wgsl_body, char* err, size_t err_len);
/* Comma-separated names of the declarations the installed helper layer overrides or adds ("" = stock module). */
int linon_module_overrides(linon_view* v, char* names, size_t names_len);
/* MainView::render (:831-849), render_high_accuracy (:778-802), render_outline_rays (:804-829) by flags.
 * Enqueues: [outline search | click positions] -> march (+ sample rays on a second stream) -> overlay. */
int linon_render(linon_view* v, uint32_t flags);
/* Waits for everything enqueued by the handle. */
int linon_sync(linon_view* v);

/* Copies a whole buffer to host memory (synchronises the handle's stream first). */
int linon_read(linon_view* v, int buffer, void* dst, size_t dst_bytes);
/* Size in bytes of a buffer at the current resolution. */
size_t linon_buffer_bytes(const linon_view* v, int buffer);
/* Device pointer of a buffer (for CUDA interop, e.g. NCCL gather of row tiles or display hand-off). */
void* linon_device_ptr(linon_view* v, int buffer);
/* cudaStream_t the handle launches on (so callers can record their own events on it). */
void* linon_stream(linon_view* v);

/* Row-block sharding for multi-GPU frames: the rows are cut into blocks of rows_per_block; block b belongs to
 * rank b % world.  The march then only renders this rank's rows (buffers stay full-frame; other rows untouched). */
int linon_set_shard(linon_view* v, uint32_t rank, uint32_t world, uint32_t rows_per_block);
/* Peer stores - a sharded frame without a gather.  linon_redirect_output makes the march of this view store `buffer`
 * (LINON_BUF_RAYCAST_RGBA8, LINON_BUF_ENDPOINT_F32X4 or LINON_BUF_HIT_ID_I32) through device_ptr instead of into the
 * view's own buffer: the frame buffer of the ROOT shard's view, of the same frame size.  A shard writes only its own rows
 * (linon_set_shard), so the root's buffers hold the assembled frame as soon as every shard's march has completed - the
 * 24 bytes per ray travel over NVLink while the march runs, and the overlay then runs on the root as before.
 * NULL restores the view's own buffer; linon_resize drops every redirection.
 *   same process:     device_ptr = linon_device_ptr(root_view, buffer), with peer access enabled between the devices;
 *   one process/GPU:  the root exports a 64-byte CUDA IPC handle (linon_ipc_export), ships it to the other ranks (any
 *                     host channel, e.g. a torch.distributed broadcast), and they map it with linon_ipc_open (the
 *                     mapping lives until linon_destroy).
 * The caller orders the root's overlay after the other shards' marches (stream sync on each rank + a barrier). */
int linon_redirect_output(linon_view* v, int buffer, void* device_ptr);
/* ---- display hand-off without a copy ---------------------------------------------------------------------
 * The reference SHOWS two of the path's outputs: `texture` is registered with egui as a native texture
 * (src/main_view.rs:114-115, re-registered on resize :569-570) and `ray_samples_buffer` is the vertex buffer of the reference
 * view's second draw (src/reference_view.rs:381-388, src/application.rs:514-518).  linon_export_buffer_fd moves
 * LINON_BUF_FINAL_RGBA8 or LINON_BUF_RAY_SAMPLES into an exportable device allocation (CUDA virtual memory management, POSIX
 * file descriptor; the contents are kept) and returns a NEW file descriptor for it each call (the caller closes it) together
 * with the size of the allocation (>= linon_buffer_bytes, rounded up to the allocation granularity).  A Vulkan / wgpu-hal host
 * imports the descriptor as VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT memory and binds a linear buffer (ray samples: vertex
 * buffer, stride 32) or copies buffer-to-image on its own queue (final image: tightly packed rgba8, row 0 = top); the frames
 * then never cross PCIe.  The buffer keeps its place across renders; linon_resize reallocates the frame buffers, after which
 * the final image must be exported again (as the reference re-registers its texture).  Synchronisation: linon_sync (or a CUDA
 * external semaphore of the host's own) before the graphics queue reads. */
int linon_export_buffer_fd(linon_view* v, int buffer, int* fd, size_t* allocation_bytes);
/* The consumer side for a CUDA host (and the tests): maps an exported descriptor in the calling process / context and returns the
 * device pointer; the descriptor can be closed afterwards.  linon_release_import unmaps.  linon_copy_to_host reads through it. */
int linon_import_buffer_fd(int device, int fd, size_t allocation_bytes, void** device_ptr, void** import_handle);
int linon_release_import(void* import_handle);
int linon_copy_to_host(int device, const void* device_ptr, void* dst, size_t bytes);
int linon_ipc_export(linon_view* v, int buffer, void* handle64);
int linon_ipc_open(linon_view* v, const void* handle64, void** device_ptr);

/* ---- one frame on several B200s of a node, driven by ONE host thread (create_sharded + render_sharded) --------
 * The reference renders on one device (src/application.rs:67-89); this is the north star's "image shards naturally by row
 * tiles across the 8 B200s".  A linon_sharded owns one linon_view per device; rows are cut into blocks of rows_per_block and
 * dealt round-robin (block b -> shard b % n); shard 0 is the ROOT: its buffers hold the assembled frame, it runs the sample
 * rays and the overlay, and every reader (linon_read, linon_device_ptr, linon_get_counters ...) is used on
 * linon_sharded_view(s, 0).  The setters below are MainView's (update_model, update_camera, update_settings, reload_shader),
 * applied to every shard.  How the rows reach the root: */
enum {
    LINON_SHARD_PEER_STORES = 0, /* default: every shard's march stores its rows straight into the root's buffers over NVLink
                                    (peer access between the devices, linon_redirect_output); the root's overlay is ordered
                                    behind the other shards' marches by CUDA events.  No copy, no collective, no host sync */
    LINON_SHARD_NCCL_GATHER = 1  /* every shard marches into its own buffers, then ONE grouped ncclSend / ncclRecv exchange
                                    moves each non-root row block to the same rows of the root's buffers (libnccl.so.2,
                                    single-process communicators; needs distinct devices) */
};
typedef struct linon_sharded linon_sharded;
/* devices: CUDA ordinals, devices[0] is the root.  The same ordinal may appear more than once (several shards on one
 * device: peer-store mode only; how the multi-shard logic is tested on a single-GPU box). */
int linon_sharded_create(const int* devices, uint32_t n_devices, uint32_t width, uint32_t height, uint32_t rows_per_block,
                         int mode, linon_sharded** out);
void linon_sharded_destroy(linon_sharded* s);
uint32_t linon_sharded_world(const linon_sharded* s);
linon_view* linon_sharded_view(linon_sharded* s, uint32_t rank);
int linon_sharded_set_mesh(linon_sharded* s, const float* verts, uint32_t n_verts, const uint32_t* faces, uint32_t n_faces);
int linon_sharded_set_camera(linon_sharded* s, const linon_camera* camera);
int linon_sharded_set_settings(linon_sharded* s, const linon_settings* settings);
/* On a compile error no shard changes its function (the body is validated on the root first). */
int linon_sharded_set_field_function(linon_sharded* s, const char* wgsl_body, char* err, size_t err_len);
int linon_sharded_set_option(linon_sharded* s, int option, int value);
int linon_sharded_resize(linon_sharded* s, uint32_t width, uint32_t height);
/* MainView::render for the whole frame: every shard marches its rows, the rows are assembled on the root (see the modes),
 * the root runs the overlay.  Asynchronous like linon_render; flags as for linon_render. */
int linon_sharded_render(linon_sharded* s, uint32_t flags);
int linon_sharded_sync(linon_sharded* s);
/* Last error of the handle (or of the failed linon_sharded_create when s is NULL). */
const char* linon_sharded_last_error(const linon_sharded* s);

int linon_set_option(linon_view* v, int option, int value);
int linon_get_counters(linon_view* v, linon_counters* out);
int linon_get_timings(linon_view* v, linon_timings* out);
/* FNV-1a (64 bit) over the machine code of the march kernel that ran last (0 before the first march): the section
 * .text.linon_march_tiles of the cubin for the default tile kernel - other kernels of the module, e.g. the sample rays, may change
 * without changing it -, every .text section for the opt-in march kernels.  Line tables are left out, so an edit that only
 * moves lines keeps the identity.  A committed ncu capture names the hash of the kernel it profiled; bench.py refuses its counters
 * when the running kernel is a different binary. */
int linon_march_cubin_hash(linon_view* v, uint64_t* out);
/* The NVRTC that compiles the field functions (the reference's counterpart: naga + the driver's compiler behind
 * create_shader_module, src/main_view.rs:96-102, 662-665).  The library opens the toolkit's libnvrtc privately by the absolute
 * path recorded at build time, so a host process that already holds another libnvrtc.so.12 (torch ships one) gets the same
 * kernels as every other process; LINON_NVRTC_LIB overrides the path.  Returns the file in use and its version. */
int linon_nvrtc_info(char* path, size_t path_len, int* major, int* minor);
/* Last error message of the handle (or of the failed linon_create when v is NULL). */
const char* linon_last_error(const linon_view* v);

/* Translator alone (no device needed): WGSL body -> CUDA C++ body. Returns LINON_ERR_COMPILE with the reason in err. */
int linon_translate_wgsl(const char* wgsl_body, char* out, size_t out_len, char* err, size_t err_len);
/* Full JIT without a device: translate + NVRTC compile the march module for sm_100a; returns cubin size in
 * *cubin_bytes and the registers/thread of the persistent kernel are not known without loading (set to 0). */
int linon_compile_check(const char* wgsl_body, uint32_t flags, int math_mode, uint32_t n_faces, size_t* cubin_bytes,
                        char* err, size_t err_len);
/* The same with the helper layer of an edited main_view.wgsl (see linon_reload_shader_source). */
int linon_compile_check_module(const char* wgsl_module_source, const char* wgsl_body, uint32_t flags, int math_mode,
                               uint32_t n_faces, size_t* cubin_bytes, char* err, size_t err_len);
/* Module translator alone: the CUDA member functions generated for the helper layer (out) and the names emitted. */
int linon_translate_wgsl_module(const char* wgsl_module_source, char* out, size_t out_len, char* names, size_t names_len,
                                char* err, size_t err_len);
/* Top-level declarations of a WGSL module, one per line: "<kind> <name> <token hash> <line> <hash with the numeric literals
 * masked> <n> <the n numeric literals>" (used to generate the tables of the reference's own declarations,
 * tools/gen_module_hashes.py). */
int linon_scan_wgsl_module(const char* wgsl_module_source, char* out, size_t out_len, char* err, size_t err_len);

/* Parity/debug entry points: evaluate the CURRENT field function / helper library / hit_triangle on arrays
 * (host pointers; n elements; vec3 arrays are n x 3 f32). `which` numbering for helpers:
 * 0 rotateX 1 rotateY 2 rotateZ 3 translate 4 refraction 5 point_plane_distance 6 sigmoid 7 refraction_index
 * 8 reflect 9 normalize 10 cross. */
int linon_debug_eval_field(linon_view* v, uint32_t n, const float* p_prev, const float* p, const float* v0,
                           const float* vel, const float* t, float* out);
/* One Runge-Kutta velocity v = (k1 + 2 k2 + 2 k3 + k4) / 6 (src/main_view.wgsl:246-251) per element for the CURRENT field
 * function, through the device function the march kernels call (in fast math mode: the whole-step restatement of the
 * predefined mirage bodies).  p, dir, v0: n x 3 f32; t, h: n f32. */
int linon_debug_eval_rk4(linon_view* v, uint32_t n, const float* p, const float* dir, const float* v0, const float* t,
                         const float* h, float* out);
int linon_debug_eval_helper(linon_view* v, int which, uint32_t n, const float* vec, const float* nrm, const float* p0,
                            const float* a, const float* b, const float* c, float* out);
int linon_debug_hit_triangle(linon_view* v, uint32_t n, const float* tri9, const float* origin, const float* dir,
                             float* t_out);

/* Raw counter block of the last march (16 x u64): [0..4] accepted, pre-hit, rejected, face tests, culling checks; then section timers of the persistent
 * kernels when LINON_OPT_PROFILE = 1: [5..9] phase A, [10..14] phase B = refill cycles, step cycles, warp iterations,
 * live lanes summed over iterations, warp lifetime cycles. */
int linon_debug_profile(linon_view* v, uint64_t* out16);

/* The structural match that selects the hand-fused restatement of the mirage family (csrc/host/mirage_match.hpp): *kind = 1..4
 * (MirageSpherical, MiragePlane, MirageSphericalSigmoid, MiragePlaneSigmoid) when wgsl_body is one of the predefined mirage
 * bodies of src/functions.rs:36-94 up to comments, white space and the values of its literals, 0 otherwise; params8 (optional)
 * receives t_env, t_src, max_dist, centre / plane_p0 x y z, and the sigmoid's K and C as found in the text.  No device needed. */
int linon_debug_match_mirage(const char* wgsl_body, int* kind, double* params8);

/* FFMA-chain microbenchmark on `device`: the measured FP32 peak used as the roofline denominator. */
int linon_measure_fp32_peak(int device, double* tflops, int* sm_count);

#ifdef __cplusplus
}
#endif
#endif /* LINON_B200_H */
