// march_kernels.cuh — the JIT kernel template of the nonlinear ray march (sm_100a, NVRTC).
//
// The module NVRTC compiles is:   wgsl_prelude.cuh  +  field_function (translated WGSL body)  +  this file.
// It replaces the `main_view` dispatch of the reference (src/main_view.wgsl:339-375 with
// nonlinear_ray_color :229-279, ray_color :177-220, hit_triangle :150-175) and sample_rays (:281-314).
//
// Variant macros (-D on the NVRTC command line; the host keys its module cache on them):
//   LINON_H_INITIAL      initial RK4 step (0.1f, or 0.001f for "Enhance", src/main_view.rs:785-787)
//   LINON_ADAPTIVE       adaptive sampling on/off (src/main_view.wgsl:227,252-255,267-269).  The host passes 0
//                        when H_INITIAL <= 0.002: the rejection test needs h > 0.002 and the restore test
//                        needs h < h_initial, so neither can fire and h is a compile-time constant.
//   LINON_N_FACES        faces of the bound mesh (arrayLength(&faces.data), src/main_view.wgsl:182)
//   LINON_FACE_UNROLL    unroll factor of the face loop
//   LINON_BLOCK          threads per CTA of the march kernels
//   LINON_REFILL_MIN     idle lanes a warp tolerates before it refills (persistent kernel)
//   LINON_CULL / LINON_N_GROUPS  plane culling of the face loop (see intersect_mesh) and the number of plane groups
//   LINON_PROFILE        1 adds clock64 section timers to the persistent kernels (counters[5..14])
//   LINON_BVH            1 for meshes beyond the shared-memory stage (OBJ models, src/main.rs:139-157): the face loop
//                        becomes an exact traversal of a bounding-volume hierarchy (see intersect_mesh)
//   LINON_PAIRS          1 also builds linon_march_pairs: two rays per thread on the packed fp32 instructions of
//                        sm_100 (FFMA2 / FMUL2 / FADD2), LinonField<f32x2>.  The host retries with 0 when a user
//                        body does not instantiate for f32x2 (data-dependent control flow).
//   LINON_PAIR_MIN_BLOCKS  resident CTAs / SM requested for linon_march_pairs (register cap)
//   LINON_QUIET_ZONE     1: fixed-step loops take straight steps (and adaptive attempts identity steps) where the spliced fused
//                        mirage body declares itself the identity (mirage_fused.cuh: quiet_steps); LINON_QUIET_RADIUS_SIGMOID its
//                        radius for the sigmoid profiles
//   LINON_WINDOW / LINON_WINDOW_CHECKS / LINON_WINDOW_REFINE   quiet runs: a lane about to reach a wall resolves its hit at plan
//                        time over this many steps / with this many checks per plan / after asking the directional bound again
//   LINON_COLLAPSE_RUNS / LINON_COLLAPSE_MIN   quiet runs of at least MIN steps are taken binade by binade (advance_coord)
//   LINON_QUIET_AHEAD    regular quiet steps after a plan that found no run, before the next plan
//   LINON_EPS, LINON_BACKFACE_CULLING, LINON_USE_LIGHTING, LINON_LINEAR_MODE, LINON_LIGHT_*, LINON_AMBIENT, LINON_SHININESS,
//   LINON_FOCAL_LENGTH, LINON_VIEWPORT_HEIGHT, LINON_LINEAR_MAX_DIST (literals of main_view), LINON_SAMPLE_COLORS, LINON_USER_SHADE,
//   LINON_T_MAX, LINON_SAMPLE_H, LINON_SAMPLE_STEPS ...   the march-level constants of an edited main_view.wgsl (hot reload)
//
// The march kernels share every per-ray device function, so their per-pixel results are bit-identical in strict mode:
//   linon_march_tiles       one thread per pixel, 8x4 pixel tiles per warp (the straightforward mapping)
//   linon_march_pairs       one thread per PAIR of horizontally adjacent pixels, 16x4 pixel tiles per warp.  The RK4
//                           arithmetic of both rays is issued as packed f32x2 instructions (half the issue slots for
//                           the same FMA-pipe work: the tile kernel is issue-bound at 89 % of the slots); MUFU,
//                           min/max, compares and the rare mesh check stay per ray.
//   linon_march_phase_a/_b  persistent CTAs (grid = k * #SM).  Every warp keeps 32 rays in flight; a lane whose
//                           ray leaves its phase is refilled through a warp-aggregated atomicAdd.  Phase A = rays
//                           that have not hit yet (RK4 + all-face intersection every step); on a hit the lane
//                           writes colour + hit id, pushes {p, dir, t, h, last_v, last_diff} to a device queue and
//                           takes a fresh ray.  Phase B (second launch) = queued rays (RK4 only,
//                           src/main_view.wgsl:259 skips the intersection once has_color) until t > 5.  Rays never
//                           stop early in the reference, so the divergence to remove is pre-hit vs
//                           post-hit lanes, and this split makes every warp homogeneous.
//
// Mesh: staged once per CTA into shared memory as {v0, e1 = v1-v0, e2 = v2-v0} (3 x float4 per face).
// All lanes read the same face at the same time, so every LDS.128 is a broadcast (no bank conflicts).
// Precomputing e1/e2 performs the same two fp32 subtractions the reference does per test, once.

#ifndef LINON_H_INITIAL
#define LINON_H_INITIAL 0.1f
#endif
#ifndef LINON_ADAPTIVE
#define LINON_ADAPTIVE 1
#endif
#ifndef LINON_N_FACES
#define LINON_N_FACES 36
#endif
#ifndef LINON_FACE_UNROLL
#define LINON_FACE_UNROLL 4
#endif
#ifndef LINON_BLOCK
#define LINON_BLOCK 256
#endif
#ifndef LINON_REFILL_MIN
#define LINON_REFILL_MIN 1
#endif
#ifndef LINON_REFILL_CHUNK
#define LINON_REFILL_CHUNK 64u  // ray slots a warp of linon_march_refill takes from the global cursor at once (2 tiles)
#endif
#ifndef LINON_GROUP_UNROLL
#define LINON_GROUP_UNROLL 5
#endif
#ifndef LINON_MIN_BLOCKS
#define LINON_MIN_BLOCKS (1024 / LINON_BLOCK)  // ask for 32 resident warps / SM: caps the kernels at 64 registers
#endif
#ifndef LINON_CULL
#define LINON_CULL 1      // 1: skip faces whose plane the step segment cannot reach (exact: never changes a result)
#endif
#ifndef LINON_N_GROUPS
#define LINON_N_GROUPS LINON_N_FACES  // plane groups: runs of consecutive coplanar faces (built by the host)
#endif
#ifndef LINON_N_P1
#define LINON_N_P1 LINON_N_GROUPS  // first pass of a culling check: planes of the clusters that enclose the scene ...
#endif
#ifndef LINON_N_BOXES
#define LINON_N_BOXES 0            // ... and one bounding box for every other cluster (host: mesh_tables.hpp)
#endif
#ifndef LINON_N_SEGS
#define LINON_N_SEGS 1             // runs of consecutive plane groups that share a bound (one run of all groups = no boxes)
#endif
#ifndef LINON_PROFILE
#define LINON_PROFILE 0  // 1: the persistent kernel also accumulates per-section clock64 cycles (developer aid)
#endif
#ifndef LINON_PAIRS
#define LINON_PAIRS 0
#endif
#ifndef LINON_BVH
#define LINON_BVH 0  // 1: the mesh is too large for shared memory (> 1024 faces): faces + a bounding-volume hierarchy in global memory
#endif
#ifndef LINON_FW_ONE
#define LINON_FW_ONE 0  // 1: settings.field_weight == 1 (host-checked): the direction blend of src/main_view.wgsl:256 is cur_dir = v
#endif
#ifndef LINON_STAGE1_IDENTITY
#define LINON_STAGE1_IDENTITY 0  // 1: the field body guarantees f(p_prev = p, v) = v, so Runge-Kutta stage 1 is not evaluated
#endif
#ifndef LINON_FUSED_RK4
#define LINON_FUSED_RK4 0  // 1: the field body provides LinonField<S>::rk4, one whole Runge-Kutta step as a single expression
#endif
#ifndef LINON_QUIET_ZONE
#define LINON_QUIET_ZONE 0  // 1: fixed-step loops take straight steps where the spliced body declares itself the identity
#endif
#if LINON_QUIET_ZONE && !defined(LINON_HAS_QUIET_ZONE)
#undef LINON_QUIET_ZONE
#define LINON_QUIET_ZONE 0  // the body does not define LinonField<float>::quiet_steps (generic translations never do)
#endif
#ifndef LINON_COLLAPSE_RUNS
#define LINON_COLLAPSE_RUNS 0   // quiet runs: n straight steps per coordinate binade by binade instead of one by one (advance_coord)
#endif
#ifndef LINON_COLLAPSE_MIN
#define LINON_COLLAPSE_MIN 160   // ... from this run length on (sweep at C3: 32 .. 512 within 2 %, 96 - 256 best)
#endif
#ifndef LINON_WINDOW
#define LINON_WINDOW 64         // quiet-zone runs: steps ahead within which a lane resolves its wall hit at plan time (0: off)
#endif
#ifndef LINON_WINDOW_REFINE
#define LINON_WINDOW_REFINE 1   // ... after asking the directional bound once more from where the window starts
#endif
#ifndef LINON_WINDOW_CHECKS
#define LINON_WINDOW_CHECKS 2   // ... and the checks it may spend on it per plan (sweep at C3: 2 -> 8.39 ms, 4 -> 8.48, 8 -> 8.9)
#endif
#ifndef LINON_QUIET_AHEAD
#define LINON_QUIET_AHEAD 3  // small steps: regular quiet steps taken after a plan that found no run, before the next plan
#endif
#ifndef LINON_WARP_STATS
#define LINON_WARP_STATS 0  // 1 (developer aid; tests/emu sets it): warp-level event counts of the fixed-step tile loops in
#endif                      // counters[5..11]: plans, runs, full steps, quiet single steps, check events, directional bounds, straight warp-steps
#if LINON_WARP_STATS
#define LINON_WARP_EVENT(k, n) do { if ((threadIdx.x & 31) == 0) atomicAdd(&P.counters[k], (unsigned long long)(n)); } while (0)
#else
#define LINON_WARP_EVENT(k, n) do { } while (0)
#endif
#ifndef LINON_STRICT_MATH
#define LINON_STRICT_MATH 0  // 1: the bit-comparable build (--fmad=false ...): no explicit FMA either
#endif
#ifndef LINON_PAIR_MIN_BLOCKS
#define LINON_PAIR_MIN_BLOCKS (512 / LINON_BLOCK)  // 16 resident warps / SM = 32 rays per scheduler slot pair: <= 128 registers
#endif

// ---- the march-level constants of the reference's module (src/main_view.wgsl:65-76, :227, :241-267, :281-313) -------------
// A user who edits main_view.wgsl (the file watcher of src/main.rs:206-222 recompiles the whole shader) can change these;
// linon_reload_shader_source maps such edits to the macros below (csrc/host/wgsl_translate.cpp: march_edits) instead of
// refusing the module.  The defaults are the reference's literals.
#ifndef LINON_EPS
#define LINON_EPS 0.0000001f           // const eps, src/main_view.wgsl:74
#endif
#ifndef LINON_BACKFACE_CULLING
#define LINON_BACKFACE_CULLING 0       // const backface_culling, :65 (hit_triangle rejects a < eps instead of |a| < eps)
#endif
#ifndef LINON_USE_LIGHTING
#define LINON_USE_LIGHTING 1           // const use_lighting, :73 (0: colour = abs(normal), alpha 1)
#endif
#ifndef LINON_LINEAR_MODE
#define LINON_LINEAR_MODE 0            // const linear_mode, :72 (1: one straight ray_color(origin, dir, 100.0) per pixel, no march)
#endif
#ifndef LINON_FOCAL_LENGTH
#define LINON_FOCAL_LENGTH 0.035f      // `let focal_length` of main_view, :355
#endif
#ifndef LINON_VIEWPORT_HEIGHT
#define LINON_VIEWPORT_HEIGHT 0.025f   // `let viewport_height` of main_view, :356
#endif
#ifndef LINON_LINEAR_MAX_DIST
#define LINON_LINEAR_MAX_DIST 100.0f   // the reach of linear_mode's one ray_color call, :369
#endif
#ifndef LINON_AMBIENT
#define LINON_AMBIENT 0.01f            // const ambient_strength, :68
#endif
#ifndef LINON_SHININESS
#define LINON_SHININESS 64.0f          // const shininess, :69
#endif
#ifndef LINON_LIGHT_R
#define LINON_LIGHT_R 1.0f             // const light_color, :67
#define LINON_LIGHT_G 1.0f
#define LINON_LIGHT_B 1.0f
#endif
#ifndef LINON_T_MAX
#define LINON_T_MAX 5.0f               // the horizon of `for (; t <= 5.0;)`, :244
#endif
#ifndef LINON_REJECT_FACTOR
#define LINON_REJECT_FACTOR 10.0f      // diff > 10.0 * last_diff, :252
#endif
#ifndef LINON_REJECT_MIN_H
#define LINON_REJECT_MIN_H 0.002f      // ... && h > 0.002, :252
#endif
#ifndef LINON_SMALL_H
#define LINON_SMALL_H 0.001f           // h = 0.001 after a rejection, :253
#endif
#ifndef LINON_RESTORE_FACTOR
#define LINON_RESTORE_FACTOR 2.0f      // 2.0 * diff < last_diff, :267
#endif
#ifndef LINON_SAMPLE_H
#define LINON_SAMPLE_H 0.001f          // sample_rays: let h = 0.001, :287
#endif
#ifndef LINON_SAMPLE_STEPS
#define LINON_SAMPLE_STEPS 5000        // sample_rays: let steps: i32 = 5000, :288 (a multiple of the 100 samples per ray)
#endif
#define LINON_DIFF_FLOOR 1e-7f  // resolution of the reference's own velocity evaluation (its rounding noise)

struct MarchParams {
    float cam_origin[4];          // Camera.origin          (src/main_view.wgsl:7-12)
    float cam_view[4];            // Camera.view_direction
    float cam_up[4];              // Camera.up
    WgslSettings settings;        // 24 B, src/main_view.rs:43-51
    int width, height;            // full frame (texture) size
    int local_rows;               // rows this shard renders
    int shard_rank, shard_world;  // row blocks are dealt round-robin: block b -> rank b % world
    int rows_per_block;
    const float4* tris;           // LINON_N_FACES x {v0, e1, e2}, then LINON_N_GROUPS x {unit normal, plane offset}, then the
                                  // first-pass bounds: LINON_N_P1 x {unit normal, offset}, LINON_N_BOXES x {lo, hi}
    const int* group_range;       // LINON_N_GROUPS x (first face | face count << 16), then LINON_N_SEGS x
                                  // (first group | group count << 12 | (box + 1) << 24)
    uchar4* raycast;              // [height][width]  ray_casting_texture (rgba8unorm); a shard writes only its rows
    float4* endpoint;             // [height][width]  mapping_texture (rgba32float)
    int* hit_id;                  // [height][width]  face index of the first hit, -1 = none
    unsigned long long* counters; // [0] accepted steps  [1] steps that ran the face loop  [2] rejected attempts
                                  // [3] ray-triangle tests actually performed (after culling)
                                  // [4] pre-hit steps that ran the plane-group loop (all of them without culling)
                                  // LINON_PROFILE: phase A [5..9], phase B [10..14] = refill cycles, step cycles,
                                  // warp iterations, live lanes summed over iterations, warp lifetime cycles
    // persistent kernel only
    unsigned int* work;           // [0] next fresh ray slot  [1] queue tail (phase A pushes)  [2] queue head (phase B pops)
    int* q_pixel;                 // queue: pixel index (texture space) of the entry
    float4* q_state;              // queue: 3 x float4 per entry, SoA planes of q_capacity
    unsigned int q_capacity;
    unsigned int n_ray_slots;     // tiles * 32 (index space of fresh rays incl. out-of-frame padding lanes)
    unsigned int n_pair_slots;    // 16x4 tiles * 32 (linon_march_pairs: one slot = two horizontally adjacent pixels)
    float plane_w_max;            // max |plane offset| over the plane groups / max |coordinate| of a BVH mesh
                                  // (rounding bound of SafeDistance)
    const float4* bvh;            // LINON_BVH: 2 x float4 per node {lo.xyz, left child | first slot}, {hi.xyz, 0 | face count}
    int n_fixed_steps;            // !LINON_ADAPTIVE: iterations of the reference's `t <= 5` loop for h = LINON_H_INITIAL (the
                                  // host runs the same fp32 recurrence once: 5001 for 0.001, 51 for 0.1)
};

template <class S> struct RayT {
    vec3t<S> p, dir, sdir, last_v;
    S t, h, last_diff;
};
typedef RayT<float> Ray;

// the spliced body, instantiated per scalar type (LinonField is emitted by the host around the translated text)
template <class S>
__device__ __forceinline__ vec3t<S> field_function(const vec3t<S> p_prev, const vec3t<S> p, const vec3t<S> v0, const vec3t<S> v,
                                                   const S t, const WgslSettings& settings, const WgslCamera& camera) {
    return LinonField<S>::eval(p_prev, p, v0, v, t, settings, camera);
}

__device__ __forceinline__ vec3f ld3(const float* a) { return vec3f(a[0], a[1], a[2]); }

// local row -> texture row for the round-robin row-block sharding (identity when world == 1)
__device__ __forceinline__ int texture_row(const MarchParams& P, int local_row) {
    const int lb = local_row / P.rows_per_block;
    const int off = local_row - lb * P.rows_per_block;
    return (lb * P.shard_world + P.shard_rank) * P.rows_per_block + off;
}

// src/main_view.wgsl:347-366: pinhole ray through pixel (gid.x, gid.y), gid.y counted from the bottom row.
template <class S>
__device__ __forceinline__ vec3t<S> camera_ray_dir(const MarchParams& P, S px, S py_from_bottom, bool normalised_pos) {
    const float width = (float)P.width;
    const float height = (float)P.height;
    const float aspect_ratio = width / height;
    const vec3f view_direction = ld3(P.cam_view);
    const vec3f up = ld3(P.cam_up);
    const float focal_length = LINON_FOCAL_LENGTH;
    const float viewport_height = LINON_VIEWPORT_HEIGHT;
    const float viewport_width = aspect_ratio * viewport_height;
    const vec3f w = -view_direction;
    const vec3f horizontal = w_cross(up, w);
    const vec3f vertical = w_cross(w, horizontal);
    S u, v;
    if (normalised_pos) {  // sample rays, src/main_view.wgsl:475-476
        u = px * viewport_width - 0.5f * viewport_width;
        v = py_from_bottom * viewport_height - 0.5f * viewport_height;
    } else {               // pixels, src/main_view.wgsl:363-364
        u = px / (width - 1.0f) * viewport_width - 0.5f * viewport_width;
        v = py_from_bottom / (height - 1.0f) * viewport_height - 0.5f * viewport_height;
    }
    const vec3t<S> s = u * w_normalize(horizontal) + v * w_normalize(vertical) + focal_length * view_direction;
    return w_normalize(s);
}

__device__ __forceinline__ WgslCamera make_camera(const MarchParams& P) {
    WgslCamera c;
    c.origin = vec4f(P.cam_origin[0], P.cam_origin[1], P.cam_origin[2], P.cam_origin[3]);
    c.view_direction = vec4f(P.cam_view[0], P.cam_view[1], P.cam_view[2], P.cam_view[3]);
    c.up = vec4f(P.cam_up[0], P.cam_up[1], P.cam_up[2], P.cam_up[3]);
    return c;
}

// One Runge-Kutta evaluation, src/main_view.wgsl:246-250 (p_prev = cur_point in all four stages).
// h is a float (compile-time constant or uniform) or per-ray (adaptive sampling in the pair kernel).
template <class S, class H>
__device__ __forceinline__ vec3t<S> rk4_velocity(const RayT<S>& r, H h, const WgslSettings& st, const WgslCamera& cam) {
#if LINON_FUSED_RK4
    return LinonField<S>::rk4(r.p, r.dir, h);  // the spliced body restates the whole step (mirage_fused.cuh)
#endif
#if LINON_STAGE1_IDENTITY
    const vec3t<S> k1 = r.dir;  // the spliced body declares field(p, p, v0, v, t) == v (mirage_fused.cuh)
#else
    const vec3t<S> k1 = field_function<S>(r.p, r.p, r.sdir, r.dir, r.t, st, cam);
#endif
    const vec3t<S> k2 = field_function<S>(r.p, r.p + 0.5f * h * k1, r.sdir, k1, r.t + 0.5f * h, st, cam);
    const vec3t<S> k3 = field_function<S>(r.p, r.p + 0.5f * h * k2, r.sdir, k2, r.t + 0.5f * h, st, cam);
    const vec3t<S> k4 = field_function<S>(r.p, r.p + h * k3, r.sdir, k3, r.t + h, st, cam);
    return (k1 + 2.0f * k2 + 2.0f * k3 + k4) / 6.0f;
}

// hit_triangle for one face (src/main_view.wgsl:150-175) + the closest-hit update of ray_color (:193-197).
// Branch-free restatement of the reference's early returns (same comparisons, same NaN behaviour).
__device__ __forceinline__ void test_face(const float4* __restrict__ s_tri, int i, vec3f origin, vec3f direction,
                                          float max_dist, float& t_best, int& id_best) {
#if LINON_BVH
    const float4 a = __ldg(s_tri + 3 * i), b = __ldg(s_tri + 3 * i + 1), c = __ldg(s_tri + 3 * i + 2);
#else
    const float4 a = s_tri[3 * i], b = s_tri[3 * i + 1], c = s_tri[3 * i + 2];
#endif
    const vec3f v0(a.x, a.y, a.z), edge1(b.x, b.y, b.z), edge2(c.x, c.y, c.z);
    const vec3f h = w_cross(direction, edge2);
    const float det = w_dot(edge1, h);
    const float f = 1.0f / det;
    const vec3f s = origin - v0;
    const float u = f * w_dot(s, h);
    const vec3f q = w_cross(s, edge1);
    const float v = f * w_dot(direction, q);
    const float t = f * w_dot(edge2, q);
    // `if ((backface_culling || a > -eps) && a < eps) { return -1.0; }`, src/main_view.wgsl:156
    const bool ok = !((LINON_BACKFACE_CULLING || det > -LINON_EPS) && det < LINON_EPS) && !(u < 0.0f || u > 1.0f) &&
                    !(v < 0.0f || u + v > 1.0f) && (t > LINON_EPS);
    const float t_new = ok ? t : -1.0f;
#if LINON_BVH
    // the hierarchy visits faces out of index order: the reference's strict `<` over ascending indices (lowest index
    // wins an exact tie) becomes a lexicographic comparison on (t, original face index); a.w carries that index
    if (t_new > 0.0f && t_new < max_dist) {
        const bool better = t_best < 0.0f || t_new < t_best ||
                            (t_new == t_best && __float_as_int(a.w) < __float_as_int(__ldg(s_tri + 3 * id_best).w));
        if (better) {
            t_best = t_new;
            id_best = i;  // slot in the hierarchy's face order
        }
    }
#else
    if (t_new > 0.0f && t_new < max_dist && (t_best < 0.0f || t_new < t_best)) {
        t_best = t_new;
        id_best = i;
    }
#endif
}

// ray_color's loop over every face (src/main_view.wgsl:182-198), keeping the closest hit in (0, max_dist).
//
// LINON_CULL, level 1 (here): a face can only report t in (eps, max_dist) if the step segment reaches its plane, i.e.
// if the distance of the segment origin to the plane is at most max_dist (direction is a unit vector).  The host groups
// runs of consecutive coplanar faces; one signed distance (3 FFMA) + one compare per group decides whether the group's
// faces are tested at all.  The margin (2e-3 relative on max_dist plus 1e-5 (1 + |origin|_1) absolute) is orders of
// magnitude above the rounding error of either computation, so a skipped face is one the full test would have
// rejected: results are bit-identical with LINON_CULL = 0 (tests/test_gpu_parity.py checks it).  Faces are still
// visited in index order, so the strict `t_new < t` tie-break of the reference (lowest index wins) is preserved.
// The smallest plane distance is returned for level 2 (see SafeDistance).
__device__ __forceinline__ void intersect_mesh(const float4* __restrict__ s_tri, const int* __restrict__ s_range,
                                               const float4* __restrict__ bvh, vec3f origin, vec3f direction,
                                               float max_dist, float origin_l1, float& t_best, int& id_best,
                                               float& min_plane_dist, unsigned int& face_tests, bool boxes_clear = false,
                                               float box_floor = 0.0f) {
    // boxes_clear (two-pass checks of the shared-memory path): the caller's first pass has already found every cluster box
    // farther than an upper bound of this step's reach (box_floor = the smallest box distance), so the runs of groups those
    // boxes bound are skipped without looking at them - by the argument of the box test below, from the same distances (every
    // face of a cluster lies inside its inflated box; a box farther than 1.002 max_dist + 1e-5 (1 + |origin|_1) holds no face
    // that can report t < max_dist).  On the final approach to a wall of the room this is the common case: the check then
    // costs the room's planes and the faces of the one wall in reach instead of all plane groups and every box.
    t_best = -1.0f;
    id_best = -1;
#if LINON_BVH
    // Exact traversal.  A face can only report t in (eps, max_dist) if its (slightly inflated) bounding box is within
    // max_dist of the segment origin, so a subtree whose box is farther than `reach` (same margins as the plane test
    // below) cannot contain the closest hit: skipping it never changes a result (tests compare with the oracle's loop
    // over every face).  The smallest box distance seen - of skipped subtrees and of visited leaves - is a lower bound on
    // the distance to any face and feeds SafeDistance like the plane distances do.  s_tri = faces in hierarchy order
    // (global memory, L2 resident), s_range = unused, bvh = P.bvh.
    {
        const float reach = 1.002f * max_dist + 1e-5f * (1.0f + origin_l1);
        const float reach2 = reach * reach * 1.0001f;
        float nearest2 = 3.0e38f;
        int stack[40];
        int sp = 0, node = 0;
        for (;;) {
            const float4 lo = __ldg(bvh + 2 * node), hi = __ldg(bvh + 2 * node + 1);
            const float dx = fmaxf(fmaxf(lo.x - origin.x, origin.x - hi.x), 0.0f);
            const float dy = fmaxf(fmaxf(lo.y - origin.y, origin.y - hi.y), 0.0f);
            const float dz = fmaxf(fmaxf(lo.z - origin.z, origin.z - hi.z), 0.0f);
            const float d2 = dx * dx + dy * dy + dz * dz;
            bool descend = false;
            if (!(d2 <= reach2)) {
                nearest2 = fminf(nearest2, d2);  // skipped subtree: every face in it is at least this far
            } else {
                const int count = __float_as_int(hi.w), first = __float_as_int(lo.w);
                if (count > 0) {
                    nearest2 = fminf(nearest2, d2);  // visited leaf
                    for (int i = first; i < first + count; ++i) {
                        test_face(s_tri, i, origin, direction, max_dist, t_best, id_best);
                        ++face_tests;
                    }
                } else {
                    stack[sp++] = first + 1;  // children are adjacent: left = first, right = first + 1
                    node = first;
                    descend = true;
                }
            }
            if (!descend) {
                if (sp == 0) break;
                node = stack[--sp];
            }
        }
        min_plane_dist = sqrtf(nearest2);
    }
#elif LINON_CULL
    const float reach = 1.002f * max_dist + 1e-5f * (1.0f + origin_l1);
    const float4* __restrict__ s_plane = s_tri + 3 * LINON_N_FACES;
    float nearest = 3.0e38f;
    // The groups are walked in index order (the reference's tie-break), one SEGMENT at a time: a run of consecutive groups
    // whose cluster is bounded by a box (mesh_tables.hpp) is skipped as a whole when that box is farther than the reach -
    // every face of the run lies inside the (inflated) box, so none can report t in (eps, max_dist); same argument and
    // margins as the hierarchy traversal above.  A skipped run feeds its box distance into the lower bound; a visited
    // one the larger of its box distance and its smallest plane distance (both bound the distance to its faces).
#if LINON_N_BOXES > 0
    const float reach2 = reach * reach * 1.0001f;
    const float4* __restrict__ s_box = s_plane + LINON_N_GROUPS + LINON_N_P1;
#endif
#pragma unroll 1
    for (int sg = 0; sg < LINON_N_SEGS; ++sg) {
        const int seg = s_range[LINON_N_GROUPS + sg];
        const int g_first = seg & 0xfff, g_last = g_first + ((seg >> 12) & 0xfff);
        float box_dist = 0.0f;
#if LINON_N_BOXES > 0
        const int box = (seg >> 24) - 1;
        if (box >= 0 && boxes_clear) {
            nearest = fminf(nearest, box_floor);  // min over the skipped boxes of their distances: what the loop would accumulate
            continue;
        }
        if (box >= 0) {
            const float4 lo = s_box[2 * box], hi = s_box[2 * box + 1];
            const float dx = fmaxf(fmaxf(lo.x - origin.x, origin.x - hi.x), 0.0f);
            const float dy = fmaxf(fmaxf(lo.y - origin.y, origin.y - hi.y), 0.0f);
            const float dz = fmaxf(fmaxf(lo.z - origin.z, origin.z - hi.z), 0.0f);
            const float d2 = dx * dx + dy * dy + dz * dz;
            box_dist = sqrtf(d2);
            if (d2 > reach2) {  // NaN compares false -> the run is tested
                nearest = fminf(nearest, box_dist);
                continue;
            }
        }
#endif
        float seg_nearest = 3.0e38f;
        // partial unroll on purpose: a fully unrolled loop lets the compiler hoist all plane loads out of the time loop
        // (60 live registers, spills); with a loop-carried index the planes stay in shared memory (broadcast LDS.128)
#pragma unroll LINON_GROUP_UNROLL
        for (int g = g_first; g < g_last; ++g) {
            const float4 pl = s_plane[g];
            const float sigma = __fmaf_rn(origin.x, pl.x, __fmaf_rn(origin.y, pl.y, __fmaf_rn(origin.z, pl.z, -pl.w)));
            const float dist = fabsf(sigma);
            seg_nearest = fminf(seg_nearest, dist);
            if (dist <= reach) {
                // Second look, only for groups in reach: the segment must also CROSS the plane.  With long steps (h = 0.1)
                // most planes in reach are approached obliquely or left behind; their face tests ran with one or two
                // active lanes each and made default-step frames three times dearer than the arithmetic (ncu: 10.7 of 32
                // lanes active).  sigma is the signed distance at the origin, sigma_end at the end of the segment
                // (direction is a unit vector); same sign and both farther than the margin -> no face of the group can
                // report t in (eps, max_dist).  The margin is 20x the absolute slack of the reach test.
                const float sigma_end = __fmaf_rn(__fmaf_rn(direction.x, pl.x, __fmaf_rn(direction.y, pl.y, direction.z * pl.z)),
                                                  max_dist, sigma);
                const bool same_side = sigma * sigma_end > 0.0f && fminf(dist, fabsf(sigma_end)) > 2e-4f * (1.0f + origin_l1);
                if (!same_side) {
                    const int range = s_range[g];
                    const int first = range & 0xffff, last = first + (range >> 16);
#pragma unroll 1
                    for (int i = first; i < last; ++i) {
                        test_face(s_tri, i, origin, direction, max_dist, t_best, id_best);
                        ++face_tests;
                    }
                }
            }
        }
        nearest = fminf(nearest, fmaxf(seg_nearest, box_dist));
    }
    min_plane_dist = nearest;
#else
#pragma unroll LINON_FACE_UNROLL
    for (int i = 0; i < LINON_N_FACES; ++i) test_face(s_tri, i, origin, direction, max_dist, t_best, id_best);
    face_tests += LINON_N_FACES;
    min_plane_dist = 0.0f;
#endif
}

// First pass of a culling check: a lower bound of the distance from `origin` to any triangle, and nothing else.
// The mesh is clustered on the host (mesh_tables.hpp): the clusters that enclose the scene (the room of the Cornell box)
// contribute the planes of their groups - 5 instructions per plane, the same sums as intersect_mesh - and every other
// cluster (light, short box, tall box) ONE inflated bounding box (13 instructions), because the planes of the rotated
// inner boxes cut through the whole room and past the camera: with plane distances alone a ray paid ~20 checks per
// crossing of such an extension for a box it never came near.  Every triangle lies in its plane and inside its
// cluster's box, so the value bounds the distance to the geometry from below; the margins of SafeDistance and of the
// reach (orders of magnitude above the rounding of either evaluation, the same assumption level 1 makes) do the rest.
// When the bound exceeds the reach no face can be hit on this step and the caller skips intersect_mesh - including
// normalize(step_dir) and length(step_dir), which only feed it.
#if LINON_CULL && !LINON_BVH
__device__ __forceinline__ float nearest_geometry_distance(const float4* __restrict__ s_tri, vec3f origin, float& box_nearest) {
    const float4* __restrict__ s_plane = s_tri + 3 * LINON_N_FACES + LINON_N_GROUPS;
    float nearest = 3.0e38f;
#pragma unroll LINON_GROUP_UNROLL
    for (int g = 0; g < LINON_N_P1; ++g) {
        const float4 pl = s_plane[g];
        const float sigma = __fmaf_rn(origin.x, pl.x, __fmaf_rn(origin.y, pl.y, __fmaf_rn(origin.z, pl.z, -pl.w)));
        nearest = fminf(nearest, fabsf(sigma));
    }
#if LINON_N_BOXES > 0
    const float4* __restrict__ s_box = s_plane + LINON_N_P1;
    float nearest2 = 3.0e38f;
#pragma unroll 3
    for (int b = 0; b < LINON_N_BOXES; ++b) {
        const float4 lo = s_box[2 * b], hi = s_box[2 * b + 1];
        const float dx = fmaxf(fmaxf(lo.x - origin.x, origin.x - hi.x), 0.0f);
        const float dy = fmaxf(fmaxf(lo.y - origin.y, origin.y - hi.y), 0.0f);
        const float dz = fmaxf(fmaxf(lo.z - origin.z, origin.z - hi.z), 0.0f);
        nearest2 = fminf(nearest2, dx * dx + dy * dy + dz * dz);
    }
    box_nearest = sqrtf(nearest2);
    nearest = fminf(nearest, box_nearest);
#else
    box_nearest = 3.0e38f;
#endif
    return nearest;
}
#endif

// LINON_CULL, level 2: a per-ray budget that stays positive only while NO plane group can pass the level-1 test, so
// the whole face loop - including normalize(step_dir) and length(step_dir), which only feed it - is skipped and a
// pre-hit step costs 6 instructions instead of ~110.
//
// At a full check at p0 the loop returns nearest0 = min over groups of the COMPUTED plane distance.  For a later point
// p of the same ray, with d_i the displacement of step i and e(.) <= 3 2^-24 (|p|_1 + |w|) the rounding of the computed
// distance (three FMAs on a unit normal and the plane offset w):
//     computed_dist(p, g) >= nearest0 - e(p0) - e(p) - sum_i |d_i|_2 ,
//     |d_i|_2 <= |step_i|_1 + 1.5 2^-23 |p|_inf        (p + step is rounded once per component).
// Level 1 tests a group only if computed_dist <= 1.002 |step|_2 + 1e-5 (1 + |p|_1).  While the budget below is positive
// the travelled distance is below nearest0, hence |p|_1 <= M = |p0|_1 + nearest0, and with C = 1 + M + max_g |w_g|
//     budget = nearest0 - 1.1e-5 C - sum_{i <= now} (1.003 |step_i|_1 + 2e-7 C)  >  0
// implies computed_dist(p, g) > reach for every group (1e-5 C covers the absolute term of the level-1 test, 1e-6 C both
// e(.) terms, 2e-7 C the rounding of each position update, 1.003 the L1 >= L2 slack and the rounding of |step|_2).
// All sums are rounded against the skip (additions up, the budget down), NaNs compare false -> check.  A skipped loop
// is therefore one in which level 1 would have skipped every group: results are bit-identical with LINON_CULL = 0.
struct SafeDistance {
    float budget, drift;
    __device__ __forceinline__ void reset() { budget = 0.0f; drift = 0.0f; }
    // what this step takes from the budget
    __device__ __forceinline__ float cost(vec3f step_dir) const {
        const float travel = __fadd_ru(__fadd_ru(fabsf(step_dir.x), fabsf(step_dir.y)), fabsf(step_dir.z));
        return __fmaf_ru(1.003f, travel, drift);
    }
    // true: the face loop must run for the step p -> p + step_dir
    __device__ __forceinline__ bool must_check(float step_cost) const {
#if LINON_CULL
        return !(__fsub_rd(budget, step_cost) > 0.0f);
#else
        return true;
#endif
    }
    __device__ __forceinline__ void checked(float min_plane_dist, float p_l1, float plane_w_max) {
        const float c = __fadd_ru(__fadd_ru(__fadd_ru(1.0f, p_l1), min_plane_dist), plane_w_max);
        budget = __fmaf_rd(-1.1e-5f, c, min_plane_dist);
        drift = __fmul_ru(2e-7f, c);
    }
    __device__ __forceinline__ void moved(float step_cost) { budget = __fsub_rd(budget, step_cost); }
    // Fixed-step loops: the same cost from the direction's max-norm.  The displacement of a step is h v (one FMA per
    // component, or the rounded product in the strict build: <= h |v| (1 + 2^-24) either way), |h v|_2 <= sqrt(3) h |v|_inf,
    // and kh = 1.7321 * 1.003 * h exceeds 1.003 sqrt(3) h after its own two roundings (1.7321 / sqrt(3) = 1 + 2.8e-5):
    // kh |v|_inf + drift >= 1.003 |step|_2 + drift, which is all the proof above uses of |step|_1.  Three instructions
    // (FMNMX3 on |v|, FFMA.RP) instead of nine, and step_dir is not formed on steps that skip the face loop.
    __device__ __forceinline__ float cost_linf(float v_inf, float kh) const { return __fmaf_ru(kh, v_inf, drift); }
    // an upper bound of the level-1 reach 1.002 |step_dir|_2 + 1e-5 (1 + |p|_1) that needs no square root
    static __device__ __forceinline__ float reach_bound(float v_inf, float kh, float p_l1) {
        return __fmaf_ru(kh, v_inf, __fmul_ru(1e-5f, __fadd_ru(1.0f, p_l1)));
    }
    // what the budget will be after this step; the face loop must run when it is not positive (same test as must_check)
    __device__ __forceinline__ float left(float step_cost) const { return __fsub_rd(budget, step_cost); }
    __device__ __forceinline__ void retire() { budget = __int_as_float(0x7f800000); }  // the ray has hit: never check again
};
__device__ __forceinline__ float l1_norm(vec3f p) { return fabsf(p.x) + fabsf(p.y) + fabsf(p.z); }

#if LINON_CULL && !LINON_BVH && LINON_QUIET_ZONE
// LINON_QUIET_ZONE: the number of steps of a STRAIGHT ray - step j runs from p + j d to p + (j + 1) d, d = h dir unchanged -
// during which no face can be hit, from the same first-pass bounds as nearest_geometry_distance but along the direction of
// travel: the safe-distance budget shrinks with every step whatever the direction, so a ray running parallel to a wall
// pays a culling check every few steps although it never approaches the wall.
//   plane   a face of the group can report t in (eps, max_dist) on step j only if the segment reaches its plane, i.e. if
//           the distance D_j at the start of the step is at most the distance r the step covers towards it.  D_j = D - j r
//           while approaching, so steps 0 .. n - 1 are clear when D - n r > margin: n = floor((D - margin) / r).  A ray
//           moving away from a plane never reaches it.
//   box     every face of the cluster lies inside its box; the segment must be inside all three slabs of the box at once,
//           so no step before the LATEST slab entry can reach a face: n = floor(max_k gap_k / rate_k).
// Rounding: the k-th position differs from p + k d by at most k 2e-7 C per coordinate (C = 1 + |p|_1 + the longest possible
// travel + max |w|: the bound SafeDistance uses, with the travel added because a run can be long), so every approach rate is
// increased by delta = 2e-7 C, and margin = 2e-5 C covers the rounding of the distance evaluations and of hit_triangle's
// own t.  A step declared clear is one on which the face loop would find nothing: skipping its check never changes a result
// (tests/test_emulation_cpu.py compares with LINON_CULL = 0).
__device__ __forceinline__ int straight_clear_steps(const float4* __restrict__ s_tri, vec3f p, vec3f d, float travel_max,
                                                    float plane_w_max) {
    const float4* __restrict__ s_plane = s_tri + 3 * LINON_N_FACES + LINON_N_GROUPS;
    const float C = __fadd_ru(__fadd_ru(__fadd_ru(1.0f, l1_norm(p)), travel_max), plane_w_max);
    const float margin = 2e-5f * C, delta = 2e-7f * C;
    float clear = 3.0e8f;
#pragma unroll LINON_GROUP_UNROLL
    for (int g = 0; g < LINON_N_P1; ++g) {
        const float4 pl = s_plane[g];
        const float sigma = __fmaf_rn(p.x, pl.x, __fmaf_rn(p.y, pl.y, __fmaf_rn(p.z, pl.z, -pl.w)));
        const float rate = __fmaf_rn(d.x, pl.x, __fmaf_rn(d.y, pl.y, d.z * pl.z));  // change of sigma per step
        // approaching when sigma and rate have opposite signs: towards = -rate sign(sigma)
        const float towards = fmaxf(sigma > 0.0f ? -rate : rate, 0.0f);
        const float n = __fdividef(fabsf(sigma) - margin, towards + delta);  // <= 0 within the margin; NaN -> 0 below
        clear = fminf(clear, n);
    }
#if LINON_N_BOXES > 0
    const float4* __restrict__ s_box = s_plane + LINON_N_P1;
#pragma unroll 2
    for (int b = 0; b < LINON_N_BOXES; ++b) {
        const float4 lo = s_box[2 * b], hi = s_box[2 * b + 1];
        float latest = 0.0f;
        {
            const float below = lo.x - margin - p.x, above = p.x - hi.x - margin;
            const float gap = fmaxf(fmaxf(below, above), 0.0f);
            const float towards = fmaxf(below > 0.0f ? d.x : -d.x, 0.0f);
            latest = fmaxf(latest, __fdividef(gap, towards + delta));
        }
        {
            const float below = lo.y - margin - p.y, above = p.y - hi.y - margin;
            const float gap = fmaxf(fmaxf(below, above), 0.0f);
            const float towards = fmaxf(below > 0.0f ? d.y : -d.y, 0.0f);
            latest = fmaxf(latest, __fdividef(gap, towards + delta));
        }
        {
            const float below = lo.z - margin - p.z, above = p.z - hi.z - margin;
            const float gap = fmaxf(fmaxf(below, above), 0.0f);
            const float towards = fmaxf(below > 0.0f ? d.z : -d.z, 0.0f);
            latest = fmaxf(latest, __fdividef(gap, towards + delta));
        }
        clear = fminf(clear, latest);
    }
#endif
    // 0.999: the division; NaN (a NaN position) and negative values give 0
    return (int)fmaxf(clear * 0.999f, 0.0f);
}
#endif


// Blinn-Phong with the camera as light, src/main_view.wgsl:199-217.
__device__ __forceinline__ vec3f shade(const float4* __restrict__ s_tri, int id, vec3f direction) {
    const float4 b = s_tri[3 * id + 1], c = s_tri[3 * id + 2];
    const vec3f normal = w_normalize(w_cross(vec3f(b.x, b.y, b.z), vec3f(c.x, c.y, c.z)));
#if !LINON_USE_LIGHTING
    return w_abs(normal);  // `if (!use_lighting) { return vec4<f32>(abs(normal), 1.0); }`, src/main_view.wgsl:201-204
#else
    const vec3f light_color(LINON_LIGHT_R, LINON_LIGHT_G, LINON_LIGHT_B);
    const vec3f ambient = LINON_AMBIENT * light_color;
    float intensity = w_max(w_dot(normal, -direction), 0.0f);
    if (intensity == 0.0f) intensity = w_max(w_dot(-normal, -direction), 0.0f);
    const vec3f diffuse = intensity * light_color;
    const vec3f specular = w_pow(intensity, LINON_SHININESS) * light_color;
    return (ambient + diffuse + specular) * w_abs(normal);
#endif
}

// A hit the shader accepts: `result.color = ray_color(..); has_color = result.color.a > 0.0`, src/main_view.wgsl:262-263.  The
// stock shading block returns alpha = t > 0 (or 1 without lighting), so every geometric hit is accepted and this is shade().
// LINON_USER_SHADE (linon_reload_shader with an edited shading block of ray_color, src/main_view.wgsl:199-218): the block is
// the user's text, translated to LinonField<float>::linon_ray_color_shade (wgsl_translate.cpp), and decides by its own alpha -
// a hit it declines (alpha <= 0, or no return: ray_color's final `return vec4(0)`) leaves the ray marching, like the reference.
__device__ __forceinline__ bool shade_hit(const float4* __restrict__ s_tri, int id, vec3f origin, vec3f direction, float max_dist,
                                          float t_hit, vec3f& color, float& alpha) {
#ifdef LINON_USER_SHADE
    const float4 b = s_tri[3 * id + 1], c = s_tri[3 * id + 2];
    const vec4f r = LinonField<float>::linon_ray_color_shade(origin, direction, max_dist, t_hit, vec3f(b.x, b.y, b.z),
                                                             vec3f(c.x, c.y, c.z));
    alpha = r.w;
    color = vec3f(r.x, r.y, r.z);
    return r.w > 0.0f;
#else
    color = shade(s_tri, id, direction);
    alpha = LINON_USE_LIGHTING ? t_hit : 1.0f;
    return true;
#endif
}
__device__ __forceinline__ bool shade_hit(const float4* __restrict__ s_tri, int id, vec3f origin, vec3f direction, float max_dist,
                                          float t_hit, vec3f& color) {
    float alpha;
#ifdef LINON_USER_SHADE
    vec3f c;
    if (!shade_hit(s_tri, id, origin, direction, max_dist, t_hit, c, alpha)) return false;  // `color` keeps its value: only
    color = c;                                                                               // has_color makes it count
    return true;
#else
    return shade_hit(s_tri, id, origin, direction, max_dist, t_hit, color, alpha);
#endif
}

// rgba8unorm store conversion: clamp to [0,1], x255, round to nearest even; NaN -> 0.
__device__ __forceinline__ unsigned char unorm8(float x) {
    x = fminf(fmaxf(x, 0.0f), 1.0f);  // fmaxf(NaN, 0) = 0
    return (unsigned char)__float2int_rn(x * 255.0f);
}
__device__ __forceinline__ uchar4 pack_color(vec3f c) { return make_uchar4(unorm8(c.x), unorm8(c.y), unorm8(c.z), 255); }

#define LINON_MESH_F4 (3 * LINON_N_FACES + LINON_N_GROUPS + LINON_N_P1 + 2 * LINON_N_BOXES)
__device__ __forceinline__ void stage_mesh(float4* s_tri, int* s_range, const MarchParams& P) {
    for (int i = threadIdx.x; i < LINON_MESH_F4; i += blockDim.x) s_tri[i] = P.tris[i];
    for (int i = threadIdx.x; i < LINON_N_GROUPS + LINON_N_SEGS; i += blockDim.x) s_range[i] = P.group_range[i];
    __syncthreads();
}
#if LINON_BVH
// large mesh: faces (in hierarchy order, original index in .w of the first float4) stay in global memory
#define LINON_MESH_PROLOGUE(P) const float4* const s_tri = (P).tris; const int* const s_range = nullptr;
#define LINON_FACE_ID(tris, slot) __float_as_int(__ldg((tris) + 3 * (slot)).w)
#else
#define LINON_MESH_PROLOGUE(P)                   \
    __shared__ float4 s_tri[LINON_MESH_F4];      \
    __shared__ int s_range[LINON_N_GROUPS + LINON_N_SEGS]; \
    stage_mesh(s_tri, s_range, P);
#define LINON_FACE_ID(tris, slot) (slot)
#endif

// 32 consecutive ray slots form an 8x4 pixel tile; returns false for padding lanes outside the shard.
__device__ __forceinline__ bool slot_to_pixel(const MarchParams& P, unsigned int slot, int& x, int& local_row) {
    const unsigned int tiles_x = (unsigned int)(P.width + 7) >> 3;
    const unsigned int tile = slot >> 5, lane = slot & 31u;
    const unsigned int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    x = (int)(tx * 8u + (lane & 7u));
    local_row = (int)(ty * 4u + (lane >> 3));
    return x < P.width && local_row < P.local_rows;
}

__device__ __forceinline__ void init_ray(const MarchParams& P, int x, int local_row, Ray& r) {
    const int gid_y = P.height - 1 - texture_row(P, local_row);  // src/main_view.wgsl:342
    r.p = ld3(P.cam_origin);
    r.sdir = camera_ray_dir<float>(P, (float)x, (float)gid_y, false);
    r.dir = r.sdir;
    r.t = 0.0f;
    r.last_v = vec3f(0.0f, 0.0f, 0.0f);
    r.last_diff = -1.0f;
    r.h = LINON_H_INITIAL;
}

struct StepCounters {
    unsigned int accepted, prehit, rejected, face_tests, checks;
#if LINON_QUIET_ZONE
    unsigned int quiet;  // accepted steps taken as straight steps (velocity = direction, no field evaluation)
#endif
};

__device__ __forceinline__ void flush_counters(const MarchParams& P, StepCounters c) {
    // per-lane 32-bit counts are widened before the reduction (a persistent lane can exceed 2^32 face tests)
    unsigned long long a = c.accepted, b = c.prehit, r = c.rejected, f = c.face_tests, k = c.checks;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
        r += __shfl_xor_sync(0xffffffffu, r, o);
        f += __shfl_xor_sync(0xffffffffu, f, o);
        k += __shfl_xor_sync(0xffffffffu, k, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&P.counters[0], a);
        atomicAdd(&P.counters[1], b);
        atomicAdd(&P.counters[2], r);
        atomicAdd(&P.counters[3], f);
        atomicAdd(&P.counters[4], k);
    }
#if LINON_QUIET_ZONE
    unsigned long long q = c.quiet;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&P.counters[15], q);
#endif
}

// One loop iteration of nonlinear_ray_color up to (not including) the intersection, src/main_view.wgsl:246-258.
// Returns false when the adaptive test rejects the attempt (h <- 0.001, state unchanged).
// quiet (LINON_QUIET_ZONE): every stage point of this attempt lies outside the body's zone of influence, where the field is
// the identity: the velocity is the direction and the field is not evaluated (a real branch: a warp whose lanes are all
// quiet skips the evaluation, a mixed warp pays for it once either way).
__device__ __forceinline__ bool rk4_attempt(Ray& r, const WgslSettings& st, const WgslCamera& cam, vec3f& v, float& diff,
                                            vec3f& step_dir, bool quiet = false) {
#if LINON_ADAPTIVE
    const float h = r.h;
#else
    const float h = LINON_H_INITIAL;
#endif
#if LINON_QUIET_ZONE
    if (quiet) v = r.dir;
    else
#endif
    v = rk4_velocity(r, h, st, cam);
#if LINON_ADAPTIVE
    diff = w_length(v - r.last_v);
#if LINON_FUSED_RK4
    // The fused restatements are exact where the reference is noisy: outside a mirage's zone they return the direction
    // unchanged (diff == 0) and in the tail of a profile they resolve changes of 1e-10, where the reference's evaluation adds
    // ~1e-7 of rounding noise to every velocity.  The reference's step control is relative (diff > 10 last_diff
    // rejects, 2 diff < last_diff restores), so that noise is what it compares in calm regions - and what restores its step
    // there, on about every other step.  With exact values the restore test never holds again (0 < 0) and a ray leaving a
    // plane mirage crawled at h = 0.001 to the end of the frame: 795 instead of 128 accepted steps per ray for
    // MiragePlaneSigmoid at 1080p, measured.  The same resolution is therefore applied here: a difference below the noise
    // floor counts as the floor (never rejects on its own) and as calm (restores the step, rk4_advance).
    diff = fmaxf(diff, LINON_DIFF_FLOOR);
#endif
    if (r.last_diff >= 0.0f && diff > LINON_REJECT_FACTOR * r.last_diff && h > LINON_REJECT_MIN_H) {
        r.h = LINON_SMALL_H;
        return false;
    }
#else
    diff = 0.0f;
#endif
#if LINON_FW_ONE
    r.dir = v;  // field_weight == 1: (1 - 1) * cur_dir + 1 * v
#else
    r.dir = (1.0f - st.field_weight) * r.dir + st.field_weight * v;
#endif
    step_dir = r.dir * h;
    return true;
}

// The tail of an accepted iteration, src/main_view.wgsl:264-270.
__device__ __forceinline__ void rk4_advance(Ray& r, vec3f v, float diff, vec3f step_dir) {
    r.p = r.p + step_dir;
#if LINON_ADAPTIVE
    r.t = r.t + r.h;
    r.last_v = v;
#if LINON_FUSED_RK4
    if ((LINON_RESTORE_FACTOR * diff < r.last_diff || diff <= LINON_DIFF_FLOOR) && r.h < LINON_H_INITIAL) r.h = LINON_H_INITIAL;
#else
    if (LINON_RESTORE_FACTOR * diff < r.last_diff && r.h < LINON_H_INITIAL) r.h = LINON_H_INITIAL;
#endif
    r.last_diff = diff;
#else
    r.t = r.t + LINON_H_INITIAL;
#endif
}

// ------------------------------------------------------------------------------------------------------------
// Kernel 1: one thread per pixel.
// ------------------------------------------------------------------------------------------------------------
// src/main_view.wgsl:264 in the production builds is one FFMA per component; the strict build rounds step_dir first,
// like the reference's two statements (:257, :264).
__device__ __forceinline__ vec3f advance_point(vec3f p, vec3f dir, float h) {
#if LINON_STRICT_MATH
    return p + dir * h;
#else
    return vec3f(__fmaf_rn(dir.x, h, p.x), __fmaf_rn(dir.y, h, p.y), __fmaf_rn(dir.z, h, p.z));
#endif
}

#if !LINON_STRICT_MATH
// n straight steps of ONE coordinate, p <- fma(d, h, p) n times, bit for bit - without executing them one by one.
// While p stays inside one binade every p is a multiple of the binade's ulp u, so fl(p + a), a = d h exactly, is
// p + rn_u(a): the rounded increment delta is the SAME for every step that starts and ends in the binade (a tie, a exactly
// half-way between two multiples of u, is the exception: its rounding depends on the parity of p).  One real step gives
// q = fl(p + a) and delta = q - p (exact: same binade); then k further steps are q + k delta, which is representable (a
// multiple of u inside the binade) and hence exactly what one FMA returns.  k is limited by the binade: a step that starts at
// magnitude m must end at m + |delta| <= 2^(E+1) when the magnitude grows (above 2^(E+1) the grid is coarser, and the rounded
// sum lands on 2^(E+1) or below either way) and at m - |delta| >= 2^E + 2u when it shrinks (below 2^E the grid is finer).
// Near zero, across a binade boundary, on a (near-)tie, for NaN / inf / tiny p: single steps, the loop always advances.
// The rays of a quiet run do exactly this n times per coordinate (advance_point); a run of 700 steps that crosses two binade
// boundaries costs three real steps and three jumps per coordinate.  Pinned against the step-by-step loop on random and
// adversarial inputs (tests/test_emulation_cpu.py: test_collapsed_runs_equal_the_step_by_step_chain) and by the frames that
// must equal the kernel without runs (culling off) bit for bit.
__device__ __forceinline__ float advance_coord(float p, const float d, const float h, int n) {
    while (n > 0) {
        const float q = __fmaf_rn(d, h, p);  // one real step
        --n;
        if (q == p) return p;                // a rounds to nothing here: every further step returns the same value
        const int bp = __float_as_int(p), bq = __float_as_int(q);
        const int e = bp & 0x7f800000;
        if (n > 0 && ((bp ^ bq) & 0xff800000) == 0 && e >= (64 << 23) && e < (254 << 23)) {  // same sign and binade, plainly normal
            const float delta = q - p;                                // exact
            const float u = __int_as_float(e - (23 << 23));          // the binade's ulp
            const float res = __fmaf_rn(d, h, -delta);               // a - delta, |.| <= u / 2 (rounded: only its size matters)
            if (fabsf(res) < 0.499999f * u) {                        // not a tie, not even nearly
                const float lo = __int_as_float(e);                  // 2^E
                const float mq = fabsf(q), ad = fabsf(delta);
                const bool growing = ((__float_as_int(delta) ^ bp) & 0x80000000) == 0;
                const float room = growing ? (2.0f * lo - mq) : (mq - lo - 2.0f * u);  // exact differences
                // k further steps need k |delta| <= room.  The approximate quotient is corrected by the exact test: k |delta| and
                // room are multiples of u below 2^(E+1), so fma(k, |delta|, -room) is exact; a quotient beyond the remaining
                // steps (up to 2^24 |delta| / |delta|) is clamped first, where the estimate's error no longer matters
                float kf = fminf(floorf(__fdividef(room, ad)), (float)n);
                kf = fmaxf(kf, 0.0f);
                if (__fmaf_rn(kf, ad, -room) > 0.0f) kf -= 1.0f;   // estimate one too high
                if (__fmaf_rn(kf, ad, -room) > 0.0f) kf -= 1.0f;   // (two ulps of the division at most)
                const int k = (int)fmaxf(kf, 0.0f);
                p = __fmaf_rn((float)k, delta, q);                   // exact: a multiple of u inside the binade
                n -= k;
                continue;
            }
        }
        p = q;
    }
    return p;
}
#endif

// n straight steps of a quiet run: advance_point n times.  LINON_COLLAPSE_RUNS: runs of at least LINON_COLLAPSE_MIN steps are
// taken binade by binade (advance_coord: ~40 instructions per real step + jump and coordinate, a few of them per run, against
// 3.4 instructions per step of the loop - measured at C3: 8.43 -> 7.12 ms; the 51-step frames and the 50-step stretches between
// two samples of a sample ray are faster one by one: 0.27 vs 0.35 ms at 1080p).  n must be warp-uniform here.
__device__ __forceinline__ vec3f advance_point_n(vec3f p, vec3f dir, float h, int n) {
#if LINON_COLLAPSE_RUNS && !LINON_STRICT_MATH
    if (n >= LINON_COLLAPSE_MIN)
        return vec3f(advance_coord(p.x, dir.x, h, n), advance_coord(p.y, dir.y, h, n), advance_coord(p.z, dir.z, h, n));
#endif
#pragma unroll 8
    for (int k = 0; k < n; ++k) p = advance_point(p, dir, h);
    return p;
}

extern "C" __global__ void __launch_bounds__(LINON_BLOCK, LINON_MIN_BLOCKS) linon_march_tiles(const MarchParams P) {
    LINON_MESH_PROLOGUE(P)

    const unsigned int slot = blockIdx.x * LINON_BLOCK + threadIdx.x;
    int x, local_row;
    StepCounters cnt = {0u, 0u, 0u, 0u, 0u};
#if LINON_LINEAR_MODE
    // `const linear_mode: bool = true` (src/main_view.wgsl:72, :368-370): one straight ray per pixel,
    // textureStore(ttarget, coords, ray_color(origin, dir, 100.0)) - the colour's alpha is the hit distance t (0 and a black
    // pixel on a miss), and `mapping` is NOT written: endpoints keep whatever the previous frame left there.
    if (slot < P.n_ray_slots && slot_to_pixel(P, slot, x, local_row)) {
        Ray r;
        init_ray(P, x, local_row, r);
        float t_hit, nearest;
        int id;
        intersect_mesh(s_tri, s_range, P.bvh, r.p, r.dir, LINON_LINEAR_MAX_DIST, l1_norm(r.p), t_hit, id, nearest, cnt.face_tests);
        const size_t pix = (size_t)texture_row(P, local_row) * P.width + x;
        uchar4 c = make_uchar4(0, 0, 0, 0);
        int hit = -1;
        if (t_hit > 0.0f && t_hit < LINON_LINEAR_MAX_DIST) {
#ifdef LINON_USER_SHADE
            vec3f rgb;
            float alpha;
            shade_hit(s_tri, id, r.p, r.dir, LINON_LINEAR_MAX_DIST, t_hit, rgb, alpha);  // stored whatever its alpha says, :369-370
            c = pack_color(rgb);
            c.w = unorm8(alpha);
#else
            const vec3f rgb = shade(s_tri, id, r.dir);
            c = pack_color(rgb);
#if LINON_USE_LIGHTING
            c.w = unorm8(t_hit);  // vec4<f32>(result, t), src/main_view.wgsl:217
#endif
#endif
            hit = LINON_FACE_ID(s_tri, id);
        }
        P.raycast[pix] = c;
        P.hit_id[pix] = hit;
    }
    flush_counters(P, cnt);
    return;
#endif
#if !LINON_ADAPTIVE
    // Fixed step (Enhance, or adaptive sampling switched off): every ray runs the same P.n_fixed_steps iterations, so
    // the whole warp stays converged from the first step to the last.  Padding lanes outside the frame repeat the nearest
    // pixel and store nothing.  That makes every vote a full-mask vote and splits the loop by WARP phase:
    //   searching   some lane has not hit yet.  All lanes keep the safe-distance budget (a lane that has hit carries
    //               +inf), ONE vote per step decides whether the face loop runs, and the common path has no divergent
    //               branch at all (no BSSY/BSYNC pair, no per-lane has_color test);
    //   coasting    every lane has hit: the loop is the Runge-Kutta step and nothing else.
    // A ray's arithmetic is the same in both phases (advance_point, rk4_velocity), so its result does not depend on its
    // neighbours; only the counters of executed tests do, as before.
    if (P.local_rows <= 0) return;
    const bool valid = slot_to_pixel(P, slot, x, local_row) && slot < P.n_ray_slots;
    x = min(x, P.width - 1);
    local_row = min(local_row, P.local_rows - 1);
    {
        const WgslCamera cam = make_camera(P);
        const WgslSettings st = P.settings;
        const float h = LINON_H_INITIAL;
        Ray r;
        init_ray(P, x, local_row, r);
        bool has_color = false;
        SafeDistance sd;
        sd.reset();
        vec3f color(0.0f, 0.0f, 0.0f);
        int hit = -1;
        const int n_steps = P.n_fixed_steps;
        const float kh = 1.7321f * 1.003f * h;  // SafeDistance::cost_linf
        int i = 0;
        bool searching = true;  // warp-uniform
#if LINON_QUIET_ZONE
        int quiet_ahead = 0;    // warp-uniform: steps from here known to be quiet for every lane (see below)
        int clear_until = 0;    // window (below): steps before this index are verified clear of the geometry for this lane
#endif
        for (; i < n_steps && searching; ++i) {
#if LINON_QUIET_ZONE
            // Quiet zone (mirage_fused.cuh): while every lane of the warp is outside the body's zone of influence the field
            // is the identity, a Runge-Kutta step is p += h dir with dir unchanged, and the cost a step takes from the
            // safe-distance budget is a constant.  Plan a run of such steps - as many as the lane with the nearest zone
            // entry and the lane nearest to the geometry allow (its budget or, when that is what binds, the directional bound
            // of straight_clear_steps) - and take them back to back: every step is still executed (one FFMA per component),
            // none needs a field evaluation, a vote or a culling check.  The plan is warp-uniform (REDUX.MIN); lanes that
            // have hit carry an infinite budget.  A run that the geometry cuts short is followed by ONE quiet step through
            // the regular check logic below (velocity = direction), which refreshes the budget.
            bool quiet = false;  // warp-uniform: this step's velocity is the direction
            if (quiet_ahead > 0) {
                // the last plan found no run (some lane is about to reach a face) but at least this many quiet steps for
                // every lane: take them one by one through the check logic without planning again - on the final approach
                // to a wall the answer would be "no run" again, and a plan costs as much as a full Runge-Kutta step
                --quiet_ahead;
                quiet = true;
            } else if (!__any_sync(0xffffffffu, !(LinonField<float>::quiet_clearance(r.p) > 0.0f))) {
                LINON_WARP_EVENT(5, 1);
                const int nq = LinonField<float>::quiet_steps(r.p, r.dir, h);
                const float q_inf = fmaxf(fmaxf(fabsf(r.dir.x), fabsf(r.dir.y)), fabsf(r.dir.z));
                // 1.000001: the run takes n * cost_up from the budget in one rounding instead of n roundings of cost
                const float cost_up = __fmul_ru(sd.cost_linf(q_inf, kh), 1.000001f);
                // steps whose budget stays positive (0.9999 covers the division and keeps the budget after the run > 0)
                // (fmaxf first: a NaN - 0 / 0 for a ray that does not move - and a negative budget give 0)
                int nb = (int)fminf(fmaxf(__fdividef(sd.budget * 0.9999f, cost_up), 0.0f), 268435456.0f);
                int n_clear = nb;  // steps this lane can take without a culling check
#if LINON_CULL && !LINON_BVH
                // the budget is what limits the run (a wall nearby, or a run that just ended): ask how far the straight ray
                // really is from the geometry along its direction
                if (__any_sync(0xffffffffu, n_clear < nq && !has_color)) {
                    LINON_WARP_EVENT(10, 1);
                    const vec3f d = r.dir * h;
                    const float travel_max = (float)(n_steps - i) * l1_norm(d);
                    if (!has_color) n_clear = max(n_clear, straight_clear_steps(s_tri, r.p, d, travel_max, P.plane_w_max));
                }
#endif
#if LINON_CULL && !LINON_BVH && LINON_WINDOW > 0
                // Window (small steps only): a lane whose run the geometry cuts short within the next LINON_WINDOW steps
                // resolves what happens at its wall NOW instead of stopping the whole warp there.  While a ray is quiet its
                // positions are a known chain - p_(j+1) = fma(h, dir, p_j), the very FFMAs the run executes - so the lane walks
                // a private copy to its first step that is not known to be clear and performs the regular culling check of
                // that step, and of the next ones, until a face is hit or LINON_WINDOW_CHECKS steps are verified clear one by
                // one.  Every check is the check the step-by-step path would perform at that step, at the same position, with
                // the same direction: results are bit-identical, only their order in time changes (a hit is known a few steps
                // early, which nothing else depends on - the ray marches on either way).  All lanes of the warp that are about
                // to reach a wall do this together, each at its own step, so the final approach to a wall costs a few check
                // rounds with most lanes active instead of one round per step with one to three (round 2 ncu: the checks
                // were 19 % of the instructions and 30 % of the stall samples of the C3 frame).
                if (LINON_H_INITIAL < 0.01f) {
                    n_clear = max(n_clear, clear_until - i);  // what an earlier window has already verified step by step
                    const int horizon = min(min(__reduce_min_sync(0xffffffffu, nq), n_steps - i), LINON_WINDOW);  // straight for all
                    const bool limited = !has_color && n_clear < horizon;
                    if (__any_sync(0xffffffffu, limited)) {
                        LINON_WARP_EVENT(12, 1);
                        vec3f pj = r.p;
                        int jc = limited ? n_clear : 0;  // steps [0, jc) of this lane are known to be clear
                        const int jmax = __reduce_max_sync(0xffffffffu, jc);
                        for (int k = 0; k < jmax; ++k)
                            if (k < jc) pj = advance_point(pj, r.dir, h);
                        const vec3f w_step = r.dir * h;
                        const float w_max_dist = w_length(w_step);
                        const vec3f w_dirn = w_normalize(w_step);
#if LINON_WINDOW_REFINE
                        // n_clear was asked from far away (0.1 % of thousands of steps is several steps): ask the directional
                        // bound again from the walked position, where the same 0.1 % is nothing, and walk on to the step that
                        // reaches the wall - the first check then is the one that finds the face
                        {
                            const int more = limited ? min(straight_clear_steps(s_tri, pj, w_step, (float)(n_steps - i) * l1_norm(w_step),
                                                                               P.plane_w_max), max(horizon - jc, 0)) : 0;
                            const int more_max = __reduce_max_sync(0xffffffffu, more);
                            for (int k = 0; k < more_max; ++k)
                                if (k < more) pj = advance_point(pj, r.dir, h);
                            jc += more;
                        }
#endif
#pragma unroll 1
                        for (int k = 0; k < LINON_WINDOW_CHECKS; ++k) {
                            const bool act = limited && !has_color && jc < horizon;
                            if (!__any_sync(0xffffffffu, act)) break;
                            LINON_WARP_EVENT(13, 1);
                            if (act) {
                                const float p_l1 = l1_norm(pj);
                                ++cnt.checks;
                                float box_nearest = 0.0f;
                                float nearest = nearest_geometry_distance(s_tri, pj, box_nearest);
                                const float reach_up = SafeDistance::reach_bound(q_inf, kh, p_l1);
                                if (!(nearest > reach_up)) {
                                    float t_hit, plane_nearest;
                                    int id;
                                    intersect_mesh(s_tri, s_range, P.bvh, pj, w_dirn, w_max_dist, p_l1, t_hit, id, plane_nearest,
                                                   cnt.face_tests, box_nearest > reach_up, box_nearest);
                                    nearest = fmaxf(nearest, plane_nearest);
                                    if (t_hit > 0.0f && t_hit < w_max_dist && shade_hit(s_tri, id, pj, w_dirn, w_max_dist, t_hit, color)) {
                                        hit = LINON_FACE_ID(s_tri, id);
                                        has_color = true;
                                        cnt.prehit = i + jc + 1;  // the hit belongs to step i + jc
                                        sd.budget = __int_as_float(0x7f800000);  // a lane that has hit never asks again
                                    }
                                }
                                if (!has_color) {
                                    // the checked step is clear, and so are the steps the safe-distance budget of THIS check
                                    // covers (SafeDistance: the same refresh and the same cost the step-by-step path applies
                                    // after a check at this position) - inside a cluster's box, where the directional bound
                                    // says "now", that is what carries the lane to the face
                                    SafeDistance wsd;
                                    wsd.checked(nearest, p_l1, P.plane_w_max);
                                    const float wcost = __fmul_ru(wsd.cost_linf(q_inf, kh), 1.000001f);
                                    const int covered = (int)fminf(fmaxf(__fdividef(wsd.budget * 0.9999f, wcost), 1.0f), 1048576.0f);
                                    const int skip = min(covered, horizon - jc);  // >= 1: the step just checked
                                    for (int q = 0; q < skip; ++q) pj = advance_point(pj, r.dir, h);
                                    jc += covered;  // may pass the horizon: the lane is then done with this window
                                }
                            }
                        }
                        if (limited) {
                            n_clear = has_color ? 0x7fffffff : jc;
                            clear_until = i + jc;
                            if (has_color) nb = 268435456;  // its budget is infinite now: the run below must not spend it
                        }
                        searching = __any_sync(0xffffffffu, !has_color);  // the last lane may have hit in here: the run below still
                                                                          // executes, then the loop hands over to the coasting loop
                    }
                }
#endif
                const int n_lane = LINON_CULL ? min(nq, n_clear) : 0;  // without culling every pre-hit step runs the face loop
                const int n = min(__reduce_min_sync(0xffffffffu, n_lane), n_steps - i);
                if (n > 0) {
                    r.p = advance_point_n(r.p, r.dir, h, n);
                    r.t = __fmaf_rn((float)n, h, r.t);
                    // within the budget: take the run from it (+inf stays +inf); beyond it (the directional bound allowed
                    // more): the budget is spent, the next regular step runs a check and refreshes it
                    sd.budget = (n <= nb) ? __fmaf_rd(-(float)n, cost_up, sd.budget) : 0.0f;
                    LINON_WARP_EVENT(6, 1);
                    LINON_WARP_EVENT(11, n);
                    cnt.quiet += (unsigned int)n;
                    i += n - 1;  // the loop header adds the last one
                    continue;
                }
                const int nq_warp = __reduce_min_sync(0xffffffffu, nq);
                quiet = nq_warp > 0;
                // (small steps only: with the default step a ray crosses a tenth of the room per step and the situation
                // has changed by the next one)
                constexpr int ahead = LINON_H_INITIAL < 0.01f ? LINON_QUIET_AHEAD : 1;
                quiet_ahead = min(nq_warp, ahead) - 1;  // <= 0: the next step plans again
            }
            if (quiet) ++cnt.quiet;
            LINON_WARP_EVENT(quiet ? 8 : 7, 1);
            const vec3f v = quiet ? r.dir : rk4_velocity(r, h, st, cam);
#else
            LINON_WARP_EVENT(7, 1);
            const vec3f v = rk4_velocity(r, h, st, cam);
#endif
#if LINON_FW_ONE
            r.dir = v;  // field_weight == 1: (1 - 1) * cur_dir + 1 * v
#else
            r.dir = (1.0f - st.field_weight) * r.dir + st.field_weight * v;
#endif
            const float v_inf = fmaxf(fmaxf(fabsf(r.dir.x), fabsf(r.dir.y)), fabsf(r.dir.z));
            const float step_cost = sd.cost_linf(v_inf, kh);
            float left = sd.left(step_cost);
            // Warp-coherent checks (see the adaptive loop below): when one lane's budget expires every searching lane of
            // the warp refreshes its budget with it.
#if LINON_CULL
            if (__any_sync(0xffffffffu, !(left > 0.0f)))
#endif
            {
                LINON_WARP_EVENT(9, 1);
                if (!has_color) {
                    const float p_l1 = l1_norm(r.p);
                    ++cnt.checks;
#if LINON_CULL && !LINON_BVH
                    // Small steps (Enhance): two passes - plane distances first; the face loop (and the normalisation of
                    // step_dir that only it needs) runs only when some plane is within the reach of this step, which is
                    // rare.  With the default step (0.1, a tenth of the room) most checks do find a plane in reach and the
                    // first pass would be paid on top (measured: +30 % on the 51-step frames), so it is compiled out.
                    constexpr bool two_pass = LINON_H_INITIAL < 0.01f;
                    float box_nearest = 0.0f;
                    float nearest = two_pass ? nearest_geometry_distance(s_tri, r.p, box_nearest) : 0.0f;
                    const float reach_up = SafeDistance::reach_bound(v_inf, kh, p_l1);
                    const bool boxes_clear = two_pass && box_nearest > reach_up;  // NaN compares false: every run is looked at
                    if (!two_pass || !(nearest > reach_up))
#else
                    float nearest = 0.0f;
                    const float box_nearest = 0.0f;
                    const bool boxes_clear = false;
#endif
                    {
                        const vec3f step_dir = r.dir * h;
                        const float max_dist = w_length(step_dir);
                        const vec3f dirn = w_normalize(step_dir);
                        float t_hit, plane_nearest;
                        int id;
                        intersect_mesh(s_tri, s_range, P.bvh, r.p, dirn, max_dist, p_l1, t_hit, id, plane_nearest, cnt.face_tests,
                                       boxes_clear, box_nearest);
                        // both are lower bounds of the distance to the geometry: keep the better one (inside a cluster's
                        // box the first-pass bound is 0 and the plane distances take over)
                        nearest = fmaxf(nearest, plane_nearest);
                        if (t_hit > 0.0f && t_hit < max_dist && shade_hit(s_tri, id, r.p, dirn, max_dist, t_hit, color)) {
                            hit = LINON_FACE_ID(s_tri, id);
                            has_color = true;        // result.color.a = t > 0
                            cnt.prehit = i + 1;      // accepted steps that ran the pre-hit path, this one included
                        }
                    }
                    sd.checked(nearest, p_l1, P.plane_w_max);
                    left = has_color ? __int_as_float(0x7f800000) : sd.left(step_cost);  // +inf: a lane that has hit never asks again
                }
                searching = __any_sync(0xffffffffu, !has_color);
            }
            sd.budget = left;
            r.p = advance_point(r.p, r.dir, h);
            r.t = r.t + h;
        }
        for (; i < n_steps; ++i) {
#if LINON_QUIET_ZONE
            // coasting through the quiet zone: the same plan without a budget
            if (!__any_sync(0xffffffffu, !(LinonField<float>::quiet_clearance(r.p) > 0.0f))) {
                LINON_WARP_EVENT(5, 1);
                const int n = min(__reduce_min_sync(0xffffffffu, LinonField<float>::quiet_steps(r.p, r.dir, h)), n_steps - i);
                if (n > 0) {
                    LINON_WARP_EVENT(6, 1);
                    LINON_WARP_EVENT(11, n);
                    r.p = advance_point_n(r.p, r.dir, h, n);
                    r.t = __fmaf_rn((float)n, h, r.t);
                    cnt.quiet += (unsigned int)n;
                    i += n - 1;
                    continue;
                }
            }
#endif
            LINON_WARP_EVENT(7, 1);
            const vec3f v = rk4_velocity(r, h, st, cam);
#if LINON_FW_ONE
            r.dir = v;
#else
            r.dir = (1.0f - st.field_weight) * r.dir + st.field_weight * v;
#endif
            r.p = advance_point(r.p, r.dir, h);
            r.t = r.t + h;
        }
        cnt.accepted = n_steps;
        if (!has_color) cnt.prehit = n_steps;
        if (valid) {
            const size_t pix = (size_t)texture_row(P, local_row) * P.width + x;
            P.raycast[pix] = pack_color(color);  // black when !has_color, alpha = 1 (src/main_view.wgsl:274-277)
            P.endpoint[pix] = make_float4(r.p.x, r.p.y, r.p.z, 1.0f);
            P.hit_id[pix] = hit;
        } else {
            cnt.accepted = cnt.prehit = cnt.rejected = cnt.face_tests = cnt.checks = 0u;
#if LINON_QUIET_ZONE
            cnt.quiet = 0u;
#endif
        }
    }
#else
    const bool valid = slot < P.n_ray_slots && slot_to_pixel(P, slot, x, local_row);
    if (valid) {
        const WgslCamera cam = make_camera(P);
        const WgslSettings st = P.settings;
        Ray r;
        init_ray(P, x, local_row, r);
        bool has_color = false;
        SafeDistance sd;
        sd.reset();
        vec3f color(0.0f, 0.0f, 0.0f);
        int hit = -1;
        for (; r.t <= LINON_T_MAX;) {
            vec3f v, step_dir;
            float diff;
#if LINON_QUIET_ZONE
            // Adaptive frames: the same zone argument per attempt (the step size is the lane's own).  Outside the zone v ==
            // last_v exactly, so diff is 0 there: the reference's rejections and restorations outside the zone are driven by
            // the rounding noise of its evaluation and have no counterpart; entering the zone raises diff from 0,
            // which rejects the attempt and drops h to 0.001 like the reference does a little further in.
            const bool quiet = LinonField<float>::quiet_steps(r.p, r.dir, r.h) > 0;
            if (quiet) ++cnt.quiet;
            if (!rk4_attempt(r, st, cam, v, diff, step_dir, quiet)) {
#else
            if (!rk4_attempt(r, st, cam, v, diff, step_dir)) {
#endif
                ++cnt.rejected;
                continue;
            }
            if (!has_color) {
                const float step_cost = sd.cost(step_dir);
                // Warp-coherent checks: when one pre-hit lane's budget expires, every pre-hit lane of the warp runs the
                // face loop with it.  SIMT pays for the loop once either way; refreshing all budgets together makes the
                // next expiry the warp's minimum instead of 32 independent ones (a warp of 8x4 neighbouring pixels
                // approaches the same planes): 62 % of the warp steps ran a mostly idle face loop before, see
                // profiles/.  A check with budget left finds no group in reach (proof at SafeDistance), so results
                // do not depend on which lanes join - only the counters of executed tests do.
                if (__any_sync(__activemask(), sd.must_check(step_cost))) {
                    const float max_dist = w_length(step_dir);
                    const vec3f dirn = w_normalize(step_dir);
                    const float p_l1 = l1_norm(r.p);
                    float t_hit, nearest;
                    int id;
                    intersect_mesh(s_tri, s_range, P.bvh, r.p, dirn, max_dist, p_l1, t_hit, id, nearest, cnt.face_tests);
                    ++cnt.checks;
                    sd.checked(nearest, p_l1, P.plane_w_max);
                    if (t_hit > 0.0f && t_hit < max_dist && shade_hit(s_tri, id, r.p, dirn, max_dist, t_hit, color)) {
                        hit = LINON_FACE_ID(s_tri, id);
                        has_color = true;           // result.color.a = t > 0
                        cnt.prehit = cnt.accepted + 1;  // accepted steps that ran the pre-hit path, this one included
                    }
                }
                sd.moved(step_cost);
            }
            rk4_advance(r, v, diff, step_dir);
            ++cnt.accepted;
        }
        if (!has_color) cnt.prehit = cnt.accepted;
        const size_t pix = (size_t)texture_row(P, local_row) * P.width + x;
        P.raycast[pix] = pack_color(color);  // black when !has_color, alpha = 1 (src/main_view.wgsl:274-277)
        P.endpoint[pix] = make_float4(r.p.x, r.p.y, r.p.z, 1.0f);
        P.hit_id[pix] = hit;
    }
#endif
    flush_counters(P, cnt);
}

#if LINON_PAIRS && !LINON_BVH
// ------------------------------------------------------------------------------------------------------------
// Kernel 1b: one thread per pair of horizontally adjacent pixels (packed f32x2 arithmetic).
//
// 32 consecutive pair slots form a 16x4 pixel tile: lane l owns pixels (16 tx + 2 (l & 7) + {0, 1}, 4 ty + (l >> 3)),
// so a warp's stores are as contiguous as the tile kernel's.  The RK4 stages, the direction blend and the position
// update of both rays are packed instructions; the safe-distance bookkeeping and the (rare) mesh check run per ray on
// the unpacked component, through the same device functions as the tile kernel.  A ray's result does not depend on its
// partner: in strict mode (component-wise f32x2 arithmetic) the frame is bit-identical with linon_march_tiles.
// An odd frame width leaves the last column's partner outside the frame: it is integrated (harmless) and not stored.
// ------------------------------------------------------------------------------------------------------------
typedef RayT<f32x2> Ray2;
__device__ __forceinline__ vec3f lane_lo(const vec3t<f32x2>& a) { return vec3f(a.x.lo, a.y.lo, a.z.lo); }
__device__ __forceinline__ vec3f lane_hi(const vec3t<f32x2>& a) { return vec3f(a.x.hi, a.y.hi, a.z.hi); }

struct PixelState {  // what one ray of the pair carries besides its packed RK4 state
    SafeDistance sd;
    vec3f color;
    int hit;
    bool has_color;
    unsigned int prehit;
};

// ray_color's face loop for one ray of a pair.  Out of line: it runs on ~3 % of the pre-hit steps (SafeDistance) and
// inlining it twice per pair would only enlarge the hot loop.  Everything travels in registers: 8 floats in,
// {t_hit, face id, nearest plane distance, tests performed} out as one float4 (no per-ray state in local memory).
__device__ __noinline__ float4 pair_intersect(const float4* __restrict__ s_tri, const int* __restrict__ s_range, float px,
                                              float py, float pz, float dx, float dy, float dz, float max_dist, float p_l1) {
    float t_hit, nearest;
    int id;
    unsigned int tests = 0u;
    intersect_mesh(s_tri, s_range, nullptr, vec3f(px, py, pz), vec3f(dx, dy, dz), max_dist, p_l1, t_hit, id, nearest, tests);
    return make_float4(t_hit, __int_as_float(id), nearest, __uint_as_float(tests));
}

// The pre-hit part of one accepted step for one ray, src/main_view.wgsl:259-262.
__device__ __forceinline__ void pair_prehit(const float4* __restrict__ s_tri, const int* __restrict__ s_range,
                                            float plane_w_max, vec3f p, vec3f step_dir, unsigned int step_index,
                                            PixelState& px, StepCounters& cnt) {
    const float step_cost = px.sd.cost(step_dir);
    if (px.sd.must_check(step_cost)) {
        const float max_dist = w_length(step_dir);
        const vec3f dirn = w_normalize(step_dir);
        const float p_l1 = l1_norm(p);
        const float4 r = pair_intersect(s_tri, s_range, p.x, p.y, p.z, dirn.x, dirn.y, dirn.z, max_dist, p_l1);
        cnt.face_tests += __float_as_uint(r.w);
        ++cnt.checks;
        px.sd.checked(r.z, p_l1, plane_w_max);
        const int id = __float_as_int(r.y);
        if (r.x > 0.0f && r.x < max_dist && shade_hit(s_tri, id, p, dirn, max_dist, r.x, px.color)) {
            px.hit = id;
            px.has_color = true;
            px.prehit = step_index + 1;  // accepted steps that ran the pre-hit path, this one included
        }
    }
    px.sd.moved(step_cost);
}

extern "C" __global__ void __launch_bounds__(LINON_BLOCK, LINON_PAIR_MIN_BLOCKS) linon_march_pairs(const MarchParams P) {
    LINON_MESH_PROLOGUE(P)

    const unsigned int slot = blockIdx.x * LINON_BLOCK + threadIdx.x;
    const unsigned int tiles_x = (unsigned int)(P.width + 15) >> 4;
    const unsigned int tile = slot >> 5, lane = slot & 31u;
    const unsigned int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const int x0 = (int)(tx * 16u + 2u * (lane & 7u));
    const int local_row = (int)(ty * 4u + (lane >> 3));
    const bool valid0 = slot < P.n_pair_slots && x0 < P.width && local_row < P.local_rows;
    const bool valid1 = valid0 && x0 + 1 < P.width;
    StepCounters cnt = {0u, 0u, 0u, 0u, 0u};
    if (valid0) {
        const WgslCamera cam = make_camera(P);
        const WgslSettings st = P.settings;
        const int tex_row = texture_row(P, local_row);
        const float gid_y = (float)(P.height - 1 - tex_row);  // src/main_view.wgsl:342
        Ray2 r;
        r.p = vec3t<f32x2>(ld3(P.cam_origin));
        r.sdir = camera_ray_dir<f32x2>(P, f32x2((float)x0, (float)(x0 + 1)), f32x2(gid_y), false);
        r.dir = r.sdir;
        r.t = f32x2(0.0f);
        r.last_v = vec3t<f32x2>(0.0f, 0.0f, 0.0f);
        r.last_diff = f32x2(-1.0f);
        r.h = f32x2(LINON_H_INITIAL);
        PixelState px0, px1;
        px0.sd.reset(); px1.sd.reset();
        px0.color = px1.color = vec3f(0.0f, 0.0f, 0.0f);
        px0.hit = px1.hit = -1;
        px0.has_color = false;
        px1.has_color = !valid1;  // the out-of-frame partner never searches the mesh
        px0.prehit = px1.prehit = 0u;
        unsigned int acc0 = 0u, acc1 = 0u;  // accepted steps per ray
#if LINON_ADAPTIVE
        // src/main_view.wgsl:244-271 with a per-ray `continue`: a rejected or finished ray keeps its state (selects)
        // while its partner goes on
        bool act0 = true, act1 = valid1;
        while (act0 || act1) {
            const f32x2 h = r.h;
            const vec3t<f32x2> v = rk4_velocity(r, h, st, cam);
#if LINON_FUSED_RK4
            const f32x2 diff = w_max(w_length(v - r.last_v), f32x2(LINON_DIFF_FLOOR));  // the noise floor of rk4_attempt
#else
            const f32x2 diff = w_length(v - r.last_v);
#endif
            const bool2 rej = (r.last_diff >= 0.0f) && (diff > LINON_REJECT_FACTOR * r.last_diff) && (h > LINON_REJECT_MIN_H);
            const bool2 rejected(act0 && rej.lo, act1 && rej.hi);
            const bool2 accepted(act0 && !rej.lo, act1 && !rej.hi);
            cnt.rejected += (rejected.lo ? 1u : 0u) + (rejected.hi ? 1u : 0u);
            r.h = s_sel(rejected, f32x2(LINON_SMALL_H), r.h);
#if LINON_FW_ONE
            const vec3t<f32x2> new_dir = v;
#else
            const vec3t<f32x2> new_dir = (1.0f - st.field_weight) * r.dir + st.field_weight * v;
#endif
            const vec3t<f32x2> step_dir = new_dir * h;
            r.dir = w_select(r.dir, new_dir, accepted);
            if (accepted.lo && !px0.has_color) pair_prehit(s_tri, s_range, P.plane_w_max, lane_lo(r.p), lane_lo(step_dir), acc0, px0, cnt);
            if (accepted.hi && !px1.has_color) pair_prehit(s_tri, s_range, P.plane_w_max, lane_hi(r.p), lane_hi(step_dir), acc1, px1, cnt);
            r.p = w_select(r.p, r.p + step_dir, accepted);
            r.t = s_sel(accepted, r.t + h, r.t);
            r.last_v = w_select(r.last_v, v, accepted);
#if LINON_FUSED_RK4
            const bool2 restore = accepted && ((LINON_RESTORE_FACTOR * diff < r.last_diff) || (diff <= LINON_DIFF_FLOOR)) && (r.h < LINON_H_INITIAL);
#else
            const bool2 restore = accepted && (LINON_RESTORE_FACTOR * diff < r.last_diff) && (r.h < LINON_H_INITIAL);
#endif
            r.h = s_sel(restore, f32x2(LINON_H_INITIAL), r.h);
            r.last_diff = s_sel(accepted, diff, r.last_diff);
            acc0 += accepted.lo ? 1u : 0u;
            acc1 += accepted.hi ? 1u : 0u;
            act0 = act0 && (r.t.lo <= LINON_T_MAX);
            act1 = act1 && (r.t.hi <= LINON_T_MAX);
        }
#else
        // fixed step: both rays are in lockstep (same t), every attempt is accepted
        for (; r.t.lo <= LINON_T_MAX;) {
            const vec3t<f32x2> v = rk4_velocity(r, LINON_H_INITIAL, st, cam);
#if LINON_FW_ONE
            r.dir = v;
#else
            r.dir = (1.0f - st.field_weight) * r.dir + st.field_weight * v;
#endif
            const vec3t<f32x2> step_dir = r.dir * LINON_H_INITIAL;
            if (!px0.has_color) pair_prehit(s_tri, s_range, P.plane_w_max, lane_lo(r.p), lane_lo(step_dir), acc0, px0, cnt);
            if (!px1.has_color) pair_prehit(s_tri, s_range, P.plane_w_max, lane_hi(r.p), lane_hi(step_dir), acc0, px1, cnt);
            r.p = r.p + step_dir;
            r.t = r.t + LINON_H_INITIAL;
            ++acc0;
        }
        acc1 = valid1 ? acc0 : 0u;
#endif
        if (!px0.has_color) px0.prehit = acc0;
        if (valid1 && px1.hit < 0) px1.prehit = acc1;
        cnt.accepted = acc0 + acc1;
        cnt.prehit = px0.prehit + (valid1 ? px1.prehit : 0u);
        const size_t pix = (size_t)tex_row * P.width + x0;
        P.raycast[pix] = pack_color(px0.color);  // black when !has_color, alpha = 1 (src/main_view.wgsl:274-277)
        P.endpoint[pix] = make_float4(r.p.x.lo, r.p.y.lo, r.p.z.lo, 1.0f);
        P.hit_id[pix] = px0.hit;
        if (valid1) {
            P.raycast[pix + 1] = pack_color(px1.color);
            P.endpoint[pix + 1] = make_float4(r.p.x.hi, r.p.y.hi, r.p.z.hi, 1.0f);
            P.hit_id[pix + 1] = px1.hit;
        }
    }
    flush_counters(P, cnt);
}
#endif  // LINON_PAIRS

// ------------------------------------------------------------------------------------------------------------
// Kernels 2a/2b: persistent warps with lane refill, split by phase.
//
//   linon_march_phase_a   every lane carries a ray that has NOT hit yet (RK4 + face loop each step).  A lane whose
//                         ray hits writes colour + hit id, appends {p, dir, t, h, last_v, last_diff, pixel} to the
//                         device queue (warp-aggregated atomicAdd on the tail) and is refilled with a fresh ray from
//                         the cursor (warp-aggregated atomicAdd).  A ray that reaches t > 5 without a hit writes
//                         black / -1 / endpoint.
//   linon_march_phase_b   launched after phase A on the same stream, so the queue is complete and immutable: lanes
//                         pop entries with a plain atomicAdd on the head (no CAS, no publication flags, no fences)
//                         and integrate RK4 only until t > 5, then write the endpoint.
// All waiting is removed by construction: neither kernel ever spins on another warp's progress.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int lane_rank(unsigned int mask) { return __popc(mask & ((1u << (threadIdx.x & 31)) - 1u)); }

#if LINON_PROFILE
#define PF_DECL unsigned long long pf_refill = 0, pf_step = 0, pf_iters = 0, pf_lanes = 0; const long long pf_start = clock64();
#define PF_BEGIN(name) const long long name = clock64();
#define PF_END(acc, name) acc += (unsigned long long)(clock64() - name);
#define PF_FLUSH(base)                                                                   \
    if ((threadIdx.x & 31) == 0) {                                                       \
        atomicAdd(&P.counters[(base) + 0], pf_refill);                                   \
        atomicAdd(&P.counters[(base) + 1], pf_step);                                     \
        atomicAdd(&P.counters[(base) + 2], pf_iters);                                    \
        atomicAdd(&P.counters[(base) + 3], pf_lanes);                                    \
        atomicAdd(&P.counters[(base) + 4], (unsigned long long)(clock64() - pf_start));  \
    }
#else
#define PF_DECL
#define PF_BEGIN(name)
#define PF_END(acc, name)
#define PF_FLUSH(base)
#endif

// ------------------------------------------------------------------------------------------------------------
// Kernel 3: persistent warps, ONE phase.  Every lane carries a ray from its first step to t > 5 (pre-hit and post-hit
// in the same lane, like the tile kernel); a lane whose ray finishes takes the next ray slot from a global cursor
// (one warp-aggregated atomicAdd per refill, and only once LINON_REFILL_MIN lanes are idle).  Made for ADAPTIVE frames:
// with adaptive sampling on, accepted + rejected attempts per ray spread from 52 to several thousand between
// neighbouring pixels (rounding-noise driven), so a tile warp idles most of its lanes while the slowest ray
// finishes.  Fixed-step frames have no such spread and stay on the tile kernel.
// ------------------------------------------------------------------------------------------------------------
extern "C" __global__ void __launch_bounds__(LINON_BLOCK, LINON_MIN_BLOCKS) linon_march_refill(const MarchParams P) {
    LINON_MESH_PROLOGUE(P)

    const unsigned int lane = threadIdx.x & 31;
    const WgslCamera cam = make_camera(P);
    const WgslSettings st = P.settings;
    unsigned int* const w_next = P.work + 0;

    Ray r;
    r.p = r.dir = r.sdir = r.last_v = vec3f(0.0f, 0.0f, 0.0f);
    r.t = 0.0f; r.h = LINON_H_INITIAL; r.last_diff = -1.0f;
    int pix = -1;
    bool alive = false, has_color = false;
    bool fresh_left = true;  // warp-uniform
    SafeDistance sd;
    sd.reset();
    vec3f color(0.0f, 0.0f, 0.0f);
    int hit = -1;
    unsigned int steps = 0u;  // accepted steps of the ray this lane carries
    unsigned int chunk_next = 0u, chunk_end = 0u;  // warp-uniform: the slots this warp still owns
    StepCounters cnt = {0u, 0u, 0u, 0u, 0u};

    for (;;) {
        const unsigned int idle = __ballot_sync(0xffffffffu, !alive);
        if (fresh_left && (__popc(idle) >= LINON_REFILL_MIN)) {
            // ray slots come in chunks of 2 tiles: ONE global atomicAdd per 64 rays (a single cursor address serialises
            // at ~10 ns per atomic in L2 - with one atomic per refill that alone cost more than the march), then the warp
            // hands the chunk out from registers
            if (chunk_next >= chunk_end) {
                unsigned int base = 0;
                if (lane == 0) base = atomicAdd(w_next, LINON_REFILL_CHUNK);
                chunk_next = __shfl_sync(0xffffffffu, base, 0);
                chunk_end = min(chunk_next + LINON_REFILL_CHUNK, P.n_ray_slots);
                if (chunk_next >= P.n_ray_slots) {
                    fresh_left = false;
                    chunk_end = chunk_next;
                }
            }
            const unsigned int rank = lane_rank(idle);
            if (!alive && chunk_next + rank < chunk_end) {
                const unsigned int slot = chunk_next + rank;
                int x, local_row;
                if (slot_to_pixel(P, slot, x, local_row)) {
                    init_ray(P, x, local_row, r);
                    sd.reset();
                    pix = texture_row(P, local_row) * P.width + x;
                    alive = true;
                    has_color = false;
                    color = vec3f(0.0f, 0.0f, 0.0f);
                    hit = -1;
                    steps = 0u;
                }
            }
            chunk_next = min(chunk_next + (unsigned int)__popc(idle), chunk_end);
        }
        if (__ballot_sync(0xffffffffu, alive) == 0u) {
            if (fresh_left) continue;
            break;
        }
        if (alive) {
            vec3f v, step_dir;
            float diff;
            if (rk4_attempt(r, st, cam, v, diff, step_dir)) {
                if (!has_color) {
                    const float step_cost = sd.cost(step_dir);
                    if (__any_sync(__activemask(), sd.must_check(step_cost))) {
                        const float max_dist = w_length(step_dir);
                        const vec3f dirn = w_normalize(step_dir);
                        const float p_l1 = l1_norm(r.p);
                        float t_hit, nearest;
                        int id;
                        intersect_mesh(s_tri, s_range, P.bvh, r.p, dirn, max_dist, p_l1, t_hit, id, nearest, cnt.face_tests);
                        ++cnt.checks;
                        sd.checked(nearest, p_l1, P.plane_w_max);
                        if (t_hit > 0.0f && t_hit < max_dist && shade_hit(s_tri, id, r.p, dirn, max_dist, t_hit, color)) {
                            hit = LINON_FACE_ID(s_tri, id);
                            has_color = true;
                            cnt.prehit += steps + 1;
                        }
                    }
                    sd.moved(step_cost);
                }
                rk4_advance(r, v, diff, step_dir);
                ++steps;
                ++cnt.accepted;
                if (!(r.t <= LINON_T_MAX)) {
                    if (!has_color) cnt.prehit += steps;
                    P.raycast[pix] = pack_color(color);
                    P.endpoint[pix] = make_float4(r.p.x, r.p.y, r.p.z, 1.0f);
                    P.hit_id[pix] = hit;
                    alive = false;
                }
            } else {
                ++cnt.rejected;
            }
        }
    }
    flush_counters(P, cnt);
}

extern "C" __global__ void __launch_bounds__(LINON_BLOCK, LINON_MIN_BLOCKS) linon_march_phase_a(const MarchParams P) {
    LINON_MESH_PROLOGUE(P)

    const unsigned int lane = threadIdx.x & 31;
    const WgslCamera cam = make_camera(P);
    const WgslSettings st = P.settings;
    unsigned int* const w_next = P.work + 0;
    unsigned int* const w_tail = P.work + 1;
    const size_t plane = P.q_capacity;

    Ray r;
    r.p = r.dir = r.sdir = r.last_v = vec3f(0.0f, 0.0f, 0.0f);
    r.t = 0.0f; r.h = LINON_H_INITIAL; r.last_diff = -1.0f;
    int pix = -1;            // pixel index of the ray this lane carries
    bool alive = false;
    bool fresh_left = true;  // warp-uniform
    SafeDistance sd;
    sd.reset();
    StepCounters cnt = {0u, 0u, 0u, 0u, 0u};
    PF_DECL

    for (;;) {
        const unsigned int idle = __ballot_sync(0xffffffffu, !alive);
        if (fresh_left && (__popc(idle) >= LINON_REFILL_MIN)) {
            PF_BEGIN(pf_r0)
            const unsigned int need = __popc(idle);
            unsigned int base = 0;
            if (lane == 0) base = atomicAdd(w_next, need);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (!alive) {
                const unsigned int slot = base + lane_rank(idle);
                int x, local_row;
                if (slot < P.n_ray_slots && slot_to_pixel(P, slot, x, local_row)) {
                    init_ray(P, x, local_row, r);
                    sd.reset();
                    pix = texture_row(P, local_row) * P.width + x;
                    alive = true;
                }
            }
            if (base + need >= P.n_ray_slots) fresh_left = false;
            PF_END(pf_refill, pf_r0)
        }
        const unsigned int live = __ballot_sync(0xffffffffu, alive);
        if (live == 0u) {
            if (fresh_left) continue;
            break;
        }
#if LINON_PROFILE
        ++pf_iters;
        pf_lanes += __popc(live);
#endif
        PF_BEGIN(pf_s0)
        bool push = false;
        if (alive) {
            vec3f v, step_dir;
            float diff;
            if (rk4_attempt(r, st, cam, v, diff, step_dir)) {
                const float step_cost = sd.cost(step_dir);
                bool hit_now = false;
                if (sd.must_check(step_cost)) {
                    const float max_dist = w_length(step_dir);
                    const vec3f dirn = w_normalize(step_dir);
                    const float p_l1 = l1_norm(r.p);
                    float t_hit, nearest;
                    int id;
                    intersect_mesh(s_tri, s_range, P.bvh, r.p, dirn, max_dist, p_l1, t_hit, id, nearest, cnt.face_tests);
                    ++cnt.checks;
                    sd.checked(nearest, p_l1, P.plane_w_max);
                    vec3f rgb;
                    hit_now = t_hit > 0.0f && t_hit < max_dist && shade_hit(s_tri, id, r.p, dirn, max_dist, t_hit, rgb);
                    if (hit_now) {
                        P.raycast[pix] = pack_color(rgb);
                        P.hit_id[pix] = LINON_FACE_ID(s_tri, id);
                    }
                }
                sd.moved(step_cost);
                ++cnt.prehit;
                rk4_advance(r, v, diff, step_dir);
                ++cnt.accepted;
                const bool finished = !(r.t <= LINON_T_MAX);
                if (finished) {
                    P.endpoint[pix] = make_float4(r.p.x, r.p.y, r.p.z, 1.0f);
                    if (!hit_now) {  // never hit: black, alpha 1 (src/main_view.wgsl:275-277)
                        P.raycast[pix] = make_uchar4(0, 0, 0, 255);
                        P.hit_id[pix] = -1;
                    }
                }
                push = hit_now && !finished;
                if (finished || hit_now) alive = false;
            } else {
                ++cnt.rejected;
            }
        }
        const unsigned int push_mask = __ballot_sync(0xffffffffu, push);
        if (push_mask) {
            unsigned int base = 0;
            if (lane == 0) base = atomicAdd(w_tail, (unsigned int)__popc(push_mask));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (push) {
                const unsigned int e = base + lane_rank(push_mask);
                P.q_state[e] = make_float4(r.p.x, r.p.y, r.p.z, r.t);
                P.q_state[plane + e] = make_float4(r.dir.x, r.dir.y, r.dir.z, r.h);
                P.q_state[2 * plane + e] = make_float4(r.last_v.x, r.last_v.y, r.last_v.z, r.last_diff);
                P.q_pixel[e] = pix;
            }
        }
        PF_END(pf_step, pf_s0)
    }
    flush_counters(P, cnt);
    PF_FLUSH(5)
}

extern "C" __global__ void __launch_bounds__(LINON_BLOCK, LINON_MIN_BLOCKS) linon_march_phase_b(const MarchParams P) {
    const unsigned int lane = threadIdx.x & 31;
    const WgslCamera cam = make_camera(P);
    const WgslSettings st = P.settings;
    unsigned int* const w_head = P.work + 2;
    const unsigned int total = P.work[1];  // final queue length: phase A has completed
    const size_t plane = P.q_capacity;

    Ray r;
    r.p = r.dir = r.sdir = r.last_v = vec3f(0.0f, 0.0f, 0.0f);
    r.t = 0.0f; r.h = LINON_H_INITIAL; r.last_diff = -1.0f;
    int pix = -1;
    bool alive = false;
    bool queue_left = true;  // warp-uniform
    StepCounters cnt = {0u, 0u, 0u, 0u, 0u};
    PF_DECL

    for (;;) {
        const unsigned int idle = __ballot_sync(0xffffffffu, !alive);
        if (queue_left && (__popc(idle) >= LINON_REFILL_MIN)) {
            PF_BEGIN(pf_r0)
            const unsigned int need = __popc(idle);
            unsigned int base = 0;
            if (lane == 0) base = atomicAdd(w_head, need);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (!alive) {
                const unsigned int e = base + lane_rank(idle);
                if (e < total) {
                    const float4 a = P.q_state[e], b = P.q_state[plane + e], c = P.q_state[2 * plane + e];
                    pix = P.q_pixel[e];
                    r.p = vec3f(a.x, a.y, a.z); r.t = a.w;
                    r.dir = vec3f(b.x, b.y, b.z); r.h = b.w;
                    r.last_v = vec3f(c.x, c.y, c.z); r.last_diff = c.w;
                    const int row = pix / P.width;  // texture row
                    r.sdir = camera_ray_dir<float>(P, (float)(pix - row * P.width), (float)(P.height - 1 - row), false);
                    alive = true;
                }
            }
            if (base + need >= total) queue_left = false;
            PF_END(pf_refill, pf_r0)
        }
        const unsigned int live = __ballot_sync(0xffffffffu, alive);
        if (live == 0u) {
            if (queue_left) continue;
            break;
        }
#if LINON_PROFILE
        ++pf_iters;
        pf_lanes += __popc(live);
#endif
        PF_BEGIN(pf_s0)
        if (alive) {
            vec3f v, step_dir;
            float diff;
            if (rk4_attempt(r, st, cam, v, diff, step_dir)) {
                rk4_advance(r, v, diff, step_dir);
                ++cnt.accepted;
                if (!(r.t <= LINON_T_MAX)) {
                    P.endpoint[pix] = make_float4(r.p.x, r.p.y, r.p.z, 1.0f);
                    alive = false;
                }
            } else {
                ++cnt.rejected;
            }
        }
        PF_END(pf_step, pf_s0)
    }
    flush_counters(P, cnt);
    PF_FLUSH(10)
}

// ------------------------------------------------------------------------------------------------------------
// sample_rays: the 8 reference-view rays, src/main_view.wgsl:281-314 + :469-479.  One thread per ray.
// pos16 = 8 x (x, y) normalised image positions (click ring or outline search results).
// ------------------------------------------------------------------------------------------------------------
struct SampleParams {
    MarchParams march;
    const float* pos16;
    float4* ray_samples;  // 800 x {position, colour}
};

#ifndef LINON_SAMPLE_COLORS  // `var<private> sample_colors`, src/main_view.wgsl:326-335 (an edited module passes its own values)
#define LINON_SAMPLE_COLORS                                                            \
    {1.0f, 0.0f, 0.0f}, {0.5f, 0.5f, 0.0f}, {0.0f, 1.0f, 0.0f}, {0.5f, 1.0f, 0.0f}, \
    {1.0f, 1.0f, 0.0f}, {0.5f, 0.5f, 0.5f}, {0.0f, 0.0f, 1.0f}, {0.5f, 0.0f, 0.5f}
#endif
__constant__ float c_sample_colors[8][3] = {LINON_SAMPLE_COLORS};

extern "C" __global__ void __launch_bounds__(32) linon_sample_rays(const SampleParams S) {
    // ONE RAY PER CTA (grid = 8, one warp each; every lane carries the same ray, lane 0 stores): the reference marches the eight
    // rays on threads gid.x < 8 of one workgroup (src/main_view.wgsl:281-314), and in one warp their control flow interleaves -
    // whenever ANY of the eight is inside a mirage's zone of influence all eight pay a full Runge-Kutta evaluation and a zone
    // query per step.  Eight warps on eight SMs follow their own ray each; the arithmetic of a ray is unchanged.
    const int k = blockIdx.x;
    if (k >= 8) return;
    const bool writer = threadIdx.x == 0;
    const MarchParams& P = S.march;
    const WgslCamera cam = make_camera(P);
    const WgslSettings st = P.settings;
    Ray r;
    r.p = ld3(P.cam_origin);
    r.sdir = camera_ray_dir<float>(P, S.pos16[2 * k], S.pos16[2 * k + 1], true);
    r.dir = r.sdir;
    r.t = 0.0f;
    r.last_v = vec3f(0.0f, 0.0f, 0.0f);
    r.last_diff = -1.0f;
    r.h = LINON_SAMPLE_H;
    const float h = LINON_SAMPLE_H;
    const float4 colour = make_float4(c_sample_colors[k][0], c_sample_colors[k][1], c_sample_colors[k][2], 0.5f);
#if LINON_QUIET_ZONE && LINON_FW_ONE
    // LINON_QUIET_ZONE: outside the body's zone of influence the Runge-Kutta velocity is the direction itself, so the ray asks
    // once how many steps it stays outside (quiet_steps: exact entry of the straight line into the zone) and takes them as
    // straight steps, up to the next sampled step at a time - every one of the 5000 steps is accounted for bit by bit
    // (advance_point_n) and every 50th is still sampled.
    int quiet_left = 0;
#endif
    constexpr int sample_step_size = LINON_SAMPLE_STEPS / 100;  // steps / sample_steps, src/main_view.wgsl:292-293
    for (int i = 0; i < LINON_SAMPLE_STEPS; ++i) {
        if (i % sample_step_size == 0 && writer) {  // the sample is the position BEFORE the step's advance (:304-309)
            const int index = k * 100 + i / sample_step_size;
            S.ray_samples[2 * index] = make_float4(r.p.x, r.p.y, r.p.z, 1.0f);
            S.ray_samples[2 * index + 1] = colour;
        }
#if LINON_QUIET_ZONE && LINON_FW_ONE
        if (quiet_left == 0) quiet_left = min(LinonField<float>::quiet_steps(r.p, r.dir, h), LINON_SAMPLE_STEPS);
        if (quiet_left > 0) {
            // straight steps up to (not including) the next sampled step, the end of the quiet run or the last step
            const int to_sample = sample_step_size - i % sample_step_size;
            const int n = min(min(quiet_left, to_sample), LINON_SAMPLE_STEPS - i);
            r.p = advance_point_n(r.p, r.dir, h, n);
            r.t = __fmaf_rn((float)n, h, r.t);
            quiet_left -= n;
            i += n - 1;  // the loop header adds the last one
            continue;
        }
#endif
        const vec3f v = rk4_velocity(r, h, st, cam);
#if LINON_FW_ONE
        r.dir = v;
#else
        r.dir = (1.0f - st.field_weight) * r.dir + st.field_weight * v;
#endif
        const vec3f step_dir = r.dir * h;
        r.p = r.p + step_dir;
        r.t = r.t + h;
    }
}

// ------------------------------------------------------------------------------------------------------------
// Debug / parity entry points (unit tests call these through linon_debug_*; not used by render).
// ------------------------------------------------------------------------------------------------------------
extern "C" __global__ void linon_eval_field(int n, const float* p_prev, const float* p, const float* v0, const float* v,
                                            const float* t, WgslSettings st, float* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    WgslCamera cam;
    const vec3f r = field_function<float>(ld3(p_prev + 3 * i), ld3(p + 3 * i), ld3(v0 + 3 * i), ld3(v + 3 * i), t[i], st, cam);
    out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z;
}

// One Runge-Kutta velocity (src/main_view.wgsl:246-251) per element, through the SAME rk4_velocity the march kernels call
// (generic four-stage evaluation, or the whole-step restatement of mirage_fused.cuh): v0 = start direction, h per element.
extern "C" __global__ void linon_eval_rk4(int n, const float* p, const float* dir, const float* v0, const float* t,
                                          const float* h, WgslSettings st, float* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    WgslCamera cam;
    Ray r;
    r.p = ld3(p + 3 * i);
    r.dir = ld3(dir + 3 * i);
    r.sdir = ld3(v0 + 3 * i);
    r.t = t[i];
    r.h = h[i];
    r.last_v = vec3f(0.0f, 0.0f, 0.0f);
    r.last_diff = -1.0f;
    const vec3f v = rk4_velocity<float, float>(r, r.h, st, cam);
    out[3 * i] = v.x; out[3 * i + 1] = v.y; out[3 * i + 2] = v.z;
}

// which: 0 rotateX 1 rotateY 2 rotateZ 3 translate 4 refraction 5 point_plane_distance 6 sigmoid 7 refraction_index
//        8 reflect 9 normalize 10 cross        (same numbering as the oracle's oracle_helper)
extern "C" __global__ void linon_eval_helper(int which, int n, const float* v, const float* nrm, const float* p0,
                                             const float* a, const float* b, const float* c, float* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const vec3f V = ld3(v + 3 * i), N = ld3(nrm + 3 * i), P0 = ld3(p0 + 3 * i);
    vec3f r(0.0f, 0.0f, 0.0f);
    switch (which) {
        case 0: r = rotateX(V, a[i]); break;
        case 1: r = rotateY(V, a[i]); break;
        case 2: r = rotateZ(V, a[i]); break;
        case 3: r = translate(V, a[i], b[i], c[i]); break;
        case 4: r = refraction(a[i], b[i], V, N); break;
        case 5: r.x = point_plane_distance(V, N, P0); break;
        case 6: r.x = sigmoid(a[i]); break;
        case 7: r.x = refraction_index(a[i]); break;
        case 8: r = w_reflect(V, N); break;
        case 9: r = w_normalize(V); break;
        case 10: r = w_cross(V, N); break;
    }
    out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z;
}

// hit_triangle on n independent (triangle, origin, direction) triples -> t or -1 (src/main_view.wgsl:150-175)
extern "C" __global__ void linon_eval_hit_triangle(int n, const float* tri, const float* origin, const float* dir, float* t_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const vec3f v0 = ld3(tri + 9 * i), v1 = ld3(tri + 9 * i + 3), v2 = ld3(tri + 9 * i + 6);
    const vec3f edge1 = v1 - v0, edge2 = v2 - v0;
    const vec3f o = ld3(origin + 3 * i), d = ld3(dir + 3 * i);
    const vec3f h = w_cross(d, edge2);
    const float det = w_dot(edge1, h);
    const float f = 1.0f / det;
    const vec3f s = o - v0;
    const float u = f * w_dot(s, h);
    const vec3f q = w_cross(s, edge1);
    const float v = f * w_dot(d, q);
    const float t = f * w_dot(edge2, q);
    const bool ok = !(det > -LINON_EPS && det < LINON_EPS) && !(u < 0.0f || u > 1.0f) && !(v < 0.0f || u + v > 1.0f) &&
                    (t > LINON_EPS);
    t_out[i] = ok ? t : -1.0f;
}
