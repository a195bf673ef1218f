// emu_driver.cpp — TEST INFRASTRUCTURE: runs linon_march_tiles, compiled for the host from the exact NVRTC source text of
// a variant (-DLINON_EMU_SOURCE="path", dumped by LINON_DUMP_SOURCE), one warp per CTA through cuda_shim.hpp.
// Mesh tables come from the product's own host code (mesh_tables.hpp), the launch geometry mirrors linon_render.
#include "cuda_shim.hpp"

#include LINON_EMU_SOURCE

#undef threadIdx
#undef blockIdx
#undef blockDim
#undef gridDim
#include <vector>

#include "mesh_tables.hpp"

static_assert(LINON_BLOCK == 32, "the emulation runs one warp per CTA");

extern "C" {

struct EmuFrame {
    int32_t width, height;
    float cam_origin[4], cam_view[4], cam_up[4];
    float settings[6];  // 24 bytes, src/main_view.rs:43-51 (two of them are i32: copied bytewise)
    const float* verts;
    const uint32_t* faces;
    int32_t n_verts, n_faces;
    uint8_t* rgba8;       // [H][W][4]
    float* endpoint;      // [H][W][4]
    int32_t* hit_id;      // [H][W]
    uint64_t* counters;   // [16]
    int32_t n_threads;
    int32_t tile_row_step;  // > 1: only every n-th row of 8x4 tiles is run (sampled counters of a large frame)
    int32_t kernel;         // LINON_OPT_KERNEL: 0 tiles, 1 persistent (phase A, phase B), 2 pairs, 3 persistent (one phase)
    int32_t grid;           // persistent kernels: CTAs (0 = one per host thread)
    const float* pos16;     // emu_sample_rays: 8 normalised start positions
    float* ray_samples;     // emu_sample_rays: 800 x {position, colour}
};

static int fixed_step_count(float h) {  // capi.cpp: fixed_step_count
    if (!(h > 0.0f)) return 0;
    volatile float t = 0.0f;
    int n = 0;
    while (t <= LINON_T_MAX && n < (1 << 24)) {  // the module's horizon (march_kernels.cuh; 5.0 as shipped)
        t = t + h;
        ++n;
    }
    return n;
}

static void kernel_entry(void* arg) { linon_march_tiles(*(const MarchParams*)arg); }
#if LINON_PAIRS && !LINON_BVH
static void pairs_entry(void* arg) { linon_march_pairs(*(const MarchParams*)arg); }
#endif
static void refill_entry(void* arg) { linon_march_refill(*(const MarchParams*)arg); }
static void phase_a_entry(void* arg) { linon_march_phase_a(*(const MarchParams*)arg); }
static void phase_b_entry(void* arg) { linon_march_phase_b(*(const MarchParams*)arg); }
static void sample_entry(void* arg) { linon_sample_rays(*(const SampleParams*)arg); }

static int run_grid(void (*entry)(void*), void* arg, int n_blocks, int n_threads, int tile_row_step, int tiles_x) {
    int bad = 0;
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
    for (int b = 0; b < n_blocks; ++b) {
        if (tile_row_step > 1 && (b / tiles_x) % tile_row_step != 0) continue;
        emu::g_grid_dim = {(unsigned int)n_blocks, 1, 1};
        if (!emu::run_warp(entry, arg, (unsigned int)b)) {
#pragma omp atomic write
            bad = 1;
        }
    }
    return bad;
}

#if !LINON_STRICT_MATH
// advance_coord (march_kernels.cuh: n straight steps of one coordinate without executing them one by one) against the chain it
// replaces, p <- fma(d, h, p) n times with the host's correctly rounded fmaf
void emu_advance_coord(int cases, const float* p, const float* d, float h, const int* n, float* jumped, float* stepped) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < cases; ++i) {
        jumped[i] = advance_coord(p[i], d[i], h, n[i]);
        float x = p[i];
        for (int k = 0; k < n[i]; ++k) x = fmaf(d[i], h, x);
        stepped[i] = x;
    }
}
#endif

#if LINON_QUIET_ZONE
// the two planning functions of the quiet zone on arrays of rays (property tests): steps outside the body's zone of
// influence and, for the start-up tables of `verts / faces`, steps of a straight ray that cannot reach a face
void emu_quiet_plan(int n, const float* p, const float* dir, float h, const float* verts, int n_verts, const uint32_t* faces,
                    int n_faces, int n_steps_left, int* quiet_steps, int* clear_steps) {
    linon::MeshTables mt;
    std::vector<float> tris;
    if (clear_steps) {
        linon::build_mesh_tables(verts, (uint32_t)n_verts, faces, (uint32_t)n_faces, mt);
        tris = mt.tris;
        tris.insert(tris.end(), mt.planes.begin(), mt.planes.end());
        tris.insert(tris.end(), mt.p1_planes.begin(), mt.p1_planes.end());
        tris.insert(tris.end(), mt.boxes.begin(), mt.boxes.end());
    }
    for (int i = 0; i < n; ++i) {
        const vec3f pp(p[3 * i], p[3 * i + 1], p[3 * i + 2]), dd(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]);
        if (quiet_steps) quiet_steps[i] = LinonField<float>::quiet_steps(pp, dd, h);
#if LINON_CULL && !LINON_BVH
        if (clear_steps) {
            const vec3f d = dd * h;
            clear_steps[i] = straight_clear_steps((const float4*)tris.data(), pp, d, (float)n_steps_left * l1_norm(d), mt.coord_max);
        }
#endif
    }
}
#endif

// returns 0 on success, 1 when the mesh tables do not match the variant, 2 when a warp collective was abandoned
int emu_march(const EmuFrame* f) {
#if LINON_BVH
    linon::BvhTables bt;  // large meshes: faces in hierarchy order + the hierarchy, like set_mesh_bvh
    if (linon::build_bvh_tables(f->verts, (uint32_t)f->n_verts, f->faces, (uint32_t)f->n_faces, bt) != linon::BVH_OK) return 1;
    const std::vector<float>& tris = bt.tris;
    std::vector<int> ranges(1, 0);
    const float plane_w_max = bt.extent;
#else
    linon::MeshTables mt;
    if (!linon::build_mesh_tables(f->verts, (uint32_t)f->n_verts, f->faces, (uint32_t)f->n_faces, mt)) return 1;
    if ((int)mt.n_groups() != LINON_N_GROUPS || (int)mt.n_p1() != LINON_N_P1 || (int)mt.n_boxes() != LINON_N_BOXES ||
        (int)mt.n_segments() != LINON_N_SEGS || f->n_faces != LINON_N_FACES)
        return 1;
    std::vector<float> tris = mt.tris;
    tris.insert(tris.end(), mt.planes.begin(), mt.planes.end());
    tris.insert(tris.end(), mt.p1_planes.begin(), mt.p1_planes.end());
    tris.insert(tris.end(), mt.boxes.begin(), mt.boxes.end());
    std::vector<int> ranges = mt.ranges;
    ranges.insert(ranges.end(), mt.segments.begin(), mt.segments.end());
    const float plane_w_max = mt.coord_max;
#endif

    MarchParams P;
    memset(&P, 0, sizeof(P));
    memcpy(P.cam_origin, f->cam_origin, 16);
    memcpy(P.cam_view, f->cam_view, 16);
    memcpy(P.cam_up, f->cam_up, 16);
    memcpy(&P.settings, f->settings, 24);
    P.width = f->width;
    P.height = f->height;
    P.local_rows = f->height;
    P.shard_rank = 0;
    P.shard_world = 1;
    P.rows_per_block = 32;
    P.tris = (const float4*)tris.data();
    P.group_range = ranges.data();
    P.raycast = (uchar4*)f->rgba8;
    P.endpoint = (float4*)f->endpoint;
    P.hit_id = f->hit_id;
    P.counters = (unsigned long long*)f->counters;
    const unsigned int tiles_x = (f->width + 7) / 8, tiles_y = (f->height + 3) / 4;
    P.n_ray_slots = tiles_x * tiles_y * 32u;
    P.plane_w_max = plane_w_max;
#if LINON_BVH
    P.bvh = (const float4*)bt.nodes.data();
#endif
    P.n_fixed_steps = LINON_ADAPTIVE ? 0 : fixed_step_count(LINON_H_INITIAL);
    P.n_pair_slots = (((unsigned int)f->width + 15) / 16) * tiles_y * 32u;
    const int n_threads = f->n_threads > 0 ? f->n_threads : 8;
    const int grid = f->grid > 0 ? f->grid : n_threads;
    std::vector<unsigned int> work(16, 0u);
    std::vector<int> q_pixel;
    std::vector<float4> q_state;
    P.work = work.data();
    int bad = 0;
    if (f->ray_samples) {  // linon_sample_rays: eight CTAs of one warp, one ray each
        SampleParams S;
        S.march = P;
        S.pos16 = f->pos16;
        S.ray_samples = (float4*)f->ray_samples;
        bad = run_grid(sample_entry, &S, 8, n_threads, 1, 1);
    } else if (f->kernel == 0) {
        bad = run_grid(kernel_entry, &P, (int)(P.n_ray_slots / 32u), n_threads, f->tile_row_step, (int)tiles_x);
    } else if (f->kernel == 2) {
#if LINON_PAIRS && !LINON_BVH
        bad = run_grid(pairs_entry, &P, (int)(P.n_pair_slots / 32u), n_threads, 1, 1);
#else
        return 3;
#endif
    } else if (f->kernel == 3) {
        bad = run_grid(refill_entry, &P, grid, n_threads, 1, 1);
    } else if (f->kernel == 1) {
        const size_t cap = (size_t)f->width * f->height;
        q_pixel.assign(cap, 0);
        q_state.assign(3 * cap, float4{0, 0, 0, 0});
        P.q_pixel = q_pixel.data();
        P.q_state = q_state.data();
        P.q_capacity = (unsigned int)cap;
        bad = run_grid(phase_a_entry, &P, grid, n_threads, 1, 1);
        if (!bad) bad = run_grid(phase_b_entry, &P, grid, n_threads, 1, 1);  // phase B sees the complete queue
    } else {
        return 3;
    }
    return bad ? 2 : 0;
}
}
