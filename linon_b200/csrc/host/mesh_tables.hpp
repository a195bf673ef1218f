// mesh_tables.hpp — what the march kernels stage into shared memory for a mesh of up to 1024 faces (host side, pure).
//
//   faces     n_faces x {v0, e1 = v1 - v0, e2 = v2 - v0} (3 x float4): the two subtractions hit_triangle performs
//             (src/main_view.wgsl:152-153), done once
//   planes    one {unit normal, offset} per PLANE GROUP = run of consecutive coplanar faces (level 1 of the exact culling,
//             see intersect_mesh in device/march_kernels.cuh); `ranges` holds (first face | count << 16)
//   bounds    the first pass of a culling check (nearest_geometry_distance): a lower bound of the distance from a point
//             to any triangle.  A plane group that spans the scene along two axes (floor, ceiling, walls of the Cornell box)
//             keeps its plane - near it the plane IS the geometry.  The other groups are clustered by connectivity (faces
//             sharing a vertex position) and each cluster (the short box, the tall box) contributes ONE axis-aligned
//             bounding box instead of its planes: the planes of
//             the rotated inner boxes cut through the whole room and past the camera, so with plane distances alone every
//             ray crossing one of those extensions paid ~20 checks for a box it never came near.
//             Boxes are inflated outwards (1e-6 relative + absolute) so rounding of the distance evaluation stays inside the
//             margins of SafeDistance.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <numeric>
#include <tuple>
#include <vector>

namespace linon {

struct MeshTables {
    std::vector<float> tris;     // 12 floats per face
    std::vector<float> planes;   // 4 floats per group
    std::vector<int> ranges;     // per group: first face | face count << 16
    std::vector<int> segments;   // runs of consecutive groups with the same bound: first group | count << 12 | (box + 1) << 24
    std::vector<float> p1_planes;  // 4 floats per first-pass plane
    std::vector<float> boxes;      // 8 floats per first-pass box {lo.xyz, 0, hi.xyz, 0}
    float coord_max = 0.0f;        // max |plane offset| and |box coordinate| (rounding bound of SafeDistance)
    uint32_t n_groups() const { return (uint32_t)ranges.size(); }
    uint32_t n_p1() const { return (uint32_t)(p1_planes.size() / 4); }
    uint32_t n_boxes() const { return (uint32_t)(boxes.size() / 8); }
    uint32_t n_segments() const { return (uint32_t)segments.size(); }
};

constexpr uint32_t kMaxBoundBoxes = 24;

// Returns false when a face indexes past the vertex array (bad_face = its index).
inline bool build_mesh_tables(const float* verts, uint32_t n_verts, const uint32_t* faces, uint32_t n_faces, MeshTables& out,
                              uint32_t* bad_face = nullptr) {
    out = MeshTables();
    out.tris.assign((size_t)n_faces * 12, 0.0f);
    for (uint32_t i = 0; i < n_faces; ++i) {
        const uint32_t ia = faces[3 * i], ib = faces[3 * i + 1], ic = faces[3 * i + 2];
        if (ia >= n_verts || ib >= n_verts || ic >= n_verts) {
            if (bad_face) *bad_face = i;
            return false;
        }
        const float* a = verts + 3 * ia; const float* b = verts + 3 * ib; const float* c = verts + 3 * ic;
        float* t = out.tris.data() + 12 * i;
        for (int k = 0; k < 3; ++k) {
            t[k] = a[k];
            t[4 + k] = b[k] - a[k];
            t[8 + k] = c[k] - a[k];
        }
    }
    // ---- plane groups: runs of consecutive faces that lie in one plane.  A face joins the current group when its three
    // vertices are within 1e-6 (scene units are O(1) after normalisation) of the group's plane; degenerate faces get a
    // group of their own that is never culled (zero normal and offset: distance 0 <= reach, always tested).
    auto plane_of = [&](uint32_t i, double n[3], double& d0) {
        const float* t = out.tris.data() + 12 * i;
        const double e1[3] = {t[4], t[5], t[6]}, e2[3] = {t[8], t[9], t[10]};
        n[0] = e1[1] * e2[2] - e1[2] * e2[1];
        n[1] = e1[2] * e2[0] - e1[0] * e2[2];
        n[2] = e1[0] * e2[1] - e1[1] * e2[0];
        const double len = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        if (!(len > 1e-20)) return false;
        for (int k = 0; k < 3; ++k) n[k] /= len;
        d0 = n[0] * t[0] + n[1] * t[1] + n[2] * t[2];
        return true;
    };
    std::vector<uint32_t> group_of_face(n_faces, 0);
    std::vector<bool> degenerate;
    for (uint32_t i = 0; i < n_faces;) {
        double n[3], d0 = 0.0;
        const bool ok = plane_of(i, n, d0);
        uint32_t j = i + 1;
        if (ok) {
            for (; j < n_faces && j - i < 64; ++j) {
                const float* t = out.tris.data() + 12 * j;
                bool coplanar = true;
                for (int vtx = 0; vtx < 3 && coplanar; ++vtx) {
                    double p[3];
                    for (int k = 0; k < 3; ++k) p[k] = (double)t[k] + (vtx == 1 ? t[4 + k] : vtx == 2 ? t[8 + k] : 0.0f);
                    if (std::fabs(n[0] * p[0] + n[1] * p[1] + n[2] * p[2] - d0) > 1e-6) coplanar = false;
                }
                if (!coplanar) break;
            }
            out.planes.insert(out.planes.end(), {(float)n[0], (float)n[1], (float)n[2], (float)d0});
        } else {
            out.planes.insert(out.planes.end(), {0.0f, 0.0f, 0.0f, 0.0f});
        }
        degenerate.push_back(!ok);
        for (uint32_t f = i; f < j; ++f) group_of_face[f] = (uint32_t)out.ranges.size();
        out.ranges.push_back((int)i | (int)((j - i) << 16));
        i = j;
    }
    const uint32_t G = out.n_groups();
    for (uint32_t g = 0; g < G; ++g) out.coord_max = std::max(out.coord_max, std::fabs(out.planes[4 * g + 3]));

    // ---- big groups keep their plane; small groups are clustered by connectivity and bounded by one box per cluster ----
    struct Box { float lo[3], hi[3]; bool any = false; };
    std::vector<Box> gbox(G);
    Box scene;
    auto grow = [](Box& b, const float* p) {
        for (int k = 0; k < 3; ++k) {
            b.lo[k] = b.any ? std::min(b.lo[k], p[k]) : p[k];
            b.hi[k] = b.any ? std::max(b.hi[k], p[k]) : p[k];
        }
        b.any = true;
    };
    for (uint32_t f = 0; f < n_faces; ++f)
        for (int vtx = 0; vtx < 3; ++vtx) {
            const float* p = verts + 3 * faces[3 * f + vtx];
            grow(gbox[group_of_face[f]], p);
            grow(scene, p);
        }
    auto extent = [](const Box& b, int k) { return (double)b.hi[k] - (double)b.lo[k]; };
    // "big": the group spans at least 75 % of the scene along two axes (a wall, the floor, the ceiling).  Near such a group
    // its plane IS the geometry; a box around it would hold the whole room.
    std::vector<bool> as_planes(G, false);
    for (uint32_t g = 0; g < G; ++g) {
        int wide = 0;
        for (int k = 0; k < 3; ++k)
            if (extent(scene, k) > 0.0 && extent(gbox[g], k) >= 0.75 * extent(scene, k)) ++wide;
        as_planes[g] = degenerate[g] || wide >= 2;
    }
    // union-find over the vertex positions of the SMALL groups (the Cornell box repeats its corners per quad, so
    // connectivity is by position, quantised to 1e-5; the boxes stand on the floor, which is why big groups stay out)
    std::vector<uint32_t> parent(G);
    std::iota(parent.begin(), parent.end(), 0u);
    auto find = [&](uint32_t x) {
        while (parent[x] != x) x = parent[x] = parent[parent[x]];
        return x;
    };
    std::map<std::tuple<long long, long long, long long>, uint32_t> owner;
    for (uint32_t f = 0; f < n_faces; ++f) {
        const uint32_t g = group_of_face[f];
        if (as_planes[g]) continue;
        for (int vtx = 0; vtx < 3; ++vtx) {
            const float* p = verts + 3 * faces[3 * f + vtx];
            const auto key = std::make_tuple(std::llround((double)p[0] * 1e5), std::llround((double)p[1] * 1e5), std::llround((double)p[2] * 1e5));
            auto it = owner.find(key);
            if (it == owner.end()) owner.emplace(key, g);
            else parent[find(g)] = find(it->second);
        }
    }
    std::vector<Box> box(G);
    std::vector<bool> cluster_has_box_member(G, false);
    uint32_t n_boxes = 0;
    for (uint32_t g = 0; g < G; ++g) {
        if (as_planes[g]) continue;
        const uint32_t r = find(g);
        grow(box[r], gbox[g].lo);
        grow(box[r], gbox[g].hi);
        if (!cluster_has_box_member[r]) {
            cluster_has_box_member[r] = true;
            ++n_boxes;
        }
    }
    // too many clusters (a mesh of many small parts): plane distances only - the hierarchy path serves such meshes better
    if (n_boxes > kMaxBoundBoxes) {
        std::fill(as_planes.begin(), as_planes.end(), true);
        std::fill(cluster_has_box_member.begin(), cluster_has_box_member.end(), false);
    }
    for (uint32_t g = 0; g < G; ++g)
        if (as_planes[g]) out.p1_planes.insert(out.p1_planes.end(), out.planes.begin() + 4 * g, out.planes.begin() + 4 * g + 4);
    std::vector<int> box_index(G, -1);
    for (uint32_t r = 0; r < G; ++r) {
        if (!cluster_has_box_member[r]) continue;
        box_index[r] = (int)(out.boxes.size() / 8);
        const Box& b = box[r];
        float lo[3], hi[3];
        for (int k = 0; k < 3; ++k) {
            const float pad = 1e-6f * (1.0f + std::max(std::fabs(b.lo[k]), std::fabs(b.hi[k])));
            lo[k] = std::nextafter(b.lo[k] - pad, -INFINITY);
            hi[k] = std::nextafter(b.hi[k] + pad, INFINITY);
            out.coord_max = std::max(out.coord_max, std::max(std::fabs(lo[k]), std::fabs(hi[k])));
        }
        out.boxes.insert(out.boxes.end(), {lo[0], lo[1], lo[2], 0.0f, hi[0], hi[1], hi[2], 0.0f});
    }
    // segments: intersect_mesh walks the groups in index order (the reference's tie-break) and skips a whole run of groups
    // when the box of their cluster is out of reach
    for (uint32_t g = 0; g < G;) {
        const int b = as_planes[g] ? -1 : box_index[find(g)];
        uint32_t e = g + 1;
        while (e < G && e - g < 4095 && (as_planes[e] ? -1 : box_index[find(e)]) == b) ++e;
        out.segments.push_back((int)g | (int)((e - g) << 12) | ((b + 1) << 24));
        g = e;
    }
    return true;
}


// ---- bounding-volume hierarchy for meshes beyond the shared-memory stage (OBJ models, src/main.rs:139-157) --------
// Binary tree, median split of the face centroids along the longest axis, <= 4 faces per leaf, the two children of a
// node in adjacent slots.  Node = 2 x float4: {lo.xyz, int bits: left child (inner) | first face slot (leaf)},
// {hi.xyz, int bits: 0 (inner) | face count (leaf)}.  Boxes are inflated by a few ulps of the mesh extent and rounded
// outwards, so a point of a face as the kernel sees it ({v0, e1, e2} in fp32) is always inside its leaf box.
struct BvhBuild {
    struct Box { float lo[3], hi[3]; };
    std::vector<Box> face_box;
    std::vector<float> centroid;  // 3 per face
    std::vector<uint32_t> order;  // face slots in hierarchy order -> original face index
    std::vector<float> nodes;     // 8 floats per node
    uint32_t depth = 0;

    static void put_int(float* dst, int32_t x) { memcpy(dst, &x, 4); }

    void build(uint32_t node, uint32_t b, uint32_t e, uint32_t level) {
        depth = std::max(depth, level);
        Box box = face_box[order[b]];
        float clo[3], chi[3];
        for (int k = 0; k < 3; ++k) clo[k] = chi[k] = centroid[3 * order[b] + k];
        for (uint32_t i = b + 1; i < e; ++i) {
            const Box& f = face_box[order[i]];
            for (int k = 0; k < 3; ++k) {
                box.lo[k] = std::min(box.lo[k], f.lo[k]);
                box.hi[k] = std::max(box.hi[k], f.hi[k]);
                clo[k] = std::min(clo[k], centroid[3 * order[i] + k]);
                chi[k] = std::max(chi[k], centroid[3 * order[i] + k]);
            }
        }
        float* n = nodes.data() + 8 * (size_t)node;
        for (int k = 0; k < 3; ++k) { n[k] = box.lo[k]; n[4 + k] = box.hi[k]; }
        if (e - b <= 4) {
            put_int(n + 3, (int32_t)b);
            put_int(n + 7, (int32_t)(e - b));
            return;
        }
        int axis = 0;
        for (int k = 1; k < 3; ++k)
            if (chi[k] - clo[k] > chi[axis] - clo[axis]) axis = k;
        const uint32_t mid = b + (e - b) / 2;
        std::nth_element(order.begin() + b, order.begin() + mid, order.begin() + e, [&](uint32_t x, uint32_t y) {
            const float cx = centroid[3 * x + axis], cy = centroid[3 * y + axis];
            return cx < cy || (cx == cy && x < y);
        });
        const uint32_t left = (uint32_t)(nodes.size() / 8);
        nodes.resize(nodes.size() + 16);
        n = nodes.data() + 8 * (size_t)node;  // resize may have moved the storage
        put_int(n + 3, (int32_t)left);
        put_int(n + 7, 0);
        build(left, b, mid, level + 1);
        build(left + 1, mid, e, level + 1);
    }
};

struct BvhTables {
    std::vector<float> tris;   // 12 floats per face in hierarchy order; .w of the first float4 = original face index (int bits)
    std::vector<float> nodes;  // 8 floats per node
    uint32_t depth = 0;
    float extent = 0.0f;       // max |coordinate| (rounding bound of SafeDistance)
};
enum BvhStatus { BVH_OK = 0, BVH_BAD_FACE, BVH_NON_FINITE, BVH_TOO_DEEP };

inline BvhStatus build_bvh_tables(const float* verts, uint32_t n_verts, const uint32_t* faces, uint32_t n_faces, BvhTables& out,
                                  uint32_t* bad_face = nullptr) {
    std::vector<float> tris((size_t)n_faces * 12, 0.0f);
    BvhBuild bb;
    bb.face_box.resize(n_faces);
    bb.centroid.resize((size_t)3 * n_faces);
    bb.order.resize(n_faces);
    float extent = 0.0f;
    for (uint32_t i = 0; i < n_faces; ++i) {
        const uint32_t ia = faces[3 * i], ib = faces[3 * i + 1], ic = faces[3 * i + 2];
        if (ia >= n_verts || ib >= n_verts || ic >= n_verts) {
            if (bad_face) *bad_face = i;
            return BVH_BAD_FACE;
        }
        const float* a = verts + 3 * ia; const float* b = verts + 3 * ib; const float* c = verts + 3 * ic;
        float* t = tris.data() + 12 * (size_t)i;
        for (int k = 0; k < 3; ++k) {  // {v0, e1, e2}: the two subtractions hit_triangle performs (src/main_view.wgsl:152-153)
            t[k] = a[k];
            t[4 + k] = b[k] - a[k];
            t[8 + k] = c[k] - a[k];
            // the triangle the kernel tests is (v0, v0 + e1, v0 + e2); bound it together with the input vertices
            const double p1 = (double)t[k] + (double)t[4 + k], p2 = (double)t[k] + (double)t[8 + k];
            const double lo = std::min({(double)a[k], (double)b[k], (double)c[k], p1, p2});
            const double hi = std::max({(double)a[k], (double)b[k], (double)c[k], p1, p2});
            bb.face_box[i].lo[k] = (float)lo;
            bb.face_box[i].hi[k] = (float)hi;
            bb.centroid[3 * (size_t)i + k] = (float)((lo + hi) * 0.5);
            extent = std::max({extent, std::fabs((float)lo), std::fabs((float)hi)});
        }
        bb.order[i] = i;
    }
    if (!std::isfinite(extent)) return BVH_NON_FINITE;
    const float pad = 4e-7f * (1.0f + extent);  // a few ulps of the largest coordinate
    for (auto& f : bb.face_box)
        for (int k = 0; k < 3; ++k) {
            f.lo[k] = std::nextafter(f.lo[k] - pad, -INFINITY);
            f.hi[k] = std::nextafter(f.hi[k] + pad, INFINITY);
        }
    bb.nodes.assign(8, 0.0f);
    bb.build(0, 0, n_faces, 1);
    if (bb.depth > 38) return BVH_TOO_DEEP;
    // faces in hierarchy order; .w of the first float4 carries the original face index (the hit id and the tie-break)
    std::vector<float> sorted((size_t)n_faces * 12);
    for (uint32_t slot = 0; slot < n_faces; ++slot) {
        memcpy(sorted.data() + 12 * (size_t)slot, tris.data() + 12 * (size_t)bb.order[slot], 48);
        BvhBuild::put_int(sorted.data() + 12 * (size_t)slot + 3, (int32_t)bb.order[slot]);
    }
    out.tris = std::move(sorted);
    out.nodes = std::move(bb.nodes);
    out.depth = bb.depth;
    out.extent = extent;
    return BVH_OK;
}

}  // namespace linon
