// scene.hpp — host-side scene inputs of the march path (host mirror, C++).
//   cornell_box()          src/cornell_box.rs:2-163   the Cornell box as 18 quads -> 72 vertices / 36 faces
//   normalize_vertices()   src/vertices.rs:3-17       one global min/max over all coordinates
//   get_center()           src/vertices.rs:19-57      bounding-box centre
//   default_scene()        src/application.rs:105-123 negate x and z, normalise, centre
// fp32 throughout (build with -ffp-contract=off so every operation rounds like the Rust original).
#pragma once
#include <array>
#include <cstdint>
#include <vector>

namespace linon {

struct Mesh {
    std::vector<float> vertices;   // 3 per vertex
    std::vector<uint32_t> indices; // 3 per face
};

// Each quad (a, b, c, d) yields the faces (a, b, c) and (c, d, a): the pattern of INDICES in the reference.
inline Mesh cornell_box() {
    static const float quads[18][4][3] = {
        // floor, short-block footprint, tall-block footprint
        {{552.8f, 0, 0}, {0, 0, 0}, {0, 0, 559.2f}, {549.6f, 0, 559.2f}},
        {{130, 0, 65}, {82, 0, 225}, {240, 0, 272}, {290, 0, 114}},
        {{423, 0, 247}, {265, 0, 296}, {314, 0, 456}, {472, 0, 406}},
        // ceiling, light
        {{556, 548.8f, 0}, {556, 548.8f, 559.2f}, {0, 548.8f, 559.2f}, {0, 548.8f, 0}},
        {{343, 548.8f, 227}, {343, 548.8f, 332}, {213, 548.8f, 332}, {213, 548.8f, 227}},
        // back wall, right wall, left wall
        {{549.6f, 0, 559.2f}, {0, 0, 559.2f}, {0, 548.8f, 559.2f}, {556, 548.8f, 559.2f}},
        {{0, 0, 559.2f}, {0, 0, 0}, {0, 548.8f, 0}, {0, 548.8f, 559.2f}},
        {{552.8f, 0, 0}, {549.6f, 0, 559.2f}, {556, 548.8f, 559.2f}, {556, 548.8f, 0}},
        // short block: top, four sides
        {{130, 165, 65}, {82, 165, 225}, {240, 165, 272}, {290, 165, 114}},
        {{290, 0, 114}, {290, 165, 114}, {240, 165, 272}, {240, 0, 272}},
        {{130, 0, 65}, {130, 165, 65}, {290, 165, 114}, {290, 0, 114}},
        {{82, 0, 225}, {82, 165, 225}, {130, 165, 65}, {130, 0, 65}},
        {{240, 0, 272}, {240, 165, 272}, {82, 165, 225}, {82, 0, 225}},
        // tall block: top, four sides
        {{423, 330, 247}, {265, 330, 296}, {314, 330, 456}, {472, 330, 406}},
        {{423, 0, 247}, {423, 330, 247}, {472, 330, 406}, {472, 0, 406}},
        {{472, 0, 406}, {472, 330, 406}, {314, 330, 456}, {314, 0, 456}},
        {{314, 0, 456}, {314, 330, 456}, {265, 330, 296}, {265, 0, 296}},
        {{265, 0, 296}, {265, 330, 296}, {423, 330, 247}, {423, 0, 247}},
    };
    Mesh m;
    for (uint32_t q = 0; q < 18; ++q) {
        for (int k = 0; k < 4; ++k)
            for (int c = 0; c < 3; ++c) m.vertices.push_back(quads[q][k][c]);
        const uint32_t a = 4 * q;
        const uint32_t f[6] = {a, a + 1, a + 2, a + 2, a + 3, a};
        m.indices.insert(m.indices.end(), f, f + 6);
    }
    return m;
}

inline void normalize_vertices(std::vector<float>& v) {
    float mx = -3.40282347e+38f, mn = 3.40282347e+38f;
    for (float x : v) {
        if (x > mx) mx = x;
        if (x < mn) mn = x;
    }
    for (float& x : v) x = (x - mn) / (mx - mn) * 2.0f - 1.0f;
}

inline std::array<float, 3> get_center(const std::vector<float>& v) {
    float lo[3] = {v[0], v[1], v[2]}, hi[3] = {v[0], v[1], v[2]};
    for (size_t i = 1; i < v.size() / 3; ++i)
        for (int c = 0; c < 3; ++c) {
            const float x = v[3 * i + c];
            if (x < lo[c]) lo[c] = x;
            if (x > hi[c]) hi[c] = x;
        }
    return {(lo[0] + hi[0]) / 2.0f, (lo[1] + hi[1]) / 2.0f, (lo[2] + hi[2]) / 2.0f};
}

// The scene Application::new uploads at start-up.  `negate_xz = false` gives load_default_model (the `r` key,
// src/application.rs:194-197), which skips the axis inversion.
inline Mesh default_scene(std::array<float, 3>* center, bool negate_xz = true) {
    Mesh m = cornell_box();
    if (negate_xz)
        for (size_t i = 0; i < m.vertices.size() / 3; ++i) {
            m.vertices[3 * i] = -m.vertices[3 * i];
            m.vertices[3 * i + 2] = -m.vertices[3 * i + 2];
        }
    normalize_vertices(m.vertices);
    if (center) *center = get_center(m.vertices);
    return m;
}

}  // namespace linon
