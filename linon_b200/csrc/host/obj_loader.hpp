// obj_loader.hpp — the OBJ drag-and-drop path of the reference (src/main.rs:139-157), host mirror.
//
// The reference calls tobj::load_obj(path, LoadOptions { triangulate: true, ..Default::default() }) (tobj 4.0.2,
// Cargo.lock) and hands models[0].mesh.positions / .indices to Application::load_model (src/application.rs:199-226),
// which normalises the positions with one global min/max (src/vertices.rs:3-17, NO x/z negation on this path) and
// recentres the camera.  tobj is not vendored under the reference mount; what is restated here is its documented
// behaviour for that option set:
//   * `v x y z [w | r g b]`, faces `f i`, `i/j`, `i//k`, `i/j/k`, negative (relative) indices, comments, `\` line joins are not used
//   * a face with n > 3 corners becomes the fan (0, i, i+1), i = 1..n-2; faces with fewer than 3 corners are dropped
//   * `o` / `g` (and `usemtl` after faces) close the current model; only the FIRST model that has faces is returned
//   * default `single_index: false`: positions are the position indices the model's faces use, de-duplicated in order
//     of first use; `indices` are re-numbered accordingly (texture / normal indices are parsed and ignored)
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "scene.hpp"

namespace linon {

inline bool load_obj_text(const std::string& text, Mesh& out, std::string& error) {
    std::vector<float> pos;  // every `v` of the file
    std::vector<uint32_t> model_faces;  // position indices (0-based, into pos), 3 per triangle, of the current model
    bool have_model = false;
    auto close_model = [&]() {
        if (!model_faces.empty()) have_model = true;
    };
    size_t i = 0;
    int line_no = 0;
    while (i < text.size() && !have_model) {
        size_t e = text.find('\n', i);
        if (e == std::string::npos) e = text.size();
        std::string line = text.substr(i, e - i);
        i = e + 1;
        ++line_no;
        const size_t hash = line.find('#');
        if (hash != std::string::npos) line.resize(hash);
        while (!line.empty() && (line.back() == '\r' || line.back() == ' ' || line.back() == '\t')) line.pop_back();
        size_t s = line.find_first_not_of(" \t");
        if (s == std::string::npos) continue;
        const size_t kw_end = line.find_first_of(" \t", s);
        const std::string kw = line.substr(s, kw_end == std::string::npos ? std::string::npos : kw_end - s);
        const char* rest = kw_end == std::string::npos ? "" : line.c_str() + kw_end;
        if (kw == "v") {
            char* end = nullptr;
            float xyz[3];
            for (int k = 0; k < 3; ++k) {
                xyz[k] = std::strtof(rest, &end);
                if (end == rest) {
                    error = "line " + std::to_string(line_no) + ": `v` needs three coordinates";
                    return false;
                }
                rest = end;
            }
            pos.insert(pos.end(), xyz, xyz + 3);
        } else if (kw == "f") {
            std::vector<uint32_t> corners;
            const char* p = rest;
            for (;;) {
                while (*p == ' ' || *p == '\t') ++p;
                if (!*p) break;
                char* end = nullptr;
                const long idx = std::strtol(p, &end, 10);
                if (end == p) {
                    error = "line " + std::to_string(line_no) + ": malformed face corner";
                    return false;
                }
                const long n_pos = (long)(pos.size() / 3);
                const long zero_based = idx > 0 ? idx - 1 : n_pos + idx;  // negative indices count back from the last vertex
                if (idx == 0 || zero_based < 0 || zero_based >= n_pos) {
                    error = "line " + std::to_string(line_no) + ": face references vertex " + std::to_string(idx) + " of " + std::to_string(n_pos);
                    return false;
                }
                corners.push_back((uint32_t)zero_based);
                p = end;
                while (*p && *p != ' ' && *p != '\t') ++p;  // skip /vt/vn
            }
            for (size_t c = 1; c + 1 < corners.size(); ++c) {
                model_faces.push_back(corners[0]);
                model_faces.push_back(corners[c]);
                model_faces.push_back(corners[c + 1]);
            }
        } else if (kw == "o" || kw == "g" || kw == "usemtl") {
            close_model();
        }
    }
    if (model_faces.empty()) {
        error = "no faces in the OBJ text";
        return false;
    }
    out.vertices.clear();
    out.indices.clear();
    std::unordered_map<uint32_t, uint32_t> remap;
    for (uint32_t v : model_faces) {
        auto it = remap.find(v);
        if (it == remap.end()) {
            it = remap.emplace(v, (uint32_t)remap.size()).first;
            out.vertices.insert(out.vertices.end(), pos.begin() + 3 * (size_t)v, pos.begin() + 3 * (size_t)v + 3);
        }
        out.indices.push_back(it->second);
    }
    return true;
}

// Application::load_model, src/application.rs:199-226: normalise in place, return the bounding-box centre.
inline std::array<float, 3> prepare_loaded_model(Mesh& m) {
    normalize_vertices(m.vertices);
    return get_center(m.vertices);
}

}  // namespace linon
