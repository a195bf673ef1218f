// functions.hpp — the predefined field functions (host mirror of src/functions.rs).
//
// The bodies are DATA: WGSL text the reference splices into its shader (src/main_view.rs:852-861).  They are
// kept byte-for-byte (including the non-ASCII degree signs in comments and the trailing commas of the attractor
// constructors) because the translator must accept exactly what the reference's GUI hands over;
// tests/test_field_functions.py compares them with /root/reference/src/functions.rs whenever that mount exists.
#pragma once
#include <string>

namespace linon {

enum class PredefinedFunction : int {  // enum order of src/functions.rs:2-13
    Custom = 0,
    MirageSpherical = 1,
    MiragePlane = 2,
    MirageSphericalSigmoid = 3,
    MiragePlaneSigmoid = 4,
    TranslationX = 5,
    TranslationZ = 6,
    Rotation = 7,
    LorenzAttractor = 8,
    RoesslerAttractor = 9,
};
constexpr int kNumPredefinedFunctions = 10;

inline const char* to_string(PredefinedFunction f) {  // src/functions.rs:15-30
    switch (f) {
        case PredefinedFunction::Custom: return "Custom";
        case PredefinedFunction::MirageSpherical: return "Mirage (spherical)";
        case PredefinedFunction::MiragePlane: return "Mirage (plane)";
        case PredefinedFunction::MirageSphericalSigmoid: return "Mirage (spherical sigmoid)";
        case PredefinedFunction::MiragePlaneSigmoid: return "Mirage (plane sigmoid)";
        case PredefinedFunction::TranslationX: return "Translation (x-axis)";
        case PredefinedFunction::TranslationZ: return "Translation (z-axis)";
        case PredefinedFunction::Rotation: return "Rotation";
        case PredefinedFunction::LorenzAttractor: return "Lorenz attractor";
        case PredefinedFunction::RoesslerAttractor: return "Roessler attractor";
    }
    return "";
}

namespace detail {
// shared text blocks of the four mirage bodies (src/functions.rs:36-94)
inline std::string mirage_head(const char* t_src, const char* max_dist) {
    return std::string("let t_env = 15.0; // \xC2\xB0"
                       "C\nlet t_src = ") + t_src + "; // \xC2\xB0" "C\nlet max_dist = " + max_dist + ";\n";
}
inline std::string mirage_tail(const char* arg_in, const char* arg_out, const char* normal) {
    return std::string("let part_in = clamp(0.0, 1.0, ") + arg_in + ");\n"
           "let part_out = clamp(0.0, 1.0, " + arg_out + ");\n"
           "let t_in = part_in * t_env + (1.0 - part_in) * t_src;\n"
           "let t_out = part_out * t_env + (1.0 - part_out) * t_src;\n"
           "\n"
           "return refraction(t_in, t_out, v, " + normal + ");";
}
inline std::string spherical(bool sig) {
    return mirage_head("200.0", "0.25") +
           "let center = vec3<f32>(-0.5, 0.5, -0.5);\n"
           "\n"
           "let center_dest = p - center;\n"
           "let normal = normalize(center_dest);\n"
           "let dist_in = length(p_prev - center);\n"
           "let dist_out = length(center_dest);\n" +
           (sig ? mirage_tail("sigmoid(dist_in / max_dist * 12.0 - 6.0)", "sigmoid(dist_out / max_dist * 12.0 - 6.0)", "normal")
                : mirage_tail("dist_in / max_dist", "dist_out / max_dist", "normal"));
}
inline std::string plane(bool sig) {
    return mirage_head("30.0", "0.01") +
           "let plane_p0 = vec3<f32>(0.0, 0.1, 0.0);\n"
           "let plane_n = vec3<f32>(0.0, 1.0, 0.0);\n"
           "\n"
           "let dist_in = point_plane_distance(p_prev, plane_n, plane_p0);\n"
           "let dist_out = point_plane_distance(p, plane_n, plane_p0);\n" +
           (sig ? mirage_tail("sigmoid(dist_in / max_dist * 12.0 - 6.0)", "sigmoid(dist_out / max_dist * 12.0 - 6.0)", "plane_n")
                : mirage_tail("dist_in / max_dist", "dist_out / max_dist", "plane_n"));
}
}  // namespace detail

// PredefinedFunction::to_code, src/functions.rs:33-128
inline std::string to_code(PredefinedFunction f) {
    switch (f) {
        case PredefinedFunction::Custom: return "";
        case PredefinedFunction::MirageSpherical: return detail::spherical(false);
        case PredefinedFunction::MiragePlane: return detail::plane(false);
        case PredefinedFunction::MirageSphericalSigmoid: return detail::spherical(true);
        case PredefinedFunction::MiragePlaneSigmoid: return detail::plane(true);
        case PredefinedFunction::TranslationX:
            return "let dx = 0.5 * t;\nlet dy = 0.0;\nlet dz = 0.0;\nreturn translate(v0, dx, dy, dz);";
        case PredefinedFunction::TranslationZ:
            return "let dx = 0.0;\nlet dy = 0.0;\nlet dz = 0.5 * t;\nreturn translate(v0, dx, dy, dz);";
        case PredefinedFunction::Rotation: return "return normalize(v0 + rotateZ(v0, PI * t));";
        case PredefinedFunction::LorenzAttractor:
            return "let rho = 28.0;\nlet sigma = 10.0;\nlet beta = 8.0 / 3.0;\nreturn vec3<f32>(\n"
                   "    sigma * (p.y - p.x),\n    p.x * (rho - p.z) - p.y,\n    p.x * p.y - beta * p.z,\n);";
        case PredefinedFunction::RoesslerAttractor:
            return "let a = 0.1;\nlet b = 0.1;\nlet c = 14.0;\nreturn vec3<f32>(\n"
                   "    -p.y - p.z,\n    p.x + a * p.y,\n    b + p.z * (p.x - c),\n);";
    }
    return "";
}

}  // namespace linon
