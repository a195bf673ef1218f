// wgsl_translate.h — WGSL field-function body -> CUDA C++ (see wgsl_translate.cpp).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace linon {

// Translates the body text that the reference splices into `fn field_function(...)`
// (src/main_view.rs:852-861).  Returns false and fills `error` when the body uses WGSL outside the
// supported subset; line structure is preserved so compiler diagnostics map to the user's lines.
bool translate_wgsl_body(const std::string& body, std::string& cuda, std::string& error);
// The same with the names an installed helper layer provides (ModuleLayer::names, comma-separated): the body may call them.
// Every other identifier must be a WGSL keyword / type / builtin, a helper of main_view.wgsl, a parameter of field_function,
// PI / settings / camera, or declared in the body - nothing else reaches the CUDA compiler (see pass_sanitize).
bool translate_wgsl_body(const std::string& body, const std::string& module_names, std::string& cuda, std::string& error);

// ---- whole-shader hot reload (src/main.rs:206-222 -> Application::reload_compute_shader, src/application.rs:228-232
//      -> MainView::reload_shader(device, Some(new_src), field_function), src/main_view.rs:649-690) ------------------
// One top-level declaration of a WGSL module.
struct ModuleItem {
    std::string kind;  // "fn" "const" "struct" "var" "other"
    std::string name;
    uint64_t hash;     // FNV-1a over the item's tokens (comments and white space do not count)
    int line;          // 1-based line of the item's first token
    uint64_t masked_hash = 0;      // the same with every numeric literal hashed as one placeholder: equal for two items that
                                   // differ in the VALUES of their literals only
    std::vector<double> literals;  // the numeric literals in order of appearance (magnitudes: a sign is its own token)
};
bool scan_wgsl_module(const std::string& src, std::vector<ModuleItem>& items, std::string& error);

// The part of an edited main_view.wgsl that the CUDA path can honour: the helper layer the field function calls.
struct ModuleLayer {
    std::string members;    // CUDA member functions of LinonField<S> (new helpers + overridden library helpers)
    std::string preamble;   // module-scope constants, re-declared at the top of every emitted function and of the body
    std::string names;      // comma-separated names emitted (introspection / tests)
    bool empty() const { return members.empty() && preamble.empty(); }  // no helper layer (march edits do not count)
    // Edits to the march-level constants of the module (src/main_view.wgsl:65-76 and the literals of nonlinear_ray_color /
    // sample_rays): `#define LINON_...` lines for the kernel template (csrc/device/march_kernels.cuh), plus the two
    // module-scope switches the reference itself toggles by string replacement (src/main_view.rs:752-829).
    std::string march_defines;
    int adaptive_sampling = -1;    // -1 = as shipped (true); 0 / 1 = `const adaptive_sampling` of the edited module
    int sample_outline_rays = -1;  // -1 = as shipped (false)
    double h_initial = -1.0;       // > 0: `let h_initial` of the edited nonlinear_ray_color (then Enhance's replacement no
                                   // longer matches the text, like in the reference, and the flag has no effect)
    double t_max = -1.0;           // > 0: the edited horizon of `for (; t <= 5.0;)`
    double reject_min_h = -1.0;    // >= 0: the edited threshold of `h > 0.002`
    bool linear_mode = false;      // `const linear_mode: bool = true`: one straight ray per pixel, no march
};
// Returns false with `error` set when the module cannot be honoured; `unsupported` tells a change to the kernel's own
// functions / bindings / constants (named in the message) from a syntax error.
bool translate_wgsl_module(const std::string& src, ModuleLayer& layer, std::string& error, bool& unsupported);

}  // namespace linon
