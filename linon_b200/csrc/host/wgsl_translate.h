// wgsl_translate.h — WGSL field-function body -> CUDA C++ (see wgsl_translate.cpp).
#pragma once
#include <string>

namespace linon {

// Translates the body text that the reference splices into `fn field_function(...)`
// (src/main_view.rs:852-861).  Returns false and fills `error` when the body uses WGSL outside the
// supported subset; line structure is preserved so compiler diagnostics map to the user's lines.
bool translate_wgsl_body(const std::string& body, std::string& cuda, std::string& error);

}  // namespace linon
