// wgsl_translate.cpp — turns the BODY of a linon field function (WGSL) into CUDA C++ that compiles
// against device/wgsl_prelude.cuh.
//
// Reference contract: src/main_view.rs:852-861 splices the body text into
//   fn field_function(p_prev: vec3<f32>, p: vec3<f32>, v0: vec3<f32>, v: vec3<f32>, t: f32) -> vec3<f32> { <body> }
// The nine predefined bodies are src/functions.rs:36-126; the subset that must work is SURVEY Appendix A.
//
// The rewrite is token based and line preserving (every output line is the same source line, so NVRTC
// diagnostics under `#line 1 "field_function.wgsl"` cite the user's own line numbers):
//   * comments are dropped (they hold non-ASCII text in the predefined bodies), newlines kept
//   * float literals get an `f` suffix (WGSL abstract floats concretise to f32; without it the whole
//     expression would silently become fp64), `1f` -> `1.0f`, i32 suffix dropped, u32 suffix kept
//   * a `let`/`const` initialiser made only of literals and + - * / ( ) is emitted as (float)(double expr):
//     naga folds such expressions in abstract float (f64) and then concretises (e.g. `8.0 / 3.0`)
//   * let/const/var declarations -> const auto / auto / typed C++ declarations; `var x: T;` zero-initialises
//   * vecN<f32> / vecNf / vecN( -> vecNf, matCxR<f32> -> matCxRf, f32 -> wg_f32, i32/u32 -> int/unsigned
//     (vecNf, matCxRf and wg_f32 are re-declared per instantiation by LinonField<S>: float or two rays at once)
//   * WGSL builtins called as functions -> w_<name>
//   * multi-component swizzles .xyz/.rgb -> .xyz(); .r/.g/.b/.a -> .x/.y/.z/.w
//   * `if cond {` / `while cond {` get their parentheses; a trailing comma before `)` is removed
//   * identifiers that are C++ keywords get a trailing underscore
//   * literal-only sub-expressions, including calls of const-evaluable builtins (`tan(0.1)`, `x + 1.0 / 3.0`), are
//     evaluated in f64 on the host like naga's constant evaluator and emitted as one f32 literal (pass_fold)
// Anything outside the subset is passed through and reported by NVRTC against the user's line.

#include "wgsl_translate.h"

#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>
#include <string>
#include <vector>

namespace linon {
namespace {

enum Kind { WS, NL, IDENT, NUMBER, PUNCT };

struct Tok {
    Kind kind;
    std::string text;
};

const char* const kBuiltins[] = {
    "abs", "min", "max", "clamp", "saturate", "sqrt", "inverseSqrt", "exp", "exp2", "log", "log2", "pow", "sin", "cos",
    "tan", "asin", "acos", "atan", "atan2", "sinh", "cosh", "tanh", "floor", "ceil", "round", "trunc", "fract", "sign",
    "step", "mix", "smoothstep", "fma", "degrees", "radians", "select", "dot", "cross", "length", "distance",
    "normalize", "reflect", "faceForward", "transpose", "determinant"};

const char* const kCppKeywords[] = {
    "auto", "char", "class", "delete", "double", "extern", "float", "friend", "goto", "inline", "int", "long",
    "mutable", "namespace", "new", "operator", "register", "short", "signed", "sizeof", "static", "struct",
    "template", "this", "typedef", "union", "unsigned", "virtual", "void", "volatile", "asm", "and", "or", "not",
    "xor", "catch", "try", "throw", "explicit", "export", "public", "private", "protected", "typename", "using",
    "wchar_t", "nullptr", "constexpr", "decltype", "noexcept", "alignas", "alignof", "concept", "requires"};

bool is_ident_start(char c) { return std::isalpha((unsigned char)c) || c == '_'; }
bool is_ident_char(char c) { return std::isalnum((unsigned char)c) || c == '_'; }

struct Lexer {
    const std::string& s;
    size_t i = 0;
    std::vector<Tok> out;
    std::string err;
    explicit Lexer(const std::string& src) : s(src) {}

    bool number() {
        size_t j = i;
        bool is_float = false, is_hex = false;
        if (s[j] == '0' && j + 1 < s.size() && (s[j + 1] == 'x' || s[j + 1] == 'X')) {
            is_hex = true;
            j += 2;
            while (j < s.size() && std::isxdigit((unsigned char)s[j])) ++j;
        } else {
            while (j < s.size() && std::isdigit((unsigned char)s[j])) ++j;
            if (j < s.size() && s[j] == '.' && !(j + 1 < s.size() && is_ident_start(s[j + 1]) && s[j + 1] != 'e' &&
                                                 s[j + 1] != 'E' && s[j + 1] != 'f' && s[j + 1] != 'h')) {
                is_float = true;
                ++j;
                while (j < s.size() && std::isdigit((unsigned char)s[j])) ++j;
            }
            if (j < s.size() && (s[j] == 'e' || s[j] == 'E')) {
                size_t k = j + 1;
                if (k < s.size() && (s[k] == '+' || s[k] == '-')) ++k;
                if (k < s.size() && std::isdigit((unsigned char)s[k])) {
                    is_float = true;
                    j = k;
                    while (j < s.size() && std::isdigit((unsigned char)s[j])) ++j;
                }
            }
        }
        std::string body = s.substr(i, j - i);
        char suffix = 0;
        if (j < s.size() && (s[j] == 'f' || s[j] == 'h' || s[j] == 'i' || s[j] == 'u') &&
            !(j + 1 < s.size() && is_ident_char(s[j + 1]))) {
            suffix = s[j];
            ++j;
        }
        if (suffix == 'h') {
            err = "f16 literals are not supported";
            return false;
        }
        std::string text;
        if (is_hex) {
            text = body + (suffix == 'u' ? "u" : "");
        } else if (is_float || suffix == 'f') {
            if (body.find('.') == std::string::npos && body.find('e') == std::string::npos &&
                body.find('E') == std::string::npos)
                body += ".0";
            if (body[0] == '.') body = "0" + body;
            text = body + "f";
        } else {
            text = body + (suffix == 'u' ? "u" : "");
        }
        out.push_back({NUMBER, text});
        i = j;
        return true;
    }

    bool run() {
        static const char* const multi[] = {"<<=", ">>=", "->", "==", "!=", "<=", ">=", "&&", "||", "+=", "-=", "*=",
                                            "/=",  "%=",  "&=", "|=", "^=", "++", "--", "<<", ">>"};
        while (i < s.size()) {
            const char c = s[i];
            if (c == '\n') {
                out.push_back({NL, "\n"});
                ++i;
            } else if (c == '\r') {
                ++i;
            } else if (c == ' ' || c == '\t') {
                size_t j = i;
                while (j < s.size() && (s[j] == ' ' || s[j] == '\t')) ++j;
                out.push_back({WS, s.substr(i, j - i)});
                i = j;
            } else if (c == '/' && i + 1 < s.size() && s[i + 1] == '/') {
                while (i < s.size() && s[i] != '\n') ++i;
            } else if (c == '/' && i + 1 < s.size() && s[i + 1] == '*') {
                int depth = 1;  // WGSL block comments nest
                i += 2;
                while (i < s.size() && depth > 0) {
                    if (s[i] == '/' && i + 1 < s.size() && s[i + 1] == '*') {
                        ++depth;
                        i += 2;
                    } else if (s[i] == '*' && i + 1 < s.size() && s[i + 1] == '/') {
                        --depth;
                        i += 2;
                    } else {
                        if (s[i] == '\n') out.push_back({NL, "\n"});
                        ++i;
                    }
                }
                if (depth > 0) {
                    err = "unterminated block comment";
                    return false;
                }
                out.push_back({WS, " "});
            } else if (is_ident_start(c)) {
                size_t j = i;
                while (j < s.size() && is_ident_char(s[j])) ++j;
                out.push_back({IDENT, s.substr(i, j - i)});
                i = j;
            } else if (std::isdigit((unsigned char)c) ||
                       (c == '.' && i + 1 < s.size() && std::isdigit((unsigned char)s[i + 1]))) {
                if (!number()) return false;
            } else if ((unsigned char)c >= 0x80) {
                err = "non-ASCII character outside a comment";
                return false;
            } else {
                bool matched = false;
                for (const char* m : multi) {
                    const size_t n = std::strlen(m);
                    if (s.compare(i, n, m) == 0) {
                        out.push_back({PUNCT, m});
                        i += n;
                        matched = true;
                        break;
                    }
                }
                if (!matched) {
                    out.push_back({PUNCT, std::string(1, c)});
                    ++i;
                }
            }
        }
        return true;
    }
};

struct Rewriter {
    std::vector<Tok> t;
    std::string err;
    std::set<std::string> builtins, keywords;

    Rewriter() {
        for (const char* b : kBuiltins) builtins.insert(b);
        for (const char* k : kCppKeywords) keywords.insert(k);
    }

    // index of the next / previous token that is not whitespace or a newline (-1 if none)
    int next_sig(int i) const {
        for (int j = i + 1; j < (int)t.size(); ++j)
            if (t[j].kind != WS && t[j].kind != NL) return j;
        return -1;
    }
    int prev_sig(int i) const {
        for (int j = i - 1; j >= 0; --j)
            if (t[j].kind != WS && t[j].kind != NL) return j;
        return -1;
    }
    bool is_p(int i, const char* p) const { return i >= 0 && t[i].kind == PUNCT && t[i].text == p; }
    void blank(int i) { t[i] = {WS, ""}; }

    // Parses a type starting at token i; returns the C++ spelling and sets `end` to the last token used.
    bool parse_type(int i, std::string& cpp, int& end) {
        if (i < 0 || t[i].kind != IDENT) return false;
        const std::string& n = t[i].text;
        end = i;
        auto scalar = [](const std::string& s, std::string& o) {
            if (s == "f32") o = "wg_f32";  // float, or the two-ray scalar of the instantiation (LinonField<S>)
            else if (s == "i32") o = "int";
            else if (s == "u32") o = "unsigned";
            else if (s == "bool") o = "bool";
            else return false;
            return true;
        };
        if (scalar(n, cpp)) return true;
        const bool is_vec = n == "vec2" || n == "vec3" || n == "vec4";
        const bool is_mat = n == "mat2x2" || n == "mat3x3";
        if (is_vec || is_mat) {
            const int a = next_sig(i);
            if (is_p(a, "<")) {
                const int b = next_sig(a), c = b >= 0 ? next_sig(b) : -1;
                if (b < 0 || t[b].kind != IDENT || !is_p(c, ">")) return false;
                if (t[b].text != "f32") {
                    err = n + "<" + t[b].text + "> is not supported (only f32 vectors and matrices)";
                    return false;
                }
                end = c;
            }
            cpp = n + "f";
            return true;
        }
        if (n == "vec2f" || n == "vec3f" || n == "vec4f" || n == "mat2x2f" || n == "mat3x3f") {
            cpp = n;
            return true;
        }
        return false;
    }

    // ---- pass 1: types (vec3<f32> -> vec3f, f32( -> float( ...) --------------------------------------
    bool pass_types() {
        for (int i = 0; i < (int)t.size(); ++i) {
            if (t[i].kind != IDENT) continue;
            const int p = prev_sig(i);
            if (is_p(p, ".")) continue;  // member name
            std::string cpp;
            int end = i;
            const std::string before = err;
            if (parse_type(i, cpp, end)) {
                const std::string& n = t[i].text;
                const bool templ = end != i;
                const bool bare_vec = (n == "vec2" || n == "vec3" || n == "vec4" || n == "mat2x2" || n == "mat3x3");
                // a bare `vec3` is only a type when used as a constructor: vec3(...)
                if (bare_vec && !templ && !is_p(next_sig(i), "(")) continue;
                // f32/i32/... only as type position or cast; identifiers named e.g. `bool` cannot occur in WGSL
                t[i].text = cpp;
                for (int j = i + 1; j <= end; ++j) blank(j);
            } else if (err != before) {
                return false;
            } else if (t[i].text == "array" || t[i].text == "ptr" || t[i].text == "atomic" ||
                       t[i].text == "mat4x4" || t[i].text == "mat4x4f" || t[i].text == "bitcast") {
                err = "'" + t[i].text + "' is not supported in a field function body";
                return false;
            }
        }
        return true;
    }

    bool literal_only(int a, int b) const {  // tokens (a, b) exclusive, at least one operator and one number
        bool has_op = false, has_num = false;
        for (int j = a + 1; j < b; ++j) {
            if (t[j].kind == WS || t[j].kind == NL) continue;
            if (t[j].kind == NUMBER) {
                if (t[j].text.back() != 'f') return false;  // only float literals
                has_num = true;
            } else if (t[j].kind == PUNCT && (t[j].text == "+" || t[j].text == "-" || t[j].text == "*" ||
                                              t[j].text == "/" || t[j].text == "(" || t[j].text == ")")) {
                if (t[j].text != "(" && t[j].text != ")") has_op = true;
            } else {
                return false;
            }
        }
        return has_op && has_num;
    }

    // ---- pass 2: declarations ------------------------------------------------------------------------------
    bool pass_decls() {
        for (int i = 0; i < (int)t.size(); ++i) {
            if (t[i].kind != IDENT) continue;
            const std::string kw = t[i].text;
            if (kw != "let" && kw != "var" && kw != "const") continue;
            if (is_p(prev_sig(i), ".")) continue;
            int n = next_sig(i);
            if (kw == "var" && is_p(n, "<")) {  // var<function> / var<private>
                int j = n;
                while (j >= 0 && !is_p(j, ">")) {
                    blank(j);
                    j = next_sig(j);
                }
                if (j < 0) {
                    err = "malformed var<...> declaration";
                    return false;
                }
                blank(j);
                n = next_sig(j);
            }
            if (n < 0 || t[n].kind != IDENT) {
                err = "expected a name after '" + kw + "'";
                return false;
            }
            const int name = n;
            int after = next_sig(name);
            std::string type;
            if (is_p(after, ":")) {
                const int ty = next_sig(after);
                int end = ty;
                // types were already rewritten by pass_types: a single IDENT token
                if (ty < 0 || t[ty].kind != IDENT) {
                    err = "expected a type after ':' in the declaration of '" + t[name].text + "'";
                    return false;
                }
                type = t[ty].text;
                blank(after);
                for (int j = ty; j <= end; ++j) blank(j);
                after = next_sig(end);
            }
            const bool is_const = kw != "var";
            if (is_p(after, "=")) {
                t[i].text = std::string(is_const ? "const " : "") + (type.empty() ? "auto" : type);
                // literal-only initialiser: fold in double like naga's abstract-float const evaluation
                int semi = after;
                int depth = 0;
                for (int j = after + 1; j < (int)t.size(); ++j) {
                    if (is_p(j, "(")) ++depth;
                    if (is_p(j, ")")) --depth;
                    if (depth == 0 && (is_p(j, ";") || is_p(j, ","))) {
                        semi = j;
                        break;
                    }
                }
                if (semi > after && type.empty() && literal_only(after, semi)) {
                    for (int j = after + 1; j < semi; ++j)
                        if (t[j].kind == NUMBER) t[j].text.pop_back();  // drop the f: double literal
                    const int first = next_sig(after);
                    t[first].text = "(float)(" + t[first].text;
                    const int last = prev_sig(semi);
                    t[last].text += ")";
                }
            } else if (is_p(after, ";")) {
                if (type.empty()) {
                    err = "declaration of '" + t[name].text + "' needs a type or an initialiser";
                    return false;
                }
                t[i].text = type;
                t[name].text += " = " + type + "()";  // WGSL zero-initialises
            } else {
                err = "malformed declaration of '" + t[name].text + "'";
                return false;
            }
        }
        return true;
    }

    // ---- pass 3: control flow parentheses --------------------------------------------------------------------
    bool pass_control() {
        for (int i = 0; i < (int)t.size(); ++i) {
            if (t[i].kind != IDENT) continue;
            const std::string& kw = t[i].text;
            if (kw == "loop" || kw == "continuing" || kw == "switch" || kw == "fn" || kw == "discard") {
                if (!is_p(prev_sig(i), ".")) {
                    err = "'" + kw + "' is not supported in a field function body (use if / for / while)";
                    return false;
                }
            }
            if (kw != "if" && kw != "while") continue;
            const int first = next_sig(i);
            if (first < 0) {
                err = "dangling '" + kw + "'";
                return false;
            }
            int depth = 0, brace = -1;
            for (int j = first; j < (int)t.size(); ++j) {
                if (is_p(j, "(")) ++depth;
                if (is_p(j, ")")) --depth;
                if (depth == 0 && is_p(j, "{")) {
                    brace = j;
                    break;
                }
                if (depth == 0 && is_p(j, ";")) break;
            }
            if (brace < 0) {
                err = "expected '{' after the condition of '" + kw + "'";
                return false;
            }
            const int last = prev_sig(brace);
            t[first].text = "(" + t[first].text;
            t[last].text += ")";
        }
        return true;
    }

    // ---- pass 5: constant evaluation of literal-only sub-expressions -----------------------------------------------
    // WGSL literals without suffix are AbstractFloats: an expression made only of them - including calls of the
    // const-evaluable builtins, e.g. `tan(0.1) + sqrt(2.0)`, `pow(2.0, 0.5) * x`, `x + 1.0 / 3.0` - is evaluated by
    // naga's constant evaluator in f64 and converted to f32 ONCE where it meets a concrete operand.  Evaluating the
    // pieces in f32 at run time differs from that in the last bit.  This pass finds the literal-only SUBTREES of the
    // expression tree on the token stream (unary minus, parentheses, calls, then * / before + -, left-associative: a
    // run is only folded where WGSL's grammar makes it a subtree), evaluates them in double on the host and replaces
    // them by one f32 literal.  A whole literal-only `let` initialiser was already turned into (float)(double
    // expression) by pass_decls (its numbers lost their `f` and are not candidates here).  Newlines are kept.
    struct FoldElem {
        int tok;          // index into t of a raw token, or of the first token of a folded literal
        int last;         // last token covered
        bool lit;         // a literal value
        bool nontrivial;  // involves an operator or a call (worth re-emitting)
        double v;
    };
    static bool fold_call(const std::string& f, const std::vector<double>& a, double& r) {
        const size_t n = a.size();
        auto is = [&](const char* name, size_t arity) { return f == name && n == arity; };
        if (is("w_sqrt", 1)) r = std::sqrt(a[0]);
        else if (is("w_inverseSqrt", 1)) r = 1.0 / std::sqrt(a[0]);
        else if (is("w_exp", 1)) r = std::exp(a[0]);
        else if (is("w_exp2", 1)) r = std::exp2(a[0]);
        else if (is("w_log", 1)) r = std::log(a[0]);
        else if (is("w_log2", 1)) r = std::log2(a[0]);
        else if (is("w_pow", 2)) r = std::pow(a[0], a[1]);
        else if (is("w_sin", 1)) r = std::sin(a[0]);
        else if (is("w_cos", 1)) r = std::cos(a[0]);
        else if (is("w_tan", 1)) r = std::tan(a[0]);
        else if (is("w_asin", 1)) r = std::asin(a[0]);
        else if (is("w_acos", 1)) r = std::acos(a[0]);
        else if (is("w_atan", 1)) r = std::atan(a[0]);
        else if (is("w_atan2", 2)) r = std::atan2(a[0], a[1]);
        else if (is("w_sinh", 1)) r = std::sinh(a[0]);
        else if (is("w_cosh", 1)) r = std::cosh(a[0]);
        else if (is("w_tanh", 1)) r = std::tanh(a[0]);
        else if (is("w_abs", 1)) r = std::fabs(a[0]);
        else if (is("w_floor", 1)) r = std::floor(a[0]);
        else if (is("w_ceil", 1)) r = std::ceil(a[0]);
        else if (is("w_trunc", 1)) r = std::trunc(a[0]);
        else if (is("w_round", 1)) r = std::nearbyint(a[0]);  // ties to even
        else if (is("w_fract", 1)) r = a[0] - std::floor(a[0]);
        else if (is("w_sign", 1)) r = a[0] > 0.0 ? 1.0 : (a[0] < 0.0 ? -1.0 : 0.0);
        else if (is("w_min", 2)) r = a[1] < a[0] ? a[1] : a[0];
        else if (is("w_max", 2)) r = a[0] < a[1] ? a[1] : a[0];
        else if (is("w_clamp", 3)) { const double m = a[0] < a[1] ? a[1] : a[0]; r = a[2] < m ? a[2] : m; }
        else if (is("w_saturate", 1)) { const double m = a[0] < 0.0 ? 0.0 : a[0]; r = 1.0 < m ? 1.0 : m; }
        else if (is("w_step", 2)) r = a[0] <= a[1] ? 1.0 : 0.0;
        else if (is("w_radians", 1)) r = a[0] * 0.017453292519943295;
        else if (is("w_degrees", 1)) r = a[0] * 57.29577951308232;
        else return false;
        return true;
    }
    bool pass_fold() {
        std::vector<FoldElem> e;
        for (int i = 0; i < (int)t.size(); ++i) {
            if (t[i].kind == WS || t[i].kind == NL) continue;
            FoldElem x{i, i, false, false, 0.0};
            if (t[i].kind == NUMBER && t[i].text.size() > 1 && t[i].text.back() == 'f' &&
                t[i].text.find_first_of("xX") == std::string::npos) {
                x.lit = true;
                x.v = std::strtod(t[i].text.substr(0, t[i].text.size() - 1).c_str(), nullptr);
            }
            e.push_back(x);
        }
        auto punct = [&](int k, const char* p) { return k >= 0 && k < (int)e.size() && !e[k].lit && is_p(e[k].tok, p); };
        auto value_end = [&](int k) {  // can the element end an operand (so that a following + - is binary)?
            if (k < 0) return false;
            if (e[k].lit) return true;
            const Tok& x = t[e[k].tok];
            return x.kind == IDENT || x.kind == NUMBER || (x.kind == PUNCT && (x.text == ")" || x.text == "]"));
        };
        auto mul_op = [&](int k) { return punct(k, "*") || punct(k, "/") || punct(k, "%"); };
        auto add_op = [&](int k) { return (punct(k, "+") || punct(k, "-")) && value_end(k - 1); };
        auto merge = [&](int a, int b, double v) {  // elements [a, b] -> one nontrivial literal
            if (!std::isfinite(v) || !std::isfinite((double)(float)v)) return false;  // naga rejects these: leave as written
            FoldElem x{e[a].tok, e[b].last, true, true, v};
            e.erase(e.begin() + a, e.begin() + b + 1);
            e.insert(e.begin() + a, x);
            return true;
        };
        bool changed = true;
        while (changed) {
            changed = false;
            for (int k = 0; k < (int)e.size() && !changed; ++k) {
                // unary minus
                if (punct(k, "-") && !value_end(k - 1) && k + 1 < (int)e.size() && e[k + 1].lit) {
                    FoldElem x{e[k].tok, e[k + 1].last, true, e[k + 1].nontrivial, -e[k + 1].v};
                    e.erase(e.begin() + k, e.begin() + k + 2);
                    e.insert(e.begin() + k, x);
                    changed = true;
                    break;
                }
                // parentheses that are not a call's
                if (punct(k, "(") && k + 2 < (int)e.size() && e[k + 1].lit && punct(k + 2, ")") &&
                    !(k > 0 && !e[k - 1].lit && t[e[k - 1].tok].kind == IDENT)) {
                    FoldElem x{e[k].tok, e[k + 2].last, true, e[k + 1].nontrivial, e[k + 1].v};
                    e.erase(e.begin() + k, e.begin() + k + 3);
                    e.insert(e.begin() + k, x);
                    changed = true;
                    break;
                }
                // call of a const-evaluable builtin with literal arguments
                if (!e[k].lit && t[e[k].tok].kind == IDENT && punct(k + 1, "(") && !punct(k - 1, ".")) {
                    std::vector<double> args;
                    int j = k + 2;
                    bool ok = true;
                    while (ok && j < (int)e.size()) {
                        if (!e[j].lit) { ok = false; break; }
                        args.push_back(e[j].v);
                        ++j;
                        if (punct(j, ",")) { ++j; if (punct(j, ")")) break; continue; }
                        break;
                    }
                    double r;
                    if (ok && punct(j, ")") && !args.empty() && fold_call(t[e[k].tok].text, args, r) && merge(k, j, r)) {
                        changed = true;
                        break;
                    }
                }
                if (!e[k].lit || k + 2 >= (int)e.size() || !e[k + 2].lit) continue;
                // L * L, L / L: a subtree unless the left literal belongs to an earlier multiplicative chain
                if ((punct(k + 1, "*") || punct(k + 1, "/")) && !mul_op(k - 1)) {
                    const double r = punct(k + 1, "*") ? e[k].v * e[k + 2].v : e[k].v / e[k + 2].v;
                    if (merge(k, k + 2, r)) { changed = true; break; }
                }
                // L + L, L - L: a subtree unless either literal is claimed by a tighter or earlier operator
                if ((punct(k + 1, "+") || punct(k + 1, "-")) && !mul_op(k - 1) && !add_op(k - 1) && !mul_op(k + 3) &&
                    !(punct(k - 1, "-") && !value_end(k - 2))) {
                    const double r = punct(k + 1, "+") ? e[k].v + e[k + 2].v : e[k].v - e[k + 2].v;
                    if (merge(k, k + 2, r)) { changed = true; break; }
                }
            }
        }
        for (const FoldElem& x : e) {
            if (!x.lit || !x.nontrivial) continue;
            char buf[64];
            std::snprintf(buf, sizeof buf, "%.9g", (double)(float)x.v);
            std::string lit = buf;
            if (lit.find_first_of(".e") == std::string::npos) lit += ".0";
            lit += "f";
            if (lit[0] == '-') lit = "(" + lit + ")";
            t[x.tok] = {NUMBER, lit};
            for (int j = x.tok + 1; j <= x.last; ++j)
                if (t[j].kind != NL) blank(j);
        }
        return true;
    }

    // ---- pass 4: builtins, swizzles, keywords, trailing commas --------------------------------------------------
    bool pass_names() {
        for (int i = 0; i < (int)t.size(); ++i) {
            if (t[i].kind == PUNCT && t[i].text == ",") {
                if (is_p(next_sig(i), ")")) blank(i);
                continue;
            }
            if (t[i].kind != IDENT) continue;
            const int p = prev_sig(i);
            const int n = next_sig(i);
            std::string& s = t[i].text;
            if (is_p(p, ".")) {
                if (is_p(n, "(")) continue;
                const bool xyzw = s.find_first_not_of("xyzw") == std::string::npos;
                const bool rgba = s.find_first_not_of("rgba") == std::string::npos;
                if ((xyzw || rgba) && s.size() <= 4) {
                    if (rgba && !xyzw)
                        for (char& c : s) c = (c == 'r') ? 'x' : (c == 'g') ? 'y' : (c == 'b') ? 'z' : 'w';
                    if (s.size() > 1) s += "()";
                }
                continue;
            }
            if (is_p(n, "(") && builtins.count(s)) {
                s = "w_" + s;
                continue;
            }
            // runs before the other passes, so every keyword seen here is the user's own identifier
            if (keywords.count(s)) s += "_";
        }
        return true;
    }
};

}  // namespace

bool translate_wgsl_body(const std::string& body, std::string& cuda, std::string& error) {
    Lexer lx(body);
    if (!lx.run()) {
        error = "WGSL translation error: " + lx.err;
        return false;
    }
    Rewriter rw;
    rw.t = std::move(lx.out);
    // keyword-clash renaming must see the user's identifiers before passes introduce C++ keywords
    if (!rw.pass_names() || !rw.pass_types() || !rw.pass_decls() || !rw.pass_control() || !rw.pass_fold()) {
        error = "WGSL translation error: " + rw.err;
        return false;
    }
    cuda.clear();
    for (const Tok& k : rw.t) cuda += k.text;
    return true;
}


// =================================================================================================================
// Whole-shader hot reload: the helper layer of an edited main_view.wgsl
// =================================================================================================================
namespace {

// Top-level declarations of the reference's src/main_view.wgsl as (kind, name, token hash): generated by
// tools/gen_module_hashes.py from the mounted reference (hashes only, no source text).  An item whose hash matches is
// "unchanged" and keeps its hand-written CUDA counterpart.
struct Pristine {
    const char* kind;
    const char* name;
    uint64_t hash;
};
const Pristine kPristine[] = {
#include "module_items.inc"
    {"", "", 0}};

// library helpers a field function may call (README.md:37-44): an edited version overrides the prelude's
const char* const kLibraryHelpers[] = {"rotateX", "rotateY", "rotateZ", "translate", "refraction_index", "refraction",
                                       "point_plane_distance", "sigmoid"};
// names the generated struct already uses
const char* const kReservedNames[] = {"eval", "rk4", "stage", "rk4_sum", "wg_f32", "vec2f", "vec3f", "vec4f", "mat2x2f",
                                      "mat3x3f", "LinonField", "S", "settings", "camera"};

uint64_t fnv1a(uint64_t h, const std::string& s) {
    for (unsigned char c : s) {
        h ^= c;
        h *= 1099511628211ull;
    }
    h ^= ' ';
    h *= 1099511628211ull;
    return h;
}

struct RawItem {
    ModuleItem info;
    int begin, end;        // token range [begin, end)
    int name_tok = -1;
    int paren_open = -1, paren_close = -1, brace_open = -1, brace_close = -1;  // fn only
};

bool split_module(const std::vector<Tok>& t, std::vector<RawItem>& items, std::string& err) {
    auto is_p = [&](int i, const char* p) { return i >= 0 && i < (int)t.size() && t[i].kind == PUNCT && t[i].text == p; };
    auto sig = [&](int i) {  // first significant token at or after i
        while (i < (int)t.size() && (t[i].kind == WS || t[i].kind == NL)) ++i;
        return i;
    };
    std::vector<int> line_of(t.size() + 1, 1);
    for (size_t i = 0, line = 1; i < t.size(); ++i) {
        line_of[i] = (int)line;
        if (t[i].kind == NL) ++line;
    }
    auto match = [&](int open, const char* o, const char* c) {  // index of the closing token, -1 if unbalanced
        int depth = 0;
        for (int j = open; j < (int)t.size(); ++j) {
            if (is_p(j, o)) ++depth;
            if (is_p(j, c) && --depth == 0) return j;
        }
        return -1;
    };
    int i = sig(0);
    while (i < (int)t.size()) {
        RawItem it;
        it.begin = i;
        it.info.line = line_of[i];
        while (is_p(i, "@")) {  // attributes belong to the declaration that follows
            i = sig(i + 1);
            if (i >= (int)t.size() || t[i].kind != IDENT) { err = "malformed attribute"; return false; }
            i = sig(i + 1);
            if (is_p(i, "(")) {
                const int c = match(i, "(", ")");
                if (c < 0) { err = "unbalanced '(' in an attribute"; return false; }
                i = sig(c + 1);
            }
        }
        if (i >= (int)t.size()) { err = "attribute without a declaration"; return false; }
        if (is_p(i, ";")) { i = sig(i + 1); continue; }
        if (t[i].kind != IDENT) {
            err = "line " + std::to_string(line_of[i]) + ": unexpected '" + t[i].text + "' at module scope";
            return false;
        }
        const std::string kw = t[i].text;
        auto until_semicolon = [&](int from) {
            int depth = 0;
            for (int j = from; j < (int)t.size(); ++j) {
                if (is_p(j, "(") || is_p(j, "{") || is_p(j, "[")) ++depth;
                if (is_p(j, ")") || is_p(j, "}") || is_p(j, "]")) --depth;
                if (depth == 0 && is_p(j, ";")) return j;
            }
            return -1;
        };
        if (kw == "fn") {
            it.info.kind = "fn";
            it.name_tok = sig(i + 1);
            if (it.name_tok >= (int)t.size() || t[it.name_tok].kind != IDENT) { err = "expected a name after 'fn'"; return false; }
            it.paren_open = sig(it.name_tok + 1);
            if (!is_p(it.paren_open, "(")) { err = "expected '(' after 'fn " + t[it.name_tok].text + "'"; return false; }
            it.paren_close = match(it.paren_open, "(", ")");
            if (it.paren_close < 0) { err = "unbalanced '(' in the parameters of '" + t[it.name_tok].text + "'"; return false; }
            int j = it.paren_close + 1;
            while (j < (int)t.size() && !is_p(j, "{")) {
                if (is_p(j, ";")) { err = "function '" + t[it.name_tok].text + "' has no body"; return false; }
                ++j;
            }
            it.brace_open = j;
            it.brace_close = j < (int)t.size() ? match(j, "{", "}") : -1;
            if (it.brace_close < 0) { err = "unbalanced '{' in function '" + t[it.name_tok].text + "'"; return false; }
            it.end = it.brace_close + 1;
        } else if (kw == "struct") {
            it.info.kind = "struct";
            it.name_tok = sig(i + 1);
            int j = it.name_tok;
            while (j < (int)t.size() && !is_p(j, "{")) ++j;
            const int c = j < (int)t.size() ? match(j, "{", "}") : -1;
            if (c < 0) { err = "malformed struct declaration"; return false; }
            it.end = c + 1;
        } else if (kw == "var" || kw == "const" || kw == "let" || kw == "override") {
            it.info.kind = kw == "var" ? "var" : "const";
            int j = sig(i + 1);
            if (is_p(j, "<")) {  // var<uniform>, var<storage, read_write>, var<private>
                while (j < (int)t.size() && !is_p(j, ">")) ++j;
                j = sig(j + 1);
            }
            it.name_tok = j;
            const int semi = until_semicolon(i);
            if (semi < 0 || j >= (int)t.size() || t[j].kind != IDENT) { err = "malformed '" + kw + "' declaration at module scope"; return false; }
            it.end = semi + 1;
        } else {
            it.info.kind = "other";  // alias, enable, requires, diagnostic, const_assert
            it.name_tok = i;
            const int semi = until_semicolon(i);
            if (semi < 0) { err = "line " + std::to_string(line_of[i]) + ": unexpected '" + kw + "' at module scope"; return false; }
            it.end = semi + 1;
        }
        it.info.name = t[it.name_tok].text;
        uint64_t h = 1469598103934665603ull;
        for (int j = it.begin; j < it.end; ++j)
            if (t[j].kind != WS && t[j].kind != NL) h = fnv1a(h, t[j].text);
        it.info.hash = h;
        items.push_back(it);
        i = sig(it.end);
    }
    return true;
}

const Pristine* find_pristine(const std::string& kind, const std::string& name) {
    for (const Pristine* p = kPristine; p->kind[0]; ++p)
        if (kind == p->kind && name == p->name) return p;
    return nullptr;
}
template <size_t N> bool in_list(const char* const (&list)[N], const std::string& s) {
    for (const char* x : list)
        if (s == x) return true;
    return false;
}

// translate tokens [b, e) with the statement passes; newlines are kept (keep_nl) or dropped
bool translate_range(const std::vector<Tok>& t, int b, int e, bool keep_nl, std::string& out, std::string& err) {
    Rewriter rw;
    rw.t.assign(t.begin() + b, t.begin() + e);
    if (!rw.pass_names() || !rw.pass_types() || !rw.pass_decls() || !rw.pass_control() || !rw.pass_fold()) {
        err = rw.err;
        return false;
    }
    out.clear();
    for (const Tok& k : rw.t) {
        if (k.kind == NL && !keep_nl) out += ' ';
        else out += k.text;
    }
    return true;
}

// `fn name(a: T, b: U) -> R { body }` -> `static __device__ __forceinline__ R name(const T a, const U b) { preamble body }`
bool emit_function(const std::vector<Tok>& t, const RawItem& it, const std::string& preamble, std::string& out, std::string& err) {
    Rewriter rw;
    rw.t.assign(t.begin() + it.paren_open, t.begin() + it.end);
    const int off = it.paren_open;
    const int pc = it.paren_close - off, bo = it.brace_open - off;
    if (!rw.pass_names() || !rw.pass_types() || !rw.pass_decls() || !rw.pass_control() || !rw.pass_fold()) {
        err = "function '" + it.info.name + "': " + rw.err;
        return false;
    }
    // parameters: name : TYPE  ->  const TYPE name   (WGSL parameters are immutable)
    for (int i = 1; i < pc; ++i) {
        if (!rw.is_p(i, ":")) continue;
        const int name = rw.prev_sig(i), type = rw.next_sig(i);
        if (name < 1 || type < 0 || type >= pc || rw.t[name].kind != IDENT || rw.t[type].kind != IDENT) {
            err = "function '" + it.info.name + "': unsupported parameter list";
            return false;
        }
        rw.t[name].text = "const " + rw.t[type].text + " " + rw.t[name].text;
        rw.blank(i);
        rw.blank(type);
    }
    for (int i = 1; i < pc; ++i)
        if (rw.is_p(i, "@")) {
            err = "function '" + it.info.name + "': parameter attributes are not supported in a helper function";
            return false;
        }
    // return type
    std::string ret = "void";
    for (int i = pc + 1; i < bo; ++i) {
        if (rw.t[i].kind == WS || rw.t[i].kind == NL) continue;
        if (rw.is_p(i, "->")) {
            const int ty = rw.next_sig(i);
            if (ty < 0 || ty >= bo || rw.t[ty].kind != IDENT) {
                err = "function '" + it.info.name + "': unsupported return type";
                return false;
            }
            ret = rw.t[ty].text;
            rw.blank(i);
            rw.blank(ty);
            i = ty;
        } else if (!rw.t[i].text.empty()) {
            err = "function '" + it.info.name + "': unsupported return type";
            return false;
        }
    }
    out += "#line " + std::to_string(it.info.line) + " \"main_view.wgsl\"\n";
    out += "    static __device__ __forceinline__ " + ret + " " + it.info.name;
    for (int i = 0; i < (int)rw.t.size(); ++i) {
        out += rw.t[i].text;
        if (i == bo) out += " " + preamble;
    }
    out += "\n";
    return true;
}

}  // namespace

bool scan_wgsl_module(const std::string& src, std::vector<ModuleItem>& items, std::string& error) {
    Lexer lx(src);
    if (!lx.run()) {
        error = "WGSL module error: " + lx.err;
        return false;
    }
    std::vector<RawItem> raw;
    std::string err;
    if (!split_module(lx.out, raw, err)) {
        error = "WGSL module error: " + err;
        return false;
    }
    for (const RawItem& r : raw) items.push_back(r.info);
    return true;
}

bool translate_wgsl_module(const std::string& src, ModuleLayer& layer, std::string& error, bool& unsupported) {
    unsupported = false;
    layer = ModuleLayer();
    Lexer lx(src);
    if (!lx.run()) {
        error = "WGSL module error: " + lx.err;
        return false;
    }
    const std::vector<Tok>& t = lx.out;
    std::vector<RawItem> items;
    std::string err;
    if (!split_module(t, items, err)) {
        error = "WGSL module error: " + err;
        return false;
    }
    // ---- classify ---------------------------------------------------------------------------------------------------
    std::vector<int> emit_fn, emit_const, unchanged_helpers;
    for (int k = 0; k < (int)items.size(); ++k) {
        const RawItem& it = items[k];
        const std::string& kind = it.info.kind;
        const std::string& name = it.info.name;
        for (int j = 0; j < k; ++j)
            if (items[j].info.name == name && items[j].info.kind != "other" && kind != "other") {
                error = "WGSL module error: line " + std::to_string(it.info.line) + ": redefinition of '" + name + "'";
                return false;
            }
        if (kind == "fn" && name == "field_function") continue;  // the placeholder: the body is installed separately
        const Pristine* p = find_pristine(kind, name);
        if (p && p->hash == it.info.hash) {
            if (kind == "fn" && in_list(kLibraryHelpers, name)) unchanged_helpers.push_back(k);
            continue;
        }
        if (kind == "fn" && (!p || in_list(kLibraryHelpers, name))) {
            if (in_list(kReservedNames, name) || name.compare(0, 2, "w_") == 0 || name.compare(0, 6, "linon_") == 0) {
                error = "WGSL module error: line " + std::to_string(it.info.line) + ": the function name '" + name + "' is reserved on the CUDA path";
                return false;
            }
            emit_fn.push_back(k);
            continue;
        }
        if (kind == "const" && (!p || name == "PI")) {
            emit_const.push_back(k);
            continue;
        }
        unsupported = true;
        error = "main_view.wgsl line " + std::to_string(it.info.line) + ": " + (p ? "the " : "a new ") + kind + " '" + name + "' " +
                (p ? "was edited" : "was added") + "; the CUDA path hot-reloads helper functions, module constants and the field function - "
                "the march itself (hit_triangle, ray_color, nonlinear_ray_color, sample_rays, main_view, their constants, structs and "
                "bindings) is the hand-written kernel template";
        return false;
    }
    // ---- an unchanged library helper that uses an emitted name must be emitted too (it has to bind to the override) ----
    auto uses = [&](const RawItem& it, const std::string& name) {
        for (int j = it.brace_open; j < it.brace_close; ++j)
            if (t[j].kind == IDENT && t[j].text == name) return true;
        return false;
    };
    for (bool grew = true; grew;) {
        grew = false;
        for (size_t u = 0; u < unchanged_helpers.size(); ++u) {
            const RawItem& h = items[unchanged_helpers[u]];
            bool hit = false;
            for (int k : emit_fn) hit = hit || uses(h, items[k].info.name);
            for (int k : emit_const) hit = hit || uses(h, items[k].info.name);
            if (hit) {
                emit_fn.push_back(unchanged_helpers[u]);
                unchanged_helpers.erase(unchanged_helpers.begin() + u);
                grew = true;
                break;
            }
        }
    }
    // ---- constants: one line, re-declared at the top of every emitted function and of the field-function body ----
    for (int k : emit_const) {
        std::string text;
        if (!translate_range(t, items[k].begin, items[k].end, false, text, err)) {
            error = "WGSL module error: constant '" + items[k].info.name + "': " + err;
            return false;
        }
        layer.preamble += text + " ";
        layer.names += (layer.names.empty() ? "" : ",") + items[k].info.name;
    }
    for (int k : emit_fn) {
        if (!emit_function(t, items[k], layer.preamble, layer.members, err)) {
            error = "WGSL module error: " + err;
            return false;
        }
        layer.names += (layer.names.empty() ? "" : ",") + items[k].info.name;
    }
    return true;
}

}  // namespace linon
