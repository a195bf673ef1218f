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
#include <map>
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

    int line_of(int i) const {  // 1-based line of token i
        int line = 1;
        for (int j = 0; j < i && j < (int)t.size(); ++j)
            if (t[j].kind == NL) ++line;
        return line;
    }
    std::string at(int i) const { return "line " + std::to_string(line_of(i)) + ": "; }

    // ---- pass 0: the text must be WGSL, nothing else ---------------------------------------------------------------
    // The reference hands the body to naga, which sandboxes it; here the translated text becomes CUDA C++ that NVRTC
    // compiles, so whatever is not WGSL must not reach the compiler: preprocessor directives (`#include "/etc/hostname"` made
    // NVRTC read a host file), string literals, scope resolution, reserved `__` names, brackets that close the enclosing
    // function, and EVERY identifier that is not a WGSL keyword / type / builtin, a helper of main_view.wgsl, a parameter or
    // module-scope name, or declared in this text.  `extra` = names the module provides (helper functions, constants),
    // `params` = the parameters of the enclosing function.
    bool pass_sanitize(const std::set<std::string>& extra, const std::vector<std::string>& params) {
        static const char* const kw[] = {"let", "var", "const", "if", "else", "for", "while", "loop", "break", "continue",
                                         "continuing", "return", "switch", "case", "default", "true", "false", "fn", "discard",
                                         "function", "private"};
        static const char* const types[] = {"f32", "i32", "u32", "bool", "vec2", "vec3", "vec4", "vec2f", "vec3f", "vec4f",
                                            "mat2x2", "mat3x3", "mat2x2f", "mat3x3f", "array", "ptr", "atomic", "mat4x4",
                                            "mat4x4f", "bitcast", "f16", "vec2i", "vec3i", "vec4i", "vec2u", "vec3u", "vec4u"};
        static const char* const helpers[] = {"rotateX", "rotateY", "rotateZ", "translate", "refraction_index", "refraction",
                                              "point_plane_distance", "sigmoid"};
        static const char* const module_names[] = {"PI", "settings", "camera"};
        static const char* const members[] = {"field_weight", "mouse_pos", "overlay_mode", "central_difference_delta",
                                              "lyapunov_scaling", "origin", "view_direction", "up", "view_matrix"};
        std::set<std::string> known(extra);
        for (const char* k : kw) known.insert(k);
        for (const char* k : types) known.insert(k);
        for (const char* k : helpers) known.insert(k);
        for (const char* k : module_names) known.insert(k);
        for (const std::string& k : params) known.insert(k);
        for (const std::string& b : builtins) known.insert(b);
        // declarations of this text (flat: WGSL scoping is checked by the C++ compiler afterwards)
        for (int i = 0; i < (int)t.size(); ++i) {
            if (t[i].kind != IDENT || (t[i].text != "let" && t[i].text != "var" && t[i].text != "const")) continue;
            int n = next_sig(i);
            if (is_p(n, "<")) {
                while (n >= 0 && !is_p(n, ">")) n = next_sig(n);
                n = n >= 0 ? next_sig(n) : -1;
            }
            if (n >= 0 && t[n].kind == IDENT) known.insert(t[n].text);
        }
        std::vector<char> stack;
        for (int i = 0; i < (int)t.size(); ++i) {
            const Tok& k = t[i];
            if (k.kind == PUNCT) {
                const std::string& c = k.text;
                if (c == "#" || c == "\"" || c == "'" || c == "\\" || c == "@" || c == "$" || c == "`" || c == "?") {
                    err = at(i) + "'" + c + "' is not WGSL";
                    return false;
                }
                if (c == ":" && is_p(next_sig(i), ":") ) {
                    err = at(i) + "'::' is not WGSL";
                    return false;
                }
                if (c == "(" || c == "{" || c == "[") stack.push_back(c[0]);
                if (c == ")" || c == "}" || c == "]") {
                    const char open = c == ")" ? '(' : (c == "}" ? '{' : '[');
                    if (stack.empty() || stack.back() != open) {
                        err = at(i) + "unbalanced '" + c + "'";
                        return false;
                    }
                    stack.pop_back();
                }
                continue;
            }
            if (k.kind != IDENT) continue;
            if (k.text.compare(0, 2, "__") == 0) {
                err = at(i) + "'" + k.text + "': identifiers must not start with two underscores";
                return false;
            }
            if (is_p(prev_sig(i), ".")) {
                const bool swz = k.text.size() <= 4 && (k.text.find_first_not_of("xyzw") == std::string::npos ||
                                                        k.text.find_first_not_of("rgba") == std::string::npos);
                bool ok = swz;
                for (const char* m : members) ok = ok || k.text == m;
                if (!ok) {
                    err = at(i) + "unknown member '." + k.text + "'";
                    return false;
                }
                continue;
            }
            if (!known.count(k.text)) {
                err = at(i) + "unknown identifier '" + k.text + "' (not a WGSL builtin, a helper of main_view.wgsl, a parameter or a declaration of this text)";
                return false;
            }
        }
        if (!stack.empty()) {
            err = std::string("unbalanced '") + stack.back() + "'";
            return false;
        }
        // loop { body continuing { ... } }: pass_control moves the continuing block to the top of the next iteration, where
        // the declarations of the loop body are out of scope (WGSL lets a continuing block see them)
        for (int i = 0; i < (int)t.size(); ++i) {
            if (t[i].kind != IDENT || t[i].text != "loop" || is_p(prev_sig(i), ".")) continue;
            const int open = next_sig(i);
            const int close = is_p(open, "{") ? match_fwd(open) : -1;
            if (close < 0) continue;
            std::set<std::string> body_decls;
            for (int j = open + 1, depth = 0; j < close; ++j) {
                if (is_p(j, "{")) ++depth;
                if (is_p(j, "}")) --depth;
                if (t[j].kind != IDENT) continue;
                if (t[j].text == "let" || t[j].text == "var" || t[j].text == "const") {
                    const int n = next_sig(j);
                    if (n >= 0 && t[n].kind == IDENT) body_decls.insert(t[n].text);
                }
                if (depth == 0 && t[j].text == "continuing") {
                    const int co = next_sig(j), cc = is_p(co, "{") ? match_fwd(co) : -1;
                    for (int q = co + 1; q < cc; ++q)
                        if (t[q].kind == IDENT && body_decls.count(t[q].text) && !is_p(prev_sig(q), ".")) {
                            err = at(q) + "a 'continuing' block that uses '" + t[q].text + "', declared in the loop body, is not supported";
                            return false;
                        }
                    break;
                }
            }
        }
        return true;
    }

    // index of the token that opens the bracket closed at `close` (same kind), -1 if none
    int match_back(int close) const {
        const std::string c = t[close].text, o = c == ")" ? "(" : (c == "]" ? "[" : "{");
        int depth = 0;
        for (int j = close; j >= 0; --j) {
            if (is_p(j, c.c_str())) ++depth;
            if (is_p(j, o.c_str()) && --depth == 0) return j;
        }
        return -1;
    }
    int match_fwd(int open) const {
        const std::string o = t[open].text, c = o == "(" ? ")" : (o == "[" ? "]" : "}");
        int depth = 0;
        for (int j = open; j < (int)t.size(); ++j) {
            if (is_p(j, o.c_str())) ++depth;
            if (is_p(j, c.c_str()) && --depth == 0) return j;
        }
        return -1;
    }
    // first token of the postfix expression (primary with calls, members, indices) whose last token is `end`
    int postfix_start(int end) const {
        int k = end;
        for (;;) {
            int start;
            if (k < 0) return -1;
            if (is_p(k, ")") || is_p(k, "]")) {
                const int o = match_back(k);
                if (o < 0) return -1;
                const int before = prev_sig(o);
                if (is_p(k, "]")) {  // indexing: the indexed expression continues to the left
                    k = before;
                    continue;
                }
                start = (before >= 0 && t[before].kind == IDENT && !is_keyword_token(before)) ? before : o;
            } else if (t[k].kind == IDENT || t[k].kind == NUMBER) {
                start = k;
            } else {
                return -1;
            }
            const int dot = prev_sig(start);
            if (is_p(dot, ".")) {  // member of something further left
                k = prev_sig(dot);
                continue;
            }
            return start;
        }
    }
    bool is_keyword_token(int i) const {
        const std::string& s = t[i].text;
        return s == "return" || s == "if" || s == "while" || s == "for" || s == "else" || s == "switch" || s == "case";
    }
    // last token of the unary-prefixed postfix expression that starts at `first`
    int unary_end(int first) const {
        int k = first;
        while (is_p(k, "-") || is_p(k, "!") || is_p(k, "~") || is_p(k, "+")) k = next_sig(k);
        if (k < 0) return -1;
        int end;
        if (is_p(k, "(")) end = match_fwd(k);
        else if (t[k].kind == IDENT || t[k].kind == NUMBER) end = k;
        else return -1;
        for (;;) {  // postfix: call, index, member
            if (end < 0) return -1;
            const int n = next_sig(end);
            if (is_p(n, "(") && t[end].kind == IDENT) { end = match_fwd(n); continue; }
            if (is_p(n, "[")) { end = match_fwd(n); continue; }
            if (is_p(n, ".")) { end = next_sig(n); continue; }
            return end;
        }
    }
    bool unary_position(int i) const {  // is the + or - at i a sign rather than a binary operator?
        const int p = prev_sig(i);
        if (p < 0) return true;
        if (t[p].kind == NUMBER) return false;
        if (t[p].kind == IDENT) return is_keyword_token(p);
        return !(is_p(p, ")") || is_p(p, "]"));
    }

    // ---- pass: WGSL's % works on floats, C++'s does not: a % b -> w_mod(a, b), a %= b -> a = w_mod(a, b) ----------------
    bool pass_mod() {
        // rewrites insert TOKENS (never edit a bracket's text): later scans must still recognise every ( ) [ ]
        for (int i = 0; i < (int)t.size(); ++i) {
            if (t[i].kind != PUNCT || (t[i].text != "%" && t[i].text != "%=")) continue;
            const int left_end = prev_sig(i), right_first = next_sig(i);
            int start = left_end >= 0 ? postfix_start(left_end) : -1;
            if (start < 0 || right_first < 0) {
                err = at(i) + "cannot find the operands of '" + t[i].text + "'";
                return false;
            }
            if (t[i].text == "%=") {
                int semi = -1, depth = 0;
                for (int j = i + 1; j < (int)t.size(); ++j) {
                    if (is_p(j, "(") || is_p(j, "[")) ++depth;
                    if (is_p(j, ")") || is_p(j, "]")) --depth;
                    if (depth == 0 && is_p(j, ";")) { semi = j; break; }
                }
                if (semi < 0) { err = at(i) + "'%=' without a terminating ';'"; return false; }
                std::vector<Tok> head = {{PUNCT, "="}, {WS, " "}, {IDENT, "w_mod"}, {PUNCT, "("}};
                for (int j = start; j <= left_end; ++j) head.push_back(t[j]);
                head.push_back({PUNCT, ","});
                head.push_back({PUNCT, "("});
                const std::vector<Tok> tail = {{PUNCT, ")"}, {PUNCT, ")"}};
                t.insert(t.begin() + semi, tail.begin(), tail.end());
                t.erase(t.begin() + i);
                t.insert(t.begin() + i, head.begin(), head.end());
                continue;
            }
            // the left operand is the whole multiplicative chain that precedes (left-associative with * and /)
            for (;;) {
                int sgn = prev_sig(start);
                while ((is_p(sgn, "-") || is_p(sgn, "+") || is_p(sgn, "!")) && unary_position(sgn)) {
                    start = sgn;
                    sgn = prev_sig(start);
                }
                const int op = prev_sig(start);
                if (!(is_p(op, "*") || is_p(op, "/"))) break;
                const int s2 = postfix_start(prev_sig(op));
                if (s2 < 0) { err = at(i) + "cannot find the left operand of '%'"; return false; }
                start = s2;
            }
            const int end = unary_end(right_first);
            if (end < 0) { err = at(i) + "cannot find the right operand of '%'"; return false; }
            t[i] = {PUNCT, ","};
            t.insert(t.begin() + end + 1, Tok{PUNCT, ")"});
            const std::vector<Tok> head = {{IDENT, "w_mod"}, {PUNCT, "("}};
            t.insert(t.begin() + start, head.begin(), head.end());
            i += 2;  // the comma that replaced this %
        }
        return true;
    }

    // ---- pass: WGSL never converts i32 / u32 to f32 implicitly, C++ does ------------------------------------------------
    // `let a = 1; return v * a;` is a type error in the reference (naga) and would compile here.  A conservative token-level
    // check: a name declared as an integer (typed i32 / u32, or initialised by an integer literal) must not be a direct
    // operand of + - * / whose other operand is a float literal, a float parameter (p_prev p v0 v t, possibly swizzled) or a
    // name declared from a float literal / a vector constructor / with a float type.  Anything this cannot decide is left to
    // the C++ compiler (which accepts it).
    bool pass_intmix(const std::vector<std::string>& float_params) {
        std::set<std::string> ints, floats(float_params.begin(), float_params.end());
        for (int i = 0; i < (int)t.size(); ++i) {
            if (t[i].kind != IDENT || (t[i].text != "let" && t[i].text != "var" && t[i].text != "const")) continue;
            int n = next_sig(i);
            if (is_p(n, "<")) {
                while (n >= 0 && !is_p(n, ">")) n = next_sig(n);
                n = n >= 0 ? next_sig(n) : -1;
            }
            if (n < 0 || t[n].kind != IDENT) continue;
            const std::string name = t[n].text;
            int a = next_sig(n);
            if (is_p(a, ":")) {
                const int ty = next_sig(a);
                if (ty >= 0 && (t[ty].text == "i32" || t[ty].text == "u32")) ints.insert(name);
                else if (ty >= 0 && (t[ty].text == "f32" || t[ty].text.compare(0, 3, "vec") == 0)) floats.insert(name);
                continue;
            }
            if (!is_p(a, "=")) continue;
            const int v = next_sig(a), after = v >= 0 ? next_sig(v) : -1;
            if (v < 0) continue;
            if (t[v].kind == NUMBER && is_p(after, ";")) {
                const std::string& lit = t[v].text;
                (lit.back() == 'f' && lit.find_first_of("xX") == std::string::npos ? floats : ints).insert(name);
            } else if (t[v].kind == IDENT && t[v].text.compare(0, 3, "vec") == 0) {
                floats.insert(name);
            }
        }
        if (ints.empty()) return true;
        auto is_float_operand_end = [&](int k) {  // token k ends a float operand (literal, float name, or a swizzle of one)
            if (k < 0) return false;
            if (t[k].kind == NUMBER) return t[k].text.back() == 'f' && t[k].text.find_first_of("xX") == std::string::npos;
            if (t[k].kind != IDENT) return false;
            if (is_p(prev_sig(k), ".")) {
                const int base = prev_sig(prev_sig(k));
                return base >= 0 && t[base].kind == IDENT && floats.count(t[base].text) > 0;
            }
            return floats.count(t[k].text) > 0 && !is_p(next_sig(k), "(");
        };
        auto is_float_operand_start = [&](int k) {
            if (k < 0) return false;
            if (t[k].kind == NUMBER) return t[k].text.back() == 'f' && t[k].text.find_first_of("xX") == std::string::npos;
            return t[k].kind == IDENT && floats.count(t[k].text) > 0 && !is_p(next_sig(k), "(");
        };
        for (int i = 0; i < (int)t.size(); ++i) {
            if (t[i].kind != IDENT || !ints.count(t[i].text) || is_p(prev_sig(i), ".")) continue;
            const int p = prev_sig(i), n = next_sig(i);
            auto arith = [&](int k) { return is_p(k, "*") || is_p(k, "/") || is_p(k, "+") || is_p(k, "-"); };
            const bool left = arith(p) && !unary_position(p) && is_float_operand_end(prev_sig(p));
            const bool right = arith(n) && is_float_operand_start(next_sig(n));
            if (left || right) {
                err = at(i) + "'" + t[i].text + "' is an integer: WGSL does not convert i32 / u32 to f32 implicitly (write f32(" + t[i].text + "))";
                return false;
            }
        }
        return true;
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

    // ---- pass 3: control flow ----------------------------------------------------------------------------------------
    // if / while get their parentheses; `loop { ... continuing { ... } }` and `switch e { case a, b: { ... } default: { ... } }`
    // become their C++ forms.
    bool pass_control() {
        for (int i = 0; i < (int)t.size(); ++i) {
            if (t[i].kind != IDENT) continue;
            const std::string kw = t[i].text;
            if (is_p(prev_sig(i), ".")) continue;
            if (kw == "fn" || kw == "discard") {
                err = "'" + kw + "' is not supported in a field function body";
                return false;
            }
            if (kw == "loop") {
                // loop { body continuing { c; break if e; } }  ->  for (bool first_ = true;; first_ = false) { if (!first_) { c;
                // if (e) break; } body }: the continuing block runs between iterations, also after a `continue`.  It moves to
                // the top of the NEXT iteration, so it cannot see declarations of the loop body (rejected below).
                const int open = next_sig(i);
                if (!is_p(open, "{")) { err = at(i) + "expected '{' after 'loop'"; return false; }
                const int close = match_fwd(open);
                if (close < 0) { err = at(i) + "unbalanced '{' after 'loop'"; return false; }
                int cont = -1;
                for (int j = open + 1, depth = 0; j < close; ++j) {
                    if (is_p(j, "{")) ++depth;
                    if (is_p(j, "}")) --depth;
                    if (depth == 0 && t[j].kind == IDENT && t[j].text == "continuing") cont = j;
                }
                if (cont < 0) {
                    t[i].text = "for (;;)";
                    continue;
                }
                const int copen = next_sig(cont), cclose = is_p(copen, "{") ? match_fwd(copen) : -1;
                if (cclose < 0 || next_sig(cclose) != close) { err = at(cont) + "'continuing' must be the last statement of its loop"; return false; }
                std::set<std::string> body_decls;
                for (int j = open + 1; j < cont; ++j)
                    if (t[j].kind == IDENT && (t[j].text == "let" || t[j].text == "var" || t[j].text == "const")) {
                        const int n = next_sig(j);
                        if (n >= 0 && t[n].kind == IDENT) body_decls.insert(t[n].text);
                    }
                std::string c_text;
                for (int j = copen + 1; j < cclose; ++j) {
                    if (t[j].kind == IDENT && body_decls.count(t[j].text) && !is_p(prev_sig(j), ".")) {
                        err = at(j) + "a 'continuing' block that uses '" + t[j].text + "', declared in the loop body, is not supported";
                        return false;
                    }
                    if (t[j].kind == IDENT && t[j].text == "break" && t[next_sig(j)].kind == IDENT && t[next_sig(j)].text == "if") {
                        // break if e;  ->  if (e) break;
                        const int iff = next_sig(j);
                        int semi = -1;
                        for (int q = iff + 1; q < cclose; ++q)
                            if (is_p(q, ";")) { semi = q; break; }
                        if (semi < 0) { err = at(j) + "'break if' without ';'"; return false; }
                        std::string cond;
                        for (int q = iff + 1; q < semi; ++q) cond += t[q].kind == NL ? " " : t[q].text;
                        c_text += "if (" + cond + ") break;";
                        j = semi;
                        continue;
                    }
                    c_text += t[j].kind == NL ? " " : t[j].text;
                }
                for (int j = cont; j <= cclose; ++j)
                    if (t[j].kind != NL) blank(j);
                const std::string flag = "linon_first_" + std::to_string(i) + "_";
                t[i].text = "for (bool " + flag + " = true;; " + flag + " = false)";
                t[open].text = "{ if (!" + flag + ") { " + c_text + " }";
                continue;
            }
            if (kw == "continuing") {
                err = at(i) + "'continuing' outside a loop";
                return false;
            }
            if (kw == "switch") {
                const int first = next_sig(i);
                int open = -1;
                for (int j = first, depth = 0; j >= 0 && j < (int)t.size(); ++j) {
                    if (is_p(j, "(")) ++depth;
                    if (is_p(j, ")")) --depth;
                    if (depth == 0 && is_p(j, "{")) { open = j; break; }
                }
                const int close = open >= 0 ? match_fwd(open) : -1;
                if (first < 0 || close < 0) { err = at(i) + "malformed 'switch'"; return false; }
                const int last = prev_sig(open);
                t[first].text = "(" + t[first].text;
                t[last].text += ")";
                // clauses: `case a, b: {` / `case a {` / `default: {` / `default {`; WGSL never falls through: `break;` is appended
                for (int j = open + 1; j < close; ++j) {
                    if (t[j].kind != IDENT || (t[j].text != "case" && t[j].text != "default") || is_p(prev_sig(j), ".")) {
                        if (is_p(j, "{")) j = match_fwd(j);  // a nested block that is not a clause cannot occur here, be safe
                        continue;
                    }
                    int body = -1;
                    for (int q = j + 1; q < close; ++q)
                        if (is_p(q, "{")) { body = q; break; }
                    const int bclose = body >= 0 ? match_fwd(body) : -1;
                    if (bclose < 0 || bclose > close) { err = at(j) + "malformed switch clause"; return false; }
                    bool colon = false;
                    if (t[j].text == "case") {
                        // selectors separated by commas: each becomes its own label; `default` may appear among them
                        for (int q = j + 1; q < body; ++q) {
                            if (is_p(q, ",")) {
                                const int nq = next_sig(q);
                                if (nq >= body || is_p(nq, ":")) { blank(q); continue; }  // trailing comma
                                const bool dflt = t[nq].kind == IDENT && t[nq].text == "default";
                                t[q].text = dflt ? ":" : ": case";
                            }
                            if (is_p(q, ":")) colon = true;
                        }
                    } else {
                        for (int q = j + 1; q < body; ++q)
                            if (is_p(q, ":")) colon = true;
                    }
                    if (!colon) t[body].text = ": {";
                    t[bclose].text = "} break;";
                    j = bclose;
                }
                continue;
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

namespace {
const std::vector<std::string> kFieldParams = {"p_prev", "p", "v0", "v", "t"};  // src/main_view.rs:853-855
std::set<std::string> split_names(const std::string& csv) {
    std::set<std::string> out;
    size_t b = 0;
    while (b <= csv.size()) {
        const size_t e = csv.find(',', b);
        const std::string n = csv.substr(b, e == std::string::npos ? std::string::npos : e - b);
        if (!n.empty()) out.insert(n);
        if (e == std::string::npos) break;
        b = e + 1;
    }
    return out;
}
}  // namespace

bool translate_wgsl_body(const std::string& body, std::string& cuda, std::string& error) {
    return translate_wgsl_body(body, "", cuda, error);
}

bool translate_wgsl_body(const std::string& body, const std::string& module_names, std::string& cuda, std::string& error) {
    Lexer lx(body);
    if (!lx.run()) {
        error = "WGSL translation error: " + lx.err;
        return false;
    }
    Rewriter rw;
    rw.t = std::move(lx.out);
    // keyword-clash renaming must see the user's identifiers before passes introduce C++ keywords
    if (!rw.pass_sanitize(split_names(module_names), kFieldParams) || !rw.pass_intmix(kFieldParams) || !rw.pass_names() ||
        !rw.pass_types() || !rw.pass_decls() || !rw.pass_control() || !rw.pass_mod() || !rw.pass_fold()) {
        // "line N: what" -> the format of the compiler's own diagnostics for the body, so callers parse one shape
        if (rw.err.compare(0, 5, "line ") == 0 && rw.err.find(": ") != std::string::npos) {
            const size_t c = rw.err.find(": ");
            error = "field_function.wgsl(" + rw.err.substr(5, c - 5) + "): error: " + rw.err.substr(c + 2);
        } else {
            error = "WGSL translation error: " + rw.err;
        }
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
    uint64_t masked_hash;
};
const Pristine kPristine[] = {
#include "module_items.inc"
    {"", "", 0, 0}};
// the numeric literals of the reference's nonlinear_ray_color and sample_rays in order of appearance (values only)
#include "module_literals.inc"
// ray_color with the statements of its shading block removed (src/main_view.wgsl:177-220 without :200-217)
#include "module_shade.inc"

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
        uint64_t h = 1469598103934665603ull, hm = 1469598103934665603ull;
        for (int j = it.begin; j < it.end; ++j) {
            if (t[j].kind == WS || t[j].kind == NL) continue;
            h = fnv1a(h, t[j].text);
            if (t[j].kind == NUMBER) {
                hm = fnv1a(hm, "#");
                std::string num = t[j].text;
                while (!num.empty() && (num.back() == 'f' || num.back() == 'u')) num.pop_back();
                it.info.literals.push_back(std::strtod(num.c_str(), nullptr));
            } else {
                hm = fnv1a(hm, t[j].text);
            }
        }
        it.info.hash = h;
        it.info.masked_hash = hm;
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
bool translate_range(const std::vector<Tok>& t, int b, int e, bool keep_nl, const std::set<std::string>& module_names,
                     std::string& out, std::string& err) {
    Rewriter rw;
    rw.t.assign(t.begin() + b, t.begin() + e);
    if (!rw.pass_sanitize(module_names, {}) || !rw.pass_names() || !rw.pass_types() || !rw.pass_decls() || !rw.pass_control() ||
        !rw.pass_mod() || !rw.pass_fold()) {
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
bool emit_function(const std::vector<Tok>& t, const RawItem& it, const std::string& preamble, const std::set<std::string>& module_names,
                   std::string& out, std::string& err) {
    Rewriter rw;
    rw.t.assign(t.begin() + it.paren_open, t.begin() + it.end);
    const int off = it.paren_open;
    const int pc = it.paren_close - off, bo = it.brace_open - off;
    std::vector<std::string> params;  // `name : type` inside the parameter list
    for (int i = 1; i < pc; ++i)
        if (rw.is_p(i, ":") && rw.prev_sig(i) >= 1 && rw.t[rw.prev_sig(i)].kind == IDENT) params.push_back(rw.t[rw.prev_sig(i)].text);
    for (int i = 1; i < pc; ++i)
        if (rw.is_p(i, "@")) {
            err = "function '" + it.info.name + "': parameter attributes are not supported in a helper function";
            return false;
        }
    if (!rw.pass_sanitize(module_names, params) || !rw.pass_names() || !rw.pass_types() || !rw.pass_decls() || !rw.pass_control() ||
        !rw.pass_mod() || !rw.pass_fold()) {
        // "line N: what" counts from the function's first line: cite the module's own line like the compiler would
        if (rw.err.compare(0, 5, "line ") == 0 && rw.err.find(": ") != std::string::npos) {
            const size_t c = rw.err.find(": ");
            const int rel = std::atoi(rw.err.substr(5, c - 5).c_str());
            // the token range starts at the '(' of the parameter list, which sits on the line of `fn name`
            int fn_line = it.info.line;
            for (int j = it.begin; j < it.paren_open; ++j)
                if (t[j].kind == NL) ++fn_line;
            err = "main_view.wgsl(" + std::to_string(fn_line + rel - 1) + "): error: function '" + it.info.name + "': " + rw.err.substr(c + 2);
        } else {
            err = "function '" + it.info.name + "': " + rw.err;
        }
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

// ---- march-level edits ------------------------------------------------------------------------------------------------
std::string f32_literal(double x) {
    char b[64];
    std::snprintf(b, sizeof b, "%.9g", (double)(float)x);
    std::string s = b;
    if (s.find_first_of(".en") == std::string::npos) s += ".0";
    return "(" + s + "f)";
}

// value tokens of `const NAME [: TYPE] = VALUE ;` -> numbers (with their signs) and identifiers
bool const_value(const std::vector<Tok>& t, const RawItem& it, std::vector<double>& nums, std::vector<std::string>& idents) {
    int eq = -1;
    for (int j = it.begin; j < it.end; ++j)
        if (t[j].kind == PUNCT && t[j].text == "=") { eq = j; break; }
    if (eq < 0) return false;
    double sign = 1.0;
    for (int j = eq + 1; j < it.end; ++j) {
        const Tok& k = t[j];
        if (k.kind == WS || k.kind == NL) continue;
        if (k.kind == NUMBER) {
            std::string num = k.text;
            while (!num.empty() && (num.back() == 'f' || num.back() == 'u')) num.pop_back();
            nums.push_back(sign * std::strtod(num.c_str(), nullptr));
            sign = 1.0;
        } else if (k.kind == IDENT) {
            idents.push_back(k.text);
        } else if (k.text == "-") {
            sign = -sign;
        } else if (k.text != "(" && k.text != ")" && k.text != "," && k.text != "<" && k.text != ">" && k.text != ";" && k.text != "+") {
            return false;  // an expression, not a literal value
        }
    }
    return true;
}

// A pristine-named module constant with another value: true = mapped to the kernel template (layer updated); false with
// err empty = not a constant this path knows; false with err set = known but outside the supported range.
bool apply_march_const(const std::vector<Tok>& t, const RawItem& it, ModuleLayer& layer, std::string& err,
                       std::map<std::string, std::string>& values) {
    const std::string& name = it.info.name;
    std::vector<double> n;
    std::vector<std::string> id;
    if (!const_value(t, it, n, id)) return false;
    auto is_bool = [&](bool& v) {
        if (!n.empty()) return false;
        int seen = 0;
        for (const std::string& s : id) {
            if (s == "true") { v = true; ++seen; }
            if (s == "false") { v = false; ++seen; }
        }
        return seen == 1;
    };
    auto out_of_range = [&](const char* what) {
        err = "the constant '" + name + "' is outside the range the CUDA path supports (" + what + ")";
        return false;
    };
    bool b = false;
    if (name == "backface_culling" || name == "use_lighting" || name == "linear_mode" || name == "adaptive_sampling" ||
        name == "sample_outline_rays") {
        if (!is_bool(b)) return false;
        if (name == "backface_culling") layer.march_defines += std::string("#define LINON_BACKFACE_CULLING ") + (b ? "1" : "0") + "\n";
        if (name == "use_lighting") layer.march_defines += std::string("#define LINON_USE_LIGHTING ") + (b ? "1" : "0") + "\n";
        if (name == "linear_mode") {
            layer.march_defines += std::string("#define LINON_LINEAR_MODE ") + (b ? "1" : "0") + "\n";
            layer.linear_mode = b;
        }
        if (name == "adaptive_sampling") layer.adaptive_sampling = b ? 1 : 0;
        if (name == "sample_outline_rays") layer.sample_outline_rays = b ? 1 : 0;
        values[name] = b ? "true" : "false";
    } else if (name == "eps" || name == "ambient_strength" || name == "shininess") {
        if (n.size() != 1 || !std::isfinite(n[0])) return false;
        if (name == "eps") {
            if (!(n[0] >= 0.0 && n[0] <= 1e-2)) return out_of_range("0 <= eps <= 0.01");
            layer.march_defines += "#define LINON_EPS " + f32_literal(n[0]) + "\n";
        } else if (name == "ambient_strength") {
            if (!(n[0] >= 0.0 && n[0] <= 100.0)) return out_of_range("0 <= ambient_strength <= 100");
            layer.march_defines += "#define LINON_AMBIENT " + f32_literal(n[0]) + "\n";
        } else {
            if (!(n[0] >= 0.0 && n[0] <= 1e4)) return out_of_range("0 <= shininess <= 10000");
            layer.march_defines += "#define LINON_SHININESS " + f32_literal(n[0]) + "\n";
        }
        values[name] = f32_literal(n[0]);
    } else if (name == "light_color" || name == "object_color") {
        if (n.size() != 3 && n.size() != 1) return false;
        if (n.size() == 1) n = {n[0], n[0], n[0]};  // vec3<f32>(x) splats
        for (double x : n)
            if (!std::isfinite(x) || std::fabs(x) > 1e3) return out_of_range("|component| <= 1000");
        if (name == "light_color")
            layer.march_defines += "#define LINON_LIGHT_R " + f32_literal(n[0]) + "\n#define LINON_LIGHT_G " + f32_literal(n[1]) +
                                   "\n#define LINON_LIGHT_B " + f32_literal(n[2]) + "\n";
        // object_color is declared and never used by the reference's shader: any value renders the same frame
        values[name] = "vec3f(" + f32_literal(n[0]) + ", " + f32_literal(n[1]) + ", " + f32_literal(n[2]) + ")";
    } else {
        return false;
    }
    layer.names += (layer.names.empty() ? "" : ",") + name;
    return true;
}

// nonlinear_ray_color / sample_rays whose tokens equal the reference's except for the VALUES of numeric literals
bool apply_march_literals(const RawItem& it, ModuleLayer& layer, std::string& err) {
    const std::string& name = it.info.name;
    const double* pristine = nullptr;
    size_t n_pristine = 0;
    if (name == "nonlinear_ray_color") { pristine = kLit_nonlinear_ray_color; n_pristine = sizeof(kLit_nonlinear_ray_color) / sizeof(double); }
    if (name == "sample_rays") { pristine = kLit_sample_rays; n_pristine = sizeof(kLit_sample_rays) / sizeof(double); }
    if (name == "main_view") { pristine = kLit_main_view; n_pristine = sizeof(kLit_main_view) / sizeof(double); }
    if (!pristine || it.info.literals.size() != n_pristine) return false;
    struct Hole { int ordinal; const char* what; };
    // ordinals verified against the reference's text by tools/gen_module_hashes.py (kLitHoles_* in module_literals.inc)
    for (size_t k = 0; k < n_pristine; ++k) {
        const double v = it.info.literals[k];
        if (v == pristine[k]) continue;
        const char* what = nullptr;
        static const char* const names_nl[] = {"h_initial", "t_max", "reject_factor", "reject_min_h", "small_h", "restore_factor"};
        static const char* const names_sr[] = {"sample_h", "sample_steps"};
        static const char* const names_mv[] = {"focal_length", "viewport_height", "linear_max_dist"};
        const int* holes = name == "nonlinear_ray_color" ? kLitHoles_nonlinear_ray_color : (name == "sample_rays" ? kLitHoles_sample_rays : kLitHoles_main_view);
        const char* const* hole_names = name == "nonlinear_ray_color" ? names_nl : (name == "sample_rays" ? names_sr : names_mv);
        const int n_holes = name == "nonlinear_ray_color" ? 6 : (name == "sample_rays" ? 2 : 3);
        for (int q = 0; q < n_holes; ++q)
            if (holes[q] == (int)k) what = hole_names[q];
        if (!what) {
            err = "the function '" + name + "' was edited (literal " + std::to_string(k + 1) + " of " + std::to_string(n_pristine) + "); the CUDA path maps " +
                  (name == "main_view" ? "its focal length, viewport height and the reach of linear_mode"
                                       : "its step size, horizon and step-control constants") +
                  " to the kernel template, nothing else of it";
            return false;
        }
        auto bad = [&](const char* range) {
            err = std::string("'") + what + "' of '" + name + "' is outside the range the CUDA path supports (" + range + ")";
            return false;
        };
        const std::string w = what;
        if (!std::isfinite(v)) return bad("finite");
        if (w == "h_initial") {
            if (!(v >= 1e-5 && v <= 1.0)) return bad("1e-5 <= h_initial <= 1");
            layer.h_initial = v;
        } else if (w == "t_max") {
            if (!(v > 0.0 && v <= 100.0)) return bad("0 < horizon <= 100");
            layer.t_max = v;
            layer.march_defines += "#define LINON_T_MAX " + f32_literal(v) + "\n";
        } else if (w == "reject_factor") {
            if (!(v >= 1.0 && v <= 1e6)) return bad("1 <= factor <= 1e6");
            layer.march_defines += "#define LINON_REJECT_FACTOR " + f32_literal(v) + "\n";
        } else if (w == "reject_min_h") {
            if (!(v >= 0.0 && v <= 1.0)) return bad("0 <= threshold <= 1");
            layer.reject_min_h = v;
            layer.march_defines += "#define LINON_REJECT_MIN_H " + f32_literal(v) + "\n";
        } else if (w == "small_h") {
            if (!(v >= 1e-5 && v <= 1.0)) return bad("1e-5 <= h <= 1");
            layer.march_defines += "#define LINON_SMALL_H " + f32_literal(v) + "\n";
        } else if (w == "restore_factor") {
            if (!(v >= 0.0 && v <= 1e6)) return bad("0 <= factor <= 1e6");
            layer.march_defines += "#define LINON_RESTORE_FACTOR " + f32_literal(v) + "\n";
        } else if (w == "sample_h") {
            if (!(v >= 1e-5 && v <= 1.0)) return bad("1e-5 <= h <= 1");
            layer.march_defines += "#define LINON_SAMPLE_H " + f32_literal(v) + "\n";
        } else if (w == "focal_length") {
            if (!(v >= 1e-4 && v <= 10.0)) return bad("1e-4 <= focal_length <= 10");
            layer.march_defines += "#define LINON_FOCAL_LENGTH " + f32_literal(v) + "\n";
        } else if (w == "viewport_height") {
            if (!(v >= 1e-4 && v <= 10.0)) return bad("1e-4 <= viewport_height <= 10");
            layer.march_defines += "#define LINON_VIEWPORT_HEIGHT " + f32_literal(v) + "\n";
        } else if (w == "linear_max_dist") {
            if (!(v > 0.0 && v <= 1e6)) return bad("0 < reach <= 1e6");
            layer.march_defines += "#define LINON_LINEAR_MAX_DIST " + f32_literal(v) + "\n";
        } else if (w == "sample_steps") {
            const long steps = (long)v;
            if ((double)steps != v || steps < 100 || steps > 1000000 || steps % 100 != 0) return bad("a multiple of 100 in 100 ... 1 000 000: 100 samples per ray");
            layer.march_defines += "#define LINON_SAMPLE_STEPS " + std::to_string(steps) + "\n";
        }
    }
    layer.names += (layer.names.empty() ? "" : ",") + name;
    return true;
}


// ray_color edited inside its shading block only (`if (t > 0.0 && t < max_dist) { ... }`, src/main_view.wgsl:199-218 - the last
// `if` of the function): locates the block in the user's text.  The tokens outside it must hash to the reference's.
bool find_shade_block(const std::vector<Tok>& t, const RawItem& it, int& open, int& close) {
    auto is_p = [&](int i, const char* p) { return t[i].kind == PUNCT && t[i].text == p; };
    int depth = 0, last_if = -1;
    for (int j = it.brace_open + 1; j < it.brace_close; ++j) {
        if (is_p(j, "{") || is_p(j, "(") || is_p(j, "[")) ++depth;
        if (is_p(j, "}") || is_p(j, ")") || is_p(j, "]")) --depth;
        if (depth == 0 && t[j].kind == IDENT && t[j].text == "if") last_if = j;
    }
    if (last_if < 0) return false;
    int j = last_if + 1;
    while (j < it.brace_close && !is_p(j, "(")) ++j;
    for (depth = 0; j < it.brace_close; ++j) {  // the condition
        if (is_p(j, "(")) ++depth;
        if (is_p(j, ")") && --depth == 0) break;
    }
    ++j;
    while (j < it.brace_close && (t[j].kind == WS || t[j].kind == NL)) ++j;
    if (j >= it.brace_close || !is_p(j, "{")) return false;
    open = j;
    for (depth = 0; j < it.brace_close; ++j) {
        if (is_p(j, "{")) ++depth;
        if (is_p(j, "}") && --depth == 0) break;
    }
    if (j >= it.brace_close) return false;
    close = j;
    uint64_t h = 1469598103934665603ull;
    for (int k = it.begin; k < it.end; ++k) {
        if (k > open && k < close) continue;
        if (t[k].kind == WS || t[k].kind == NL) continue;
        h = fnv1a(h, t[k].text);
    }
    return h == kRayColorRestHash;
}

// The user's shading block as a member of LinonField<S>: `vec4f linon_ray_color_shade(origin, direction, max_dist, t, d1, d2)`
// (march_kernels.cuh: shade_hit calls it at every geometric hit).  The variables of ray_color that are live in the block
// (t, d1, d2: the nearest hit and its triangle's edges) arrive as values; the module's march-level constants are declared
// with the values of the user's module (const_values) or the reference's; a block that falls through returns ray_color's
// final `vec4<f32>(0.0, 0.0, 0.0, 0.0)`.
bool emit_shade_block(const std::vector<Tok>& t, const RawItem& it, int open, int close, const std::string& preamble,
                      std::set<std::string> module_names, const std::map<std::string, std::string>& const_values, std::string& out,
                      std::string& err) {
    std::string block;
    for (int j = open + 1; j < close; ++j) block += t[j].text;
    int line = it.info.line;
    for (int j = it.begin; j < open; ++j)
        if (t[j].kind == NL) ++line;
    const std::string text =
        "fn linon_ray_color_shade(origin: vec3<f32>, direction: vec3<f32>, max_dist: f32, linon_t: f32, linon_d1: vec3<f32>, "
        "linon_d2: vec3<f32>) -> vec4<f32> { { var t: f32 = linon_t; var d1: vec3<f32> = linon_d1; var d2: vec3<f32> = linon_d2; " +
        block + "\n}\nreturn vec4<f32>(0.0, 0.0, 0.0, 0.0);\n}\n";  // inner scope: t, d1, d2 may shadow a module constant, like in WGSL
    Lexer lx(text);
    if (!lx.run()) {
        err = "ray_color: " + lx.err;
        return false;
    }
    std::vector<RawItem> items;
    if (!split_module(lx.out, items, err) || items.size() != 1 || items[0].info.kind != "fn") {
        err = "ray_color: the shading block does not parse as statements" + (err.empty() ? std::string() : " (" + err + ")");
        return false;
    }
    items[0].info.line = line;  // diagnostics cite the user's main_view.wgsl lines: the header above sits on the block's `{` line
    items[0].info.name = "ray_color";
    struct Default { const char* name; const char* type; const char* value; };
    static const Default kDefaults[] = {  // src/main_view.wgsl:65-74, :227, :337
        {"backface_culling", "bool", "false"}, {"light_color", "vec3f", "vec3f(1.0f, 1.0f, 1.0f)"},
        {"ambient_strength", "wg_f32", "(0.01f)"}, {"shininess", "wg_f32", "(64.0f)"},
        {"object_color", "vec3f", "vec3f(0.5f, 0.5f, 0.5f)"}, {"linear_mode", "bool", "false"}, {"use_lighting", "bool", "true"},
        {"eps", "wg_f32", "(0.0000001f)"}, {"adaptive_sampling", "bool", "true"}, {"sample_outline_rays", "bool", "false"}};
    std::string consts;
    for (const Default& d : kDefaults) {
        const auto v = const_values.find(d.name);
        consts += std::string("const ") + d.type + " " + d.name + " = " + (v != const_values.end() ? v->second : std::string(d.value)) + "; ";
        module_names.insert(d.name);
    }
    std::string member;
    if (!emit_function(lx.out, items[0], preamble + consts, module_names, member, err)) return false;
    const size_t at = member.find(" ray_color(");
    if (at == std::string::npos) {
        err = "ray_color: internal error while emitting the shading block";
        return false;
    }
    member.replace(at, std::strlen(" ray_color("), " linon_ray_color_shade(");
    out += member;
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
    std::map<std::string, std::string> const_values;  // march-level constants the module edits, as CUDA expressions
    int shade_item = -1, shade_open = -1, shade_close = -1;
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
        // ray_color edited in its shading block only: the block is translated from the user's text (after the loop, when
        // the constants of the whole module are known)
        if (kind == "fn" && name == "ray_color" && p) {
            int bo = -1, bc = -1;
            if (find_shade_block(t, it, bo, bc)) {
                shade_item = k;
                shade_open = bo;
                shade_close = bc;
                continue;
            }
        }
        // march-level constants and literals: mapped to macros of the kernel template (march_kernels.cuh)
        {
            std::string merr;
            if (kind == "const" && p && apply_march_const(t, it, layer, merr, const_values)) continue;
            if (kind == "fn" && p && p->masked_hash == it.info.masked_hash && apply_march_literals(it, layer, merr)) continue;
            // `var<private> sample_colors` with other VALUES (:326-335): the colours the sample kernel writes next to each position
            if (kind == "var" && name == "sample_colors" && p && p->masked_hash == it.info.masked_hash) {
                std::vector<double> n;
                std::vector<std::string> id;
                if (const_value(t, it, n, id) && n.size() == 25) {  // the 8 of the constructor's type, then 8 x 3 components
                    bool ok = true;
                    std::string rows;
                    for (int c = 0; c < 8 && ok; ++c) {
                        rows += c ? ", {" : "{";
                        for (int q = 0; q < 3; ++q) {
                            const double x = n[1 + 3 * c + q];
                            ok = ok && std::isfinite(x) && std::fabs(x) <= 1e3;
                            rows += (q ? ", " : "") + f32_literal(x);
                        }
                        rows += "}";
                    }
                    if (ok) {
                        layer.march_defines += "#define LINON_SAMPLE_COLORS " + rows + "\n";
                        layer.names += (layer.names.empty() ? "" : ",") + name;
                        continue;
                    }
                    merr = "a component of 'sample_colors' is outside the range the CUDA path supports (|component| <= 1000)";
                }
            }
            if (!merr.empty()) {
                unsupported = true;
                error = "main_view.wgsl line " + std::to_string(it.info.line) + ": " + merr;
                return false;
            }
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
    // names a helper may use: every function and constant of the module text (the reference's own and the user's)
    std::set<std::string> module_names;
    for (const RawItem& it : items)
        if (it.info.kind == "fn" || it.info.kind == "const") module_names.insert(it.info.name);
    // ---- constants: one line, re-declared at the top of every emitted function and of the field-function body ----
    for (int k : emit_const) {
        std::string text;
        if (!translate_range(t, items[k].begin, items[k].end, false, module_names, text, err)) {
            error = "WGSL module error: constant '" + items[k].info.name + "': " + err;
            return false;
        }
        layer.preamble += text + " ";
        layer.names += (layer.names.empty() ? "" : ",") + items[k].info.name;
    }
    for (int k : emit_fn) {
        if (!emit_function(t, items[k], layer.preamble, module_names, layer.members, err)) {
            error = "WGSL module error: " + err;
            return false;
        }
        layer.names += (layer.names.empty() ? "" : ",") + items[k].info.name;
    }
    if (shade_item >= 0) {
        if (!emit_shade_block(t, items[shade_item], shade_open, shade_close, layer.preamble, module_names, const_values,
                              layer.members, err)) {
            error = "WGSL module error: " + err;
            return false;
        }
        layer.march_defines += "#define LINON_USER_SHADE 1\n";
        layer.names += (layer.names.empty() ? "" : ",") + std::string("ray_color");
    }
    return true;
}

}  // namespace linon
