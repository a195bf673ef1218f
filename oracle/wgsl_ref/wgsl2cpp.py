"""wgsl2cpp.py — mechanical WGSL -> C++17 transpiler for the reference's shader modules.

TEST INFRASTRUCTURE ONLY (oracle/README.md).  Never imported by linon_b200/ (the product has its own C++ translator,
linon_b200/csrc/host/wgsl_translate.cpp, which emits CUDA for field-function bodies only; the two share no code, so
each checks the other).

Purpose: the reference (niklaskorz/linon) cannot be built or run in this image (no rustc/cargo/Vulkan),
but its hot path IS its shader text.  This front end parses a whole WGSL module (tokeniser + recursive-descent parser,
no regex rewriting) and re-emits every declaration, statement and expression as C++ against oracle/wgsl_ref/
wgsl_runtime.hpp — same control flow, same expression tree (operator precedence and left-associativity of WGSL),
same literals.  Typing rules that matter for bit-exact fp32 are implemented from the WGSL specification:
  * a literal without suffix is an AbstractFloat / AbstractInt; an expression made only of those is evaluated in
    f64 / i64 (as naga's constant evaluator does) and converted to f32 / i32 where it meets a concrete operand, a
    typed declaration, a `let`/`var` without type, a function argument or a constructor with an explicit element type;
  * everything else is concrete and evaluated in its own type, one rounding per operation (-ffp-contract=off).
The generated C++ is written under oracle/_ref/ only (git-ignored); reference sources are never copied into the repo.
"""
from __future__ import annotations

import re

CPP_KEYWORDS = set("""alignas alignof and and_eq asm auto bitand bitor bool break case catch char char8_t char16_t
char32_t class compl concept const consteval constexpr constinit const_cast continue co_await co_return co_yield
decltype default delete do double dynamic_cast else enum explicit export extern false float for friend goto if inline
int long mutable namespace new noexcept not not_eq nullptr operator or or_eq private protected public register
reinterpret_cast requires return short signed sizeof static static_assert static_cast struct switch template this
thread_local throw true try typedef typeid typename union unsigned using virtual void volatile wchar_t while xor
xor_eq wgsl main errno stdin stdout stderr NULL""".split())
WGSL_OWN_KEYWORDS = {"true", "false", "if", "else", "for", "while", "return", "break", "continue", "struct", "const",
                     "switch", "case", "default"}

# builtin functions provided by wgsl_runtime.hpp (called as wgsl::name unless the module defines a function of that name)
BUILTINS = set("""abs sqrt exp exp2 log log2 sin cos tan asin acos atan atan2 sinh cosh tanh pow floor ceil round trunc
fract sign inverseSqrt min max clamp saturate mix step smoothstep fma select degrees radians any all dot cross length
distance normalize reflect transpose arrayLength textureDimensions textureStore textureLoad""".split())
# builtins that are constant-evaluated in abstract float when every argument is abstract
ABSTRACT_MATH = set("abs sqrt exp exp2 log log2 sin cos tan pow floor min max clamp".split())

SCALARS = {"f32": "float", "i32": "int32_t", "u32": "uint32_t", "bool": "bool", "f16": "float"}
SUFFIX_TYPES = {"f": "f32", "i": "i32", "u": "u32", "h": "f16"}

TOKEN_RE = re.compile(r"""
    (?P<ws>\s+)
  | (?P<lc>//[^\n]*)
  | (?P<hex>0[xX][0-9a-fA-F]+[iu]?)
  | (?P<flt>(?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?[fh]?|\d+[eE][+-]?\d+[fh]?|\d+[fh])
  | (?P<int>\d+[iu]?)
  | (?P<id>[A-Za-z_][A-Za-z0-9_]*)
  | (?P<op>->|&&|\|\||==|!=|<=|>=|<<=|>>=|<<|>>|\+=|-=|\*=|/=|%=|&=|\|=|\^=|\+\+|--|[-+*/%&|^~!<>=.,;:(){}\[\]@])
""", re.X | re.S)


class WgslError(Exception):
    pass


class Tok:
    __slots__ = ("kind", "text", "line")

    def __init__(self, kind, text, line):
        self.kind, self.text, self.line = kind, text, line

    def __repr__(self):
        return f"{self.kind}:{self.text!r}@{self.line}"


def tokenize(src: str):
    toks, pos, line = [], 0, 1
    while pos < len(src):
        if src.startswith("/*", pos):  # block comments nest in WGSL
            depth, j = 1, pos + 2
            while depth and j < len(src):
                if src.startswith("/*", j):
                    depth, j = depth + 1, j + 2
                elif src.startswith("*/", j):
                    depth, j = depth - 1, j + 2
                else:
                    j += 1
            if depth:
                raise WgslError(f"line {line}: block comment is never closed")
            line += src.count("\n", pos, j)
            pos = j
            continue
        m = TOKEN_RE.match(src, pos)
        if not m:
            raise WgslError(f"line {line}: cannot tokenise {src[pos:pos + 20]!r}")
        kind = m.lastgroup
        text = m.group()
        if kind not in ("ws", "lc"):
            toks.append(Tok(kind, text, line))
        line += text.count("\n")
        pos = m.end()
    toks.append(Tok("eof", "", line))
    return toks


def ident(name: str) -> str:
    return name + "_" if name in CPP_KEYWORDS and name not in WGSL_OWN_KEYWORDS else name


# ---- types ---------------------------------------------------------------------------------------------------------
class Type:
    """Parsed WGSL type: name + template arguments (types or integer texts)."""

    def __init__(self, name, args=()):
        self.name, self.args = name, list(args)

    def cpp(self, structs) -> str:
        n, a = self.name, self.args
        if n in SCALARS:
            return SCALARS[n]
        m = re.fullmatch(r"vec([234])([fiuh]?)", n)
        if m:
            if m.group(2):
                return f"wgsl::vec<{SCALARS[SUFFIX_TYPES[m.group(2)]]}, {m.group(1)}>"
            if not a:
                raise WgslError(f"{n} without element type is only supported as a constructor")
            return f"wgsl::vec<{a[0].cpp(structs)}, {m.group(1)}>"
        m = re.fullmatch(r"mat([234])x([234])([fh]?)", n)
        if m:
            return f"wgsl::mat<{m.group(1)}, {m.group(2)}>"
        if n == "array":
            if len(a) == 2:
                return f"wgsl::array<{a[0].cpp(structs)}, {a[1]}>"
            return f"wgsl::rt_array<{a[0].cpp(structs)}>"
        if n == "atomic":
            return a[0].cpp(structs)
        if n == "texture_storage_2d":
            return f"wgsl::texture_storage_2d<wgsl::{a[0].name}, wgsl::{a[1].name}>"
        if n == "texture_2d":
            return f"wgsl::texture_2d<{a[0].cpp(structs)}>"
        if n in structs:
            return ident(n)
        raise WgslError(f"unsupported type {n}")

    def elem_scalar(self):
        """Element scalar type of a vecN<T> / vecNf / scalar type, as a C++ type name (None if not scalar-like)."""
        if self.name in SCALARS:
            return SCALARS[self.name]
        m = re.fullmatch(r"vec([234])([fiuh]?)", self.name)
        if m:
            if m.group(2):
                return SCALARS[SUFFIX_TYPES[m.group(2)]]
            if self.args:
                return self.args[0].elem_scalar()
        if re.fullmatch(r"mat[234]x[234][fh]?", self.name):
            return "float"
        return None


TYPE_GENERATORS = re.compile(r"vec[234]|mat[234]x[234]|array|atomic|ptr|texture_\w+|bitcast")


# ---- parser ----------------------------------------------------------------------------------------------------------
class Parser:
    def __init__(self, src: str):
        self.toks = tokenize(src)
        self.i = 0

    # token helpers
    @property
    def t(self):
        return self.toks[self.i]

    def peek(self, k=1):
        return self.toks[min(self.i + k, len(self.toks) - 1)]

    def at(self, text):
        return self.t.text == text and self.t.kind in ("op", "id")

    def accept(self, text):
        if self.at(text):
            self.i += 1
            return True
        return False

    def expect(self, text):
        if not self.accept(text):
            raise WgslError(f"line {self.t.line}: expected {text!r}, found {self.t.text!r}")

    def name(self):
        if self.t.kind != "id":
            raise WgslError(f"line {self.t.line}: expected identifier, found {self.t.text!r}")
        self.i += 1
        return self.toks[self.i - 1].text

    def split_close(self):
        """Consumes one '>' of a template list, splitting '>>' / '>=' / '>>=' tokens."""
        t = self.t
        if t.kind == "op" and t.text.startswith(">"):
            if len(t.text) == 1:
                self.i += 1
            else:
                self.toks[self.i] = Tok("op", t.text[1:], t.line)
            return
        raise WgslError(f"line {t.line}: expected '>', found {t.text!r}")

    def attributes(self):
        attrs = {}
        while self.accept("@"):
            n = self.name()
            args = []
            if self.accept("("):
                while not self.at(")"):
                    args.append(self.t.text)
                    self.i += 1
                    self.accept(",")
                self.expect(")")
            attrs[n] = args
        return attrs

    def type(self) -> Type:
        n = self.name()
        args = []
        if self.at("<"):
            self.i += 1
            while True:
                if self.t.kind in ("int", "hex"):
                    args.append(self.t.text.rstrip("iu"))
                    self.i += 1
                else:
                    args.append(self.type())
                if self.accept(","):
                    if self.t.text.startswith(">"):
                        break
                    continue
                break
            self.split_close()
        return Type(n, args)

    # ---- expressions (WGSL precedence: unary > multiplicative > additive > shift > relational > & ^ | > && > ||) ----
    def expr(self):
        return self.binary(0)

    LEVELS = [("||",), ("&&",), ("|",), ("^",), ("&",), ("==", "!=", "<", ">", "<=", ">="), ("<<", ">>"), ("+", "-"),
              ("*", "/", "%")]

    def binary(self, lvl):
        if lvl == len(self.LEVELS):
            return self.unary()
        lhs = self.binary(lvl + 1)
        while self.t.kind == "op" and self.t.text in self.LEVELS[lvl]:
            op = self.t.text
            self.i += 1
            rhs = self.binary(lvl + 1)
            lhs = ("bin", op, lhs, rhs)
        return lhs

    def unary(self):
        if self.t.kind == "op" and self.t.text in ("-", "!", "~", "&", "*"):
            op = self.t.text
            self.i += 1
            return ("un", op, self.unary())
        return self.postfix()

    def call_args(self):
        args = []
        self.expect("(")
        while not self.at(")"):
            args.append(self.expr())
            if not self.accept(","):
                break
        self.expect(")")
        return args

    def primary(self):
        t = self.t
        if t.kind in ("flt", "int", "hex"):
            self.i += 1
            return ("num", t.kind, t.text)
        if t.kind == "id":
            if t.text in ("true", "false"):
                self.i += 1
                return ("bool", t.text)
            # templated constructor / bitcast:  vec3<f32>(...)   array<vec3<f32>, 3>(...)
            if TYPE_GENERATORS.fullmatch(t.text) and self.peek().text == "<":
                ty = self.type()
                return ("ctor", ty, self.call_args())
            self.i += 1
            if self.at("("):
                return ("call", t.text, self.call_args())
            return ("id", t.text)
        if self.accept("("):
            e = self.expr()
            self.expect(")")
            return ("paren", e)
        raise WgslError(f"line {t.line}: unexpected {t.text!r} in expression")

    def postfix(self):
        e = self.primary()
        while True:
            if self.accept("."):
                e = ("member", e, self.name())
            elif self.accept("["):
                idx = self.expr()
                self.expect("]")
                e = ("index", e, idx)
            else:
                return e

    # ---- statements ----------------------------------------------------------------------------------------------
    def block(self):
        self.expect("{")
        body = []
        while not self.at("}"):
            body.append(self.statement())
        self.expect("}")
        return ("block", body)

    def var_decl(self):
        """let/var/const declaration without the trailing ';'."""
        kw = self.name()
        if kw == "var" and self.at("<"):
            self.i += 1
            while not self.t.text.startswith(">"):
                self.i += 1
            self.split_close()
        n = self.name()
        ty = self.type() if self.accept(":") else None
        init = self.expr() if self.accept("=") else None
        return ("decl", kw, n, ty, init, self.toks[self.i - 1].line)

    def simple_statement(self):
        """Assignment, increment, call or declaration (no trailing ';'): the forms allowed in a for header."""
        if self.t.kind == "id" and self.t.text in ("let", "var", "const"):
            return self.var_decl()
        if self.at("_"):
            self.i += 1
            self.expect("=")
            return ("phony", self.expr())
        lhs = self.unary()
        if self.t.kind == "op" and self.t.text in ("=", "+=", "-=", "*=", "/=", "%=", "&=", "|=", "^=", "<<=", ">>="):
            op = self.t.text
            self.i += 1
            return ("assign", op, lhs, self.expr())
        if self.t.kind == "op" and self.t.text in ("++", "--"):
            op = self.t.text
            self.i += 1
            return ("incr", op, lhs)
        if lhs[0] != "call":
            raise WgslError(f"line {self.t.line}: expression statement is not a call")
        return ("expr", lhs)

    def statement(self):
        t = self.t
        if self.at("{"):
            return self.block()
        if self.accept(";"):
            return ("empty",)
        if t.kind == "id":
            if t.text == "if":
                self.i += 1
                cond = self.expr()
                then = self.block()
                other = None
                if self.accept("else"):
                    other = self.statement() if self.at("if") else self.block()
                return ("if", cond, then, other)
            if t.text == "for":
                self.i += 1
                self.expect("(")
                init = None if self.at(";") else self.simple_statement()
                self.expect(";")
                cond = None if self.at(";") else self.expr()
                self.expect(";")
                upd = None if self.at(")") else self.simple_statement()
                self.expect(")")
                return ("for", init, cond, upd, self.block())
            if t.text == "while":
                self.i += 1
                cond = self.expr()
                return ("while", cond, self.block())
            if t.text == "loop":
                # loop { statements [continuing { statements [break if e;] }] }   (WGSL spec 9.4.2, 9.4.5)
                self.i += 1
                self.expect("{")
                body, cont, break_if = [], None, None
                while not self.at("}"):
                    if self.t.kind == "id" and self.t.text == "continuing":
                        self.i += 1
                        self.expect("{")
                        cont = []
                        while not self.at("}"):
                            if self.t.kind == "id" and self.t.text == "break" and self.peek().text == "if":
                                self.i += 2
                                break_if = self.expr()
                                self.expect(";")
                                if not self.at("}"):
                                    raise WgslError(f"line {self.t.line}: 'break if' must end the continuing block")
                                break
                            cont.append(self.statement())
                        self.expect("}")
                        if not self.at("}"):
                            raise WgslError(f"line {self.t.line}: 'continuing' must be the last statement of its loop")
                        break
                    body.append(self.statement())
                self.expect("}")
                if cont is None:
                    return ("while", ("bool", "true"), ("block", body))
                return ("loop", ("block", body), ("block", cont), break_if, t.line)
            if t.text == "switch":
                # switch e { case s, s [:] { } default [:] { } }   (WGSL spec 9.4.1: no fall-through)
                self.i += 1
                sel = self.expr()
                self.expect("{")
                clauses = []
                while not self.at("}"):
                    labels = []
                    if self.accept("default"):
                        labels.append(None)
                    else:
                        self.expect("case")
                        while True:
                            labels.append(None if self.accept("default") else self.expr())
                            if not self.accept(","):
                                break
                            if self.at(":") or self.at("{"):
                                break
                    self.accept(":")
                    clauses.append((labels, self.block()))
                self.expect("}")
                return ("switch", sel, clauses)
            if t.text in ("continuing", "discard"):
                raise WgslError(f"line {t.line}: '{t.text}' is not supported by the reference transpiler")
            if t.text == "return":
                self.i += 1
                e = None if self.at(";") else self.expr()
                self.expect(";")
                return ("return", e)
            if t.text in ("break", "continue"):
                self.i += 1
                self.expect(";")
                return (t.text,)
        s = self.simple_statement()
        self.expect(";")
        return s

    # ---- module ----------------------------------------------------------------------------------------------------
    def module(self):
        decls = []
        while self.t.kind != "eof":
            if self.accept(";"):
                continue
            attrs = self.attributes()
            t = self.t
            if t.text == "struct":
                self.i += 1
                n = self.name()
                self.expect("{")
                fields = []
                while not self.at("}"):
                    self.attributes()
                    fn = self.name()
                    self.expect(":")
                    fields.append((fn, self.type()))
                    if not self.accept(","):
                        break
                self.expect("}")
                decls.append(("struct", n, fields))
            elif t.text in ("var", "const", "let", "override"):
                if t.text == "var" and self.peek().text == "<":
                    j = self.i + 2
                    space = self.toks[j].text
                else:
                    space = "handle" if t.text == "var" else t.text
                d = self.var_decl()
                self.expect(";")
                decls.append(("global", space, attrs) + d[1:])
            elif t.text == "fn":
                self.i += 1
                n = self.name()
                self.expect("(")
                params = []
                while not self.at(")"):
                    pattrs = self.attributes()
                    pn = self.name()
                    self.expect(":")
                    params.append((pn, self.type(), pattrs))
                    if not self.accept(","):
                        break
                self.expect(")")
                ret = None
                if self.accept("->"):
                    self.attributes()
                    ret = self.type()
                decls.append(("fn", n, params, ret, self.block(), attrs))
            elif t.text in ("alias", "enable", "requires", "diagnostic", "const_assert"):
                raise WgslError(f"line {t.line}: '{t.text}' is not supported by the reference transpiler")
            else:
                raise WgslError(f"line {t.line}: unexpected {t.text!r} at module scope")
        return decls


# ---- emitter -----------------------------------------------------------------------------------------------------------
AF, AI, CC = "AF", "AI", "C"  # abstract float, abstract int, concrete


class Emitter:
    def __init__(self, decls):
        self.decls = decls
        self.structs = {d[1]: d[2] for d in decls if d[0] == "struct"}
        self.functions = {d[1]: d for d in decls if d[0] == "fn"}
        self.field_names = {f for fields in self.structs.values() for f, _ in fields}
        self.abstract_consts = {}  # module-scope const without type and abstract initialiser -> kind
        self.entry_points = {}

    # ---- expressions: returns (text, kind) ------------------------------------------------------------------------
    def conc(self, e, want=None):
        """C++ text of expression e converted to a concrete value (AbstractFloat -> f32, AbstractInt stays an int
        literal unless `want` names the C++ scalar type it must take)."""
        text, kind = self.expr(e)
        if kind == AF:
            return f"{want or 'float'}({text})"
        if kind == AI and want:
            return f"{want}({text})"
        return text

    def number(self, kind, text):
        if kind == "flt":
            if text[-1] in "fh":
                body = text[:-1]
                if not re.search(r"[.eE]", body):
                    body += ".0"
                return f"{body}f", CC
            return (text if re.search(r"[.eE]", text) else text + ".0"), AF
        if text[-1] == "u":
            return text, CC
        if text[-1] == "i":
            return f"int32_t({text[:-1]})", CC
        return text, AI

    def expr(self, e):
        k = e[0]
        if k == "num":
            return self.number(e[1], e[2])
        if k == "bool":
            return e[1], CC
        if k == "id":
            if e[1] in self.abstract_consts:
                return ident(e[1]), self.abstract_consts[e[1]]
            return ident(e[1]), CC
        if k == "paren":
            text, kind = self.expr(e[1])
            return f"({text})", kind
        if k == "un":
            text, kind = self.expr(e[2])
            if e[1] == "&":
                return f"(&{text})", CC
            if e[1] == "*":
                return f"(*{text})", CC
            return f"({e[1]}{text})", kind
        if k == "bin":
            op = e[1]
            (lt, lk), (rt, rk) = self.expr(e[2]), self.expr(e[3])
            if lk != CC and rk != CC:  # constant evaluation in the abstract types
                kind = AF if AF in (lk, rk) else AI
                if op in ("==", "!=", "<", ">", "<=", ">=", "&&", "||"):
                    kind = CC
                if kind == AF:  # abstract-int operand of an abstract-float operation
                    lt = f"double({lt})" if lk == AI else lt
                    rt = f"double({rt})" if rk == AI else rt
                if op == "%":
                    return f"wgsl::mod({lt}, {rt})", kind
                return f"({lt} {op} {rt})", kind
            if lk == AF:
                lt = f"float({lt})"
            if rk == AF:
                rt = f"float({rt})"
            if op == "%":  # C++ has no % on floats: wgsl::mod dispatches on the operand types
                return f"wgsl::mod({lt}, {rt})", CC
            return f"({lt} {op} {rt})", CC
        if k == "member":
            text, _ = self.expr(e[1])
            n = e[2]
            if n not in self.field_names and (re.fullmatch(r"[xyzw]{2,4}", n) or re.fullmatch(r"[rgba]{2,4}", n)):
                idx = ", ".join(str("xyzw".index(c) if c in "xyzw" else "rgba".index(c)) for c in n)
                return f"wgsl::swz<{idx}>({text})", CC
            return f"{text}.{ident(n)}", CC
        if k == "index":
            text, _ = self.expr(e[1])
            return f"{text}[{self.conc(e[2])}]", CC
        if k == "ctor":
            return self.constructor(e[1], e[2]), CC
        if k == "call":
            return self.call(e[1], e[2])
        raise WgslError(f"unhandled expression node {k}")

    def constructor(self, ty: Type, args):
        cpp = ty.cpp(self.structs)
        if ty.name == "array":
            el = ty.args[0].elem_scalar() if ty.args[0].name in SCALARS else None
            return f"{cpp}({', '.join(self.conc(a, el) for a in args)})"
        el = ty.elem_scalar()
        return f"{cpp}({', '.join(self.conc(a, el) for a in args)})"

    def call(self, name, args):
        # scalar conversions f32(x) i32(x) u32(x) bool(x)
        if name in SCALARS and name not in self.functions:
            return f"{SCALARS[name]}({self.conc(args[0])})", CC
        # vec3f(...) / vec3(...) / mat3x3f(...): constructors without template list
        m = re.fullmatch(r"vec([234])([fiuh]?)", name)
        if m and name not in self.functions:
            if m.group(2):
                el = SCALARS[SUFFIX_TYPES[m.group(2)]]
            else:  # element type inferred from the arguments: abstract ints only -> i32, otherwise f32
                kinds = [self.expr(a)[1] for a in args]
                el = "int32_t" if kinds and all(kk == AI for kk in kinds) else "float"
            return f"wgsl::vec<{el}, {m.group(1)}>({', '.join(self.conc(a, el) for a in args)})", CC
        m = re.fullmatch(r"mat([234])x([234])([fh]?)", name)
        if m and name not in self.functions:
            return f"wgsl::mat<{m.group(1)}, {m.group(2)}>({', '.join(self.conc(a, 'float') for a in args)})", CC
        if name in self.structs:
            fields = self.structs[name]
            parts = [self.conc(a, fields[i][1].elem_scalar() if fields[i][1].name in SCALARS else None)
                     for i, a in enumerate(args)]
            return f"{ident(name)}{{{', '.join(parts)}}}", CC
        if name in self.functions:
            params = self.functions[name][2]
            parts = [self.conc(a, params[i][1].elem_scalar() if params[i][1].name in SCALARS else None)
                     for i, a in enumerate(args)]
            return f"{ident(name)}({', '.join(parts)})", CC
        if name in BUILTINS:
            pairs = [self.expr(a) for a in args]
            if name in ABSTRACT_MATH and pairs and all(kk != CC for _, kk in pairs):
                if all(kk == AI for _, kk in pairs) and name in ("abs", "min", "max", "clamp"):
                    return f"wgsl::{name}({', '.join(f'(long long)({t})' for t, _ in pairs)})", AI
                return f"wgsl::{name}({', '.join(f'double({t})' for t, _ in pairs)})", AF
            parts = [f"float({t})" if kk == AF else t for t, kk in pairs]
            return f"wgsl::{name}({', '.join(parts)})", CC
        raise WgslError(f"call of unknown function {name!r}")

    # ---- statements ---------------------------------------------------------------------------------------------------
    def decl(self, d, module_scope=False, space=None):
        _, kw, n, ty, init = d[:5]
        prefix = ""
        if module_scope:
            prefix = "static thread_local " if space == "private" else "static "
        if kw in ("let", "const", "override"):
            if ty is None:
                text, kind = self.expr(init)
                if kind == AF:
                    if module_scope:  # stays abstract (WGSL const without type)
                        self.abstract_consts[n] = AF
                        return f"{prefix}const double {ident(n)} = {text};"
                    return f"{prefix}const float {ident(n)} = float({text});"
                if kind == AI:
                    if module_scope:
                        self.abstract_consts[n] = AI
                        return f"{prefix}const long long {ident(n)} = {text};"
                    return f"{prefix}const int32_t {ident(n)} = {text};"
                return f"{prefix}const auto {ident(n)} = {text};"
            return f"{prefix}const {ty.cpp(self.structs)} {ident(n)} = {self.conc(init, ty.elem_scalar() if ty.name in SCALARS else None)};"
        # var
        if ty is None:
            text, kind = self.expr(init)
            if kind == AF:
                return f"{prefix}float {ident(n)} = float({text});"
            if kind == AI:
                return f"{prefix}int32_t {ident(n)} = {text};"
            return f"{prefix}auto {ident(n)} = {text};"
        if init is None:
            return f"{prefix}{ty.cpp(self.structs)} {ident(n)}{{}};"
        return f"{prefix}{ty.cpp(self.structs)} {ident(n)} = {self.conc(init, ty.elem_scalar() if ty.name in SCALARS else None)};"

    def simple(self, s):
        k = s[0]
        if k == "decl":
            return self.decl(s).rstrip(";")
        if k == "phony":
            return f"(void)({self.conc(s[1])})"
        if k == "assign":
            lhs, _ = self.expr(s[2])
            if s[1] != "=":  # WGSL defines `a op= b` as `a = a op b`
                text, _ = self.expr(("bin", s[1][:-1], s[2], s[3]))
                return f"{lhs} = {text}"
            return f"{lhs} = {self.conc(s[3])}"
        if k == "incr":
            lhs, _ = self.expr(s[2])
            return f"{lhs}{s[1]}"
        if k == "expr":
            return self.expr(s[1])[0]
        raise WgslError(f"unhandled simple statement {k}")

    def stmt(self, s, out, ind):
        pad = "    " * ind
        k = s[0]
        if k == "block":
            out.append(pad + "{")
            for x in s[1]:
                self.stmt(x, out, ind + 1)
            out.append(pad + "}")
        elif k == "empty":
            out.append(pad + ";")
        elif k in ("decl", "phony", "assign", "incr", "expr"):
            out.append(pad + self.simple(s) + ";")
        elif k == "if":
            out.append(f"{pad}if ({self.conc(s[1])})")
            self.stmt(s[2], out, ind)
            if s[3] is not None:
                out.append(pad + "else")
                self.stmt(s[3], out, ind)
        elif k == "for":
            init = self.simple(s[1]) if s[1] is not None else ""
            cond = self.conc(s[2]) if s[2] is not None else ""
            upd = self.simple(s[3]) if s[3] is not None else ""
            out.append(f"{pad}for ({init}; {cond}; {upd})")
            self.stmt(s[4], out, ind)
        elif k == "while":
            out.append(f"{pad}while ({self.conc(s[1])})")
            self.stmt(s[2], out, ind)
        elif k == "return":
            out.append(pad + ("return;" if s[1] is None else f"return {self.conc(s[1])};"))
        elif k in ("break", "continue"):
            out.append(pad + k + ";")
        elif k == "loop_break":  # a `break` of a loop whose body is wrapped for its continuing block (see "loop")
            out.append(f"{pad}{{ {s[1]} = true; break; }}")
        elif k == "loop":
            # for (;;) { bool brk = false; do { body } while (false); if (brk) break; continuing; if (break_if) break; }
            # `continue` inside the do-while leaves it and so reaches the continuing block; a `break` of THIS loop sets brk.
            flag = f"wgsl_loop_break_{s[4]}_{len(out)}"
            out.append(f"{pad}for (;;)")
            out.append(pad + "{")
            out.append(f"{pad}    bool {flag} = false;")
            out.append(f"{pad}    do")
            self.stmt(self.rebind_breaks(s[1], flag), out, ind + 1)
            out.append(f"{pad}    while (false);")
            out.append(f"{pad}    if ({flag}) break;")
            self.stmt(s[2], out, ind + 1)
            if s[3] is not None:
                out.append(f"{pad}    if ({self.conc(s[3])}) break;")
            out.append(pad + "}")
        elif k == "switch":
            out.append(f"{pad}switch ({self.conc(s[1])})")
            out.append(pad + "{")
            for labels, body in s[2]:
                for lab in labels:
                    out.append(f"{pad}    " + ("default:" if lab is None else f"case {self.conc(lab)}:"))
                self.stmt(body, out, ind + 2)
                out.append(f"{pad}        break;")
            out.append(pad + "}")
        else:
            raise WgslError(f"unhandled statement {k}")

    def rebind_breaks(self, s, flag):
        """`break` statements that belong to the enclosing loop (not to a nested loop or switch) -> ("loop_break", flag)."""
        k = s[0]
        if k == "break":
            return ("loop_break", flag)
        if k == "block":
            return ("block", [self.rebind_breaks(x, flag) for x in s[1]])
        if k == "if":
            return ("if", s[1], self.rebind_breaks(s[2], flag), None if s[3] is None else self.rebind_breaks(s[3], flag))
        return s  # for / while / loop / switch bind their own breaks; everything else has none

    # ---- module ------------------------------------------------------------------------------------------------------------
    def signature(self, d):
        _, n, params, ret, _, _ = d
        ps = ", ".join(f"{ty.cpp(self.structs)} {ident(pn)}" for pn, ty, _ in params)
        return f"static {ret.cpp(self.structs) if ret else 'void'} {ident(n)}({ps})"

    def emit(self) -> str:
        # WGSL module scope is order-independent: types first, then module variables / constants, then every function
        # prototype, then the function bodies - each group in source order.
        out = []
        for d in self.decls:
            if d[0] == "struct":
                out.append(f"struct {ident(d[1])} {{")
                for fn, ty in d[2]:
                    out.append(f"    {ty.cpp(self.structs)} {ident(fn)}{{}};")
                out.append("};")
        for d in self.decls:
            if d[0] == "global":
                out.append(self.decl(("decl",) + d[3:], module_scope=True, space=d[1]))
        for f in self.functions.values():
            out.append(self.signature(f) + ";")
        for d in self.decls:
            if d[0] == "fn":
                attrs = d[5]
                if "workgroup_size" in attrs:
                    ws = (attrs["workgroup_size"] + ["1", "1", "1"])[:3]
                    self.entry_points[d[1]] = ws
                    out.append(f"static const int {ident(d[1])}_workgroup_size[3] = {{{', '.join(ws)}}};")
                out.append(self.signature(d))
                self.stmt(d[4], out, 0)
        return "\n".join(out) + "\n"


def transpile(src: str) -> str:
    """WGSL module text -> C++ declarations (to be placed inside a namespace, after #include "wgsl_runtime.hpp")."""
    return Emitter(Parser(src).module()).emit()


if __name__ == "__main__":
    import sys
    sys.stdout.write(transpile(open(sys.argv[1]).read()))
