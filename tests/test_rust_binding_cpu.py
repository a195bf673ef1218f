"""The Rust `-sys` crate a maintainer adds to the reference is uncompiled source in this image (no rustc), so
its agreement with include/linon_b200.h is checked textually: every constant it declares has the header's value, every option
key of the header is declared, and every function it binds exists in the header with the same number of arguments."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _read(p):
    with open(os.path.join(ROOT, p)) as f:
        return f.read()


def _strip_c_comments(s):
    return re.sub(r"/\*.*?\*/", " ", s, flags=re.S)


def _header_constants(h):
    h = _strip_c_comments(h)
    out = {}
    for m in re.finditer(r"\b(LINON_[A-Z0-9_]+)\s*=\s*(-?\d+)\b", h):
        out[m.group(1)] = int(m.group(2))
    for m in re.finditer(r"#define\s+(LINON_[A-Z0-9_]+)\s+(-?\d+)u?\b", h):
        out[m.group(1)] = int(m.group(2))
    return out


def _header_functions(h):
    h = _strip_c_comments(h)
    out = {}
    for m in re.finditer(r"\b(linon_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", h, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return out


def test_rust_constants_carry_the_headers_values():
    hc = _header_constants(_read("include/linon_b200.h"))
    rs = _read("rust/linon-cuda-sys/src/lib.rs")
    seen = {}
    for m in re.finditer(r"pub const (LINON_[A-Z0-9_]+)\s*:\s*[a-z0-9_]+\s*=\s*(-?\d+)\s*;", rs):
        seen[m.group(1)] = int(m.group(2))
    assert len(seen) > 30
    for name, value in seen.items():
        assert name in hc, f"{name} is not in include/linon_b200.h"
        assert hc[name] == value, f"{name}: header {hc[name]}, rust {value}"
    for name in hc:
        if name.startswith(("LINON_OPT_", "LINON_BUF_", "LINON_ERR_", "LINON_RENDER_")):
            assert name in seen, f"{name} is missing from the Rust binding"


def test_rust_functions_exist_in_the_header_with_the_same_arity():
    hf = _header_functions(_read("include/linon_b200.h"))
    rs = re.sub(r"//[^\n]*", "", _read("rust/linon-cuda-sys/src/lib.rs"))
    bound = {}
    for m in re.finditer(r"pub fn (linon_[a-z0-9_]+)\s*\(([^)]*)\)", rs, flags=re.S):
        args = m.group(2).strip().rstrip(",")
        bound[m.group(1)] = 0 if not args else len(args.split(","))
    assert len(bound) >= 40
    for name, n in bound.items():
        assert name in hf, f"{name} is bound in Rust but not declared in the header"
        assert hf[name] == n, f"{name}: {hf[name]} arguments in the header, {n} in Rust"
    # the path's own entry points (everything but the test hooks) are all bound
    for name in hf:
        if not name.startswith(("linon_debug_", "linon_compile_check", "linon_translate_", "linon_scan_")):
            assert name in bound, f"{name} has no Rust binding"


_C2RS = {"float": "f32", "int32_t": "i32", "uint32_t": "u32", "uint64_t": "u64", "int": "c_int"}


def test_rust_structs_have_the_headers_fields_in_order():
    h = _strip_c_comments(_read("include/linon_b200.h"))
    rs = re.sub(r"//[^\n]*", "", _read("rust/linon-cuda-sys/src/lib.rs"))
    for name in ("linon_camera", "linon_settings", "linon_counters", "linon_timings"):
        body = re.search(r"typedef struct %s\s*\{(.*?)\}\s*%s\s*;" % (name, name), h, flags=re.S).group(1)
        c_fields = []
        for m in re.finditer(r"\b([a-z0-9_]+)\s+([a-z_0-9]+)((?:\[\d+\])*)\s*;", body):
            dims = re.findall(r"\[(\d+)\]", m.group(3))
            count = 1
            for d in dims:
                count *= int(d)
            c_fields.append((m.group(2), _C2RS[m.group(1)], count))
        rbody = re.search(r"pub struct %s\s*\{(.*?)\n\}" % name, rs, flags=re.S).group(1)
        r_fields = []
        for m in re.finditer(r"pub ([a-z_0-9]+)\s*:\s*([^,\n]+),", rbody):  # [[f32; 4]; 4] == float[16]: same bytes
            ty = m.group(2).strip()
            count = 1
            for d in re.findall(r";\s*(\d+)\]", ty):
                count *= int(d)
            r_fields.append((m.group(1), ty.replace("[", "").split(";")[0].strip(), count))
        assert c_fields and c_fields == r_fields, f"{name}: header {c_fields}, rust {r_fields}"
