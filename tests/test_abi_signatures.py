"""Static cross-check of the three statements of the C ABI: the prototypes in include/rfb200.h, the ctypes table the
Python host layer (and every GPU test) calls through, and the `ccall`s of the Julia shim -- which cannot be executed
here (no Julia runtime), so a wrong argument count or an Int64 passed where the library expects a pointer would
otherwise only show up on a maintainer's machine.  Each parameter is reduced to its kind (pointer / 64-bit integer /
32-bit integer / size_t) and the three lists must agree, position by position."""
import ctypes as C
import os
import re

import rfb200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "rfb200.h")).read()
SHIM = open(os.path.join(ROOT, "recursivefactorization.jl_b200", "julia", "RecursiveFactorizationB200.jl")).read()


def c_kind(decl):
    decl = decl.strip()
    if "*" in decl or "[" in decl:
        return "ptr"
    if re.search(r"\bint64_t\b", decl):
        return "i64"
    if re.search(r"\bsize_t\b", decl):
        return "size"
    if re.search(r"\b(int|int32_t)\b", decl):
        return "i32"
    raise AssertionError(f"unclassified C parameter: {decl!r}")


def header_prototypes():
    code = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    code = re.sub(r"//[^\n]*", "", code)
    protos = {}
    for ret, name, args in re.findall(r"\b(int|const char \*)\s*(rfb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", code, flags=re.S):
        args = " ".join(args.split())
        params = [] if args in ("", "void") else [c_kind(a) for a in args.split(",")]
        protos[name] = ("ptr" if "*" in ret else "i32", params)
    return protos


def ctypes_kind(t):
    if t is C.c_void_p or t is C.c_char_p or hasattr(t, "contents") or (isinstance(t, type) and issubclass(t, C._Pointer)):
        return "ptr"
    return {C.c_int64: "i64", C.c_int: "i32", C.c_int32: "i32", C.c_size_t: "size"}[t]


def test_ctypes_table_matches_the_header_parameter_by_parameter():
    protos = header_prototypes()
    assert len(protos) >= 80 and set(protos) == set(rfb200._lib.SIGNATURES)
    for name, (res, args) in rfb200._lib.SIGNATURES.items():
        want_ret, want = protos[name]
        assert ctypes_kind(res) == want_ret, name
        assert [ctypes_kind(a) for a in args] == want, (name, [ctypes_kind(a) for a in args], want)


def julia_kind(t):
    t = t.strip()
    if t.startswith("Ptr{") or t in ("Cstring",):
        return "ptr"
    return {"Int64": "i64", "Cint": "i32", "Int32": "i32", "Csize_t": "size"}[t]


def split_top(s):
    """split a comma list at nesting depth 0"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return [x.strip() for x in out]


def julia_ccalls():
    """(symbol, return kind, [parameter kinds]) of every ccall in the shim; loop variables of the `for (T, sym, ...) in
    (...)` / `Symbol("rfb_x_", suf)` code generators are resolved for their Float64 instance."""
    code = "\n".join(line.split("#", 1)[0] for line in SHIM.splitlines())
    # variables bound to symbols by the @eval loops: first tuple of each `for (vars) in ((...), (...))` header; a ccall
    # is resolved against the nearest header above it
    headers = []
    for m in re.finditer(r"for \(([A-Za-z_, ]+)\) in \(\(([^()]*)\)", code):
        b = {}
        for v, val in zip(split_top(m.group(1)), split_top(m.group(2))):
            if val.startswith(":rfb_"):
                b[v] = val[1:]
            elif val.startswith('"'):
                b[v] = val.strip('"')
        headers.append((m.start(), b))
    for m in re.finditer(r"([a-z_]+) = QuoteNode\(Symbol\(\"(rfb_[a-z0-9_]+)\", ([a-z]+)\)\)", code):
        b = [h for pos, h in headers if pos < m.start()][-1]
        b[m.group(1)] = m.group(2) + b[m.group(3)]

    def bound_at(pos):
        return [h for p, h in headers if p < pos][-1]

    calls = []
    for m in re.finditer(r"ccall\(\(", code):
        i = m.end()
        depth, j = 1, i
        while depth:                                   # end of the (symbol, library) tuple
            depth += {"(": 1, ")": -1}.get(code[j], 0); j += 1
        sym = split_top(code[i:j - 1])[0]
        mm = re.match(r":(rfb_[a-z0-9_]+)$", sym) or re.match(r"\$\(QuoteNode\(([a-z_]+)\)\)$", sym) or re.match(r"\$([a-z_]+)$", sym)
        assert mm, sym
        name = mm.group(1) if sym.startswith(":") else bound_at(m.start())[mm.group(1)]
        rest = code[j:]
        mret = re.match(r"\s*,\s*([A-Za-z]+)\s*,\s*\(", rest)
        assert mret, (name, rest[:60])
        k = j + mret.end()
        depth, e = 1, k
        while depth:
            depth += {"(": 1, ")": -1}.get(code[e], 0); e += 1
        types = [t.replace("$T", "Float64") for t in split_top(code[k:e - 1])]
        calls.append((name, julia_kind(mret.group(1)), [julia_kind(t) for t in types]))
    return calls


def test_julia_ccalls_match_the_header_parameter_by_parameter():
    protos = header_prototypes()
    calls = julia_ccalls()
    assert len(calls) >= 20
    seen = set()
    for name, ret, kinds in calls:
        assert name in protos, name
        want_ret, want = protos[name]
        assert ret == want_ret, (name, ret, want_ret)
        assert kinds == want, (name, kinds, want)
        seen.add(name)
    # the shim's whole-path, solve, butterfly, batched, multi-GPU and kernel-level bindings were all inspected
    for must in ("rfb_create", "rfb_lu_f64", "rfb_solve_f64", "rfb_butterfly_solve_f64", "rfb_lu_batched_f64", "rfb_mg_lu_f64",
                 "rfb_lu_range_f64", "rfb_laswp_range_f64", "rfb_trsm_llnu_f64", "rfb_gemm_nn_sub_f64", "rfb_solve_kept_f64",
                 "rfb_set_early_download", "rfb_perm_buffers", "rfb_memset"):
        assert must in seen, must
