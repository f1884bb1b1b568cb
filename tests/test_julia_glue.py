"""julia/PNB200.jl cannot be executed here (no julia binary in the image), so it is kept in
lock-step with include/pnb200.h mechanically: every `ccall((:sym, libpnb200), Ret, (ArgTypes...), ...)`
is parsed and its symbol, return type, arity and C types are compared with the C declaration."""
import os
import re

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JL = os.path.join(REPO, "pointneighbors.jl_b200", "julia", "PNB200.jl")
HDR = os.path.join(REPO, "include", "pnb200.h")


def _split_top(s):
    """split at top-level commas (parentheses / braces nest)"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _balanced(text, start):
    """text[start] == '(' -> index one past the matching ')'"""
    depth = 0
    for k in range(start, len(text)):
        if text[k] == "(":
            depth += 1
        elif text[k] == ")":
            depth -= 1
            if depth == 0:
                return k + 1
    raise ValueError("unbalanced")


def julia_ccalls():
    src = open(JL).read()
    src = re.sub(r"#[^\n]*", "", src)
    calls = []
    for m in re.finditer(r"\bccall\(", src):
        end = _balanced(src, m.end() - 1)
        args = _split_top(src[m.end():end - 1])
        sym = re.match(r"\(\s*:(\w+)\s*,\s*libpnb200\s*\)", args[0])
        assert sym, f"unexpected ccall target: {args[0]}"
        ret = args[1]
        tup = args[2].strip()
        assert tup.startswith("(") and tup.endswith(")"), tup
        types = _split_top(tup[1:-1])
        calls.append((sym.group(1), ret, types, len(args) - 3))
    return calls


def header_decls():
    h = open(HDR).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    decls = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(pnb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", h, flags=re.S):
        ret, name, params = m.group(1).strip(), m.group(2), m.group(3).strip()
        plist = [] if params in ("", "void") else [re.sub(r"\s+", " ", p.strip()) for p in params.split(",")]
        decls[name] = (ret, plist)
    return decls


# C parameter -> set of acceptable Julia ccall types
def _julia_types_for(cparam):
    p = cparam.replace("const ", "").strip()
    ptr = p.count("*")
    base = re.sub(r"\*", " ", p).split()
    # drop the parameter name (last token) unless the type is a single token like `void`
    tname = " ".join(base[:-1]) if len(base) > 1 else base[0]
    if ptr >= 2:
        return {"Ref{Ptr{Cvoid}}", "Ptr{Ptr{Cvoid}}"}
    if ptr == 1:
        generic = {"Ptr{Cvoid}"}
        typed = {"float": {"Ptr{Cfloat}", "Ref{Cfloat}"}, "double": {"Ptr{Cdouble}", "Ref{Cdouble}"},
                 "int64_t": {"Ptr{Int64}", "Ref{Int64}"}, "int32_t": {"Ptr{Int32}", "Ref{Int32}", "Ptr{Cint}"},
                 "int": {"Ptr{Cint}", "Ref{Cint}"}, "uint32_t": {"Ptr{UInt32}"}, "uint8_t": {"Ptr{UInt8}"},
                 "pnb_wcsph_params": {"Ref{WcsphParams}", "Ptr{WcsphParams}"},
                 "pnb_tlsph_params": {"Ref{TlsphParams}", "Ptr{TlsphParams}"},
                 "pnb_wcsph_params_f64": {"Ref{WcsphParams64}", "Ptr{WcsphParams64}"},
                 "pnb_slab_arrays": {"Ref{SlabArrays}"},
                 "char": {"Cstring", "Ptr{UInt8}"}}
        return generic | typed.get(tname, set())
    return {"int": {"Cint"}, "float": {"Cfloat"}, "double": {"Cdouble"}, "int64_t": {"Int64"},
            "int32_t": {"Int32", "Cint"}, "pnb_status": {"Cint"}}.get(tname, {tname})


_RET = {"pnb_status": {"Cint"}, "int": {"Cint"}, "void": {"Cvoid"}, "int64_t": {"Int64"},
        "const char *": {"Cstring"}, "const char*": {"Cstring"}}


def test_every_ccall_matches_the_header():
    calls = julia_ccalls()
    decls = header_decls()
    assert len(calls) >= 30 and len(decls) >= 60
    for sym, ret, types, n_values in calls:
        assert sym in decls, f"PNB200.jl calls {sym}, which include/pnb200.h does not declare"
        cret, cparams = decls[sym]
        assert ret in _RET[cret], f"{sym}: Julia return type {ret}, C returns {cret}"
        assert len(types) == len(cparams), \
            f"{sym}: {len(types)} ccall argument types, {len(cparams)} C parameters"
        assert n_values == len(types), f"{sym}: {n_values} values passed for {len(types)} types"
        for k, (jt, cp) in enumerate(zip(types, cparams)):
            assert jt in _julia_types_for(cp), f"{sym} argument {k + 1}: Julia {jt} vs C `{cp}`"


def test_glue_covers_the_hot_path_entry_points():
    """every north_star call of the path has a ccall in the glue"""
    called = {c[0] for c in julia_ccalls()}
    for sym in ("pnb_grid_create_padded_f32", "pnb_grid_build_f32", "pnb_count_neighbors_f32",
                "pnb_nbody_f32", "pnb_wcsph_interact_f32", "pnb_nlist_build_f32",
                "pnb_nlist_export_dvov", "pnb_tlsph_deformation_grad_f32", "pnb_malloc",
                "pnb_memcpy_h2d", "pnb_memcpy_d2h", "pnb_last_error"):
        assert sym in called


def _c_struct_fields(name):
    """[(ctype, field), ...] of `typedef struct name { ... } name;` in the header"""
    h = open(HDR).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    m = re.search(r"typedef\s+struct\s+" + name + r"\s*\{(.*?)\}\s*" + name + r"\s*;", h, flags=re.S)
    assert m, name
    out = []
    for stmt in m.group(1).split(";"):
        stmt = stmt.strip()
        if not stmt:
            continue
        ctype, names = stmt.split(None, 1)
        out += [(ctype, n.strip()) for n in names.split(",")]
    return out


def _julia_struct_fields(name):
    src = open(JL).read()
    m = re.search(r"\bstruct\s+" + name + r"\b[^\n]*\n(.*?)\nend", src, flags=re.S)
    assert m, name
    out = []
    for line in m.group(1).splitlines():
        line = re.sub(r"#.*", "", line).strip()
        if line:
            f, t = [x.strip() for x in line.split("::")]
            out.append((t, f))
    return out


def test_parameter_structs_have_the_same_layout():
    """The by-reference parameter blocks: field order and scalar types of the C struct, the Julia
    struct and the ctypes Structure of the Python mirror agree (a silent mismatch would shift
    every physical parameter)."""
    import ctypes as C
    import sys
    sys.path.insert(0, os.path.join(REPO, "pointneighbors.jl_b200"))
    from pnb200 import _lib
    jl_of = {"float": "Cfloat", "double": "Cdouble", "int32_t": "Int32", "int": "Cint"}
    ct_of = {"float": C.c_float, "double": C.c_double, "int32_t": C.c_int32, "int": C.c_int}
    for cname, jname, pyname in (("pnb_wcsph_params", "WcsphParams", "WcsphParams"),
                                 ("pnb_wcsph_params_f64", "WcsphParams64", "WcsphParams64"),
                                 ("pnb_tlsph_params", "TlsphParams", "TlsphParams")):
        cf = _c_struct_fields(cname)
        jf = _julia_struct_fields(jname)
        assert [f for _, f in cf] == [f for _, f in jf], (cname, cf, jf)
        alias = {"Float32": "Cfloat", "Float64": "Cdouble", "Int32": "Int32", "Cint": "Cint"}
        assert [jl_of[t] for t, _ in cf] == [alias.get(t, t) for t, _ in jf], (cname, cf, jf)
        py = getattr(_lib, pyname)._fields_
        assert [f for _, f in cf] == [f for f, _ in py], (cname, cf, py)
        assert [ct_of[t] for t, _ in cf] == [t for _, t in py], (cname, cf, py)
