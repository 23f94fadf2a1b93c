"""julia/AbmEngine.jl against include/abm.h, statically (Julia is not installed here, so the shim never runs in this repository's
tests): the hand-mirrored `AbmConfig` must have the fields of `struct abm_config` in the same order with types of the same size
and add up to `abm_config_size()` of the built library; every `ccall((:sym, libabm), ret, (argtypes...), ...)` must name an
exported function of the header with the same number of arguments, pointer arguments where the prototype has pointers, integer /
floating-point arguments of the prototype's width elsewhere, and the prototype's return type."""
import os
import re

from conftest import ROOT

HDR = open(os.path.join(ROOT, "include", "abm.h")).read()
JL = open(os.path.join(ROOT, "julia", "AbmEngine.jl")).read()

C_SIZES = {"int32_t": 4, "uint32_t": 4, "int64_t": 8, "uint64_t": 8, "double": 8, "float": 4, "int": 4}
JL_SIZES = {"Int32": 4, "UInt32": 4, "Int64": 8, "UInt64": 8, "Float64": 8, "Float32": 4, "Cint": 4}


def _strip_comments(src):
    return re.sub(r"/\*.*?\*/", " ", src, flags=re.S)


def _c_struct_fields():
    body = re.search(r"typedef struct abm_config \{(.*?)\} abm_config;", _strip_comments(HDR), re.S).group(1)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        ty, names = decl.split(None, 1)
        fields += [(n.strip(), ty) for n in names.split(",")]
    return fields


def _jl_struct_fields():
    body = re.search(r"mutable struct AbmConfig\n(.*?)\nend", JL, re.S).group(1)
    body = re.sub(r"#.*", "", body)
    return re.findall(r"(\w+)::(\w+)\s*=", body)


def test_abmconfig_mirrors_struct_abm_config(emu):
    c, j = _c_struct_fields(), _jl_struct_fields()
    assert [n for n, _ in c] == [n for n, _ in j]
    for (n, ct), (_, jt) in zip(c, j):
        assert C_SIZES[ct] == JL_SIZES[jt], f"{n}: {ct} vs {jt}"
        assert (ct in ("double", "float")) == jt.startswith("Float"), n
    # natural alignment, no padding needed in this layout: offsets of the 8-byte fields are multiples of 8
    off = 0
    for n, ct in c:
        assert off % C_SIZES[ct] == 0, f"{n} would be padded"
        off += C_SIZES[ct]
    assert off == emu.fn["config_size"]() == 144  # sizeof(abm_config) of the built library; the shim asserts the same at load time


def _c_prototypes():
    protos = {}
    for m in re.finditer(r"^\s*([\w\s\*]+?)\b(abm_\w+)\s*\(([^;{]*?)\)\s*;", _strip_comments(HDR), re.M):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        argl = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        protos[name] = (ret, argl)
    return protos


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _jl_ccalls():
    calls = []
    for m in re.finditer(r"ccall\(\(:(\w+), libabm\),\s*(\w+),\s*\(", JL):
        start, depth, i = m.end(), 1, m.end()
        while depth:
            depth += {"(": 1, ")": -1}.get(JL[i], 0)
            i += 1
        calls.append((m.group(1), m.group(2), _split_top(JL[start:i - 1])))
    return calls


def _is_pointer_c(arg):
    return "*" in arg


def _is_pointer_jl(ty):
    return ty.startswith(("Ptr{", "Ref{")) or ty == "Cstring"


def test_every_ccall_matches_its_prototype():
    protos, calls = _c_prototypes(), _jl_ccalls()
    assert len(calls) >= 30 and len(protos) >= 40
    for name, ret, args in calls:
        assert name in protos, f"{name} is not declared in include/abm.h"
        cret, cargs = protos[name]
        assert len(args) == len(cargs), f"{name}: {len(args)} ccall arguments, prototype has {len(cargs)}"
        if "*" in cret:
            assert ret in ("Cstring", "Ptr{Cvoid}"), f"{name} returns a pointer"
        elif cret == "void":
            assert ret == "Cvoid", name
        else:
            assert JL_SIZES.get(ret) == C_SIZES[cret.split()[-1]], f"{name}: return {ret} vs {cret}"
        for k, (ja, ca) in enumerate(zip(args, cargs)):
            if _is_pointer_c(ca):
                assert _is_pointer_jl(ja), f"{name} argument {k}: {ca} vs {ja}"
            else:
                cty = [t for t in ca.replace("const", "").split() if t in C_SIZES][0]
                assert not _is_pointer_jl(ja) and JL_SIZES[ja] == C_SIZES[cty], f"{name} argument {k}: {ca} vs {ja}"
                assert (cty in ("double", "float")) == ja.startswith("Float"), f"{name} argument {k}"
