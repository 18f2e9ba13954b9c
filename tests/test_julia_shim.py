"""Static checks of julia/ITensorQTTB200.jl against include/qtt_b200.h (Julia is not installed in this image, so the shim
cannot be executed here): every `ccall` names an exported function and passes the right number of arguments with
ABI-compatible types, the two struct mirrors have the header's field order and types, and the DFT entry points follow the
reference's reversal conventions (src/dft.jl:121-133)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = open(os.path.join(ROOT, "include", "qtt_b200.h")).read()
JL = open(os.path.join(ROOT, "julia", "ITensorQTTB200.jl")).read()


def strip_comments(c):
    return re.sub(r"/\*.*?\*/", " ", c, flags=re.S)


def c_prototypes():
    protos = {}
    for m in re.finditer(r"QTT_API\s+([\w\s\*]+?)\s*\b(qtt_\w+)\s*\(([^;]*?)\)\s*;", strip_comments(HDR), flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        arglist = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        protos[name] = (ret, arglist)
    return protos


def c_kind(decl):
    """ABI class of a C parameter / return declaration."""
    d = decl.strip()
    if "*" in d:
        return "ptr"
    base = re.sub(r"\b(const|unsigned)\b", "", d).split()
    t = base[0]
    return {"int": "i32", "int32_t": "i32", "int64_t": "i64", "double": "f64", "float": "f32", "void": "void",
            "size_t": "i64"}[t]


def jl_kind(t):
    t = t.strip()
    if t.startswith(("Ptr{", "Ref{")) or t in ("Cstring",):
        return "ptr"
    return {"Cint": "i32", "Int32": "i32", "Int64": "i64", "Clonglong": "i64", "Cdouble": "f64", "Float64": "f64",
            "Cfloat": "f32", "Cvoid": "void", "Csize_t": "i64"}[t]


def split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        if ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def jl_ccalls():
    calls = []
    for m in re.finditer(r"ccall\(\(:(qtt_\w+),\s*libqtt\),\s*", JL):
        i = m.end()
        depth, j = 1, i           # scan to the matching ')' of ccall(
        while depth:
            depth += JL[j] in "({["
            depth -= JL[j] in ")}]"
            j += 1
        parts = split_top(JL[i:j - 1])
        ret, argt, args = parts[0], parts[1], parts[2:]
        assert argt.startswith("(") and argt.endswith(")"), (m.group(1), argt)
        types = split_top(argt[1:-1])
        calls.append((m.group(1), ret, types, args))
    return calls


def test_every_ccall_matches_the_header():
    protos = c_prototypes()
    calls = jl_ccalls()
    assert len(calls) >= 15
    for name, ret, types, args in calls:
        assert name in protos, f"{name} is not declared in include/qtt_b200.h"
        cret, cargs = protos[name]
        assert jl_kind(ret) == c_kind(cret), (name, ret, cret)
        assert len(types) == len(cargs), (name, types, cargs)
        assert len(args) == len(types), (name, "argument count differs from the type tuple", args, types)
        for t, c in zip(types, cargs):
            assert jl_kind(t) == c_kind(c), (name, t, c)


def c_struct_fields(name):
    h = strip_comments(HDR)
    end = re.search(r"\}\s*" + name + r"\s*;", h).start()
    start = h.rfind("typedef struct", 0, end)
    body = h[h.index("{", start) + 1:end]
    fields = []
    for stmt in body.split(";"):
        stmt = stmt.strip()
        if not stmt:
            continue
        m = re.match(r"^(.*?[\s\*])(\w+(?:\s*,\s*\w+)*)$", stmt, flags=re.S)
        typ, names = m.group(1), m.group(2)
        for n in names.split(","):
            fields.append((n.strip(), c_kind(typ + " x")))
    return fields


def jl_struct_fields(name):
    body = re.search(r"struct " + name + r"\n(.*?)\nend", JL, flags=re.S).group(1)
    return [(l.split("::")[0].strip(), jl_kind(l.split("::")[1])) for l in body.splitlines() if "::" in l]


def test_struct_mirrors_have_the_header_layout():
    for cname, jname in (("qtt_sweep_params", "QttSweepParams"), ("qtt_sweep_info", "QttSweepInfo")):
        assert c_struct_fields(cname) == jl_struct_fields(jname), (cname, c_struct_fields(cname), jl_struct_fields(jname))


def test_solver_constants_match_the_enum():
    enum = dict((m.group(1), int(m.group(2))) for m in re.finditer(r"(QTT_SOLVER_\w+)\s*=\s*(\d+)", strip_comments(HDR)))
    consts = dict((m.group(1), int(m.group(2))) for m in re.finditer(r"const (QTT_SOLVER_\w+) = Int32\((\d+)\)", JL))
    assert enum == consts and len(enum) == 6


def body_of(fn):
    m = re.search(r"\nfunction " + fn + r"\(.*?\n(.*?)\nend\n", JL, flags=re.S)
    assert m, fn
    return m.group(1)


def test_dft_entry_points_follow_the_reference_reversal():
    """src/dft.jl:121-126 reverses the OUTPUT of apply(dft_mpo(s), psi); :128-133 applies idft_mpo(s) to reverse(psi) and does
    not reverse the result; neither forwards a cutoff of its own to `apply`."""
    fwd, inv = body_of("apply_dft_mpo"), body_of("apply_idft_mpo")
    assert re.search(r"phi = apply_mpo\(F, psi; kwargs\.\.\.\)", fwd) and "reverse([phi[j]" in fwd
    assert "reverse([psi[j]" in inv and re.search(r"return apply_mpo\(F, rpsi; kwargs\.\.\.\)", inv)
    assert "reverse([phi" not in inv


def test_mpo_cores_reads_the_actual_primed_index():
    """dft_mpo relabels the primed indices (reverse_output_bits, src/dft.jl:76-79,98): tensor j does not carry s_j'."""
    body = body_of("mpo_cores")
    assert "mpo_site_inds(A, j)" in body and "s'" not in body
    assert "plev(i) == 1" in body_of("mpo_site_inds")
