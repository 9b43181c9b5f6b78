"""julia/FGQCuda.jl cannot be executed in this image (no Julia), so it must not drift from the C ABI: every
`ccall((:symbol, libfgq), Ret, (ArgTypes...), ...)` in it is parsed and checked against the prototype of that
symbol in include/fgq.h — symbol exists, same arity, compatible C types, same return type."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# Julia type in a ccall tuple -> admissible C parameter types (after normalisation: const dropped, names removed)
JL2C = {
    "Int64": {"int64_t"}, "Cint": {"int", "int32_t"}, "Float64": {"double"}, "Cfloat": {"float"},
    "Ptr{Float64}": {"double*"}, "Ptr{Int32}": {"int32_t*"}, "Ptr{Int64}": {"int64_t*"}, "Ptr{Cint}": {"int*"},
    "Ptr{Cfloat}": {"float*"}, "Ref{Int64}": {"int64_t*"}, "Ref{Cint}": {"int*"}, "Ref{Cfloat}": {"float*"},
    "Ptr{Cvoid}": {"void*", "double*"}, "Ptr{FGQShard}": {"fgq_shard*"}, "Ptr{Ptr{Float64}}": {"double**"},
    "Cstring": {"char*"}, "Ptr{Ptr{Cvoid}}": {"void**"},
}
RET = {"Cint": "int", "Int64": "int64_t", "Cstring": "char*", "Cvoid": "void"}


def _prototypes():
    txt = open(os.path.join(ROOT, "include", "fgq.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    protos = {}
    for ret, name, args in re.findall(r"\b((?:const\s+)?[a-z_0-9]+\s*\*?)\s*(fgq_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", txt):
        params = []
        args = args.strip()
        if args and args != "void":
            for a in args.split(","):
                a = re.sub(r"\bconst\b", "", a).strip()
                stars = a.count("*")
                base = a.replace("*", " ").split()
                ctype = " ".join(base[:-1]) if len(base) > 1 else base[0]   # drop the parameter name
                if ctype == "unsigned long long":
                    ctype = "uint64_t"
                params.append(ctype + "*" * stars)
        protos[name] = (re.sub(r"\bconst\b", "", ret).replace(" ", ""), params)
    return protos


def _split_tuple(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch == "{":
            depth += 1
        if ch == "}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _ccalls():
    src = open(os.path.join(ROOT, "julia", "FGQCuda.jl")).read()
    src = re.sub(r"#.*", "", src)
    calls = []
    # direct symbols
    for m in re.finditer(r"ccall\(\(\s*:(fgq_[a-z0-9_]+)\s*,\s*libfgq\)\s*,\s*([A-Za-z0-9{}]+)\s*,\s*\(", src):
        start = m.end()
        depth, i = 1, start
        while depth:
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        calls.append((m.group(1), m.group(2), _split_tuple(src[start:i - 1])))
    # @eval loops: ccall(($(QuoteNode(sym)), libfgq), ...) with sym drawn from a literal list of :fgq_ symbols
    for m in re.finditer(r"for \(f, sym\) in \((.*?)\)\n(.*?)\nend", src, flags=re.S):
        syms = re.findall(r":(fgq_[a-z0-9_]+)", m.group(1))
        mm = re.search(r"ccall\(\(\$\(QuoteNode\(sym\)\), libfgq\)\s*,\s*([A-Za-z0-9{}]+)\s*,\s*\(([^)]*)\)", m.group(2), flags=re.S)
        assert mm, "unparsed @eval ccall"
        for sname in syms:
            calls.append((sname, mm.group(1), _split_tuple(mm.group(2))))
    return calls


def test_every_ccall_matches_its_prototype():
    protos = _prototypes()
    calls = _ccalls()
    assert len(calls) >= 25, f"only {len(calls)} ccalls parsed"
    for name, ret, args in calls:
        assert name in protos, f"{name} is ccall'ed in julia/FGQCuda.jl but not declared in include/fgq.h"
        cret, cargs = protos[name]
        assert RET.get(ret) == cret, f"{name}: Julia return {ret} vs C {cret}"
        assert len(args) == len(cargs), f"{name}: {len(args)} ccall argument types vs {len(cargs)} C parameters"
        for i, (ja, ca) in enumerate(zip(args, cargs)):
            assert ja in JL2C, f"{name}: unknown Julia type {ja}"
            assert ca in JL2C[ja], f"{name}: argument {i}: Julia {ja} vs C {ca}"


def test_reference_entry_points_are_all_wrapped():
    """Every exported function of the reference on the path (src/FastGaussQuadrature.jl:7-19) has a wrapper,
    including the type-first methods (gausshermite.jl:28, gausslaguerre.jl:1, gaussradau.jl:31, gausslobatto.jl:31,
    gausschebyshev.jl:22,51,80,109) and the multi-GPU calls."""
    src = open(os.path.join(ROOT, "julia", "FGQCuda.jl")).read()
    for f in ("gausslegendre", "gaussjacobi", "gausshermite", "gausslaguerre", "gaussradau", "gausslobatto",
              "gausschebyshevt", "gausschebyshevu", "gausschebyshevv", "gausschebyshevw", "approx_besselroots"):
        assert re.search(rf"\b{f}\(", src), f
    for f in ("gausshermite", "gausslaguerre", "gaussradau", "gausslobatto", "gausschebyshevt", "gausschebyshevu",
              "gausschebyshevv", "gausschebyshevw", "gausslegendre"):
        assert re.search(rf"^{f}\(::Type\{{", src, flags=re.M), f"type-first method of {f} missing"
    for f in ("gausslegendre_sharded", "allgather", "gausslegendre_replicated", "gausslegendre_multi", "free_shards!"):
        assert f in src
    # the struct mirrors fgq_shard field for field
    body = re.search(r"struct FGQShard\n(.*?)\nend", src, flags=re.S).group(1)
    fields = [ln.split("::") for ln in body.split("\n")]
    assert [(a.strip(), b.strip()) for a, b in fields] == [("device", "Int32"), ("owner", "Int32"), ("lo", "Int64"),
                                                           ("hi", "Int64"), ("x", "Ptr{Float64}"), ("w", "Ptr{Float64}")]
