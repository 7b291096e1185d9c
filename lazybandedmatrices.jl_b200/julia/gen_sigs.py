#!/usr/bin/env python
"""Derive the Julia `ccall` signature of every entry point of include/lbm_b200.h.

`python gen_sigs.py` prints the `@sig` table that LBMB200.jl carries verbatim; tests/test_abi.py re-derives it from the
header and compares it with the shim line by line, so the two cannot drift apart (Julia cannot run in this image).

C type -> Julia ccall type:
  lbm_handle_t / lbm_multi_t / lbm_shard_t / lbm_svec_t, void *      -> Ptr{Cvoid}
  char -> Cchar, int -> Cint, int64_t -> Int64, uint64_t -> UInt64, size_t -> Csize_t, double -> Float64, float -> Float32
  [const] T *  -> Ptr{T};  const T *const * -> Ptr{Ptr{T}};  lbm_X_t * -> Ptr{Ptr{Cvoid}};  void ** -> Ptr{Ptr{Cvoid}}
  const char * -> Cstring
"""
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.normpath(os.path.join(HERE, "..", "..", "include", "lbm_b200.h"))

SCALARS = {"char": "Cchar", "int": "Cint", "int64_t": "Int64", "uint64_t": "UInt64", "size_t": "Csize_t", "double": "Float64",
           "float": "Float32", "void": "Cvoid"}
OPAQUE = ("lbm_handle_t", "lbm_multi_t", "lbm_shard_t", "lbm_svec_t")


def jl_type(ctype: str) -> str:
    t = ctype.replace("const", " ").strip()
    t = re.sub(r"\s+", " ", t)
    stars = t.count("*")
    base = t.replace("*", "").strip()
    if base in OPAQUE:
        out = "Ptr{Cvoid}"
    elif base == "char" and stars == 1 and "const" in ctype:
        return "Cstring"
    elif base in SCALARS:
        out = SCALARS[base]
    else:
        raise ValueError(f"unknown C type {ctype!r}")
    for _ in range(stars):
        out = f"Ptr{{{out}}}"
    return out


def header_signatures(path: str = HEADER):
    """{name: (julia return type, [julia argument types])} for every function the header declares."""
    src = open(path, encoding="utf-8").read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    out = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(lbm_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if "typedef" in ret:
            continue
        ret = ret.replace("extern", "").strip()
        argl = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                a = re.sub(r"\b[A-Za-z_]\w*$", "", a).strip() if not a.endswith("*") else a   # drop the parameter name
                argl.append(jl_type(a))
        out[name] = (jl_type(ret), argl)
    return out


def sig_lines():
    sigs = header_signatures()
    lines = []
    for name in sorted(sigs):
        ret, args = sigs[name]
        lines.append(f"@sig {name} {ret} ({', '.join(args)}{',' if len(args) == 1 else ''})")
    return lines


def rewrite_shim(path=os.path.join(HERE, "LBMB200.jl")):
    """Replace the table between the BEGIN / END SIGNATURES markers of the shim with the current one."""
    src = open(path, encoding="utf-8").read()
    a = src.index("#= BEGIN SIGNATURES =#") + len("#= BEGIN SIGNATURES =#\n")
    b = src.index("#= END SIGNATURES =#")
    open(path, "w", encoding="utf-8").write(src[:a] + "\n".join(sig_lines()) + "\n" + src[b:])


if __name__ == "__main__":
    if "--write" in sys.argv:
        rewrite_shim()
    else:
        sys.stdout.write("\n".join(sig_lines()) + "\n")
