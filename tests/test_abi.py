"""The C-ABI shared library loads and exports every symbol include/lbm_b200.h declares
(no compute calls: this file runs without a GPU)."""
import ctypes as C

import pytest


def test_exports_every_declared_symbol(L):
    from lbm_b200._lib import SIGNATURES, header_symbols
    lib = L.lib()
    declared = header_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/lbm_b200.h but not exported"
    assert set(declared) == set(SIGNATURES), "python binding and header out of sync"


def test_version_and_structure_helper(L):
    lib = L.lib()
    assert lib.lbm_version() >= 10000
    kl = (C.c_int64 * 2)(1, 1)
    ku = (C.c_int64 * 2)(1, 1)
    ol, ou = C.c_int64(), C.c_int64()
    assert lib.lbm_prod_bandwidths(10, 10, 2, kl, ku, C.byref(ol), C.byref(ou)) == 0
    assert (ol.value, ou.value) == (2, 2)                      # README.md:13-26
    assert lib.lbm_prod_bandwidths(2, 2, 2, kl, ku, C.byref(ol), C.byref(ou)) == 0
    assert (ol.value, ou.value) == (1, 1)                      # clipped to size-1
    assert lib.lbm_prod_bandwidths(-1, 2, 2, kl, ku, C.byref(ol), C.byref(ou)) == -1
    assert lib.lbm_prod_bandwidths(2, 2, 0, kl, ku, C.byref(ol), C.byref(ou)) == -3


def test_no_cpu_fallback(L):
    """Without a CUDA device the product must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the -m gpu tests")
    with pytest.raises(L.LbmCudaError):
        L.Handle(0)
    A = L.BandedMatrix(torch.zeros(4, 3, dtype=torch.float64), 4, 1, 1)
    with pytest.raises(L.LbmCudaError):
        A @ torch.zeros(4, dtype=torch.float64)


def test_product_never_imports_oracle():
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "lazybandedmatrices.jl_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                src = open(os.path.join(dp, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle|liblbm_oracle|oracle/", src, flags=re.M), f


# ---- the Julia shim (julia/LBMB200.jl) cannot run here: check it mechanically against the header ----------------------
def _shim():
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return open(os.path.join(root, "lazybandedmatrices.jl_b200", "julia", "LBMB200.jl"), encoding="utf-8").read()


def _gen_sigs():
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_sigs", os.path.join(root, "lazybandedmatrices.jl_b200", "julia", "gen_sigs.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _split_args(src, i):
    """src[i] == '(' : return (list of top-level argument strings, index after the matching ')')."""
    depth, j, args, cur, instr = 0, i, [], "", False
    while j < len(src):
        ch = src[j]
        if instr:
            cur += ch
            if ch == '"' and src[j - 1] != "\\":
                instr = False
        elif ch == '"':
            instr = True
            cur += ch
        elif ch in "([{":
            depth += 1
            if depth > 1:
                cur += ch
        elif ch in ")]}":
            depth -= 1
            if depth == 0:
                if cur.strip():
                    args.append(cur.strip())
                return args, j + 1
            cur += ch
        elif ch == "," and depth == 1:
            args.append(cur.strip())
            cur = ""
        else:
            cur += ch
        j += 1
    raise AssertionError("unbalanced parentheses in the shim")


def test_julia_shim_signature_table_matches_header():
    """Every entry point of include/lbm_b200.h is bound in the shim with the argument count and types the header declares."""
    import re
    gen = _gen_sigs()
    want = gen.sig_lines()
    src = _shim()
    block = src[src.index("#= BEGIN SIGNATURES =#"):src.index("#= END SIGNATURES =#")]
    have = [ln.strip() for ln in block.splitlines() if ln.strip().startswith("@sig ")]
    assert have == want, "julia/LBMB200.jl's @sig table is out of date: regenerate it with julia/gen_sigs.py"
    assert len(have) == len(gen.header_signatures()) >= 90
    # and the python binding agrees on the argument COUNT of every entry point
    from lbm_b200._lib import SIGNATURES
    for name, (_, args) in gen.header_signatures().items():
        assert len(SIGNATURES[name][1]) == len(args), name
    assert re.search(r"struct DeviceColumnMajor <: MemoryLayout end", src)      # not an AbstractColumnMajor


def test_julia_shim_calls_have_the_declared_argument_counts():
    """Every `c_lbm_*(...)` call in the shim (direct, or through the `$var(...)` interpolations of the @eval loops) passes as
    many arguments as the header declares for that entry point."""
    import re
    gen = _gen_sigs()
    sigs = gen.header_signatures()
    src = _shim()
    body = src[src.index("#= END SIGNATURES =#"):]
    checked = set()
    for m in re.finditer(r"\bc_(lbm_[a-z0-9_]+)\(", body):
        name = m.group(1)
        assert name in sigs, f"shim calls c_{name}, which the header does not declare"
        args, _ = _split_args(body, m.end() - 1)
        assert len(args) == len(sigs[name][1]), (name, len(args), len(sigs[name][1]), args)
        checked.add(name)
    # interpolated: var = Symbol(:c_lbm_, p, :suffix)  ...  $var(args...)
    var2suffix = dict(re.findall(r"(\w+) = Symbol\(:c_lbm_, p, :([a-z0-9_]+)\)", body))
    assert len(var2suffix) >= 25
    for var, suffix in var2suffix.items():
        calls = list(re.finditer(r"\$" + var + r"\(", body))
        assert calls, f"{var} (lbm_?{suffix}) is bound but never called"
        for m in calls:
            args, _ = _split_args(body, m.end() - 1)
            for pfx in ("d", "s"):
                name = f"lbm_{pfx}{suffix}"
                if name in sigs:
                    assert len(args) == len(sigs[name][1]), (name, len(args), len(sigs[name][1]))
                    checked.add(name)
    missing = sorted(set(sigs) - checked)
    assert not missing, f"entry points bound in the @sig table but never used by a method of the shim: {missing}"


def test_integration_md_index_is_the_shim():
    """INTEGRATION.md's dispatch table is generated from the shim (julia/gen_index.py): it must be up to date, and every entry
    point of the header must be reachable from some method of the shim it lists."""
    import importlib.util
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_index", os.path.join(root, "lazybandedmatrices.jl_b200", "julia", "gen_index.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    want = mod.rows()
    doc = open(os.path.join(root, "INTEGRATION.md"), encoding="utf-8").read()
    block = doc[doc.index("<!-- BEGIN INDEX -->"):doc.index("<!-- END INDEX -->")]
    have = [ln for ln in block.splitlines() if ln.startswith("| `LBMB200.jl:")]
    assert have == want, "INTEGRATION.md index is stale: python lazybandedmatrices.jl_b200/julia/gen_index.py --write"
    names = set(re.findall(r"`(lbm_[?a-z0-9_]+)`", "\n".join(want)))
    gen = _gen_sigs()
    missing = [h for h in gen.header_signatures() if h not in names and re.sub(r"^lbm_[ds]", "lbm_?", h) not in names]
    assert not missing, missing
