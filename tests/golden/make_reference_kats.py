#!/usr/bin/env python
"""Transcribe the reference's own known-answer tests for the hot path into a fixture.

The reference is Julia and cannot run in this image, so the golden vectors are the
LITERALS of its test-suite.  Each entry below names the reference file:line it was taken
from and a snippet that must appear verbatim on that line; when /root/reference is
present this script re-checks every snippet against the reference source before
writing ``reference_kats.json`` (so a stale or mistyped transcription fails here, not
silently in the tests).  The GPU box has no /root/reference: tests only read the JSON.

    python tests/golden/make_reference_kats.py
"""
import json
import os
import sys

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kats.json")

K = []  # (kind, src "file:line", must-contain snippet, payload)


def kat(kind, src, snippet, **payload):
    K.append({"kind": kind, "src": src, "snippet": snippet, **payload})


# ---- DiagTrav index map (values exact) ------------------------------------------------
kat("diagtrav", "test/test_blockkron.jl:25", "== [1, 4, 2, 7, 5, 3]",
    X=[[1, 2, 3], [4, 5, 6], [7, 8, 9]], expect=[1, 4, 2, 7, 5, 3])
kat("diagtrav", "test/test_blockkron.jl:31", "DiagTrav(A) == [1, 4, 2, 5, 3]",
    X=[[1, 2, 3], [4, 5, 6]], expect=[1, 4, 2, 5, 3])
kat("diagtrav", "test/test_blockkron.jl:33", "DiagTrav(A) == [1, 3, 2, 5, 4]",
    X=[[1, 2], [3, 4], [5, 6]], expect=[1, 3, 2, 5, 4])
kat("diagtrav", "test/test_blockkron.jl:67", "DiagTrav(A) == [1, 4, 2, 7, 5, 3, 10, 8, 6]",
    X=[[1, 2, 3], [4, 5, 6], [7, 8, 9], [10, 11, 12]], expect=[1, 4, 2, 7, 5, 3, 10, 8, 6])
kat("diagtrav", "test/test_blockkron.jl:69", "DiagTrav(A) == [1, 2, 4, 3, 5, 7, 6, 8, 10]",
    X=[[1, 4, 7, 10], [2, 5, 8, 11], [3, 6, 9, 12]], expect=[1, 2, 4, 3, 5, 7, 6, 8, 10])
kat("diagtrav_padded", "test/test_blockkron.jl:74",
    "[1; 4; 2; 7; 5; 3; 0; 8; 6; 0; 0; 0; 9; zeros(42)]",
    shape=[10, 10], block=[[1, 2, 3], [4, 5, 6], [7, 8, 9]],
    expect=[1, 4, 2, 7, 5, 3, 0, 8, 6, 0, 0, 0, 9] + [0] * 42)
kat("diagtrav_padded", "test/test_blockkron.jl:77",
    "[1; 4; 2; 7; 5; 3; 0; 8; 6; 4; 0; 0; 9; 4; 0; 0; 0; 4; 0; 0]",
    shape=[5, 6], block=[[1, 2, 3, 4], [4, 5, 6, 4], [7, 8, 9, 4]],
    expect=[1, 4, 2, 7, 5, 3, 0, 8, 6, 4, 0, 0, 9, 4, 0, 0, 0, 4, 0, 0])
kat("diagtrav_padded", "test/test_blockkron.jl:80",
    "[1; 4; 2; 7; 5; 3; 0; 8; 6; 4; 0; 0; 9; 4; 0; 0; 0; 0; 4; 0]",
    shape=[6, 5], block=[[1, 2, 3, 4], [4, 5, 6, 4], [7, 8, 9, 4]],
    expect=[1, 4, 2, 7, 5, 3, 0, 8, 6, 4, 0, 0, 9, 4, 0, 0, 0, 0, 4, 0])
kat("diagtrav", "test/test_blockkron.jl:85", "a == ones(3)",
    X=[[1, 1, 1]], expect=[1, 1, 1])
kat("diagtrav3_blocks", "test/test_blockkron.jl:57",
    "A[Block(3)] == [A.array[3,1,1], A.array[2,2,1], A.array[2,1,2],",
    n=3, block2=[[2, 1, 1], [1, 2, 1], [1, 1, 2]],
    block3=[[3, 1, 1], [2, 2, 1], [2, 1, 2], [1, 3, 1], [1, 2, 2], [1, 1, 3]])
kat("invdiagtrav", "test/test_blockkron.jl:89", "[1 2 3; 4 5 0; 7 0 0]",
    X=[[1, 2, 3], [4, 5, 6], [7, 8, 9]], expect=[[1, 2, 3], [4, 5, 0], [7, 0, 0]])

# ---- KronTrav (values exact) -------------------------------------------------------------
kat("krontrav_vec", "test/test_blockkron.jl:123", "KronTrav(a,c) == [7,8,14,16,21]",
    a=[1, 2, 3], b=[7, 8], expect=[7, 8, 14, 16, 21])
kat("krontrav_vec", "test/test_blockkron.jl:124", "KronTrav(c,a) == [7,14,8,21,16]",
    a=[7, 8], b=[1, 2, 3], expect=[7, 14, 8, 21, 16])
kat("krontrav_mul", "test/test_blockkron.jl:146",
    "K*DiagTrav(X) == K*DiagTrav(Y) == DiagTrav(B*X*A')",
    A=[[1, 2], [3, 4]], B=[[5, 6], [7, 8]], X=[[9, 10], [11, 0]], Y=[[9, 10], [11, 12]])
kat("krontrav3_matrix", "test/test_blockkron.jl:161", "K == [1*5*2 1*5*4 1*6*2 2*5*2;",
    A=[[1, 2], [3, 4]], B=[[5, 6], [7, 8]], C=[[2, 4], [6, 8]],
    expect=[[1 * 5 * 2, 1 * 5 * 4, 1 * 6 * 2, 2 * 5 * 2],
            [1 * 5 * 6, 1 * 5 * 8, 1 * 6 * 6, 2 * 5 * 6],
            [1 * 7 * 2, 1 * 7 * 4, 1 * 8 * 2, 2 * 7 * 2],
            [3 * 5 * 2, 3 * 5 * 4, 3 * 6 * 2, 4 * 5 * 2]])
kat("krontrav_laplace2d", "test/test_blockkron.jl:206", "A * DiagTrav(X) == DiagTrav(X * Δ')",
    n=4, diag=-2, off=1)
kat("krontrav_laplace2d", "test/test_blockkron.jl:207", "B * DiagTrav(X) == DiagTrav(Δ * X)",
    n=4, diag=-2, off=1)
kat("krontrav_square", "test/test_blockkron.jl:281", "A^2 == Matrix(A)^2", n=4, diag=-2, off=1)

# ---- structure (bit-exact integer tuples) -------------------------------------------------
kat("structure", "test/test_blockkron.jl:100", "blockbandwidths(A) == (1,1)",
    expr="blockkron", factors=[[1, 1], [0, 0]], blockbandwidths=[1, 1], subblockbandwidths=[0, 0],
    isblockbanded=True, isbandedblockbanded=True)
kat("structure", "test/test_blockkron.jl:105", "blockbandwidths(A) == (1,1)",
    expr="blockkron", factors=[[1, 1], [0, 0], [0, 0]], blockbandwidths=[1, 1],
    subblockbandwidths=[0, 0], isblockbanded=True)
kat("structure", "test/test_blockkron.jl:111", "blockbandwidths(B) == (0,0)",
    expr="blockkron", factors=[[0, 0], [1, 1], [0, 0]], blockbandwidths=[0, 0], isblockbanded=True)
kat("structure", "test/test_blockkron.jl:209", "blockbandwidths(A) == blockbandwidths(B) == (1,1)",
    expr="krontrav", factors=[[1, 1], [0, 0]], blockbandwidths=[1, 1], subblockbandwidths=[1, 1],
    isblockbanded=True, isbandedblockbanded=True)
kat("structure", "test/test_blockkron.jl:211", "subblockbandwidths(B) == (0,0)",
    expr="krontrav", factors=[[0, 0], [1, 1]], blockbandwidths=[1, 1], subblockbandwidths=[0, 0],
    isblockbanded=True, isbandedblockbanded=True)
kat("structure", "test/test_blockkron.jl:228",
    "blockbandwidths(A) == blockbandwidths(B) == blockbandwidths(C) == (1,1)",
    expr="krontrav", factors=[[1, 1], [0, 0], [0, 0]], blockbandwidths=[1, 1], isblockbanded=True,
    isbandedblockbanded=False)
kat("structure", "test/test_blockkron.jl:305", "blockbandwidths(Δ) == (1,1)",
    expr="broadcast_plus_blockkron", terms=[[[1, 1], [0, 0]], [[0, 0], [1, 1]]],
    blockbandwidths=[1, 1], subblockbandwidths=[1, 1])
kat("structure", "test/test_blockconcat.jl:130", "blockbandwidths(V) == (5,0)",
    expr="blockvcat", args=[{"blocksize": [4, 4], "blockbandwidths": [1, 0]},
                            {"blocksize": [4, 4], "blockbandwidths": [1, 0]}],
    blockbandwidths=[5, 0], isblockbanded=True)
kat("structure", "test/test_blockconcat.jl:276", "blockbandwidths(H) == (1,4)",
    expr="blockhcat", args=[{"blocksize": [4, 4], "blockbandwidths": [1, 0]},
                            {"blocksize": [4, 4], "blockbandwidths": [1, 0]}],
    blockbandwidths=[1, 4], isblockbanded=True)
kat("structure", "test/test_blockconcat.jl:414", "blockbandwidths(A) == (1,2)",
    expr="interlace_hvcat", p=2,
    args=[{"blockbandwidths": [1, 2], "subblockbandwidths": [0, 0]}, "zeros", "zeros",
          {"blockbandwidths": [1, 2], "subblockbandwidths": [0, 0]}],
    blockbandwidths=[1, 2], subblockbandwidths=[0, 0])
kat("structure", "test/test_blockconcat.jl:436", "blockbandwidths(C) == (2,2)",
    expr="interlace_diagonal",
    args=[{"blockbandwidths": [1, 2], "subblockbandwidths": [0, 0]},
          {"blockbandwidths": [2, 1], "subblockbandwidths": [0, 0]}],
    blockbandwidths=[2, 2], subblockbandwidths=[0, 0])
kat("structure", "test/test_lazybandedinf.jl:59", "blockbandwidths(B) == (1, 1)",
    expr="interlace_hvcat", p=2,
    args=[{"blockbandwidths": [1, 1], "subblockbandwidths": [0, 0]}, "zeros", "zeros",
          {"blockbandwidths": [1, 1], "subblockbandwidths": [0, 0]}],
    blockbandwidths=[1, 1], subblockbandwidths=[0, 0])
kat("structure", "test/test_lazybandedinf.jl:68", "subblockbandwidths(C) == (1, 1)",
    expr="interlace_hvcat", p=2,
    args=[{"blockbandwidths": [1, 1], "subblockbandwidths": [0, 0]}] * 4,
    blockbandwidths=[1, 1], subblockbandwidths=[1, 1])
kat("structure", "test/test_misc.jl:59", "blockbandwidths(unitblocks(Diagonal(1:5))) == (0,0)",
    expr="unitblocks", bandwidths=[0, 0], blockbandwidths=[0, 0], subblockbandwidths=[0, 0])
kat("structure", "README.md:16", "ApplyMatrix(*, A, A)",
    expr="apply_mul", size=[10, 10], factors=[[1, 1], [1, 1]], bandwidths=[2, 2])
kat("structure", "src/tridiag.jl:688", "bandwidths(::Tridiagonal) = (1,1)",
    expr="tridiagonal", bandwidths=[1, 1])
kat("structure", "src/tridiag.jl:687", "bandwidths(::SymTridiagonal) = (1,1)",
    expr="symtridiagonal", bandwidths=[1, 1])
kat("structure", "src/bidiag.jl:439", "A.uplo == 'U' ? (0,1) : (1,0)",
    expr="bidiagonal", bandwidths_U=[0, 1], bandwidths_L=[1, 0])

# ---- Tridiagonal / SymTridiagonal / Bidiagonal element definitions and products -----------
kat("dense_tridiagonal", "test/test_tridiag.jl:335",
    "Tridiagonal(4:5, 1:3, 1:2) == [1 1 0; 4 2 2; 0 5 3]",
    dl=[4, 5], d=[1, 2, 3], du=[1, 2], expect=[[1, 1, 0], [4, 2, 2], [0, 5, 3]])
kat("dense_symtridiagonal", "test/test_tridiag.jl:334",
    "SymTridiagonal(1:3, 1:2) == [1 1 0; 1 2 2; 0 2 3]",
    dv=[1, 2, 3], ev=[1, 2], expect=[[1, 1, 0], [1, 2, 2], [0, 2, 3]])
kat("dense_bidiagonal", "test/test_bidiag.jl:296",
    "Bidiagonal(1:3, 1:2, :U) == [1 1 0; 0 2 2; 0 0 3]",
    dv=[1, 2, 3], ev=[1, 2], uplo="U", expect=[[1, 1, 0], [0, 2, 2], [0, 0, 3]])
kat("symtridiag_cube", "test/test_tridiag.jl:319", "SymTridiagonal([1, 2], [0])^3 == [1 0; 0 8]",
    dv=[1, 2], ev=[0], expect=[[1, 0], [0, 8]])
kat("symtridiag_mv", "test/test_tridiag.jl:341", "T*x == ones(1)",
    dv=[1.0], ev=[3.0], x=[1.0], expect=[1.0])
kat("symtridiag_empty", "test/test_tridiag.jl:342",
    "SymTridiagonal(ones(0), ones(0)) * ones(0, 2) == ones(0, 2)", n=0, nrhs=2)
kat("tridiag_sum", "test/test_tridiag.jl:370", "sum(T) == 24",
    dl=[1, 2], d=[1, 2, 3], du=[7, 8], expect=24)
kat("symtridiag_sum", "test/test_tridiag.jl:371", "sum(S) == 12",
    dv=[1, 2, 3], ev=[1, 2], expect=12)


def main():
    have_ref = os.path.isdir(REF)
    bad = 0
    for e in K:
        f, ln = e["src"].rsplit(":", 1)
        if have_ref:
            with open(os.path.join(REF, f), encoding="utf-8") as fh:
                lines = fh.read().split("\n")
            window = "\n".join(lines[max(0, int(ln) - 3): int(ln) + 2])
            if e["snippet"] not in window:
                print(f"MISMATCH {e['src']}: {e['snippet']!r} not found near that line", file=sys.stderr)
                bad += 1
    if bad:
        sys.exit(1)
    with open(OUT, "w", encoding="utf-8") as fh:
        json.dump({"reference": "LazyBandedMatrices.jl v0.11.9", "verified_against_source": have_ref,
                   "kats": K}, fh, indent=1, ensure_ascii=False)
    print(f"wrote {len(K)} KATs -> {OUT} (verified against reference source: {have_ref})")


if __name__ == "__main__":
    main()
