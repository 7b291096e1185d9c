"""CPU: the oracle's restatement of `+`/`-` among structured kinds (src/special.jl:75-222) against dense
arithmetic, and the host layer's result-kind rule against the return types written in the reference."""
import itertools

import numpy as np
import pytest

KINDS = ["D", "BU", "BL", "S", "T"]

# Return kind of `A op B`, transcribed from the constructors the reference's methods call
# (/root/reference/src/special.jl; same-kind rows from src/tridiag.jl:206-207,:571-572 and src/bidiag.jl:330-346).
RETURNS = {
    ("BU", "D"): "BU", ("BL", "D"): "BL",      # special.jl:75-83   Bidiagonal(newdv, A.ev, A.uplo)
    ("D", "BU"): "BU", ("D", "BL"): "BL",      # special.jl:85-93
    ("D", "S"): "S", ("S", "D"): "S",          # special.jl:95-113  SymTridiagonal(...)
    ("T", "S"): "T", ("S", "T"): "T",          # special.jl:117-120
    ("D", "T"): "T", ("T", "D"): "T",          # special.jl:123-141
    ("BU", "T"): "T", ("BL", "T"): "T", ("T", "BU"): "T", ("T", "BL"): "T",   # special.jl:143-161
    ("BU", "S"): "T", ("BL", "S"): "T", ("S", "BU"): "T", ("S", "BL"): "T",   # special.jl:163-181
    ("S", "S"): "S", ("T", "T"): "T",          # src/tridiag.jl:206-207, :571-572
    ("BU", "BU"): "BU", ("BL", "BL"): "BL",    # src/bidiag.jl:331-332
    ("BU", "BL"): "T", ("BL", "BU"): "T",      # src/bidiag.jl:334-336
    ("D", "D"): "D",                           # LinearAlgebra
}


def make(oracle, kind, n, seed, dt=np.float64):
    d = oracle.fill_uniform(seed, n, dt) - 0.5
    e1 = oracle.fill_uniform(seed + 1, max(n - 1, 0), dt) - 0.5
    e2 = oracle.fill_uniform(seed + 2, max(n - 1, 0), dt) - 0.5
    return {"D": ("D", None, d, None), "BU": ("BU", None, d, e1), "BL": ("BL", e1, d, None),
            "S": ("S", e1, d, e1), "T": ("T", e1, d, e2)}[kind]


@pytest.mark.parametrize("n", [1, 2, 5, 20])
def test_oracle_special_pairs_match_dense(oracle, n):
    for ka, kb in itertools.product(KINDS, KINDS):
        A, B = make(oracle, ka, n, 11), make(oracle, kb, n, 23)
        for sb in (1.0, -1.0):
            R = oracle.special_axpby(1.0, A, sb, B)
            assert R[0] == RETURNS[(ka, kb)], (ka, kb)
            want = oracle.special_dense(A) + sb * oracle.special_dense(B)
            assert np.array_equal(oracle.special_dense(R), want), (ka, kb, sb)   # test/test_special.jl:154-160


def test_oracle_uniformscaling(oracle):
    n = 7
    for k in KINDS:
        A = make(oracle, k, n, 5)
        for lam in (3.0, 1.0, -0.6):
            I = ("I", lam)
            assert np.array_equal(oracle.special_dense(oracle.special_axpby(1.0, A, 1.0, I)),
                                  oracle.special_dense(A) + lam * np.eye(n))          # special.jl:183-198
            assert np.array_equal(oracle.special_dense(oracle.special_axpby(1.0, I, -1.0, A)),
                                  lam * np.eye(n) - oracle.special_dense(A))          # special.jl:208-222
            assert oracle.special_axpby(1.0, I, -1.0, A)[0] == k


def test_reference_kat_binary_ops(oracle):
    # test/test_special.jl:113-123: a = [1.0:n;], A = Diagonal(a) converted to each kind: B + C == 2A, B - C == 0
    n = 10
    a = np.arange(1.0, n + 1)
    z = np.zeros(n - 1)
    conv = {"D": ("D", None, a, None), "BU": ("BU", None, a, z), "T": ("T", z, a, z)}
    for B, Cc in itertools.product(conv.values(), conv.values()):
        assert np.array_equal(oracle.special_dense(oracle.special_axpby(1.0, B, 1.0, Cc)), np.diag(2 * a))
        assert np.array_equal(oracle.special_dense(oracle.special_axpby(1.0, B, -1.0, Cc)), np.zeros((n, n)))


def test_host_join_rule_matches_reference_return_types():
    from lbm_b200.special import _join
    for (ka, kb), want in RETURNS.items():
        assert _join(ka, kb) == want, (ka, kb)
