"""Host-side structure algebra (bandwidths / blockbandwidths / subblockbandwidths) must match the
reference BIT-EXACTLY: every structural KAT of the reference's tests."""
import pytest


class _Band:
    """A banded operand known only by its size and bandwidths (structure queries need no data)."""

    def __init__(self, n, l, u):
        self.shape = (n, n)
        self._bw = (l, u)

    def bandwidths(self):
        return self._bw


def _mk(L, bw, n=4):
    return L.Eye(n) if tuple(bw) == (0, 0) else _Band(n, *bw)


def test_structure_kats(kats, L):
    seen = 0
    for k in [k for k in kats if k["kind"] == "structure"]:
        e = k["expr"]
        seen += 1
        if e == "blockkron":
            K = L.BlockKron(*[_mk(L, bw) for bw in k["factors"]])
            assert list(K.blockbandwidths()) == k["blockbandwidths"], k["src"]
            if "subblockbandwidths" in k:
                assert list(K.subblockbandwidths()) == k["subblockbandwidths"], k["src"]
            assert K.isblockbanded() == k["isblockbanded"]
            if "isbandedblockbanded" in k:
                assert K.isbandedblockbanded() == k["isbandedblockbanded"]
        elif e == "krontrav":
            K = L.KronTrav(*[_mk(L, bw) for bw in k["factors"]])
            assert list(K.blockbandwidths()) == k["blockbandwidths"], k["src"]
            if "subblockbandwidths" in k:
                assert list(K.subblockbandwidths()) == k["subblockbandwidths"], k["src"]
            assert K.isblockbanded() == k["isblockbanded"]
            if "isbandedblockbanded" in k:
                assert K.isbandedblockbanded() == k["isbandedblockbanded"], k["src"]
        elif e == "broadcast_plus_blockkron":
            (a1, b1), (a2, b2) = k["terms"]
            S = L.KronSum.__new__(L.KronSum)
            S.A, S.B = _Band(4, *a1), _Band(4, *b2)
            assert (a1, b1, a2, b2) == ([1, 1], [0, 0], [0, 0], [1, 1])
            assert list(S.blockbandwidths()) == k["blockbandwidths"]
            assert list(S.subblockbandwidths()) == k["subblockbandwidths"]
        elif e in ("blockvcat", "blockhcat"):
            specs = [L.BlockBandedSpec(tuple(a["blocksize"]), tuple(a["blockbandwidths"])) for a in k["args"]]
            V = (L.BlockVcat if e == "blockvcat" else L.BlockHcat)(*specs)
            assert list(V.blockbandwidths()) == k["blockbandwidths"], k["src"]
            assert V.isblockbanded() == k["isblockbanded"]
        elif e == "interlace_hvcat":
            args = [L.zeros_spec((5, 5)) if a == "zeros" else
                    L.BlockBandedSpec((5, 5), tuple(a["blockbandwidths"]), tuple(a["subblockbandwidths"]))
                    for a in k["args"]]
            Bc = L.BlockBroadcastArray("hvcat", k["p"], *args)
            assert list(Bc.blockbandwidths()) == k["blockbandwidths"], k["src"]
            assert list(Bc.subblockbandwidths()) == k["subblockbandwidths"], k["src"]
        elif e == "interlace_diagonal":
            args = [L.BlockBandedSpec((5, 5), tuple(a["blockbandwidths"]), tuple(a["subblockbandwidths"]))
                    for a in k["args"]]
            Bc = L.BlockBroadcastArray("Diagonal", *args)
            assert list(Bc.blockbandwidths()) == k["blockbandwidths"]
            assert list(Bc.subblockbandwidths()) == k["subblockbandwidths"]
        elif e == "unitblocks":
            s = L.unitblocks(_Band(5, *k["bandwidths"]))
            assert list(s.blockbandwidths) == k["blockbandwidths"]
            assert list(s.subblockbandwidths) == k["subblockbandwidths"]
        elif e == "apply_mul":
            n = k["size"][0]
            M = L.ApplyMatrix("*", *[_Band(n, *bw) for bw in k["factors"]])
            assert list(M.bandwidths()) == k["bandwidths"]
        elif e in ("tridiagonal", "symtridiagonal", "bidiagonal"):
            import torch
            d, e1 = torch.zeros(3), torch.zeros(2)
            if e == "tridiagonal":
                assert list(L.Tridiagonal(e1, d, e1).bandwidths()) == k["bandwidths"]
            elif e == "symtridiagonal":
                assert list(L.SymTridiagonal(d, e1).bandwidths()) == k["bandwidths"]
                assert list(L.SymTridiagonal(d, torch.zeros(3)).bandwidths()) == k["bandwidths"]  # len(ev)==n allowed
            else:
                assert list(L.Bidiagonal(d, e1, "U").bandwidths()) == k["bandwidths_U"]
                assert list(L.Bidiagonal(d, e1, "L").bandwidths()) == k["bandwidths_L"]
        else:
            raise AssertionError(f"unhandled structure KAT {e}")
    assert seen >= 18


def test_prod_bandwidths_clip(L):
    # bandwidths(ApplyMatrix(*, ...)) = min.(size-1, sums)
    assert L.ApplyMatrix("*", _Band(3, 2, 2), _Band(3, 2, 2)).bandwidths() == (2, 2)
    assert L.ApplyMatrix("*", _Band(10, 1, 1), _Band(10, 1, 1), _Band(10, 4, 0)).bandwidths() == (6, 2)
    assert L.ApplyMatrix("*", _Band(4096, 4, 4), _Band(4096, 4, 4), _Band(4096, 4, 4)).bandwidths() == (12, 12)


def test_constructor_errors(L):
    import torch
    with pytest.raises(L.LbmArgumentError):
        L.Bidiagonal(torch.zeros(3), torch.zeros(2), "R")          # test/test_bidiag.jl:41
    with pytest.raises(L.DimensionMismatch):
        L.Bidiagonal(torch.zeros(3), torch.zeros(3), "U")
    with pytest.raises(L.LbmArgumentError):
        L.Tridiagonal(torch.zeros(1), torch.zeros(3), torch.zeros(2))  # test/test_tridiag.jl:42-45
    with pytest.raises(L.DimensionMismatch):
        L.SymTridiagonal(torch.zeros(3), torch.zeros(5))
    with pytest.raises(L.DimensionMismatch):
        L.ApplyMatrix("*", _Band(3, 1, 1), _Band(4, 1, 1))


def test_transposes_swap_structure(L):
    import torch
    dl, d, du = torch.tensor([1., 2.]), torch.tensor([3., 4., 5.]), torch.tensor([6., 7.])
    T = L.Tridiagonal(dl, d, du).T                                  # src/tridiag.jl:449-450
    assert T.dl is not None and torch.equal(T.dl, du) and torch.equal(T.du, dl)
    B = L.Bidiagonal(d, du, "U").T                                  # src/bidiag.jl:222-233
    assert B.uplo == "L" and B.bandwidths() == (1, 0)
    assert torch.equal(L.Tridiagonal(dl, d, du).to_dense(),
                       torch.tensor([[3., 6., 0.], [1., 4., 7.], [0., 2., 5.]]))
