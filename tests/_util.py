"""Shared tolerance helpers for the parity tests."""


def _eps(dtype):
    import numpy as np
    return float(np.finfo(dtype).eps)


def tol_rows(dtype, nterms, normA, normx, alpha=1.0, beta=0.0, normy=0.0):
    """north_star tolerance: 4*eps*(l+u+1)*||A||*||x|| per row (max-norms), extended with the
    alpha/beta scaling of the 5-argument mul!."""
    return 4.0 * _eps(dtype) * (nterms * abs(alpha) * normA * normx + abs(beta) * normy) + 1e-300
