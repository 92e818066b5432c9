"""The criterion behind the speculative LU panels (csrc/lu.cu), checked on the CPU against
LAPACK: for a tall panel, "every multiplier of the UNPIVOTED elimination has magnitude <= 1" holds exactly
when partial pivoting (dgetrf) performs no interchange in that panel, and then both give the same factors."""

from __future__ import annotations

import numpy as np
import pytest
from scipy.linalg.lapack import dgetrf


def unpivoted_panel(a):
    """L (unit lower trapezoid, m x nb) and U (nb x nb) of a tall panel without interchanges."""
    a = a.copy()
    m, nb = a.shape
    for j in range(nb):
        a[j + 1:, j] /= a[j, j]
        a[j + 1:, j + 1:] -= np.outer(a[j + 1:, j], a[j, j + 1:])
    return a


@pytest.mark.parametrize(("m", "nb", "boost"), [(300, 32, 0.0), (300, 32, 3.0), (300, 32, 40.0), (500, 64, 25.0),
                                                (200, 16, 1.2), (200, 16, 6.0), (128, 128, 60.0)])
def test_no_interchange_iff_all_multipliers_at_most_one(m, nb, boost):
    rng = np.random.default_rng(int(m * 1000 + nb + boost * 10))
    for trial in range(6):
        a = rng.normal(size=(m, nb))
        a[np.arange(nb), np.arange(nb)] += boost * (1 + 0.3 * trial)
        lu, piv, info = dgetrf(a)  # LAPACK partial pivoting on the m x nb panel; piv is 0-based
        assert info == 0
        no_swap = np.array_equal(piv[:nb], np.arange(nb))
        f = unpivoted_panel(a)
        mult = np.abs(np.tril(f, -1)).max() if m > 1 else 0.0
        assert (mult <= 1.0) == no_swap, (boost, trial, mult, no_swap)
        if no_swap:
            np.testing.assert_allclose(f, lu, rtol=1e-10, atol=1e-12)


def test_tie_goes_to_the_diagonal_like_idamax():
    # |a_10| == |a_00|: idamax returns the first maximum, i.e. no interchange, and |l| == 1 passes the check
    a = np.array([[2.0, 1.0], [-2.0, 3.0], [1.0, 0.5]])
    lu, piv, info = dgetrf(a)
    assert info == 0 and piv[0] == 0
    f = unpivoted_panel(a)
    assert np.abs(np.tril(f, -1))[:, 0].max() == 1.0
