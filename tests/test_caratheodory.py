"""Exact Gauss-Newton coresets (host NumPy part; mirrors the reference's tests/test_caratheodory.py:6-59:
same sizes, same bounds) plus the small-set and degenerate cases."""
import numpy as np
import pytest

from point_cloud_registration_b200.caratheodory import caratheodory, create_gn_set, fast_caratheodory


def hge(J, r, w=None):
    w = np.ones(len(r)) if w is None else w
    return J.T @ (w[:, None] * J), J.T @ (w * r), r @ (w * r)


@pytest.mark.parametrize("seed", [0, 1])
def test_exact_same_results(seed):
    rng = np.random.default_rng(seed)
    N, k, N_target = 30000, 64, 128
    J, r = rng.standard_normal((N, 6)), rng.standard_normal(N)
    P = create_gn_set(J, r)
    assert P.shape == (28, N) and N_target > P.shape[0] + 1
    _, w, idx = fast_caratheodory(P, np.ones(N), k, N_target)
    H, g, e2 = hge(J, r)
    Hs, gs, es = hge(J[idx], r[idx], w)
    assert max(np.max(np.abs(H - Hs)), np.max(np.abs(g - gs)), abs(e2 - es)) <= 1e-10
    assert len(w) <= N_target and np.all(w > 0)                  # reference test_weights_positive


def test_gn_set_layout_matches_the_record_layout():
    """Row order = the 29-double record of pcr_linearize without the count: 21 upper-triangle entries of
    J^T J row by row, 6 entries of J r, r^2."""
    rng = np.random.default_rng(3)
    J, r = rng.standard_normal((50, 6)), rng.standard_normal(50)
    P = create_gn_set(J, r)
    H = J.T @ J
    assert np.allclose(P[:21].sum(axis=1), H[np.triu_indices(6)])
    assert np.allclose(P[21:27].sum(axis=1), J.T @ r) and np.isclose(P[27].sum(), r @ r)


def test_plain_caratheodory_and_small_sets():
    rng = np.random.default_rng(5)
    P = rng.standard_normal((4, 40))
    u = rng.random(40) + 0.1
    Q, w, idx = caratheodory(P, u, 5)
    assert len(idx) == 5 and np.all(w >= 0) and np.allclose(Q, P[:, idx])
    assert np.allclose(P @ u, Q @ w) and np.isclose(u.sum(), w.sum())
    # nothing to do: the set is already small enough
    Q, w, idx = fast_caratheodory(P[:, :3], u[:3], 8, 10)
    assert np.array_equal(idx, np.arange(3)) and np.array_equal(w, u[:3])
    with pytest.raises(ValueError):
        caratheodory(P, u, 3)                                    # fewer than rows + 1 points cannot carry the sum
