"""Exact Gauss-Newton coresets (reference point_cloud_registration/caratheodory.py:23-138; method:
K. Koide, "Exact Point Cloud Downsampling for Fast and Accurate Global Trajectory Optimization",
arXiv:2307.02948; fast variant after Maalouf et al., "Fast and Accurate Least-Mean-Squares Solvers").

A point-to-plane linearisation is a SUM over correspondences of 28-vectors (upper triangle of
J_i^T J_i, J_i r_i, r_i^2).  By Caratheodory's theorem that sum equals a POSITIVELY WEIGHTED sum of at
most 29 of them, so a scan can be replaced -- exactly, at the pose of the linearisation -- by a handful
of weighted points.  ``create_gn_set`` builds the 28-vectors from (J, r) like the reference;
``gn_set_from_registration`` takes them straight from the GPU (``pcr_export_gn_rows``: the accumulate
pass' per-point terms); ``caratheodory`` / ``fast_caratheodory`` select the weighted subset (host
NumPy, as in the reference: the selection works on a few thousand cluster means, not on the scan).
"""
import numpy as np


def create_gn_set(J, r):
    """(N, D) Jacobians and (N,) residuals -> (D(D+1)/2 + D + 1, N) matrix whose column i stacks the
    upper triangle of J_i^T J_i, J_i r_i and r_i^2 (caratheodory.py:118-138)."""
    J = np.asarray(J, dtype=np.float64)
    r = np.asarray(r, dtype=np.float64)
    iu, ju = np.triu_indices(J.shape[1])
    return np.concatenate([(J[:, iu] * J[:, ju]).T, (J * r[:, None]).T, (r * r)[None, :]], axis=0)


def gn_set_from_registration(reg, cur_T, source):
    """The same matrix for a PlaneICP / VPlaneICP object at pose ``cur_T``, computed on the GPU: one
    column per scan point in the caller's order, zero columns for points without correspondence."""
    from . import _lib
    if reg.method not in (_lib.PLANE, _lib.VPLANE):
        raise ValueError("Gauss-Newton rows exist for the scalar-residual methods only (PlaneICP, VPlaneICP)")
    if not reg.is_target_set():
        raise ValueError("Target is not set.")
    reg._upload(source, sort=False)                       # storage order = caller order
    n = _lib.as_f32_points(source, "source").shape[0]
    return reg._ctx.export_gn_rows(reg.method, np.asarray(cur_T, dtype=np.float64), reg.max_dist, n).T


def _affine_dependence(P):
    """v != 0 with P v = 0 and sum(v) = 0 (exists whenever P has more than rows + 1 columns)."""
    D = P[:, 1:] - P[:, :1]
    _, _, Vt = np.linalg.svd(D, full_matrices=True)
    tail = Vt[-1]                                         # right-singular vector of the smallest singular value (0 here)
    return np.concatenate([[-tail.sum()], tail])


def caratheodory(P, u, N_target):
    """Positive weights w and indices idx with  P[:, idx] @ w == P @ u  and  len(idx) <= N_target
    (N_target >= rows + 1).  Returns (P[:, idx], w, idx)  (caratheodory.py:35-60)."""
    P = np.asarray(P, dtype=np.float64)
    w = np.asarray(u, dtype=np.float64).copy()
    idx = np.arange(P.shape[1])
    if P.shape[1] <= N_target:
        return P, w, idx
    if N_target < P.shape[0] + 1:
        raise ValueError("N_target must be at least rows + 1")
    while P.shape[1] > N_target:
        v = _affine_dependence(P)
        with np.errstate(divide="ignore", invalid="ignore"):
            ratio = np.where(v != 0.0, w / v, np.inf)
        j = int(np.argmin(np.abs(ratio)))                 # the smallest step that empties one weight keeps all others >= 0
        w = w - ratio[j] * v
        keep = np.ones(P.shape[1], dtype=bool)
        keep[j] = False
        P, w, idx = P[:, keep], w[keep], idx[keep]
    return P, w, idx


def fast_caratheodory(P, u, k, N_target):
    """The same guarantee in near-linear time: split the columns into k contiguous clusters, run
    ``caratheodory`` on the k weighted cluster means, keep only the clusters it selects (their points
    re-weighted by the cluster's new / old weight) and repeat until at most N_target columns are left
    (caratheodory.py:62-116).  Returns (P[:, idx], w, idx)."""
    P = np.asarray(P, dtype=np.float64)
    w = np.asarray(u, dtype=np.float64).copy()
    idx = np.arange(P.shape[1])
    rows = P.shape[0]
    while P.shape[1] > N_target:
        n = P.shape[1]
        kk = min(int(k), n)
        bounds = np.linspace(0, n, kk + 1).astype(np.int64)
        starts, sizes = bounds[:-1], np.diff(bounds)
        cw = np.add.reduceat(w, starts)                                   # cluster weights
        cm = np.add.reduceat(P * w, starts, axis=1) / cw                  # weighted cluster means
        n_keep = rows + 1
        if n_keep * sizes.max() < N_target:                              # room for more clusters: stop earlier, keep more points
            n_keep = N_target // sizes.max()
        _, cw_new, chosen = caratheodory(cm, cw, n_keep)
        cols = np.concatenate([np.arange(starts[c], starts[c] + sizes[c]) for c in chosen])
        scale = np.repeat(cw_new / cw[chosen], sizes[chosen])
        P, w, idx = P[:, cols], w[cols] * scale, idx[cols]
    return P, w, idx
