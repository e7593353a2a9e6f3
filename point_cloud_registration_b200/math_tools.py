"""Host-side SE(3)/SO(3) helpers with the reference's names and numerical behaviour
(reference point_cloud_registration/math_tools.py:15-113).  They act on 3-vectors and 4x4
matrices only -- the per-point work lives in the CUDA kernels (csrc/pcr_terms.cuh)."""
import numpy as np

epsilon = 1e-5     # theta^2 threshold of the first-order branch in expSO3 (math_tools.py:12)


def skew(vector):
    """3x3 cross-product matrix [v]x (math_tools.py:61-64)."""
    a, b, c = vector[0], vector[1], vector[2]
    return np.array([[0, -c, b],
                     [c, 0, -a],
                     [-b, a, 0]])


def skews(vectors):
    """Batch of cross-product matrices, (N,3) -> (N,3,3) float64 (math_tools.py:34-41)."""
    v = np.asarray(vectors)
    out = np.zeros((v.shape[0], 3, 3))
    out[:, 2, 1] = v[:, 0]
    out[:, 1, 2] = -v[:, 0]
    out[:, 0, 2] = v[:, 1]
    out[:, 2, 0] = -v[:, 1]
    out[:, 1, 0] = v[:, 2]
    out[:, 0, 1] = -v[:, 2]
    return out


def skew2(v):
    """sum_i [v_i]x^T [v_i]x from the six second moments (math_tools.py:44-58)."""
    v = np.asarray(v)
    m = np.einsum('ni,nj->ij', v, v)
    return np.trace(m) * np.eye(3, dtype=m.dtype) - m


def skew_time_vector(v1, v2):
    """Row-wise [v1_i]x v2_i = v1_i x v2_i as float64 (math_tools.py:22-31)."""
    return np.cross(np.asarray(v1, dtype=np.float64), np.asarray(v2, dtype=np.float64))


def makeT(R, t):
    """(R, t) -> homogeneous matrix (math_tools.py:67-72)."""
    n = t.shape[0]
    T = np.eye(n + 1)
    T[:n, :n] = R
    T[:n, n] = t
    return T


def makeRt(T):
    """Homogeneous matrix -> (R, t) (math_tools.py:75-77)."""
    n = T.shape[0] - 1
    return T[:n, :n], T[:n, n]


def expSO3(omega):
    """SO(3) exponential (Rodrigues).  For theta^2 <= 1e-5 the reference returns the
    first-order, non-orthonormal I + [w]x (math_tools.py:80-98) -- reproduced, because the
    Gauss-Newton iterates depend on it."""
    omega = np.asarray(omega, dtype=np.float64)
    theta2 = omega.dot(omega)
    W = skew(omega)
    if theta2 <= epsilon:
        return np.eye(3) + W
    theta = np.sqrt(theta2)
    K = W / theta
    return np.eye(3) + np.sin(theta) * K + (1 - np.cos(theta)) * K.dot(K)


def plus(T, dx):
    """SE(3) right-plus: T [+] dx = T @ [[Exp(dx[3:]), dx[:3]], [0, 1]] (math_tools.py:101-108)."""
    return T @ makeT(expSO3(dx[3:]), dx[:3])


def transform_points(T, points):
    """(R @ P^T)^T + t (math_tools.py:111-113)."""
    R, t = makeRt(T)
    return (R @ points.T).T + t


def huber_weight(r, d=1.0):
    """Huber weights (math_tools.py:15-19; unused by the registration classes)."""
    r = np.asarray(r)
    w = np.ones_like(r)
    big = r > d
    w[big] = d / r[big]
    return w
