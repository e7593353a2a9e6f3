"""ctypes binding of libpcr_b200.so (C ABI declared in include/pcr_b200.h).

There is NO CPU fallback: if the shared library is missing or no sm_100 GPU is present the
calls raise.  Build the library with ``python -c "import __graft_entry__ as g; g.build()"``
or ``make -C point_cloud_registration_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpcr_b200.so")

ICP, PLANE, VPLANE, NDT = 0, 1, 2, 3
RECORD_LEN = 29

ERR_CUDA, ERR_ARG, ERR_STATE, ERR_SINGULAR, ERR_NCCL, ERR_LIMIT = -1, -2, -3, -4, -5, -6

# every symbol include/pcr_b200.h declares: name -> (restype, argtypes)
_vp, _i, _i64, _d = C.c_void_p, C.c_int, C.c_int64, C.c_double
_pi, _pi64, _pf = C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_float)
SYMBOLS = {
    "pcr_version": (_i, []),
    "pcr_device_count": (_i, [_pi]),
    "pcr_create": (_i, [_i, C.POINTER(_vp)]),
    "pcr_destroy": (_i, [_vp]),
    "pcr_last_error": (C.c_char_p, [_vp]),
    "pcr_set_target_points": (_i, [_vp, _vp, _i64]),
    "pcr_append_target_points": (_i, [_vp, _vp, _i64]),
    "pcr_export_gn_rows": (_i, [_vp, _i, _vp, _d, _vp]),
    "pcr_build_nn_index": (_i, [_vp]),
    "pcr_build_correspondence_lists": (_i, [_vp]),
    "pcr_estimate_normals": (_i, [_vp, _i]),
    "pcr_set_normals": (_i, [_vp, _vp]),
    "pcr_get_normals": (_i, [_vp, _vp]),
    "pcr_build_voxels": (_i, [_vp, _vp, _i64, _i, _d, _i, _i]),
    "pcr_get_voxel_count": (_i, [_vp, _pi64, _pi64]),
    "pcr_get_voxels": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "pcr_set_scan": (_i, [_vp, _vp, _i64, _i]),
    "pcr_set_scan_posed": (_i, [_vp, _vp, _i64, _i, _vp, _i]),
    "pcr_linearize": (_i, [_vp, _i, _vp, _d, _vp]),
    "pcr_linearize_host": (_i, [_vp, _i, _vp, _d, _vp, _i64, _i, _vp]),
    "pcr_align": (_i, [_vp, _i, _vp, _i, _d, _d, _vp, _pi, _vp]),
    "pcr_loop_begin": (_i, [_vp, _vp]),
    "pcr_loop_step_async": (_i, [_vp, _i, _i, _d, _d, _i]),
    "pcr_loop_state": (_i, [_vp, _vp, _pi, _pi, _vp, _i]),
    "pcr_knn": (_i, [_vp, _vp, _i64, _i, _vp, _vp]),
    "pcr_voxel_query": (_i, [_vp, _vp, _i64, _vp, _vp]),
    "pcr_voxel_filter": (_i, [_vp, _vp, _i64, _i, _d, _vp, _pi64]),
    "pcr_voxel_labels": (_i, [_vp, _vp, _i64, _i, _d, _vp, _vp, _pi64]),
    "pcr_comm_unique_id": (_i, [_vp]),
    "pcr_comm_init_rank": (_i, [_vp, _i, _i, _vp]),
    "pcr_comm_destroy": (_i, [_vp]),
    "pcr_comm_move": (_i, [_vp, _vp]),
    "pcr_sync_producer": (_i, [_vp, _i, _vp]),
    "pcr_last_kernel_ms": (_i, [_vp, _pf]),
    "pcr_launch_count": (_i, [_vp, _pi64]),
    "pcr_stream": (_i, [_vp, C.POINTER(_vp)]),
    "pcr_linearize_async": (_i, [_vp, _i, _vp, _d, _i]),
    "pcr_set_shell_lists": (_i, [_vp, _i]),
    "pcr_shell_list_stats": (_i, [_vp, _vp, _vp, _vp]),
    "pcr_debug_matches": (_i, [_vp, _i, _vp]),
    "pcr_set_voxel_lists": (_i, [_vp, _i]),
    "pcr_voxel_list_stats": (_i, [_vp, _pi64, _pi64]),
    "pcr_voxel_shell_stats": (_i, [_vp, _vp, _vp, _vp]),
    "pcr_index_stats": (_i, [_vp, _i, C.POINTER(_d), _pi64, _pi64, _pi64]),
    "pcr_set_path": (_i, [_vp, _i]),
    "pcr_tile_stats": (_i, [_vp, _i, C.POINTER(_d), _pi64, _pi64, _pi64]),
    "pcr_set_record_matches": (_i, [_vp, _i]),
}


class PcrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libpcr_b200 error {code}: {msg}")
        self.code = code


_lib = None
_lock = threading.Lock()


def load():
    """Load the shared library (once) and declare all prototypes.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: the CUDA extension has not been built. "
                "Run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). "
                "point_cloud_registration_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)          # AttributeError if the .so is stale
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class DevicePoints:
    """(N,3) C-contiguous float32/float64 array already resident on the GPU (anything exposing
    ``__cuda_array_interface__``, e.g. a torch CUDA tensor).  Passed to the C ABI as a raw
    device pointer -- no host round trip."""

    def __init__(self, obj, name="points"):
        if isinstance(obj, DevicePoints):
            obj = obj.obj
        cai = obj.__cuda_array_interface__
        shape, typestr = tuple(cai["shape"]), cai["typestr"]
        if len(shape) != 2 or shape[1] != 3:
            raise ValueError(f"{name} must have shape (N, 3), got {shape}")
        if typestr not in ("<f4", "<f8"):
            raise ValueError(f"{name}: device arrays must be float32 or float64, got {typestr}")
        if cai.get("strides") not in (None, (3 * int(typestr[2]), int(typestr[2]))):
            raise ValueError(f"{name}: device arrays must be C-contiguous")
        self.obj = obj                      # keeps the memory alive
        self.stream = cai.get("stream")     # producer stream (interface v3) or None
        self.ptr = int(cai["data"][0])
        self.shape = shape
        self.dtype = np.dtype(typestr)

    @property
    def ctypes_ptr(self):
        return C.c_void_p(self.ptr)


def is_device_array(a):
    return isinstance(a, DevicePoints) or hasattr(a, "__cuda_array_interface__")


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, DevicePoints):
        return a.ctypes_ptr
    return a.ctypes.data_as(C.c_void_p)


def as_f32_points(a, name="points"):
    """(N,3) C-contiguous float32 view/copy of an array-like (never mutates the input); GPU
    arrays (``__cuda_array_interface__``) are passed through as :class:`DevicePoints`."""
    if isinstance(a, DevicePoints):
        return a
    if is_device_array(a):
        d = DevicePoints(a, name)
        if d.dtype != np.float32:
            raise ValueError(f"{name}: device arrays must be float32 here")
        return d
    a = np.asarray(a)
    if a.ndim != 2 or a.shape[1] != 3:
        raise ValueError(f"{name} must have shape (N, 3), got {a.shape}")
    return np.ascontiguousarray(a, dtype=np.float32)


class Context:
    """One libpcr_b200 context = one GPU + one stream + the device-resident structures."""

    def __init__(self, device=None):
        self._lib = load()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0")) if "PCR_USE_LOCAL_RANK" in os.environ else 0
        h = C.c_void_p()
        rc = self._lib.pcr_create(int(device), C.byref(h))
        if rc != 0:
            raise PcrError(rc, (self._lib.pcr_last_error(None) or b"").decode())
        self._h = h
        self.device = int(device)

    # -- plumbing -----------------------------------------------------------------------
    def _check(self, rc):
        if rc == 0:
            return
        msg = (self._lib.pcr_last_error(self._h) or b"").decode()
        if rc == ERR_SINGULAR:
            raise np.linalg.LinAlgError("Singular matrix")
        raise PcrError(rc, msg)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.pcr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _order_after_producer(self, *arrays):
        """GPU arrays are read on this context's own non-blocking stream: wait for whoever produced
        them first (their __cuda_array_interface__ stream if they name one, else the whole device)."""
        for a in arrays:
            if isinstance(a, DevicePoints):
                if a.stream is None:
                    self._check(self._lib.pcr_sync_producer(self._h, 0, None))
                else:
                    self._check(self._lib.pcr_sync_producer(self._h, 1, C.c_void_p(int(a.stream))))

    # -- target side --------------------------------------------------------------------
    def set_target_points(self, pts_f32):
        self._order_after_producer(pts_f32)
        self._check(self._lib.pcr_set_target_points(self._h, _ptr(pts_f32), pts_f32.shape[0]))

    def append_target_points(self, pts_f32):
        self._order_after_producer(pts_f32)
        self._check(self._lib.pcr_append_target_points(self._h, _ptr(pts_f32), pts_f32.shape[0]))

    def build_voxels_from_target(self, voxel_size, min_points, with_icov=True):
        self._check(self._lib.pcr_build_voxels(self._h, None, 0, 0, float(voxel_size), int(min_points), int(bool(with_icov))))

    def export_gn_rows(self, method, T, max_dist, n_scan):
        T = np.ascontiguousarray(T, dtype=np.float64)
        rows = np.empty((n_scan, 28), dtype=np.float64)
        self._check(self._lib.pcr_export_gn_rows(self._h, int(method), _ptr(T), float(max_dist), _ptr(rows)))
        return rows

    def build_nn_index(self):
        self._check(self._lib.pcr_build_nn_index(self._h))

    def build_correspondence_lists(self):
        self._check(self._lib.pcr_build_correspondence_lists(self._h))

    def estimate_normals(self, k):
        self._check(self._lib.pcr_estimate_normals(self._h, int(k)))

    def set_normals(self, nrm_f32):
        self._order_after_producer(nrm_f32)
        self._check(self._lib.pcr_set_normals(self._h, _ptr(nrm_f32)))

    def get_normals(self, n):
        out = np.empty((n, 3), dtype=np.float32)
        self._check(self._lib.pcr_get_normals(self._h, _ptr(out)))
        return out

    @staticmethod
    def _f32_or_f64(pts):
        if is_device_array(pts):
            d = DevicePoints(pts)
            return d, int(d.dtype == np.float64)
        pts = np.asarray(pts)
        if pts.ndim != 2 or pts.shape[1] != 3:
            raise ValueError(f"points must have shape (N, 3), got {pts.shape}")
        if pts.dtype == np.float64:
            return np.ascontiguousarray(pts), 1
        return np.ascontiguousarray(pts, dtype=np.float32), 0

    def build_voxels(self, pts, voxel_size, min_points, with_icov=True):
        arr, is64 = self._f32_or_f64(pts)
        self._order_after_producer(arr)
        self._check(self._lib.pcr_build_voxels(self._h, _ptr(arr), arr.shape[0], is64, float(voxel_size),
                                               int(min_points), int(bool(with_icov))))

    def voxel_count(self):
        a, b = C.c_int64(), C.c_int64()
        self._check(self._lib.pcr_get_voxel_count(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def get_voxels(self, with_icov=True):
        n, _ = self.voxel_count()
        mean = np.empty((n, 3))
        cov = np.empty((n, 3, 3))
        norm = np.empty((n, 3))
        icov = np.empty((n, 3, 3)) if with_icov else None
        count = np.empty(n, dtype=np.int64)
        self._check(self._lib.pcr_get_voxels(self._h, _ptr(mean), _ptr(cov), _ptr(norm), _ptr(icov), _ptr(count)))
        return mean, cov, norm, icov, count

    # -- scan side ------------------------------------------------------------------------
    def set_scan(self, pts_f32, sort=True, T=None, method=-1):
        """sort: True/1 re-order on the device (by the correspondence-grid cell of the points posed
        by T when that grid exists, else along a Morton curve); False/0 keep order; -1 keep order,
        caller promises spatial coherence."""
        Tp = None if T is None else np.ascontiguousarray(T, dtype=np.float64)
        self._order_after_producer(pts_f32)
        self._check(self._lib.pcr_set_scan_posed(self._h, _ptr(pts_f32), pts_f32.shape[0], int(sort), _ptr(Tp), int(method)))

    def set_voxel_lists(self, enable):
        self._check(self._lib.pcr_set_voxel_lists(self._h, int(bool(enable))))

    def voxel_list_stats(self):
        a, b = C.c_int64(), C.c_int64()
        self._check(self._lib.pcr_voxel_list_stats(self._h, C.byref(a), C.byref(b)))
        return dict(band_cells=a.value, entries=b.value)

    def voxel_shell_stats(self):
        a, b, m = C.c_int64(), C.c_int64(), C.c_double()
        self._check(self._lib.pcr_voxel_shell_stats(self._h, C.byref(a), C.byref(b), C.byref(m)))
        return dict(band_cells=a.value, entries=b.value, margin_cells=m.value, bytes=b.value * 17)

    def set_shell_lists(self, enable):
        self._check(self._lib.pcr_set_shell_lists(self._h, int(bool(enable))))

    def shell_list_stats(self):
        a, b, m = C.c_int64(), C.c_int64(), C.c_double()
        self._check(self._lib.pcr_shell_list_stats(self._h, C.byref(a), C.byref(b), C.byref(m)))
        return dict(band_cells=a.value, entries=b.value, margin_cells=m.value, bytes=b.value * 17)

    def set_path(self, path):
        """'lists' (default) or 'tile' (tile-stream kernel); call before set_target builds its structures."""
        self._check(self._lib.pcr_set_path(self._h, {"lists": 0, "tile": 1}[path] if isinstance(path, str) else int(path)))

    def set_record_matches(self, enable):
        self._check(self._lib.pcr_set_record_matches(self._h, int(bool(enable))))

    def tile_stats(self, which=0):
        h, nc, no, nb = C.c_double(), C.c_int64(), C.c_int64(), C.c_int64()
        self._check(self._lib.pcr_tile_stats(self._h, int(which), C.byref(h), C.byref(nc), C.byref(no), C.byref(nb)))
        return dict(cell_edge=h.value, cells=nc.value, occupied=no.value, bytes=nb.value)

    def debug_matches(self, n_scan, which=0):
        """Caller indices matched by the last linearisation, per resident scan point (storage order)."""
        idx = np.empty(n_scan, dtype=np.int64)
        self._check(self._lib.pcr_debug_matches(self._h, int(which), _ptr(idx)))
        return idx

    def linearize(self, method, T, max_dist):
        T = np.ascontiguousarray(T, dtype=np.float64)
        out = np.empty(RECORD_LEN)
        self._check(self._lib.pcr_linearize(self._h, int(method), _ptr(T), float(max_dist), _ptr(out)))
        return out

    def linearize_host(self, method, T, max_dist, pts_f32, sort=True):
        """One linearisation of a HOST scan in one call (copy chunks overlap the kernels); the scan becomes
        the resident scan."""
        T = np.ascontiguousarray(T, dtype=np.float64)
        out = np.empty(RECORD_LEN)
        self._check(self._lib.pcr_linearize_host(self._h, int(method), _ptr(T), float(max_dist), _ptr(pts_f32), pts_f32.shape[0],
                                                 int(sort), _ptr(out)))
        return out

    def linearize_async(self, method, T, max_dist, reps=1):
        T = np.ascontiguousarray(T, dtype=np.float64)
        self._check(self._lib.pcr_linearize_async(self._h, int(method), _ptr(T), float(max_dist), int(reps)))

    def align(self, method, T0, max_iter, tol, max_dist):
        T0 = np.ascontiguousarray(T0, dtype=np.float64)
        T = np.empty((4, 4))
        iters = C.c_int(0)
        trace = np.zeros(max(int(max_iter), 1))
        self._check(self._lib.pcr_align(self._h, int(method), _ptr(T0), int(max_iter), float(tol), float(max_dist),
                                        _ptr(T), C.byref(iters), _ptr(trace)))
        return T, iters.value, trace[:min(iters.value, int(max_iter))]

    def loop_begin(self, T0):
        T0 = np.ascontiguousarray(T0, dtype=np.float64)
        self._check(self._lib.pcr_loop_begin(self._h, _ptr(T0)))

    def loop_step_async(self, method, max_iter, tol, max_dist, reps=1):
        self._check(self._lib.pcr_loop_step_async(self._h, int(method), int(max_iter), float(tol), float(max_dist), int(reps)))

    def loop_state(self, trace_cap=0):
        T = np.empty((4, 4))
        iters, done = C.c_int(0), C.c_int(0)
        trace = np.zeros(max(trace_cap, 1))
        self._check(self._lib.pcr_loop_state(self._h, _ptr(T), C.byref(iters), C.byref(done), _ptr(trace), int(trace_cap)))
        return T, iters.value, done.value, trace[:min(iters.value, trace_cap)]

    # -- utilities ------------------------------------------------------------------------
    def knn(self, q_f32, k):
        m = q_f32.shape[0]
        self._order_after_producer(q_f32)
        dist = np.empty((m, k), dtype=np.float32)
        idx = np.empty((m, k), dtype=np.int64)
        self._check(self._lib.pcr_knn(self._h, _ptr(q_f32), m, int(k), _ptr(dist), _ptr(idx)))
        return dist, idx

    def voxel_query(self, q_f32):
        m = q_f32.shape[0]
        self._order_after_producer(q_f32)
        vidx = np.empty(m, dtype=np.int64)
        dist = np.empty(m, dtype=np.float64)
        self._check(self._lib.pcr_voxel_query(self._h, _ptr(q_f32), m, _ptr(vidx), _ptr(dist)))
        return dist, vidx

    def voxel_filter(self, pts, voxel_size):
        arr, is64 = self._f32_or_f64(pts)
        self._order_after_producer(arr)
        out = np.empty((arr.shape[0], 3), dtype=np.float32)
        n_out = C.c_int64(0)
        self._check(self._lib.pcr_voxel_filter(self._h, _ptr(arr), arr.shape[0], is64, float(voxel_size), _ptr(out),
                                               C.byref(n_out)))
        return out[:n_out.value].copy()

    def voxel_labels(self, pts, voxel_size):
        """(labels (N,) int64, voxel coordinates (V,3) int32): voxel membership of every point."""
        arr, is64 = self._f32_or_f64(pts)
        self._order_after_producer(arr)
        labels = np.empty(arr.shape[0], dtype=np.int64)
        coords = np.empty((arr.shape[0], 3), dtype=np.int32)
        nv = C.c_int64(0)
        self._check(self._lib.pcr_voxel_labels(self._h, _ptr(arr), arr.shape[0], is64, float(voxel_size), _ptr(labels), _ptr(coords),
                                               C.byref(nv)))
        return labels, coords[:nv.value].copy()

    # -- multi GPU ------------------------------------------------------------------------
    @staticmethod
    def comm_unique_id():
        lib = load()
        buf = C.create_string_buffer(128)
        rc = lib.pcr_comm_unique_id(buf)
        if rc != 0:
            raise PcrError(rc, (lib.pcr_last_error(None) or b"").decode())
        return bytes(buf.raw)

    def comm_init_rank(self, nranks, rank, uid):
        buf = C.create_string_buffer(bytes(uid), 128)
        self._check(self._lib.pcr_comm_init_rank(self._h, int(nranks), int(rank), buf))

    def comm_adopt(self, other):
        """Take over the communicator of another context on the same GPU (see pcr_comm_move)."""
        self._check(self._lib.pcr_comm_move(self._h, other._h))

    def comm_destroy(self):
        self._check(self._lib.pcr_comm_destroy(self._h))

    # -- instrumentation --------------------------------------------------------------------
    def last_kernel_ms(self):
        v = C.c_float(0)
        self._check(self._lib.pcr_last_kernel_ms(self._h, C.byref(v)))
        return v.value

    def launch_count(self):
        v = C.c_int64(0)
        self._check(self._lib.pcr_launch_count(self._h, C.byref(v)))
        return v.value

    def stream(self):
        v = C.c_void_p()
        self._check(self._lib.pcr_stream(self._h, C.byref(v)))
        return v.value

    def index_stats(self, which=0):
        h, nc, nb, npts = C.c_double(), C.c_int64(), C.c_int64(), C.c_int64()
        self._check(self._lib.pcr_index_stats(self._h, int(which), C.byref(h), C.byref(nc), C.byref(nb), C.byref(npts)))
        return dict(cell_edge=h.value, cells=nc.value, bricks=nb.value, points=npts.value)


def record_to_H_g_e2(rec):
    """29-double record -> (H (6,6), g (6,), e2, inlier count)."""
    H = np.zeros((6, 6))
    H[np.triu_indices(6)] = rec[:21]
    H = H + np.triu(H, 1).T
    return H, rec[21:27].copy(), float(rec[27]), int(round(rec[28]))
