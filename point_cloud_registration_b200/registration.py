"""Gauss-Newton driver with the reference's class surface
(reference point_cloud_registration/registration.py:10-113), backed by libpcr_b200.so.

``align`` keeps the reference signature and semantics (scan cast to float32, stop test BEFORE
the update, ``ValueError("Target is not set.")``, ``LinAlgError`` on a singular system) but runs
the whole loop on the GPU (``pcr_align``) unless ``verbose=True`` or ``device_loop=False``, in
which case the reference's host loop is replayed with one fused-kernel call per iteration."""
import os

import numpy as np

from . import _lib
from .math_tools import plus


class Registration:
    method = None      # _lib.ICP / PLANE / VPLANE / NDT, set by subclasses

    def __init__(self, max_iter=30, tol=1e-3):
        self.max_iter = max_iter
        self.tol = tol
        self._is_target_set = False
        self._ctx = None
        self._dist = None          # (rank, world_size) once attach_communicator() was called
        self._comm_ctx = None      # context that currently owns the NCCL communicator
        self._scan_generation = 0  # bumped by every upload: older UploadedScan handles are refused
        self._last_source = None   # the array object calc_H_g_e2 was last called with
        self.scan_is_presharded = False   # multi-GPU: scans passed in are already this rank's tile
        self.sort_scan = True      # Morton-sort the scan on upload in align()
        self.sort_scan_on_calc = os.environ.get("PCR_SORT_ON_CALC", "1") != "0"   # ... and in calc_H_g_e2(T, array)
        self.last_iterations = 0
        self.last_e2_trace = None

    # -- reference surface --------------------------------------------------------------
    def is_target_set(self):
        return self._is_target_set

    def _target_ready(self):
        """Called by every set_target() once self._ctx holds the new target structures.  A
        communicator attached earlier belongs to the PREVIOUS context: re-attach it, otherwise the
        scan would still be sharded while the records are no longer all-reduced."""
        self._is_target_set = True
        self._scan_generation += 1          # the resident scan lived in the previous context
        if self._dist is not None and self._comm_ctx is not self._ctx:
            self._ctx.comm_adopt(self._comm_ctx)      # an NCCL id cannot be used twice: move the communicator itself
            self._comm_ctx = self._ctx

    def set_target(self, target):
        raise NotImplementedError("set_target is not implemented.")

    def update_target(self, target):
        """Add points to the target map without starting over (the reference declares this hook and
        leaves it unimplemented, registration.py:36-43).  The old part of the map stays on the GPU, the
        new points are appended and the target structures are rebuilt there; the result is identical to
        ``set_target(concatenate(old, new))``.  Implemented by the four classes."""
        raise NotImplementedError("update_target is not implemented.")

    def linearize(self, cur_T, source):
        """The reference's abstract per-point (J, r, w) hook (registration.py:45-53); the four
        classes never materialise per-point Jacobians here either."""
        raise NotImplementedError("linearize is not implemented.")

    def calc_H_g_e2(self, cur_T, source):
        """One linearisation -> (H (6,6), g (6,), e2).  ``source`` is an (N,3) array (uploaded
        on every call, like the reference re-reads it) or a handle from :meth:`upload_scan`."""
        if not self._is_target_set:
            raise ValueError("Target is not set.")
        if not isinstance(source, UploadedScan):
            src = _lib.as_f32_points(source, "source")
            if isinstance(src, np.ndarray) and self._dist is None:
                # host array: upload and linearise in ONE library call (large copies overlap the kernels).  The same
                # array OBJECT as in the previous call (a Gauss-Newton loop on the host): its on-device ordering may
                # be reused -- the reference to it is kept so that no other array can take its identity meanwhile
                sort = self.sort_scan_on_calc
                if sort and source is self._last_source:
                    sort = 2
                self._last_source = source
                rec = self._ctx.linearize_host(self.method, np.asarray(cur_T, dtype=np.float64), self.max_dist, src, sort=sort)
                self._scan_generation += 1
                H, g, e2, self.last_inliers = _lib.record_to_H_g_e2(rec)
                return H, g, e2
            self._upload(source, sort=self.sort_scan_on_calc, T=cur_T)
        else:
            self._check_handle(source)
        rec = self._ctx.linearize(self.method, np.asarray(cur_T, dtype=np.float64), self.max_dist)
        H, g, e2, self.last_inliers = _lib.record_to_H_g_e2(rec)
        return H, g, e2

    def align(self, source, init_T=np.eye(4), verbose=False, device_loop=None):
        if self.is_target_set() is False:
            raise ValueError("Target is not set.")
        cur_T = np.asarray(init_T, dtype=np.float64)
        if not isinstance(source, UploadedScan):
            source = self.upload_scan(source, T=cur_T)
        else:
            self._check_handle(source)
        if device_loop is None:
            device_loop = not verbose
        if device_loop:
            T, iters, trace = self._ctx.align(self.method, cur_T, self.max_iter, self.tol, self.max_dist)
            self.last_iterations, self.last_e2_trace = iters, trace
            if verbose:
                for i, e2 in enumerate(trace):
                    print(f"iter {i}, error {e2}")
            return T
        trace = []
        for i in range(self.max_iter):
            H, g, e2 = self.calc_H_g_e2(cur_T, source)
            trace.append(e2)
            if verbose:
                print(f"iter {i}, error {e2}")
            dx = -np.linalg.solve(H, g)
            if np.linalg.norm(dx) < self.tol:
                break
            cur_T = plus(cur_T, dx)
        self.last_iterations, self.last_e2_trace = len(trace), np.array(trace)
        return cur_T

    # -- extensions ---------------------------------------------------------------------
    def upload_scan(self, source, sort=None, T=None):
        """Upload a scan once and get a handle usable with calc_H_g_e2 / align (avoids the
        per-call host->device copy the array form implies).  ``T``: the pose the iterations will
        start from (default identity); it only steers the on-device ordering of the scan."""
        self._upload(source, sort=self.sort_scan if sort is None else sort, T=T)
        return UploadedScan(self, self._scan_generation)

    def _check_handle(self, handle):
        if handle.owner is not self:
            raise ValueError("scan handle belongs to another registration object")
        if handle.generation != self._scan_generation:
            raise ValueError("stale scan handle: another scan was uploaded (or the target was replaced) after it was created")

    def _upload(self, source, sort, T=None):
        if self._ctx is None:
            raise ValueError("Target is not set.")
        src = _lib.as_f32_points(source, "source")               # registration.py:83
        if self._dist is not None and not self.scan_is_presharded:
            from .distributed import shard_bounds
            lo, hi = shard_bounds(src.shape[0], *self._dist)
            src = src[lo:hi]
        self._ctx.set_scan(src, sort=sort, T=T, method=self.method)
        self._scan_generation += 1
        self._last_source = None

    def attach_communicator(self, rank, world_size, unique_id):
        """Multi-GPU: this process owns one GPU and one contiguous tile of every scan; the
        29-double records are all-reduced with NCCL inside the library."""
        if self._ctx is None:
            raise ValueError("Target is not set.")
        self._ctx.comm_init_rank(world_size, rank, unique_id)
        self._dist = (rank, world_size)
        self._comm_ctx = self._ctx


class UploadedScan:
    """Handle to the scan resident on the GPU of one registration object.  Only the LATEST upload
    is resident: a handle is refused once another scan was uploaded after it."""
    __slots__ = ("owner", "generation")

    def __init__(self, owner, generation):
        self.owner = owner
        self.generation = generation
