"""Host replay of the shell-list scan on the C2 workload (real B-01, the GPU's cell edge and list margin):
per Gauss-Newton iteration, the groups of four list entries a query streams, the lock-step cost per warp row
of 32 cell-ordered queries, and how many queries the lists cannot settle.  CPU only (tests/hostsim).

    python tools/host_list_study.py [cell_edge=0.39728620648384094] [margin_cells=3.0]
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import bench                                                    # noqa: E402
from oracle import pcr_oracle as orc                             # noqa: E402
from hostsim.build import build as build_hostsim                 # noqa: E402


def main():
    h = float(sys.argv[1]) if len(sys.argv) > 1 else 0.39728620648384094
    margin = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
    lib = C.CDLL(build_hostsim())
    lib.hs_grid_build.restype = C.c_void_p
    lib.hs_grid_build.argtypes = [C.c_void_p, C.c_int64, C.c_double]
    lib.hs_shell_build.restype = C.c_int64
    lib.hs_shell_build.argtypes = [C.c_void_p, C.c_double]
    lib.hs_shell_study.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_void_p]
    wl = bench.WORKLOADS["c2"]
    target, note = bench.host_cloud("c2", wl)
    scan, _ = bench.host_scan("c2", wl, target)
    print(note, len(target), "cell edge", h, "margin", margin)
    o = bench.oracle_object(wl, target)
    Ts = []
    T = np.eye(4)
    for it in range(bench.MAX_ITER):
        Ts.append(T.copy())
        H, g, e2 = o.calc_H_g_e2(T, scan)
        dx = -np.linalg.solve(H, g)
        if np.linalg.norm(dx) < bench.TOL:
            break
        T = orc.se3_plus(T, dx)
    tgt = np.ascontiguousarray(target, dtype=np.float32)
    g = lib.hs_grid_build(tgt.ctypes.data, len(tgt), h)
    entries = lib.hs_shell_build(g, margin)
    print("list entries", entries)
    for it, T in enumerate(Ts):
        q = (scan.astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32)
        if it == 0:                                             # the upload orders the scan by the cell of the FIRST pose
            key = np.floor((q - tgt.min(0)) / h).astype(np.int64)
            order = np.lexsort((key[:, 0], key[:, 1], key[:, 2]))
        q = np.ascontiguousarray(q[order])
        out = np.zeros(4)
        lib.hs_shell_study(g, q.ctypes.data, len(q), float(bench.MAX_DIST), out.ctypes.data)
        n = len(q)
        print(f"iteration {it + 1}: groups/query {out[1] / n:6.1f}  lock-step groups/query {out[0] / (n / 32) :6.1f}  "
              f"list exhausted {100 * out[2] / n:5.2f} %  no list {100 * out[3] / n:5.2f} %")


if __name__ == "__main__":
    main()
