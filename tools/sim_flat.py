"""Host lock-step model of the warp loop of linearize_flat_kernel vs the nested per-lane search
(development aid: counts candidate evaluations and warp rounds; no GPU needed)."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "hostsim"))
import build as hbuild
from point_cloud_registration_b200 import datasets as ds
from oracle import pcr_oracle as orc

lib = C.CDLL(hbuild.build())
lib.hs_grid_build.restype = C.c_void_p
lib.hs_grid_build.argtypes = [C.c_void_p, C.c_int64, C.c_double]
lib.hs_grid_cells.restype = C.c_int64
lib.hs_grid_cells.argtypes = [C.c_void_p]
lib.hs_flat_warp_sim.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
lib.hs_nn.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_void_p, C.c_void_p]
ptr = lambda a: a.ctypes.data_as(C.c_void_p)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
ppc = float(sys.argv[2]) if len(sys.argv) > 2 else 24.0
target = ds.make_urban_slab(n, seed=1)
scan = ds.perturb_scan(target, seed=0)
# Morton order of the scan (as pcr_set_scan does)
lo = scan.min(0); ext = (scan.max(0) - lo).max()
q = np.minimum(((scan - lo) * (1023.999 / ext)).astype(np.uint32), 1023)
def spread(v):
    v = v.astype(np.uint64) & 0x3ff
    v = (v | (v << 16)) & 0x030000ff; v = (v | (v << 8)) & 0x0300f00f; v = (v | (v << 4)) & 0x030c30c3; v = (v | (v << 2)) & 0x09249249
    return v
key = spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)
scan = scan[np.argsort(key, kind="stable")]
# cell edge: iterate like build_point_grid until ppc is near the target
h = 0.37
for _ in range(6):
    g = lib.hs_grid_build(ptr(target), len(target), h)
    cells = lib.hs_grid_cells(g)
    got = len(target) / cells
    if abs(got / ppc - 1) < 0.15: break
    h *= (ppc / got) ** 0.5
print(f"n={n} h={h:.3f} cells={cells} ppc={got:.1f}")
# trajectory of the oracle's PlaneICP align gives the iterates
o = orc.OracleICP(max_iter=30, max_dist=2.0, tol=1e-3); o.set_target(target)
trace = []; o.align(scan, np.eye(4), trace=trace)
threads = max(32, int(n / 10.5) // 32 * 32)
CSTEP, CCAND, CROUND = 60.0, 10.0, 12.0
sel = [0, 1, 3, len(trace) - 1]
for it in sel:
    tr = trace[it]
    T = tr["T"].astype(np.float32)
    qs = np.ascontiguousarray((scan @ T[:3, :3].T + T[:3, 3]).astype(np.float32))
    idx2 = np.empty(len(qs), np.int64); dist2 = np.empty(len(qs), np.float32)
    lib.hs_nn(g, ptr(qs), len(qs), 2.0, ptr(idx2), ptr(dist2))
    first = True
    for ch in (8, 16, 32):
        for tau in (1, 4, 8, 16):
            idx = np.empty(len(qs), np.int64); dist = np.empty(len(qs), np.float32); st = np.zeros(8)
            lib.hs_flat_warp_sim(g, ptr(qs), len(qs), 2.0, ch, threads, ptr(idx), ptr(dist), ptr(st), tau, 0 if first else 1)
            bad = int(np.sum(dist != dist2))
            rounds, sumB, sumA, candF, nestMax, nestTot, laneA, nA = st[:8]
            rows = len(qs) / 32
            if first:
                print(f"it{it}: meanNN={np.mean(dist[np.isfinite(dist)]):.3f} nested: cand/query={nestTot/len(qs):.0f} rowmax-cands/row={nestMax/rows:.0f}")
                first = False
            cost = (sumA * CSTEP + sumB * CCAND + rounds * CROUND) / rows
            print(f"   ch={ch:2d} tau={tau:2d} mism={bad} rounds/row={rounds/rows:.0f} B/row={sumB/rows:.0f} (eff {candF/(sumB*32):.2f}) A/row={sumA/rows:.1f} (eff {laneA/(sumA*32):.2f}, runs {nA/rows:.0f}) model warp-instr/row={cost:.0f}")
