#!/usr/bin/env python3
"""Benchmark of the registration hot path: Gauss-Newton ICP iterations per second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c1|c2|c2i|c3|c4|c5]

A "step" is ONE Gauss-Newton iteration of the workload's registration class: SE(3) transform of
the scan, exact correspondence search, residual + Jacobian, reduction to the 6x6 normal
equations, 6x6 solve, stop test and SE(3) update.  Prints ONE JSON line (rank 0).

Workloads (BASELINE.json `configs`):
  c1   ICP, 10k-point unit cube of the reference's tests (tests/test_icp.py shape, s = 0.1)    [context line]
  c2   PlaneICP k=15 on the reference's data/B-01.pcd (1,193,011 points, data/b01_xyz.npz),
       scan = full perturbed copy                                          [default, N = 1]
  c2i  ICP on the same cloud and scan                                                          [context line]
  c3   VPlaneICP voxel 0.5 m, 10M synthetic points                                             [context line]
  c4   NDT voxel 1.0 m, 10M synthetic points                                                   [context line]
  c5   PlaneICP on ONE fixed 100M-point workload: scan tile-sharded over the N GPUs (100M / N scan
       points per GPU), 100M-point target replicated, one 29-double NCCL all-reduce per iteration
       -- strong scaling                                                               [default, N > 1]

Timed regions (b200 arm)
  value     device-resident: scan + target structures in HBM, K iterations of the on-device loop
            (correspond kernel + accumulate kernel, whose last block also solves and updates T)
            cycling through the align() trajectory; every iteration timed with its own CUDA-event
            pair on the library's stream, L2 flushed (512 MiB memset) between iterations, outside the
            event pairs; max over ranks.
  warm_l2   the same loop enqueued back to back without flushing (what align() really does).
  e2e       through the drop-in Python class with HOST buffers: every step calls
            calc_H_g_e2(T, host_scan) -> H2D copy of the scan from pinned memory, kernels, D2H of the
            29-double record, then the host 6x6 solve and update (wall clock, synchronous API).
  cpu_baseline / --impl reference: the CPU oracle (NumPy + scipy cKDTree restatement of the
            reference, oracle/pcr_oracle.py, pinned to the live reference by tests/golden) on the host
            cores, same workload or a stated sample of it.
Every line carries its own parity numbers: transform_err_vs_ref (GPU align vs oracle align on the same
arrays, bound 1e-4) and, for N > 1, multi_vs_single_T_err (the sharded solve vs rank 0 alone on the whole
scan) plus n1_same_workload (rank 0 alone, same 100M workload) so that the scaling efficiency can be
read off a single line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RESULT_OUT = sys.stdout
METRIC = "ICP iterations/sec"
UNIT = "iterations/s"
C5_TOTAL = int(os.environ.get("PCR_BENCH_C5_N", "100000000"))

WORKLOADS = {
    "c1": dict(cls="ICP", kw={}, n=10_000, bytes_per_point=24, seed=42, scan_seed=0,
               desc="ICP, 10k-pt unit cube of the reference's tests (s = 0.1)"),
    "c2": dict(cls="PlaneICP", kw=dict(k=15), n=1_193_011, bytes_per_point=36, seed=1, scan_seed=0,
               desc="PlaneICP k=15, data/B-01.pcd (1,193,011 pts), scan = full perturbed copy"),
    "c2i": dict(cls="ICP", kw={}, n=1_193_011, bytes_per_point=24, seed=1, scan_seed=0,
                desc="ICP, data/B-01.pcd (1,193,011 pts), scan = full perturbed copy"),
    "c3": dict(cls="VPlaneICP", kw=dict(voxel_size=0.5), n=10_000_000, bytes_per_point=36, seed=10, scan_seed=0,
               desc="VPlaneICP voxel 0.5 m, 10M-pt synthetic slab"),
    "c4": dict(cls="NDT", kw=dict(voxel_size=1.0), n=10_000_000, bytes_per_point=48, seed=10, scan_seed=0,
               desc="NDT voxel 1.0 m, 10M-pt synthetic slab"),
    "c5": dict(cls="PlaneICP", kw=dict(k=15), n=C5_TOTAL, bytes_per_point=36, seed=100, scan_seed=0,
               desc="PlaneICP k=15, ONE fixed workload: scan tile-sharded over the GPUs, target replicated"),
}
MAX_DIST, MAX_ITER, TOL = 2.0, 30, 1e-3
TILES_PER_RANK = int(os.environ.get("PCR_BENCH_TILES", "16"))   # c5, N > 1: Morton tiles per rank, dealt round-robin
if os.environ.get("PCR_BENCH_TEST_N"):          # test hook (tests/test_bench_contract.py): shrink every workload
    for _w in WORKLOADS.values():
        _w["n"] = int(os.environ["PCR_BENCH_TEST_N"])

_CLOUDS = {}                                     # (kind, n, seed) -> host cloud, shared by the workloads of one run


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception as e:              # nvidia-smi missing: report it, do not fail the bench
            log("clock sampler unavailable:", e)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v == "Active":
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------
# workloads on the host (shared by the GPU arm, its parity legs and the reference arm)
# ----------------------------------------------------------------------------------------------
def host_cloud(wl_name, wl, n=None):
    """(target, data description).  c2 / c2i: the reference's B-01 when data/b01_xyz.npz is there."""
    from point_cloud_registration_b200 import datasets as ds
    n = wl["n"] if n is None else n
    if wl_name in ("c2", "c2i") and n == ds.B01_POINTS:
        key = ("b01", n, 0)
        if key not in _CLOUDS:
            b = ds.load_b01()
            _CLOUDS[key] = (b, "data/B-01.pcd of the reference (xyz, CC BY 4.0: data/README.md); scan synthetic "
                               "(rigid motion + noise of SURVEY 8d)") if b is not None else None
        if _CLOUDS[key] is not None:
            return _CLOUDS[key]
    if wl_name == "c1":
        key = ("cube", n, wl["seed"])
        if key not in _CLOUDS:
            tgt, _ = ds.unit_cube_case(n, scale=0.1, seed=wl["seed"])
            _CLOUDS[key] = (tgt.astype(np.float32), "synthetic (unit cube of the reference's tests)")
        return _CLOUDS[key]
    key = ("slab", n, wl["seed"])
    if key not in _CLOUDS:
        tgt = ds.make_urban_slab(n, seed=wl["seed"])
        ds.assert_no_key_collisions(tgt, (0.5, 1.0))        # the reference's voxel hash must be injective here
        note = "synthetic (urban slab at B-01's surface density)"
        if wl_name in ("c2", "c2i"):
            note += "; NOT the reference's B-01 (data/b01_xyz.npz missing, or the workload was resized)"
        _CLOUDS[key] = (tgt, note)
    return _CLOUDS[key]


def host_scan(wl_name, wl, target, sample=None, lever_arm=False):
    from point_cloud_registration_b200 import datasets as ds
    if wl_name == "c1":
        _, src = ds.unit_cube_case(len(target), scale=0.1, seed=wl["seed"])
        return src.astype(np.float32), (0.01, 0.02, 0.03)
    so3 = (0.01, -0.02, 0.03)
    if lever_arm:
        c = 0.5 * (target.max(0).astype(np.float64) + target.min(0))
        radius = float(np.linalg.norm(0.5 * (target.max(0).astype(np.float64) - target.min(0))[:2]))
        so3 = ds.lever_arm_so3(so3, radius)
        del c
    return ds.perturb_scan(target, so3=so3, seed=wl["scan_seed"], num_points=sample), so3


def oracle_object(wl, target, normals=None):
    from oracle import pcr_oracle as orc
    cls = {"PlaneICP": orc.OraclePlaneICP, "VPlaneICP": orc.OracleVPlaneICP, "NDT": orc.OracleNDT, "ICP": orc.OracleICP}[wl["cls"]]
    o = cls(max_iter=MAX_ITER, max_dist=MAX_DIST, tol=TOL, **wl["kw"])
    if wl["cls"] == "PlaneICP" and normals is not None:
        o.set_target(target, index=orc.NNIndex(target), normals=normals)
    else:
        o.set_target(target)
    return o


def time_oracle_steps(o, scan_f32, Ts, steps, warmup):
    """Mean wall time of calc_H_g_e2 + solve + update over `steps` calls cycling through the
    iterate sequence Ts (all host threads: scipy cKDTree workers=-1, BLAS default)."""
    from oracle import pcr_oracle as orc
    times = []
    for i in range(warmup + steps):
        T = Ts[i % len(Ts)]
        t0 = time.perf_counter()
        H, g, e2 = o.calc_H_g_e2(T, scan_f32)
        dx = -np.linalg.solve(H, g)
        if np.linalg.norm(dx) >= TOL:
            orc.se3_plus(T, dx)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return float(np.mean(times))


def workload_config(wl_name, wl, world, n_target, n_scan, **extra):
    cfg = {"workload": f"{wl_name}: {wl['desc']}", "registration": wl["cls"], **wl["kw"], "scan_points": n_scan,
           "target_points": n_target, "max_dist": MAX_DIST, "tol": TOL, "max_iter": MAX_ITER,
           "parallelism": (f"scan cut into {world * TILES_PER_RANK} spatial (Morton) tiles dealt round-robin over {world} GPUs, target replicated"
                           if world > 1 else "single GPU"),
           "l2": "flushed between timed iterations (512 MiB memset outside the event pairs)"}
    cfg.update(extra)
    return cfg


# ----------------------------------------------------------------------------------------------
# reference arm: the reference's CPU path (oracle port) on this box's host cores
# ----------------------------------------------------------------------------------------------
def run_reference(args, wl_name, wl, world, rank):
    if rank != 0:
        return
    n_total = wl["n"]
    cores = os.cpu_count()
    bounded = n_total > 2_000_000
    # bounded target for the kNN-normal / voxel setup on the CPU: 4M points of the same scene family
    n_t = min(4_000_000, n_total) if bounded else n_total
    target, data_note = host_cloud(wl_name, wl, n_t)
    scan_all, so3 = host_scan(wl_name, wl, target, lever_arm=(wl_name == "c5"))
    t0 = time.perf_counter()
    o = oracle_object(wl, target)
    setup_s = time.perf_counter() - t0
    # bounded sample: calibrate the per-point cost on 50k scan points, then size the sample so that
    # the whole --steps K --warmup W run stays near REF_BUDGET_S seconds of CPU work
    rng = np.random.default_rng(0)
    cal = scan_all[rng.choice(len(scan_all), size=min(50_000, len(scan_all)), replace=False)].astype(np.float32)
    time_oracle_steps(o, cal, [np.eye(4)], 1, 1)
    per_point = time_oracle_steps(o, cal, [np.eye(4)], 2, 0) / len(cal)
    budget_s = float(os.environ.get("REF_BUDGET_S", "200"))
    n_s = int(budget_s / (per_point * (args.steps + args.warmup)))
    n_s = max(min(20_000, len(scan_all)), min(n_s, 1_000_000 if bounded else n_total, len(scan_all)))
    log(f"reference arm: {wl['cls']} target {n_t} pts, scan sample {n_s} pts ({per_point * 1e9:.0f} ns/pt calibrated), {cores} host threads")
    scan = scan_all if n_s == len(scan_all) else scan_all[np.sort(rng.choice(len(scan_all), size=n_s, replace=False))]
    trace = []
    o.max_iter = 30 if not bounded else 4
    o.align(scan, np.eye(4), trace=trace)
    Ts = [t["T"] for t in trace]
    sec = time_oracle_steps(o, scan.astype(np.float32), Ts, args.steps, args.warmup)
    scale = n_total / n_s                      # a full step costs ~ (n_total / n_s) sampled steps (the NN query is ~linear in the scan)
    extrapolated = n_s < n_total or n_t < n_total
    value = 1.0 / (sec * scale)
    sample = (f"MEASURED: {n_s}-pt scan sample vs a {n_t}-pt target, {sec * 1e3:.1f} ms/step; value = that time scaled x{scale:.1f} "
              f"to the {n_total}-pt scan (the smaller target makes the scaling optimistic for the CPU)"
              if extrapolated else f"full {n_total}-pt workload")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": data_note,
        "extrapolated": extrapolated,
        "measured": {"scan_points": n_s, "target_points": n_t, "ms_per_step": sec * 1e3, "value_on_sample": 1.0 / sec},
        "config": workload_config(wl_name, wl, world, n_total, n_total, target_points_measured=n_t, scan_points_measured=n_s,
                                  implementation="oracle port of the reference (oracle/pcr_oracle.py: NumPy + scipy cKDTree, workers=-1); "
                                                 "the reference's own pykdtree backend is not installed"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "set_target_s": setup_s},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=RESULT_OUT, flush=True)


# ----------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------
def gpu_vs_oracle(pcr, wl, target, scan, device, normals_from_gpu=True, timed_steps=0):
    """Parity leg on host arrays: GPU align vs oracle align on the SAME target / scan.  Returns
    (|T_gpu - T_ref|_F, iterations gpu, iterations oracle, oracle seconds per step or None, H rel err at T0)."""
    reg = getattr(pcr, wl["cls"])(max_iter=MAX_ITER, max_dist=MAX_DIST, tol=TOL, device=device, **wl["kw"])
    reg.set_target(target)
    T_gpu = reg.align(scan)
    it_gpu = reg.last_iterations
    Hg = reg.calc_H_g_e2(np.eye(4), scan)
    o = oracle_object(wl, target, normals=reg.normal if (wl["cls"] == "PlaneICP" and normals_from_gpu) else None)
    trace = []
    T_ref = o.align(scan, np.eye(4), trace=trace)
    Hr = o.calc_H_g_e2(np.eye(4), scan.astype(np.float32))
    sec = time_oracle_steps(o, scan.astype(np.float32), [t["T"] for t in trace], timed_steps, 1) if timed_steps else None
    del reg
    return (float(np.linalg.norm(T_gpu - T_ref)), it_gpu, len(trace), sec,
            float(np.max(np.abs(Hg[0] - Hr[0])) / np.max(np.abs(Hr[0]))))


def timed_trajectory(torch, ctx, method, T0, m, K, W, ext, flush, barrier):
    """K timed iterations (after W warm-up ones) cycling through the m-iteration trajectory from T0,
    every iteration between its own CUDA events on the library's stream, L2 flushed in between."""
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    launches0, wall0 = 0, 0.0
    barrier()
    for i in range(W + K):
        if i == W:
            barrier()
            launches0 = ctx.launch_count()
            wall0 = time.perf_counter()
        if i % m == 0:
            ctx.loop_begin(T0)
        with torch.cuda.stream(ext):
            flush.zero_()
        if i >= W:
            ev[i - W][0].record(ext)
        ctx.loop_step_async(method, MAX_ITER, TOL, MAX_DIST, 1)
        if i >= W:
            ev[i - W][1].record(ext)
    barrier()
    wall_s = time.perf_counter() - wall0
    resets = sum(1 for i in range(W, W + K) if i % m == 0)
    launches = ctx.launch_count() - launches0 - resets * (2 if os.environ.get("PCR_PATH") == "tile" else 1)   # loop_begin resets are not hot-path launches
    step_ms = np.array([a.elapsed_time(b) for a, b in ev])
    return step_ms, wall_s, launches


def run_b200(args, wl_name, wl, world, rank, local_rank, dist=None, context_only=False):
    import torch
    import point_cloud_registration_b200 as pcr
    from point_cloud_registration_b200 import datasets as ds
    from point_cloud_registration_b200.distributed import interleaved_tiles, shard_bounds

    dev = torch.device("cuda", local_rank)
    n_total = wl["n"]
    cls = getattr(pcr, wl["cls"])
    method = cls.method
    on_host = wl_name != "c5"

    # ---- workload -------------------------------------------------------------------------
    t0 = time.perf_counter()
    if on_host:
        target, data_note = host_cloud(wl_name, wl)
        scan, so3 = host_scan(wl_name, wl, target)
        target_in, scan_full = target, scan
    else:
        # 100M points: generated on the GPU (identical on every rank); rotation matched to the scene's lever arm
        target_in = ds.make_urban_slab_torch(n_total, seed=wl["seed"], device=dev)
        ds.assert_no_key_collisions(target_in[: min(n_total, 20_000_000)], (0.5, 1.0))
        ext_xy = (target_in.max(dim=0).values - target_in.min(dim=0).values)[:2]
        radius = float(torch.linalg.norm(0.5 * ext_xy))
        so3 = ds.lever_arm_so3((0.01, -0.02, 0.03), radius)
        scan_full = ds.perturb_scan_torch(target_in, so3=so3, seed=wl["scan_seed"])
        scan_full = scan_full[ds.morton_order_torch(scan_full)].contiguous()     # contiguous index ranges = spatial tiles
        data_note = ("synthetic (urban slab at B-01's surface density, generated on the GPU, scan in Morton order, cut into spatial tiles dealt round-robin to the ranks; the CPU legs use the NumPy generator of the same scene family)")
        torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0
    # this rank's part of the scan: Morton tiles dealt round-robin (c5), else one contiguous range
    tiles = interleaved_tiles(n_total, rank, world, TILES_PER_RANK) if (not on_host and world > 1) else [shard_bounds(n_total, rank, world)]
    n_local = sum(hi - lo for lo, hi in tiles)

    def local_part(x):
        return x[tiles[0][0]:tiles[0][1]] if len(tiles) == 1 else torch.cat([x[lo:hi] for lo, hi in tiles])

    # ---- set_target (once per target; timed separately, not part of the metric) -------------
    reg = cls(max_iter=MAX_ITER, max_dist=MAX_DIST, tol=TOL, device=local_rank, **wl["kw"])
    t0 = time.perf_counter()
    reg.set_target(target_in)
    set_target_s = time.perf_counter() - t0
    ctx = reg._ctx
    if world > 1:
        from point_cloud_registration_b200.distributed import attach
        attach(reg)
        reg.scan_is_presharded = True
    scan_local = local_part(scan_full)
    pinned = torch.empty((n_local, 3), dtype=torch.float32, pin_memory=True)      # host copy of this rank's tile for the end-to-end leg
    pinned.copy_(torch.as_tensor(scan_local) if on_host else scan_local)
    torch.cuda.synchronize()
    scan_host = pinned.numpy()
    t0 = time.perf_counter()
    handle = reg.upload_scan(scan_host if on_host else scan_local.contiguous(), sort=True)
    set_scan_s = time.perf_counter() - t0
    point_index = wl["cls"] in ("ICP", "PlaneICP")
    stats = ctx.index_stats(0 if point_index else 1)
    stats["lists"] = ctx.shell_list_stats() if point_index else ctx.voxel_list_stats()

    # ---- dry run: iterate sequence of one align() --------------------------------------------
    T0 = np.eye(4)
    T_gpu = reg.align(handle, init_T=T0)
    m = reg.last_iterations
    Ts = [T0]
    ctx.loop_begin(T0)
    for _ in range(m - 1):
        ctx.loop_step_async(method, MAX_ITER, TOL, MAX_DIST, 1)
        Ts.append(ctx.loop_state()[0])
    log(f"{wl_name}: n_total={n_total} n_local={n_local} align iterations={m} set_target={set_target_s:.3f}s "
        f"upload+sort={set_scan_s * 1e3:.1f}ms gen={gen_s:.1f}s index={stats}")

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- parity (+ cpu_baseline): GPU vs the CPU oracle on the same arrays ---------------------
    parity, cpu_baseline = {}, None
    cores = os.cpu_count()
    if rank == 0 and not args.no_cpu:
        t0 = time.perf_counter()
        if on_host and n_total <= 2_000_000:
            # the whole workload: oracle align() + timed oracle steps along its trajectory
            o = oracle_object(wl, target, normals=reg.normal if wl["cls"] == "PlaneICP" else None)
            trace = []
            T_ref = o.align(scan, T0, trace=trace)
            sec = time_oracle_steps(o, scan.astype(np.float32), [t["T"] for t in trace], 3 if context_only else 5, 1)
            parity = {"transform_err_vs_ref": float(np.linalg.norm(T_gpu - T_ref)), "iterations_gpu": m, "iterations_ref": len(trace),
                      "on": f"the full {n_total}-pt workload"}
            cpu_baseline = {"value": 1.0 / sec, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": f"full {n_total}-pt workload, calc_H_g_e2 + solve steps along the oracle's align() trajectory"}
        elif on_host:
            # 10M points: the full target structure on both sides, a 500k-point sample of the scan
            n_s = min(500_000, n_total)
            rng = np.random.default_rng(7)
            sample_scan = scan[np.sort(rng.choice(n_total, size=n_s, replace=False))]
            T_s = reg.align(sample_scan)
            it_s = reg.last_iterations
            Hg = reg.calc_H_g_e2(T0, sample_scan)
            o = oracle_object(wl, target)
            trace = []
            T_ref = o.align(sample_scan, T0, trace=trace)
            Hr = o.calc_H_g_e2(T0, sample_scan.astype(np.float32))
            sec = time_oracle_steps(o, sample_scan.astype(np.float32), [t["T"] for t in trace], 3, 1)
            scale = n_total / n_s
            parity = {"transform_err_vs_ref": float(np.linalg.norm(T_s - T_ref)), "iterations_gpu": it_s, "iterations_ref": len(trace),
                      "H_rel_err_at_T0": float(np.max(np.abs(Hg[0] - Hr[0])) / np.max(np.abs(Hr[0]))),
                      "on": f"a {n_s}-pt sample of the scan against the full {n_total}-pt target structure (both sides)"}
            cpu_baseline = {"value": 1.0 / (sec * scale), "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": f"MEASURED {sec * 1e3:.0f} ms/step on a {n_s}-pt scan sample vs the full {n_total}-pt target; value = time scaled x{scale:.0f}"}
            handle = reg.upload_scan(scan_host, sort=True)           # the sample replaced the resident scan
        else:
            # 100M points: a 2M-point sub-workload of the same scene family, GPU (one GPU, no communicator) vs oracle
            sub = dict(wl, n=int(os.environ.get("PCR_BENCH_SUB_N", "2000000")))
            tgt_s, _ = host_cloud("c5", sub)
            scan_s, _ = host_scan("c5", sub, tgt_s, lever_arm=True)
            err, it_g, it_r, sec, hrel = gpu_vs_oracle(pcr, sub, tgt_s, scan_s, local_rank, timed_steps=2)
            scale = n_total / sub["n"]
            parity = {"transform_err_vs_ref": err, "iterations_gpu": it_g, "iterations_ref": it_r, "H_rel_err_at_T0": hrel,
                      "on": f"a {sub['n']}-pt sub-workload of the same scene family (NumPy generator), single GPU vs oracle"}
            cpu_baseline = {"value": 1.0 / (sec * scale), "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": f"MEASURED {sec * 1e3:.0f} ms/step on the {sub['n']}-pt sub-workload; value = time scaled x{scale:.0f}"}
        log(f"{wl_name} parity: {parity} ({time.perf_counter() - t0:.1f}s)")

    # ---- timed region 1: device-resident iterations, per-iteration CUDA events ----------------
    ext = torch.cuda.ExternalStream(ctx.stream(), device=dev)
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
    K, W_req = (min(args.steps, 60), min(args.warmup, 5)) if context_only else (args.steps, args.warmup)
    W = ((W_req + m - 1) // m) * m            # whole align() trajectories: timed step j is iteration j % m
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    step_ms, wall_s, launches = timed_trajectory(torch, ctx, method, T0, m, K, W, ext, flush, barrier)
    local_ms = float(step_ms.sum())
    if dist is not None:
        t = torch.tensor([local_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    else:
        total_ms = local_ms
    ms_per_step = total_ms / K
    value = 1e3 / ms_per_step

    # ---- timed region 2: the same loop, back to back, L2 warm -----------------------------------
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(1, min(20, K // max(m, 1)))
    barrier()
    warm_ms = 0.0
    for r in range(reps):
        ctx.loop_begin(T0)
        a.record(ext)
        ctx.loop_step_async(method, MAX_ITER, TOL, MAX_DIST, m)
        b.record(ext)
        torch.cuda.synchronize()
        warm_ms += a.elapsed_time(b)
    warm = {"ms_per_step": warm_ms / (reps * m), "value": 1e3 * reps * m / warm_ms,
            "note": "iterations enqueued back to back, no L2 flush (align()'s real access pattern)"}

    # ---- timed region 3: end to end through the Python class with host buffers ------------------
    Ke = K if wl_name == "c2" else max(3, min(K, 20))
    barrier()
    for i in range(min(W, 3) + Ke):
        if i == min(W, 3):
            barrier()
            t0 = time.perf_counter()
        T = Ts[i % m]
        H, g, e2 = reg.calc_H_g_e2(T, scan_host)
        dx = -np.linalg.solve(H, g)
        if np.linalg.norm(dx) >= TOL:
            pcr.plus(T, dx)
    barrier()
    e2e_local = (time.perf_counter() - t0) / Ke
    if dist is not None:
        t = torch.tensor([e2e_local], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_local = float(t.item())
    e2e = {"value": 1.0 / e2e_local, "unit": UNIT, "ms_per_step": e2e_local * 1e3, "steps": Ke,
           "h2d_bytes_per_step": int(n_local * 12 + 128), "d2h_bytes_per_step": 29 * 8,
           "api": f"{wl['cls']}.calc_H_g_e2(T, pinned host scan) + host solve/update per step"}
    clocks = sampler.stop() if rank == 0 else None

    # ---- N > 1: the SAME workload on rank 0 alone, and the sharded solve against it ------------------
    n1_same, multi_vs_single, far_start = None, None, None
    if world > 1:
        T_multi = reg.align(reg.upload_scan(scan_local.contiguous(), sort=True), init_T=T0)
        it_multi = reg.last_iterations
        if rank == 0:
            from point_cloud_registration_b200 import _lib
            parked = _lib.Context(local_rank)
            parked.comm_adopt(ctx)                           # rank 0 works alone for a moment: park the communicator
            saved, reg._dist = reg._dist, None
            h1 = reg.upload_scan(scan_full.contiguous(), sort=True)
            T_single = reg.align(h1, init_T=T0)
            it_single = reg.last_iterations
            s_ms, _, _ = timed_trajectory(torch, ctx, method, T0, it_single, min(K, 2 * it_single), it_single, ext, flush, torch.cuda.synchronize)
            n1_same = {"value": 1e3 / float(s_ms.mean()), "ms_per_step": float(s_ms.mean()), "iterations": it_single,
                       "note": f"rank 0 alone on the whole {n_total}-pt scan (no communicator), same target, same trajectory"}
            multi_vs_single = {"T_err": float(np.linalg.norm(T_multi - T_single)), "iterations_multi": it_multi, "iterations_single": it_single}
            reg._dist = saved
            ctx.comm_adopt(parked)
            parked.close()
        barrier()
        # the section-8d rotation NOT matched to the lever arm ("far start"): the rim moves by many metres, beyond
        # max_dist -- the reference does not converge there either; first iterations only, for the record
        scan_far = ds.perturb_scan_torch(target_in, seed=wl["scan_seed"])
        scan_far = local_part(scan_far[ds.morton_order_torch(scan_far)])
        reg.upload_scan(scan_far.contiguous(), sort=True)
        f_ms, _, _ = timed_trajectory(torch, ctx, method, T0, 5, 5, 5, ext, flush, barrier)
        f_local = float(f_ms.sum())
        t = torch.tensor([f_local], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        far_start = {"so3": [0.01, -0.02, 0.03], "ms_per_step_first5": float(t.item()) / 5,
                     "step_ms_by_iteration": [float(x) for x in f_ms],
                     "note": "un-scaled section-8d rotation: rim displacement >> max_dist; first five iterations of a run that does not converge"}
        del scan_far

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
        alg_bytes = n_local * wl["bytes_per_point"]
        achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get(wl_name)
        except Exception:
            pass
        line = {
            "impl": "b200", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "warmup_requested": W_req, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": data_note,
            "config": workload_config(wl_name, wl, world, n_total, n_total, scan_points_per_gpu=n_local, so3=list(so3)),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "bytes_per_scan_point": wl["bytes_per_point"],
                         "kernel": f"correspond_kernel<{wl['cls']}> (SE(3) transform + exact correspondence by per-cell list stream) + "
                                   f"accumulate_kernel<{wl['cls']}> (gather + residual/Jacobian + reduction + GN step); achieved = "
                                   "algorithmic bytes / CUDA-event time of the pair; traffic = ncu DRAM bytes of the pair per iteration"},
            "cpu_baseline": cpu_baseline,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "warm_l2": warm,
            "transform_err_vs_ref": parity.get("transform_err_vs_ref"),
            "parity": parity,
            "align_iterations": m,
            "step_ms_by_iteration": [float(step_ms[j::m].mean()) for j in range(m)] if K >= m else None,
            "set_target_s": set_target_s, "scan_upload_sort_ms": set_scan_s * 1e3,
            "wall_ms_per_step_incl_flush": wall_s * 1e3 / K,
            "points_per_sec": value * n_total,
            "nn_index": stats,
        }
        if world > 1:
            line["n1_same_workload"] = n1_same
            line["multi_vs_single_T_err"] = multi_vs_single["T_err"] if multi_vs_single else None
            line["multi_vs_single"] = multi_vs_single
            line["speedup_vs_n1_same_workload"] = value / n1_same["value"] if n1_same else None
            line["far_start"] = far_start
    else:
        line = None
    del reg, ctx, handle, flush
    torch.cuda.empty_cache()
    return line


def main():
    # The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version
    # banner from C): keep a private handle on the real stdout for the final line and point fd 1 at
    # stderr for everything else.
    global RESULT_OUT
    sys.stdout.flush()
    RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "c1", "c2", "c2i", "c3", "c4", "c5"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the oracle parity / cpu_baseline legs")
    ap.add_argument("--no-others", action="store_true", help="skip the c1/c2i/c3/c4 context runs of the default invocation")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:
        log(f"--gpus {args.gpus} without torchrun: running the single-process N=1 path")
    args.warmup = max(args.warmup, 3)
    wl_name = args.workload if args.workload != "auto" else ("c2" if world == 1 else "c5")
    wl = WORKLOADS[wl_name]
    if args.impl == "reference":
        run_reference(args, wl_name, wl, world, rank)
        return
    import torch
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    line = run_b200(args, wl_name, wl, world, rank, local_rank, dist)
    if world == 1 and args.workload == "auto" and not args.no_others:
        # the other single-GPU configurations of BASELINE.json (shorter timed regions), each with its own
        # cpu_baseline and parity numbers
        others = []
        for name in ("c2i", "c3", "c4", "c1"):
            try:
                o = run_b200(args, name, WORKLOADS[name], world, rank, local_rank, dist, context_only=True)
                others.append({k: o[k] for k in ("value", "unit", "ms_per_step", "steps", "step_ms_by_iteration", "align_iterations",
                                                 "set_target_s", "points_per_sec", "cpu_baseline", "parity", "transform_err_vs_ref",
                                                 "gpu_launches", "data")} |
                              {"workload": o["config"]["workload"], "roofline_frac": o["roofline"]["frac"],
                               "achieved_GBps": o["roofline"]["achieved"], "e2e_value": o["e2e"]["value"], "warm_l2_value": o["warm_l2"]["value"]})
            except Exception as e:            # context only: never lose the main line
                others.append({"workload": name, "error": repr(e)})
        line["other_workloads"] = others
    if rank == 0:
        print(json.dumps(line), file=RESULT_OUT, flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
