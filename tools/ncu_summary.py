#!/usr/bin/env python3
"""Summarise an .ncu-rep (ncu --set full) as text: one column per profiled launch of the kernels
matching a substring, plus -- with --source -- the share of executed warp instructions and the
active lanes per instruction by 256-byte code region of the first and last launch.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [--match linearize] [--source] > profiles/x_summary.txt
"""
import argparse
import collections
import csv
import io
import subprocess

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "sm__cycles_active.min", "sm__cycles_active.avg", "sm__cycles_active.max",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--match", default="linearize")
    ap.add_argument("--source", action="store_true")
    a = ap.parse_args()
    rows = ncu_csv(a.report, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    ik = hdr.index("Kernel Name")
    data = [r for r in data if a.match in r[ik]]
    print(f"# {a.report}: {len(data)} launches matching '{a.match}' (ncu --set full --clock-control none; cold caches, serialised)")
    for n in sorted(set(r[ik] for r in data)):
        print("# kernel:", n)
    for m in METRICS:
        if m not in hdr:
            continue
        i = hdr.index(m)
        print(f"{m:80s} " + " | ".join(f"{r[i]:>12.12s}" for r in data) + f"  {units[i]}")
    if not a.source:
        return
    rows = ncu_csv(a.report, "source")
    blocks, cur = [], None
    for r in rows:
        if len(r) >= 2 and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    blocks = [b for b in blocks if a.match in b["name"]]
    for bi in sorted(set([0, len(blocks) - 1])):
        b = blocks[bi]
        h = b["rows"][0]
        d = [r for r in b["rows"][1:] if len(r) > 10 and r[0].startswith("0x")]
        ia, ie, it = h.index("Address"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
        base = int(d[0][ia], 16)
        tot = sum(int(r[ie]) for r in d)
        print(f"\n# launch {bi} of {b['name']}: {tot} warp instructions; code regions with > 1.5 % of them")
        reg = collections.OrderedDict()
        for r in d:
            k = (int(r[ia], 16) - base) // 0x100
            x = reg.setdefault(k, [0, 0])
            x[0] += int(r[ie]); x[1] += int(r[it])
        for k, (e, t) in reg.items():
            if e > tot * 0.015:
                print(f"  +0x{k * 0x100:05x}: {e / tot * 100:5.1f} % of warp instructions, {t / max(e, 1):5.1f} active lanes")


if __name__ == "__main__":
    main()
