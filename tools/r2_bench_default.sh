#!/bin/bash
# the driver's two N=1 invocations, timed
mkdir -p gpurun_out
t0=$(date +%s)
timeout 900 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "b200 arm rc=$? $(( $(date +%s) - t0 )) s"
t0=$(date +%s)
timeout 600 python bench.py --impl reference > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "reference arm rc=$? $(( $(date +%s) - t0 )) s"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_default.json"))
print("c2", round(d["value"],1), "it/s", "frac", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"],1), "by_it", [round(x,4) for x in d["step_ms_by_iteration"]], "parity", d["parity"], "cpu", d["cpu_baseline"]["value"], d["data"][:40])
for o in d["other_workloads"]:
    print(o.get("workload","?")[:30], {k: (round(o[k],4) if isinstance(o.get(k), float) else o.get(k)) for k in ("value","roofline_frac","e2e_value","transform_err_vs_ref","align_iterations","error") if k in o}, (o.get("cpu_baseline") or {}).get("value"), (o.get("parity") or {}).get("iterations_ref"))
r = json.load(open("gpurun_out/r2_bench_reference.json"))
print("reference", r["value"], r["ms_per_step"], r["extrapolated"], r["measured"])
PY
tail -5 gpurun_out/r2_bench_default.err
