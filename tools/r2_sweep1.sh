#!/bin/bash
# round 2 sweep 1: tile-stream tunables on C2 (and one C3/C4 line each)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "device_resident or api_edges or properties or linearize_10k or reference_fixture or align_10k or scheduling" > gpurun_out/r2_s1_tests.log 2>&1
tail -2 gpurun_out/r2_s1_tests.log
run() {  # name, workload, env...
  name=$1; wl=$2; shift 2
  env "$@" timeout 150 python bench.py --workload $wl --steps 40 --warmup 5 --no-others --no-cpu > gpurun_out/r2_s1_$name.json 2> gpurun_out/r2_s1_$name.err
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/r2_s1_{n}.json"))
    print(n, "it/s", round(d["value"], 1), "by_it", [round(x, 4) for x in d["step_ms_by_iteration"]], "warm", round(d["warm_l2"]["value"], 1), "e2e", round(d["e2e"]["value"], 1), "iters", d["align_iterations"], flush=True)
except Exception as e:
    print(n, "failed", e, flush=True)
PY
}
run c2_reuse c2 A=1
run c2_noreuse c2 PCR_ORDER_REUSE=0
run c2i_reuse c2i A=1
run c3_reuse c3 A=1
