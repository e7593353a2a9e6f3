#!/bin/bash
# quick check of the pipelined upload with kept orders: its tests, then the C2 line (device-resident value + e2e)
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipelined or reuses or api_edges or (10m and VPlane)" > gpurun_out/r2_quick_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2_quick_tests.log
timeout 80 python bench.py --workload c2 --steps 90 --warmup 9 --no-others --no-cpu > gpurun_out/r2_quick_c2.json 2> gpurun_out/r2_quick_c2.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_quick_c2.json"))
print("c2", round(d["value"], 1), "it/s  e2e", round(d["e2e"]["value"], 1), "ms", round(d["e2e"]["ms_per_step"], 4))
PY
