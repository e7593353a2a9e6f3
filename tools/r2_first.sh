#!/bin/bash
# round 2, first GPU contact of the tile-stream path: parity subset, then C2 A/B against the list path
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tile_stream or reference_fixture or linearize_10k or align_10k" > gpurun_out/r2_first_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2_first_tests.log
tail -5 gpurun_out/r2_first_tests.log
timeout 600 python bench.py --steps 40 --warmup 5 --no-others --no-cpu > gpurun_out/r2_first_c2_tile.json 2> gpurun_out/r2_first_c2_tile.err
PCR_PATH=lists timeout 600 python bench.py --steps 40 --warmup 5 --no-others --no-cpu > gpurun_out/r2_first_c2_lists.json 2> gpurun_out/r2_first_c2_lists.err
python - <<'PY'
import json
for n in ("tile","lists"):
    try:
        d=json.load(open(f"gpurun_out/r2_first_c2_{n}.json"))
        print(n, "value", round(d["value"],1), "ms", d["ms_per_step"], "by_it", [round(x,4) for x in d["step_ms_by_iteration"]], "warm", d["warm_l2"]["value"], "e2e", d["e2e"]["value"], "iters", d["align_iterations"], "set_target", d["set_target_s"])
    except Exception as e:
        print(n, "failed", e)
PY
