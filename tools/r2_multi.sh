#!/bin/bash
# 2-GPU checks: multi-GPU parity tests, then the N=2 bench line at a reduced and at the full C5 size
mkdir -p gpurun_out
if [ "${TESTS:-1}" = "1" ]; then timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2_multi_tests.log 2>&1; tail -3 gpurun_out/r2_multi_tests.log; fi
N=${NGPU:-2}
run() {  # name, env...
  name=$1; shift
  t0=$(date +%s)
  env "$@" timeout ${TMO:-900} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-40} --warmup 5 > gpurun_out/r2_multi_${name}_n$N.json 2> gpurun_out/r2_multi_${name}_n$N.err
  echo "$name rc=$? $(( $(date +%s) - t0 )) s"
  python - "${name}_n$N" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/r2_multi_{n}.json"))
    print(n, "it/s", round(d["value"], 2), "ms", round(d["ms_per_step"], 3), "by_it", [round(x, 2) for x in d["step_ms_by_iteration"]], "iters", d["align_iterations"])
    print("   n1_same", d["n1_same_workload"], "speedup", d["speedup_vs_n1_same_workload"], "multi_vs_single", d["multi_vs_single"])
    print("   parity", d["parity"], "lists", d["nn_index"]["lists"], "set_target", round(d["set_target_s"], 2), "far", d["far_start"]["ms_per_step_first5"])
except Exception as e:
    print(n, "failed", e)
    import subprocess; print(subprocess.run(["tail", "-5", f"gpurun_out/r2_multi_{n}.err"], capture_output=True, text=True).stdout)
PY
}
if [ "${SMALL:-1}" = "1" ]; then run small PCR_BENCH_C5_N=${SMALL_N:-20000000}; fi
if [ "${FULL:-1}" = "1" ]; then run full A=1; fi
