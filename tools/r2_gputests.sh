#!/bin/bash
# GPU test-suite (optionally a -k expression in KEXPR)
mkdir -p gpurun_out
t0=$(date +%s)
if [ -n "$KEXPR" ]; then
  timeout ${TMO:-1500} python -m pytest tests -m gpu -x -q -k "$KEXPR" --durations=10 > gpurun_out/r2_gputests.log 2>&1
else
  timeout ${TMO:-1500} python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r2_gputests.log 2>&1
fi
echo "rc=$? $(( $(date +%s) - t0 )) s"
tail -${TAILN:-25} gpurun_out/r2_gputests.log
