#!/bin/bash
mkdir -p gpurun_out
t0=$(date +%s)
timeout ${TMO:-1500} python -m pytest tests -m gpu -x -q ${PYTEST_ARGS} --durations=15 > gpurun_out/r2_gputests.log 2>&1
echo "rc=$? $(( $(date +%s) - t0 )) s"
tail -25 gpurun_out/r2_gputests.log
