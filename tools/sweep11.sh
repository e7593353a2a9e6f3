# GPU sweep 11: ball-first fallback for list misses (run under gpurun)
mkdir -p gpurun_out
run() { local label="$1"; shift; echo "== $label"; env "$@" WORKLOADS="${WL:-c2}" STEPS=${STEPS:-60} bash tools/sweep.sh 2>&1 | tail -n ${TAILN:-1}; }
{
echo "##### full gpu suite"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run "c2 ball-first (default)"
run "c2 ring growth" PCR_BALL_FIRST=0
WL="c3 c4" STEPS=40 TAILN=2 run "c3/c4 ball-first (default)"
WL="c3 c4" STEPS=40 TAILN=2 run "c3/c4 ring growth" PCR_BALL_FIRST=0
} 2>&1 | tee gpurun_out/sweep11.log
