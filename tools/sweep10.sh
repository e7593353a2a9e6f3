# GPU sweep 10: paired list streams (run under gpurun)
mkdir -p gpurun_out
run() {  # label, env...
  local label="$1"; shift
  echo "== $label"
  env "$@" WORKLOADS="${WL:-c2}" STEPS=${STEPS:-60} bash tools/sweep.sh 2>&1 | tail -n ${TAILN:-1}
}
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "##### full gpu suite"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "##### c2 sweep"
run "default (pairs, mb4, ppc24)"
run "pairs off mb4" PCR_PAIR_ROWS=0
run "pairs mb3" PCR_MIN_BLOCKS=3
run "pairs mb5" PCR_MIN_BLOCKS=5
run "pairs off mb5" PCR_PAIR_ROWS=0 PCR_MIN_BLOCKS=5
run "pairs mb3 ppc16" PCR_MIN_BLOCKS=3 PCR_TARGET_PPC=16
run "pairs mb4 ppc16" PCR_TARGET_PPC=16
run "pairs mb4 rows4" PCR_GRAB_ROWS=4
echo "##### c3/c4"
WL="c3 c4" STEPS=40 TAILN=2 run "default"
echo "##### ncu c2 default"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"correspond|accumulate" -c 10 -o gpurun_out/r1_c2_pairs -f python bench.py --steps 5 --warmup 5 --no-cpu --no-others > gpurun_out/ncu_c2_pairs.log 2>&1; tail -2 gpurun_out/ncu_c2_pairs.log
} 2>&1 | tee gpurun_out/sweep10.log
