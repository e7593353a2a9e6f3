# GPU sweep 8: packed shell loop + padding + levels; pick defaults (run under gpurun)
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
run "default (rows1 mb4 ppc24 dmax2)"
run "mb3" PCR_MIN_BLOCKS=3
run "mb5" PCR_MIN_BLOCKS=5
run "mb6" PCR_MIN_BLOCKS=6
run "ppc16" PCR_TARGET_PPC=16
run "ppc32" PCR_TARGET_PPC=32
run "ppc48" PCR_TARGET_PPC=48
run "ppc12 mb6" PCR_TARGET_PPC=12 PCR_MIN_BLOCKS=6
run "rows 2" PCR_GRAB_ROWS=2
run "dmax 1.5" PCR_SHELL_DMAX=1.5
run "queue off" PCR_QUEUE=0
echo "##### c3/c4"
WL="c3 c4" STEPS=40 TAILN=2 run "default"
WL="c3 c4" STEPS=40 TAILN=2 run "mb6" PCR_MIN_BLOCKS=6
echo "##### ncu c2 default"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"correspond|accumulate" -c 10 -o gpurun_out/r1_c2_final -f python bench.py --steps 5 --warmup 5 --no-cpu --no-others > gpurun_out/ncu_c2_final.log 2>&1; tail -2 gpurun_out/ncu_c2_final.log
} 2>&1 | tee gpurun_out/sweep8.log
