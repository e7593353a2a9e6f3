# GPU sweep: split kernels / occupancy / pipelined list loop (run under gpurun)
mkdir -p gpurun_out
run() {  # label, env...
  local label="$1"; shift
  echo "== $label"
  env "$@" WORKLOADS="${WL:-c2}" STEPS=${STEPS:-60} bash tools/sweep.sh 2>&1 | tail -n ${TAILN:-1}
}
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "##### pytest variants"; timeout 900 python -m pytest tests -m gpu -x -q -k "scheduling or shell or candidate or structure" 2>&1 | tail -3
echo "##### c2 sweep"
run "split mb4 (default)"
run "split mb3" PCR_MIN_BLOCKS=3
run "split mb5" PCR_MIN_BLOCKS=5
run "split mb6" PCR_MIN_BLOCKS=6
run "fused mb3" PCR_SPLIT=0
run "split mb6 ppc16" PCR_MIN_BLOCKS=6 PCR_TARGET_PPC=16
run "split mb6 ppc12" PCR_MIN_BLOCKS=6 PCR_TARGET_PPC=12
run "split mb5 ppc16" PCR_MIN_BLOCKS=5 PCR_TARGET_PPC=16
run "split mb6 ppc32" PCR_MIN_BLOCKS=6 PCR_TARGET_PPC=32
run "split mb6 dmax1.5" PCR_MIN_BLOCKS=6 PCR_SHELL_DMAX=1.5
run "split mb6 lists off" PCR_MIN_BLOCKS=6 PCR_SHELL_LISTS=0
echo "##### c3/c4"
WL="c3 c4" STEPS=40 TAILN=2 run "split mb4 (default)"
WL="c3 c4" STEPS=40 TAILN=2 run "split mb6" PCR_MIN_BLOCKS=6
WL="c3 c4" STEPS=40 TAILN=2 run "fused" PCR_SPLIT=0
WL="c3 c4" STEPS=40 TAILN=2 run "split mb6, list dilate 3 radius 4" PCR_MIN_BLOCKS=6 PCR_LIST_DILATE=3 PCR_LIST_RADIUS=4
echo "##### ncu c2 default"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"correspond|accumulate" -c 10 -o gpurun_out/r1_c2_split -f python bench.py --steps 5 --warmup 5 --no-cpu --no-others > gpurun_out/ncu_split.log 2>&1; tail -2 gpurun_out/ncu_split.log
echo "##### full gpu suite"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
} 2>&1 | tee gpurun_out/sweep6.log
