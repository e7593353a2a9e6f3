# GPU sweep: shell lists + straggler queue (run under gpurun)
mkdir -p gpurun_out
run() {  # label, env...
  local label="$1"; shift
  echo "== $label"
  env "$@" WORKLOADS="${WL:-c2}" STEPS=${STEPS:-60} bash tools/sweep.sh 2>&1 | tail -n ${TAILN:-1}
}
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "##### pytest shell-list tests"; timeout 600 python -m pytest tests -m gpu -x -q -k "shell or agree or candidate" 2>&1 | tail -3
echo "##### c2 sweep"
run "shell off (general search)" PCR_SHELL_LISTS=0
run "shell default (dmax 2.0, ppc24, queue)"
run "shell dmax 1.5" PCR_SHELL_DMAX=1.5
run "shell dmax 1.0" PCR_SHELL_DMAX=1.0
run "shell dmax 2.0 queue off" PCR_QUEUE=0
run "shell dmax 2.0 mb4" PCR_MIN_BLOCKS=4
run "shell dmax 2.0 mb2" PCR_MIN_BLOCKS=2
run "shell dmax 2.0 ppc36" PCR_TARGET_PPC=36
run "shell dmax 2.0 ppc16" PCR_TARGET_PPC=16
run "shell dmax 2.0 ppc12" PCR_TARGET_PPC=12
run "shell dmax 2.0 ppc8" PCR_TARGET_PPC=8
run "shell dmax 2.0 ppc12 mb4" PCR_TARGET_PPC=12 PCR_MIN_BLOCKS=4
echo "##### c3/c4"
WL="c3 c4" STEPS=40 TAILN=2 run "queue on (default)"
WL="c3 c4" STEPS=40 TAILN=2 run "queue off" PCR_QUEUE=0
WL="c3 c4" STEPS=40 TAILN=2 run "queue on, list dilate 3 radius 4" PCR_LIST_DILATE=3 PCR_LIST_RADIUS=4
echo "##### ncu shell c2"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:linearize_lane -c 5 -o gpurun_out/r1_c2_shell -f python bench.py --steps 5 --warmup 5 --no-cpu --no-others > gpurun_out/ncu_shell.log 2>&1; tail -2 gpurun_out/ncu_shell.log
echo "##### full gpu suite"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
} 2>&1 | tee gpurun_out/sweep5.log
