# GPU sweep: per-cell neighbour lists (run under gpurun)
mkdir -p gpurun_out
run() {  # label, env...
  local label="$1"; shift
  echo "== $label"
  env "$@" WORKLOADS="${WL:-c2}" STEPS=${STEPS:-60} bash tools/sweep.sh 2>&1 | tail -n ${TAILN:-1}
}
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "##### pytest neighbour-list tests"; timeout 600 python -m pytest tests -m gpu -x -q -k "neighbour" 2>&1 | tail -3
echo "##### c2 sweep"
run "nbr off ppc24 mb3" PCR_NBR_LISTS=0
run "nbr on ppc24 mb3"
run "nbr on ppc24 mb4" PCR_MIN_BLOCKS=4
run "nbr on ppc24 mb2" PCR_MIN_BLOCKS=2
run "nbr on ppc36" PCR_TARGET_PPC=36
run "nbr on ppc16" PCR_TARGET_PPC=16
run "nbr on ppc12" PCR_TARGET_PPC=12
run "nbr on ppc12 mb4" PCR_TARGET_PPC=12 PCR_MIN_BLOCKS=4
run "nbr on ppc8" PCR_TARGET_PPC=8
run "nbr on ppc6" PCR_TARGET_PPC=6
run "nbr on ppc4" PCR_TARGET_PPC=4
echo "##### ncu nbr c2"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:linearize_lane -c 5 -o gpurun_out/r1_c2_nbr -f python bench.py --steps 5 --warmup 5 --no-cpu --no-others > gpurun_out/ncu_nbr.log 2>&1; tail -2 gpurun_out/ncu_nbr.log
echo "##### full gpu suite"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
} 2>&1 | tee gpurun_out/sweep4.log
