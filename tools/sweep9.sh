# GPU sweep 9: cell-ordered scans, queue on/off, cell size (run under gpurun)
mkdir -p gpurun_out
run() {  # label, env...
  local label="$1"; shift
  echo "== $label"
  env "$@" WORKLOADS="${WL:-c2}" STEPS=${STEPS:-60} bash tools/sweep.sh 2>&1 | tail -n ${TAILN:-1}
}
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "##### tests"
timeout 900 python -m pytest tests -m gpu -x -q -k "scheduling or shell or golden or properties or reference_tests" 2>&1 | tail -3
echo "##### c2 sweep"
run "default (cell order, queue off, ppc24)"
run "morton order" PCR_CELL_ORDER=0
run "queue on" PCR_QUEUE=1
run "ppc32" PCR_TARGET_PPC=32
run "ppc48" PCR_TARGET_PPC=48
run "ppc64" PCR_TARGET_PPC=64
run "ppc16" PCR_TARGET_PPC=16
run "ppc32 mb5" PCR_TARGET_PPC=32 PCR_MIN_BLOCKS=5
run "ppc48 mb5" PCR_TARGET_PPC=48 PCR_MIN_BLOCKS=5
run "ppc24 mb5" PCR_MIN_BLOCKS=5
run "ppc48 morton" PCR_TARGET_PPC=48 PCR_CELL_ORDER=0
echo "##### c3/c4"
WL="c3 c4" STEPS=40 TAILN=2 run "default"
WL="c3 c4" STEPS=40 TAILN=2 run "queue on" PCR_QUEUE=1
WL="c3 c4" STEPS=40 TAILN=2 run "morton" PCR_CELL_ORDER=0
WL="c3 c4" STEPS=40 TAILN=2 run "rows 2" PCR_GRAB_ROWS=2
} 2>&1 | tee gpurun_out/sweep9.log
