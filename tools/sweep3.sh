# GPU sweep of the search-scheduling variants (run under gpurun): prints one summary line per variant
mkdir -p gpurun_out
run() {  # label, env...
  local label="$1"; shift
  echo "== $label"
  env "$@" WORKLOADS="${WL:-c2}" STEPS=${STEPS:-60} bash tools/sweep.sh 2>&1 | tail -n ${TAILN:-1}
}
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "##### pytest flat tests"; timeout 600 python -m pytest tests -m gpu -x -q -k "flat" 2>&1 | tail -3
echo "##### c2 sweep"
run "nested mb3" PCR_SEARCH=0
run "flat ch32 tau1 mb3" PCR_SEARCH=flat PCR_FLAT_CH=32 PCR_FLAT_TAU=1
run "flat ch32 tau1 mb4" PCR_SEARCH=flat PCR_FLAT_CH=32 PCR_FLAT_TAU=1 PCR_MIN_BLOCKS=4
run "flat ch32 tau1 mb2" PCR_SEARCH=flat PCR_FLAT_CH=32 PCR_FLAT_TAU=1 PCR_MIN_BLOCKS=2
run "flat ch16 tau16 mb3" PCR_SEARCH=flat PCR_FLAT_CH=16 PCR_FLAT_TAU=16
run "flat ch16 tau1 mb3" PCR_SEARCH=flat PCR_FLAT_CH=16 PCR_FLAT_TAU=1
run "flat ch8 tau8 mb3" PCR_SEARCH=flat PCR_FLAT_CH=8 PCR_FLAT_TAU=8
run "flat ch32 tau1 ppc12" PCR_SEARCH=flat PCR_FLAT_CH=32 PCR_FLAT_TAU=1 PCR_TARGET_PPC=12
run "flat ch16 tau16 ppc12" PCR_SEARCH=flat PCR_FLAT_CH=16 PCR_FLAT_TAU=16 PCR_TARGET_PPC=12
run "flat ch32 tau1 ppc16 mb4" PCR_SEARCH=flat PCR_FLAT_CH=32 PCR_FLAT_TAU=1 PCR_TARGET_PPC=16 PCR_MIN_BLOCKS=4
run "flat ch32 tau1 ppc36" PCR_SEARCH=flat PCR_FLAT_CH=32 PCR_FLAT_TAU=1 PCR_TARGET_PPC=36
echo "##### c3/c4"
WL="c3 c4" STEPS=40 TAILN=2 run "nested" PCR_SEARCH=0
WL="c3 c4" STEPS=40 TAILN=2 run "flat ch32 tau1" PCR_SEARCH=flat PCR_FLAT_CH=32 PCR_FLAT_TAU=1
WL="c3 c4" STEPS=40 TAILN=2 run "flat ch16 tau16" PCR_SEARCH=flat PCR_FLAT_CH=16 PCR_FLAT_TAU=16
echo "##### ncu flat c2"
PCR_SEARCH=flat timeout 600 ncu --set full --import-source on --clock-control none -k regex:linearize_flat -c 5 -o gpurun_out/r1_c2_flat -f python bench.py --steps 5 --warmup 5 --no-cpu --no-others > gpurun_out/ncu_flat.log 2>&1; tail -2 gpurun_out/ncu_flat.log
echo "##### full gpu suite, flat default"
PCR_SEARCH=flat timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
} 2>&1 | tee gpurun_out/sweep3.log
