# GPU sweep: dynamic row scheduling, accumulate-kernel variants (run under gpurun)
mkdir -p gpurun_out
LIB=point_cloud_registration_b200/libpcr_b200.so
cp $LIB /tmp/lib_default.so
run() {  # label, env...
  local label="$1"; shift
  echo "== $label"
  env "$@" WORKLOADS="${WL:-c2}" STEPS=${STEPS:-60} bash tools/sweep.sh 2>&1 | tail -n ${TAILN:-1}
}
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "##### pytest variants"; timeout 900 python -m pytest tests -m gpu -x -q -k "scheduling or shell or golden_align or repeat or determin" 2>&1 | tail -3
echo "##### c2 sweep (default lib: acc batch 2, minb 3)"
run "dynamic rows 2, mb4 (default)"
run "rows 1 mb4" PCR_GRAB_ROWS=1
run "rows 4 mb4" PCR_GRAB_ROWS=4
run "rows 2 mb6" PCR_MIN_BLOCKS=6
run "rows 2 mb5" PCR_MIN_BLOCKS=5
run "rows 2 mb6 ppc16" PCR_MIN_BLOCKS=6 PCR_TARGET_PPC=16
run "rows 2 mb6 ppc32" PCR_MIN_BLOCKS=6 PCR_TARGET_PPC=32
run "rows 2 mb4 ppc32" PCR_TARGET_PPC=32
run "rows 2 mb4 ppc16" PCR_TARGET_PPC=16
for v in b1m3 b2m4 b4m3 b4m4; do cp build/libpcr_b200_$v.so $LIB; run "lib $v (acc batch/minb), rows 2 mb4"; done
cp /tmp/lib_default.so $LIB
echo "##### c3/c4"
WL="c3 c4" STEPS=40 TAILN=2 run "default"
WL="c3 c4" STEPS=40 TAILN=2 run "list dilate 3 radius 4" PCR_LIST_DILATE=3 PCR_LIST_RADIUS=4
WL="c3 c4" STEPS=40 TAILN=2 run "list dilate 3 radius 5" PCR_LIST_DILATE=3 PCR_LIST_RADIUS=5
cp build/libpcr_b200_b4m4.so $LIB; WL="c3 c4" STEPS=40 TAILN=2 run "lib b4m4"; cp /tmp/lib_default.so $LIB
} 2>&1 | tee gpurun_out/sweep7.log
