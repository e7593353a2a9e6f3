for s in 0 1; do echo "PCR_SORT_ON_CALC=$s"; PCR_SORT_ON_CALC=$s WORKLOADS="c2 c4" STEPS=60 bash tools/sweep.sh; done
