#!/bin/bash
# ncu capture of the tile-stream kernel on C2 (full set, first launches = one align() trajectory)
mkdir -p gpurun_out
WL=${WL:-c2}
ncu --set full --clock-control none --import-source on -k regex:tile_linearize -c ${NCU_COUNT:-6} -o gpurun_out/r2_tile_${WL}_${TAG:-a} -f \
    python bench.py --workload $WL --steps 6 --warmup 3 --no-others --no-cpu > gpurun_out/ncu_tile_${WL}.log 2>&1
tail -3 gpurun_out/ncu_tile_${WL}.log
ls -la gpurun_out/*.ncu-rep
