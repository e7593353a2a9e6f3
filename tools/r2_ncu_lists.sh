#!/bin/bash
# round-2 ncu evidence of the default (list) path: launch list of a short bench run + full-set captures,
# summarised ON the box (the reports themselves are too large to bring back)
mkdir -p gpurun_out
if [ "${LAUNCHES:-1}" = "1" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_c2_launches.csv \
    python bench.py --workload c2 --steps 18 --warmup 9 --no-others --no-cpu > gpurun_out/ncu_launches.log 2>&1
fi
for WL in ${WLS:-c2 c2i c3 c4}; do
  ncu --set full --clock-control none --import-source on -k regex:"correspond_kernel|accumulate_kernel" -c ${NCU_COUNT:-10} -o /tmp/r2_lists_${WL} -f \
      python bench.py --workload $WL --steps 6 --warmup 3 --no-others --no-cpu > gpurun_out/ncu_lists_${WL}.log 2>&1
  python tools/ncu_summary.py /tmp/r2_lists_${WL}.ncu-rep --match correspond > gpurun_out/r2_${WL}_correspond_summary.txt 2>&1
  python tools/ncu_summary.py /tmp/r2_lists_${WL}.ncu-rep --match accumulate > gpurun_out/r2_${WL}_accumulate_summary.txt 2>&1
  python tools/ncu_lines.py /tmp/r2_lists_${WL}.ncu-rep 0 0.8 > gpurun_out/r2_${WL}_lines_launch0.txt 2>&1
  python tools/ncu_lines.py /tmp/r2_lists_${WL}.ncu-rep 8 0.8 > gpurun_out/r2_${WL}_lines_launch8.txt 2>&1
  python tools/ncu_lines.py /tmp/r2_lists_${WL}.ncu-rep 9 0.8 > gpurun_out/r2_${WL}_lines_launch9.txt 2>&1
  rm -f /tmp/r2_lists_${WL}.ncu-rep
  head -4 gpurun_out/r2_${WL}_correspond_summary.txt | cut -c1-220
done
du -sh gpurun_out
