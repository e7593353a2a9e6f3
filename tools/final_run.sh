# Final GPU validation of the round: tests, smoke, bench (both arms), ncu launch list + full captures.
# ncu reports are summarised ON THE BOX (tools/ncu_summary.py) and deleted: gpurun_out/ must stay < 64 MiB.
mkdir -p gpurun_out
cap() {  # name, match, kernel regex, launches, bench args...
  local name="$1" match="$2" regex="$3" cnt="$4"; shift 4
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:"$regex" -c "$cnt" -o /tmp/$name -f python bench.py "$@" > gpurun_out/ncu_$name.log 2>&1
  python tools/ncu_summary.py /tmp/$name.ncu-rep --match "$match" --source > gpurun_out/${name}_summary.txt 2>> gpurun_out/ncu_$name.log
  head -4 gpurun_out/${name}_summary.txt | cut -c1-200
}
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "##### pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "##### smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "##### bench (default)"; timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -3 gpurun_out/bench_final.err | cut -c1-400; cut -c1-400 gpurun_out/bench_final.json
echo "##### bench --impl reference"; timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -1 gpurun_out/bench_reference.err; cut -c1-200 gpurun_out/bench_reference.json
echo "##### ncu launch list (c2)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r1_c2_launches.csv python bench.py --steps 10 --warmup 5 --no-cpu --no-others > gpurun_out/ncu_launches.log 2>&1; wc -l gpurun_out/r1_c2_launches.csv
echo "##### ncu full: per-iteration kernels"
cap r1_c2_final "_kernel" "correspond|accumulate" 10 --steps 5 --warmup 5 --no-cpu --no-others
cap r1_c3_final "_kernel" "correspond|accumulate" 12 --workload c3 --steps 6 --warmup 6 --no-cpu
cap r1_c4_final "_kernel" "correspond|accumulate" 10 --workload c4 --steps 5 --warmup 5 --no-cpu
echo "##### ncu full: once-per-target kernels"
cap r1_c2_set_target "_kernel" "normals_kernel|shell_build|grid_fill|gather_points|cell_key|scan_to_soa" 8 --steps 5 --warmup 5 --no-cpu --no-others
cap r1_c3_set_target "_kernel" "voxel_stats|voxel_finalize|list_build|voxel_coord" 6 --workload c3 --steps 6 --warmup 6 --no-cpu
du -sh gpurun_out
} 2>&1 | tee gpurun_out/final_run.log
