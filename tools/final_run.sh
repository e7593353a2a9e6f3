# Final GPU validation of the round: tests, smoke, bench (both arms), ncu launch list + full captures.
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "##### pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "##### smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "##### bench (default)"; timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -3 gpurun_out/bench_final.err; cut -c1-600 gpurun_out/bench_final.json
echo "##### bench --impl reference"; timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -1 gpurun_out/bench_reference.err; cut -c1-300 gpurun_out/bench_reference.json
echo "##### ncu launch list (c2)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r1_c2_launches.csv python bench.py --steps 10 --warmup 5 --no-cpu --no-others > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log | cut -c1-200
echo "##### ncu full c3 / c4"
for w in c3 c4; do timeout 600 ncu --set full --clock-control none -k regex:"correspond|accumulate" -c 12 -o gpurun_out/r1_${w}_lin -f python bench.py --workload $w --steps 6 --warmup 6 --no-cpu > gpurun_out/ncu_$w.log 2>&1; tail -1 gpurun_out/ncu_$w.log | cut -c1-160; done
echo "##### ncu full set_target kernels"
timeout 600 ncu --set full --clock-control none -k regex:"normals_kernel|shell_build|grid_fill|gather_points" -c 6 -o gpurun_out/r1_c2_set_target -f python bench.py --steps 5 --warmup 5 --no-cpu --no-others > gpurun_out/ncu_st2.log 2>&1; tail -1 gpurun_out/ncu_st2.log | cut -c1-160
timeout 600 ncu --set full --clock-control none -k regex:"voxel_stats|voxel_finalize|list_build|voxel_coord" -c 6 -o gpurun_out/r1_c3_set_target -f python bench.py --workload c3 --steps 6 --warmup 6 --no-cpu > gpurun_out/ncu_st3.log 2>&1; tail -1 gpurun_out/ncu_st3.log | cut -c1-160
} 2>&1 | tee gpurun_out/final_run.log
