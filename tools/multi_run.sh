# 2-GPU validation: parity test + C5-style bench (run under gpurun --gpus 2)
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name --format=csv,noheader
echo "##### pytest multi"; timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
echo "##### bench 2 GPUs (c5: 12.5M scan points per GPU, 25M-point target replicated)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; tail -4 gpurun_out/bench_2gpu.err | cut -c1-300; cut -c1-1500 gpurun_out/bench_2gpu.json
} 2>&1 | tee gpurun_out/multi_run.log
