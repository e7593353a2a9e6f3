timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "voxel or structures or 10k or b01" 2>&1 | tail -3
WORKLOADS="c3 c4" bash tools/sweep.sh
