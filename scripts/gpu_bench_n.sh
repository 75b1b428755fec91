#!/bin/bash
# usage: bash scripts/gpu_bench_n.sh <N> <tag>: bench.py on N GPUs the way the driver launches it
N=${1:-2}; TAG=${2:-r2g}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 ) > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err; echo "bench rc=$?" >> gpurun_out/${TAG}_bench_n${N}.err
tail -c 6000 gpurun_out/${TAG}_bench_n${N}.json; tail -n 12 gpurun_out/${TAG}_bench_n${N}.err
# the NCCL / multi-device tests that a one-GPU box skips
( timeout 600 python -m pytest tests/test_multi_gpu.py tests/test_multi_device.py -m gpu -q -x ) > gpurun_out/${TAG}_pytest_multi_gpu.log 2>&1; tail -n 4 gpurun_out/${TAG}_pytest_multi_gpu.log
