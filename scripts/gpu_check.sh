#!/bin/bash
# usage: bash scripts/gpu_check.sh <tag> [--no-tests]
# Runs on the GPU box under gpurun: GPU tests, smoke, bench, ncu launch list + one full capture.
TAG=${1:-run}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpu.txt 2>&1
if [ "$2" != "--no-tests" ]; then
  timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
  timeout 600 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${TAG}_smoke.log
fi
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?" >> gpurun_out/${TAG}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_b_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stretch_kernel -s 3 -c 1 -o gpurun_out/${TAG}_prof python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_b_ncu2.log 2>&1
tail -15 gpurun_out/${TAG}_pytest_gpu.log; tail -3 gpurun_out/${TAG}_smoke.log; cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
