#!/bin/bash
# round 2: full GPU suite + bench (N=1) + launch list for the shipped (scalar) build
TAG=${1:-r2f}
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q --durations=10 ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -25 gpurun_out/${TAG}_pytest_gpu.log
( time timeout 900 python bench.py --steps 5 --warmup 3 ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?" >> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json; tail -8 gpurun_out/${TAG}_bench.err
