#!/bin/bash
TAG=${1:-r2u}
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -22 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
