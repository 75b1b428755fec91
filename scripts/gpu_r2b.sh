#!/bin/bash
# round 2: GPU test suite after the multi-device refactor + IMAD microbenchmark
mkdir -p gpurun_out
./scripts/microbench/f32x2_issue > gpurun_out/r2b_f32x2_issue.txt 2>&1
( time timeout 2400 python -m pytest tests -m gpu -q -x --durations=15 ) > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest_gpu.log
cat gpurun_out/r2b_f32x2_issue.txt; tail -40 gpurun_out/r2b_pytest_gpu.log
