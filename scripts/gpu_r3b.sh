#!/bin/bash
# long-window (HBM chirp-z) path: parity tests + timing
TAG=${TAG:-r3b}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_any_window.py -m gpu -q -x --durations=8 ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -25 gpurun_out/${TAG}_pytest_gpu.log
: > gpurun_out/${TAG}_times.jsonl
for W in 10000 44100 131072; do
  timeout 300 python scripts/kernel_time.py --tag big_$W --factor 4 --pitch 1 --window $W --seconds 120 --reps 3 >> gpurun_out/${TAG}_times.jsonl 2>> gpurun_out/${TAG}_err.log || echo "FAIL $W" >> gpurun_out/${TAG}_times.jsonl
done
cut -c1-250 gpurun_out/${TAG}_times.jsonl
