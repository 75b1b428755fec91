#!/bin/bash
TAG=${1:-r2h}
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_any_window.py tests/test_abi.py tests/test_multi_device.py -m gpu -q --durations=5 ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -30 gpurun_out/${TAG}_pytest_gpu.log
python scripts/kernel_time.py --tag w1000 --window 1000 --factor 4 --pitch 1 --seconds 60 > gpurun_out/${TAG}_times.jsonl 2>&1
python scripts/kernel_time.py --tag w1024 --window 1024 --factor 4 --pitch 1 --seconds 60 >> gpurun_out/${TAG}_times.jsonl 2>&1
python scripts/kernel_time.py --tag w4410 --window 4410 --factor 4 --pitch 1 --seconds 60 >> gpurun_out/${TAG}_times.jsonl 2>&1
python scripts/kernel_time.py --tag w4096 --window 4096 --factor 4 --pitch 1 --seconds 60 >> gpurun_out/${TAG}_times.jsonl 2>&1
cat gpurun_out/${TAG}_times.jsonl
