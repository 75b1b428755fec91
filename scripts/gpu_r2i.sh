#!/bin/bash
TAG=${1:-r2i}
mkdir -p gpurun_out
python scripts/kernel_time.py --tag f20pm3 --factor 20 --pitch -3 > gpurun_out/${TAG}_times.jsonl 2>&1
python scripts/kernel_time.py --tag f8pm2 --factor 8 --pitch -2 >> gpurun_out/${TAG}_times.jsonl 2>&1
python scripts/kernel_time.py --tag f20pm3_w1024 --factor 20 --pitch -3 --window 1024 >> gpurun_out/${TAG}_times.jsonl 2>&1
python scripts/kernel_time.py --tag f6p3 --factor 6 --pitch 3 >> gpurun_out/${TAG}_times.jsonl 2>&1
cat gpurun_out/${TAG}_times.jsonl
( time timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_full_size_configs.py tests/test_any_window.py -m gpu -q --durations=5 ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -15 gpurun_out/${TAG}_pytest_gpu.log
