#!/bin/bash
TAG=${1:-r2y}
mkdir -p gpurun_out
timeout 120 python scripts/kernel_time.py --tag cfg3_65536 --factor 4 --pitch 1 --window 65536 > gpurun_out/${TAG}_times.jsonl 2>&1 || echo "TIMEOUT/FAIL rc=$?" >> gpurun_out/${TAG}_times.jsonl
timeout 120 python scripts/kernel_time.py --tag f8p2_65536 --factor 8 --pitch 2 --window 65536 >> gpurun_out/${TAG}_times.jsonl 2>&1 || echo "TIMEOUT/FAIL rc=$?" >> gpurun_out/${TAG}_times.jsonl
timeout 120 python scripts/kernel_time.py --tag f20pm3_65536 --factor 20 --pitch -3 --window 65536 >> gpurun_out/${TAG}_times.jsonl 2>&1 || echo "TIMEOUT/FAIL rc=$?" >> gpurun_out/${TAG}_times.jsonl
cut -c1-260 gpurun_out/${TAG}_times.jsonl
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_full_size_configs.py tests/test_plugin_path.py tests/test_multi_device.py -m gpu -q -x -k "65536 or window_65536 or cpp_one or config3 or stacks" --durations=5 ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -12 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 compute-sanitizer --tool racecheck python scripts/sanitize_small.py > gpurun_out/${TAG}_racecheck.log 2>&1; tail -3 gpurun_out/${TAG}_racecheck.log
timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > gpurun_out/${TAG}_memcheck.log 2>&1; tail -2 gpurun_out/${TAG}_memcheck.log
