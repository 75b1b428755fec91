#!/bin/bash
TAG=${1:-r2k}
mkdir -p gpurun_out
python scripts/kernel_time.py --tag cfg3_65536 --factor 4 --pitch 1 --window 65536 > gpurun_out/${TAG}_times.jsonl 2>&1
python scripts/kernel_time.py --tag f8p2_65536 --factor 8 --pitch 2 --window 65536 >> gpurun_out/${TAG}_times.jsonl 2>&1
python scripts/kernel_time.py --tag f20pm3_65536 --factor 20 --pitch -3 --window 65536 >> gpurun_out/${TAG}_times.jsonl 2>&1
cat gpurun_out/${TAG}_times.jsonl
( time timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_full_size_configs.py tests/test_plugin_path.py tests/test_multi_device.py -m gpu -q -k "65536 or 65536 or window_65536 or cpp_one or config3" --durations=5 ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -15 gpurun_out/${TAG}_pytest_gpu.log
timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_small.py > gpurun_out/${TAG}_racecheck.log 2>&1; tail -5 gpurun_out/${TAG}_racecheck.log
