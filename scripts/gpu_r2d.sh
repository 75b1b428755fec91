#!/bin/bash
# round 2: packed-FP32 kernel -- timing of all configs, GPU tests, ncu capture of the headline kernel
TAG=${1:-r2d}
mkdir -p gpurun_out
python scripts/kernel_time.py --tag cfg2 > gpurun_out/${TAG}_times.jsonl 2>&1
python scripts/kernel_time.py --tag cfg3_16384 --factor 4 --pitch 1 >> gpurun_out/${TAG}_times.jsonl 2>&1
python scripts/kernel_time.py --tag cfg3_1024 --factor 4 --pitch 1 --window 1024 >> gpurun_out/${TAG}_times.jsonl 2>&1
python scripts/kernel_time.py --tag cfg3_4096 --factor 4 --pitch 1 --window 4096 >> gpurun_out/${TAG}_times.jsonl 2>&1
python scripts/kernel_time.py --tag cfg3_65536 --factor 4 --pitch 1 --window 65536 >> gpurun_out/${TAG}_times.jsonl 2>&1
python scripts/kernel_time.py --tag cfg4_10min --factor 20 --pitch -3 >> gpurun_out/${TAG}_times.jsonl 2>&1
cat gpurun_out/${TAG}_times.jsonl
( time timeout 2400 python -m pytest tests -m gpu -q -x --durations=10 ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -25 gpurun_out/${TAG}_pytest_gpu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stretch_kernel -s 3 -c 1 -o gpurun_out/${TAG}_prof python scripts/kernel_time.py --reps 1 > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
free -g > gpurun_out/${TAG}_free.txt; nproc >> gpurun_out/${TAG}_free.txt
( time timeout 900 python bench.py --steps 5 --warmup 3 ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?" >> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_free.txt
