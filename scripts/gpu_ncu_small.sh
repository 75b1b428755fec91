#!/bin/bash
TAG=${1:-r2s}
mkdir -p gpurun_out
for W in 1024 4096; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stretch_kernel -s 3 -c 1 -o gpurun_out/${TAG}_w${W}_prof python scripts/kernel_time.py --window $W --factor 4 --pitch 1 --reps 1 > gpurun_out/${TAG}_w${W}_ncu.log 2>&1
tail -1 gpurun_out/${TAG}_w${W}_ncu.log
done
