#!/bin/bash
TAG=${1:-r2m}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stretch_pair_kernel -s 3 -c 1 -o gpurun_out/${TAG}_pair_prof python scripts/kernel_time.py --window 65536 --factor 4 --pitch 1 --reps 1 > gpurun_out/${TAG}_pair_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_pair_ncu.log
