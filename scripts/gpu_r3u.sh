#!/bin/bash
# ncu captures of the full-rate modes of the W = 16384 kernel: config 4's flags (f 20, p -3) and config 3's (f 4, p 1)
TAG=${TAG:-r3u}
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stretch_kernel -s 2 -c 1 -o gpurun_out/${TAG}_pm3 python scripts/kernel_time.py --tag pm3 --window 16384 --factor 20 --pitch -3 --reps 1 > gpurun_out/${TAG}_pm3.log 2>&1
tail -n 1 gpurun_out/${TAG}_pm3.log | cut -c1-300
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stretch_kernel -s 2 -c 1 -o gpurun_out/${TAG}_p1 python scripts/kernel_time.py --tag p1 --window 16384 --factor 4 --pitch 1 --reps 1 > gpurun_out/${TAG}_p1.log 2>&1
tail -n 1 gpurun_out/${TAG}_p1.log | cut -c1-300
