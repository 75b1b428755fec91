#!/bin/bash
# why Tensor Memory tables are slower at W <= 8192: ncu of both builds at W = 8192
TAG=${TAG:-r3i}
mkdir -p gpurun_out
RCG_LIB=variants_tmp/lib_small_tables_8192.so timeout 300 ncu --set full --clock-control none --import-source on -k regex:stretch_kernel -s 2 -c 1 -o gpurun_out/${TAG}_tables python scripts/kernel_time.py --tag tables --window 8192 --factor 4 --pitch 1 --reps 1 > gpurun_out/${TAG}_a.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stretch_kernel -s 2 -c 1 -o gpurun_out/${TAG}_tmem python scripts/kernel_time.py --tag tmem --window 8192 --factor 4 --pitch 1 --reps 1 > gpurun_out/${TAG}_b.log 2>&1
tail -2 gpurun_out/${TAG}_a.log gpurun_out/${TAG}_b.log
