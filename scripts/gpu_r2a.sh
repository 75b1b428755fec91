#!/bin/bash
# round 2, first GPU call: packed-FP32 issue microbenchmark + ncu capture of the default W=65536 pair kernel
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2a_gpu.txt 2>&1
./scripts/microbench/f32x2_issue > gpurun_out/r2a_f32x2_issue.txt 2>&1
python scripts/kernel_time.py --window 65536 --factor 4 --pitch 1 --tag pair > gpurun_out/r2a_pair_time.json 2>&1
python scripts/kernel_time.py --tag cfg2 > gpurun_out/r2a_cfg2_time.json 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stretch_pair_kernel -s 3 -c 1 -o gpurun_out/r2a_pair_prof python scripts/kernel_time.py --window 65536 --factor 4 --pitch 1 --reps 1 > gpurun_out/r2a_pair_ncu.log 2>&1
cat gpurun_out/r2a_f32x2_issue.txt gpurun_out/r2a_pair_time.json gpurun_out/r2a_cfg2_time.json
