#!/bin/bash
# round 3 checkpoint on one GPU: bench line, ncu launch list of the bench, full captures of the headline kernel
# (W = 16384, config 2) and of the CTA-pair kernel (W = 65536, config 3 flags)
TAG=${1:-r3m}
mkdir -p gpurun_out
( time timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?" >> gpurun_out/${TAG}_bench.err; tail -n 4 gpurun_out/${TAG}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e-configs > gpurun_out/${TAG}_b_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stretch_kernel -s 3 -c 1 -o gpurun_out/${TAG}_prof python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/${TAG}_b_ncu2.log 2>&1
tail -n 2 gpurun_out/${TAG}_b_ncu2.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stretch_pair -s 2 -c 1 -o gpurun_out/${TAG}_pair python scripts/kernel_time.py --tag pair --window 65536 --factor 4 --pitch 1 --reps 1 > gpurun_out/${TAG}_pair.log 2>&1
tail -n 2 gpurun_out/${TAG}_pair.log
cut -c1-600 gpurun_out/${TAG}_bench.json
