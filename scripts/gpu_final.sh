#!/bin/bash
# round 2, closing measurements on one GPU: all BASELINE configs device-resident + e2e, bench line, ncu launch list of
# the bench, one full capture of the headline kernel
TAG=${1:-r2z}
mkdir -p gpurun_out
RCG_E2E=1 python scripts/config_runs.py > gpurun_out/${TAG}_configs_n1.jsonl 2> gpurun_out/${TAG}_configs.err; tail -3 gpurun_out/${TAG}_configs.err
cat gpurun_out/${TAG}_configs_n1.jsonl | cut -c1-400
( time timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?" >> gpurun_out/${TAG}_bench.err; tail -4 gpurun_out/${TAG}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e-configs > gpurun_out/${TAG}_b_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stretch_kernel -s 3 -c 1 -o gpurun_out/${TAG}_prof python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/${TAG}_b_ncu2.log 2>&1
tail -2 gpurun_out/${TAG}_b_ncu2.log
