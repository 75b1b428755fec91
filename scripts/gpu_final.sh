#!/bin/bash
# closing measurements of a round on one GPU: all BASELINE configs device-resident + e2e, bench line, ncu launch list
# of the bench, one full capture of the headline kernel, the whole GPU suite
TAG=${1:-r3z}
mkdir -p gpurun_out
RCG_E2E=1 python scripts/config_runs.py > gpurun_out/${TAG}_configs_n1.jsonl 2> gpurun_out/${TAG}_configs.err; tail -n 3 gpurun_out/${TAG}_configs.err
cut -c1-400 gpurun_out/${TAG}_configs_n1.jsonl
( time timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?" >> gpurun_out/${TAG}_bench.err; tail -n 4 gpurun_out/${TAG}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/${TAG}_b_ncu.log 2>&1
grep -c "stretch_kernel" gpurun_out/${TAG}_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stretch_kernel -s 3 -c 1 -o gpurun_out/${TAG}_prof python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/${TAG}_b_ncu2.log 2>&1
tail -n 2 gpurun_out/${TAG}_b_ncu2.log
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -n 12 gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
cut -c1-700 gpurun_out/${TAG}_bench.json
