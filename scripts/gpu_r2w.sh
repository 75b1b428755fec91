#!/bin/bash
TAG=${1:-r2w}
mkdir -p gpurun_out
RCG_CONFIGS=config5 python scripts/config_runs.py > gpurun_out/${TAG}_cfg5.jsonl 2>&1; cut -c1-330 gpurun_out/${TAG}_cfg5.jsonl
( time timeout 2400 python -m pytest tests/test_plugin_path.py tests/test_hotswap.py tests/test_multi_device.py tests/test_full_size_configs.py tests/test_cli_and_audio_host.py -m gpu -q --durations=5 ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -12 gpurun_out/${TAG}_pytest_gpu.log
