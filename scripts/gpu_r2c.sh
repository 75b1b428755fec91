#!/bin/bash
# round 2: 2-GPU box -- plugin + multi-device + NCCL + full-size config tests
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2c_gpu.txt
( time timeout 3000 python -m pytest tests/test_plugin_path.py tests/test_multi_device.py tests/test_multi_gpu.py tests/test_full_size_configs.py tests/test_hotswap.py -m gpu -q --durations=15 ) > gpurun_out/r2c_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest_gpu.log
tail -60 gpurun_out/r2c_pytest_gpu.log
