#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_general.py tests/test_gpu_rates.py -m gpu -q --maxfail=40 --tb=short --durations=10 > gpurun_out/pytest_r2d.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r2d.log
tail -70 gpurun_out/pytest_r2d.log
