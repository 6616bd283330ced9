#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_general.py tests/test_gpu_formats.py -m gpu -q --maxfail=20 --tb=short --durations=5 > gpurun_out/pytest_r2j.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r2j.log
tail -20 gpurun_out/pytest_r2j.log
timeout 900 python scripts/general_rate.py > gpurun_out/general_rate_r2j.jsonl 2>&1; tail -8 gpurun_out/general_rate_r2j.jsonl
