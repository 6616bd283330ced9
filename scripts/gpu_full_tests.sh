#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --maxfail=60 --tb=short --durations=12 > gpurun_out/pytest_gpu_full.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_full.log
tail -40 gpurun_out/pytest_gpu_full.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
