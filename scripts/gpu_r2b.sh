#!/bin/bash
# Round 2, second GPU pass: the new tests only (formats, long goldens that exist, general helpers, exact fallback),
# then ncu launch lists and --set full captures of the sim kernels and the beat kernel.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_formats.py tests/test_gpu_general.py tests/test_gpu_long.py tests/test_gpu_sim.py tests/test_gpu_helpers.py -m gpu -q --maxfail=30 --tb=short --durations=8 > gpurun_out/pytest_r2b.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r2b.log
tail -60 gpurun_out/pytest_r2b.log
