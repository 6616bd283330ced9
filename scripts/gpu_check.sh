#!/bin/bash
# Run on the GPU box (via gpurun): parity tests, smoke, a short bench.  Logs go to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=60 --tb=short > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
timeout 900 python bench.py --clips-per-gpu ${BENCH_CLIPS:-128} --steps 3 --warmup 3 > gpurun_out/bench_small.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_small.log
tail -5 gpurun_out/bench_small.log
