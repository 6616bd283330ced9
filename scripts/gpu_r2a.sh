#!/bin/bash
# Round 2, first GPU pass: full GPU test suite (with the at-size parity tests), smoke, the full bench line with the
# configs block, and the copy-only host-link probe.  Logs go to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q --maxfail=60 --tb=short --durations=15 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -45 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2a.log 2> gpurun_out/bench_r2a.err; echo "bench exit $?" >> gpurun_out/bench_r2a.err
tail -c 6000 gpurun_out/bench_r2a.log; tail -5 gpurun_out/bench_r2a.err
timeout 300 python bench.py --probe-host-link --steps 5 > gpurun_out/host_link_n1.json 2>> gpurun_out/bench_r2a.err
cat gpurun_out/host_link_n1.json
