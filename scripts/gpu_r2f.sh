#!/bin/bash
# verification of the new k_topk shape + adaptive |X|^2 planes, then ncu captures of the adaptive kernels and the sim kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_sim.py tests/test_gpu_long.py tests/test_gpu_edges.py tests/test_gpu_byproducts.py -m gpu -q --maxfail=20 --tb=short > gpurun_out/pytest_r2f.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r2f.log
tail -12 gpurun_out/pytest_r2f.log
COMMON="--no-e2e --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'k_adaptive_model|k_beat<1024>|k_stft<2, 4, false, true>|k_mask_istft' -s 4 -c 5 -f -o gpurun_out/prof_adaptive_r2f \
    python bench.py --clips-per-gpu 8 --steps 1 --warmup 3 $COMMON --configs cfg3 --cfg3-tracks 4 > gpurun_out/ncu_adaptive_r2f.log 2>&1
echo "adaptive capture exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_simgemm|k_topk|k_simmodel_large|k_frames64' -c 5 -f -o gpurun_out/prof_sim_r2f \
    python bench.py --clips-per-gpu 8 --steps 1 --warmup 3 $COMMON --configs cfg4 > gpurun_out/ncu_sim_r2f.log 2>&1
echo "sim capture exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_online_select' -c 1 -f -o gpurun_out/prof_online_r2f \
    python bench.py --clips-per-gpu 8 --steps 1 --warmup 3 $COMMON --configs cfg5 > gpurun_out/ncu_online_r2f.log 2>&1
echo "online capture exit $?"
