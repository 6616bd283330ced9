#!/bin/bash
# Round-2 ncu evidence (run under gpurun, 1 GPU).  Numbers printed by bench.py under ncu are NOT bench values.
#   1. launch list (gpu__time_duration) of the bench command: cfg2 main path
#   2. --set full of the cfg2 kernels (7 launches of one step)
#   3. --set full of the sim kernels on a 10-minute track (cfg4) and of the adaptive kernels (cfg3, 4 tracks)
mkdir -p gpurun_out
TAG=${TAG:-r2a}
CLIPS=${PROFILE_CLIPS:-128}
COMMON="--no-e2e --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --clips-per-gpu $CLIPS --steps 2 --warmup 3 $COMMON --configs none > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ -s 21 -c 7 -f -o gpurun_out/prof_${TAG} \
    python bench.py --clips-per-gpu $CLIPS --steps 1 --warmup 3 $COMMON --configs none > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full capture exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_simgemm|k_topk|k_simmodel_large|k_frames64' -c 5 -f -o gpurun_out/prof_sim_${TAG} \
    python bench.py --clips-per-gpu 8 --steps 1 --warmup 3 $COMMON --configs cfg4 > gpurun_out/ncu_sim_${TAG}.log 2>&1
echo "sim capture exit $?"
# the adaptive kernels by their demangled names (k_beat<1024>, the Vsq-writing k_stft), skipping the main arm's launches
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'k_adaptive_model|k_beat<1024>|k_stft<2, 4, false, true>|k_mask_istft' -s 4 -c 5 -f -o gpurun_out/prof_adaptive_${TAG} \
    python bench.py --clips-per-gpu 8 --steps 1 --warmup 3 $COMMON --configs cfg3 --cfg3-tracks 4 > gpurun_out/ncu_adaptive_${TAG}.log 2>&1
echo "adaptive capture exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_online_select' -c 1 -f -o gpurun_out/prof_online_${TAG} \
    python bench.py --clips-per-gpu 8 --steps 1 --warmup 3 $COMMON --configs cfg5 > gpurun_out/ncu_online_${TAG}.log 2>&1
echo "online capture exit $?"
ls -la gpurun_out/ | tail -12
