#!/bin/bash
# ncu evidence (run under gpurun, 1 GPU): per-launch device times and one --set full capture of
# every kernel of one step (7 launches per step: stft, beat, periods, certify, finalize, model, mask_istft).  Numbers printed by bench.py under ncu are NOT bench values.
mkdir -p gpurun_out
TAG=${TAG:-r1}
CLIPS=${PROFILE_CLIPS:-128}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --clips-per-gpu $CLIPS --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "launch list exit $?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_ -s 21 -c 7 -f -o gpurun_out/prof_${TAG} \
    python bench.py --clips-per-gpu $CLIPS --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out/
