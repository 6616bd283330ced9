#!/bin/bash
mkdir -p gpurun_out
# 1. adaptive with / without the |X|^2 planes: parity tests first, then cfg3
timeout 900 python -m pytest tests/test_gpu_extended_adaptive.py "tests/test_gpu_long.py::test_adaptive_long_track" tests/test_gpu_rates.py tests/test_gpu_formats.py -m gpu -q --maxfail=20 --tb=short > gpurun_out/pytest_r2e.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r2e.log
tail -15 gpurun_out/pytest_r2e.log
for V in 0 1; do
  timeout 600 python bench.py --clips-per-gpu 8 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --configs cfg3 --tune adaptive_vsq=$V > gpurun_out/cfg3_vsq$V.log 2>&1
  python - <<PY
import json
line=[l for l in open('gpurun_out/cfg3_vsq$V.log') if l.startswith('{')][-1]
d=json.loads(line)['configs']['cfg3']
print('adaptive_vsq=$V', round(d['ms_per_step'],3), {k: round(v['ms_per_step'],3) for k,v in d['kernels'].items()}, round(d['whole_path_frac_of_hbm_peak'],3))
PY
done
# 2. more k_topk shapes
for V in "128 512 1024" "256 768 1536" "192 768 1536" "384 1024 3072" "256 1024 1024"; do
  set -- $V
  REPET_EXTRA_NVCC_FLAGS="-DREPET_TOPK_THREADS=$1 -DREPET_TOPK_CAP=$2 -DREPET_TOPK_CHUNK=$3" python repet-python_b200/build.py --force > /dev/null 2>&1
  timeout 600 python bench.py --clips-per-gpu 8 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --configs cfg4 > gpurun_out/topk_$1_$2_$3.log 2>&1
  python - <<PY
import json
line=[l for l in open('gpurun_out/topk_$1_$2_$3.log') if l.startswith('{')][-1]
d=json.loads(line)['configs']['cfg4']
print('threads $1 cap $2 chunk $3:', round(d['ms_per_step'],3), 'topk', round(d['kernels']['k_topk']['ms_per_step'],3))
PY
done
