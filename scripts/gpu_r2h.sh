#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_extended_adaptive.py "tests/test_gpu_long.py::test_adaptive_long_track" tests/test_gpu_rates.py tests/test_gpu_byproducts.py tests/test_gpu_parity.py -m gpu -q --maxfail=20 --tb=short > gpurun_out/pytest_r2h.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r2h.log
tail -25 gpurun_out/pytest_r2h.log
for V in 0 1; do
  timeout 600 python bench.py --clips-per-gpu 8 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --configs cfg3 --tune adaptive_halves=$V > gpurun_out/cfg3_halves$V.log 2>&1
  python - <<PY
import json
line=[l for l in open('gpurun_out/cfg3_halves$V.log') if l.startswith('{')][-1]
d=json.loads(line)['configs']['cfg3']
print('adaptive_halves=$V', round(d['ms_per_step'],3), {k: round(v['ms_per_step'],3) for k,v in d['kernels'].items()}, round(d['whole_path_frac_of_hbm_peak'],3))
PY
done
