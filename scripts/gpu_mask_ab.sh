#!/bin/bash
# A/B of a k_mask_istft variant selected by a compile-time macro: rebuild on the box, parity tests, then the cfg2 bench
mkdir -p gpurun_out
for FLAGS in "" "$1"; do
  REPET_EXTRA_NVCC_FLAGS="$FLAGS" python repet-python_b200/build.py --force > /dev/null 2>&1
  if [ -n "$FLAGS" ]; then
    timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sim.py tests/test_gpu_extended_adaptive.py tests/test_gpu_edges.py -m gpu -q -x --tb=short 2>&1 | tail -3
  fi
  for i in 1 2; do
    timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --configs none > gpurun_out/mask_ab.log 2>&1
    python - <<PY
import json
line=[l for l in open('gpurun_out/mask_ab.log') if l.startswith('{')][-1]
d=json.loads(line)
print('flags [$FLAGS] ms/step', round(d['ms_per_step'],3), {n:round(v['ms_total']/d['steps'],3) for n,v in d['roofline']['kernels'].items()}, d['clocks']['sm_mhz'], d['clocks']['reasons'])
PY
  done
done
