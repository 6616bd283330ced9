#!/bin/bash
# experiment: run the `original` pipeline in L2-sized clip groups by capping the per-chunk workspace
mkdir -p gpurun_out
for MB in 0 4096 1024 512 256 128; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --configs none --workspace-mb $MB > gpurun_out/l2groups_$MB.log 2>&1
  python - <<PY
import json
line=[l for l in open('gpurun_out/l2groups_$MB.log') if l.startswith('{')][-1]
d=json.loads(line)
k=d['roofline']['kernels']
print('workspace_mb $MB: ms/step', round(d['ms_per_step'],3), 'launches/step', d['gpu_launches']//d['steps'], {n:round(v['ms_total']/d['steps'],3) for n,v in k.items()})
PY
done
