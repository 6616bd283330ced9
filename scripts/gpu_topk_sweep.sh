#!/bin/bash
# experiment: k_topk CTA shape (threads / candidate cap / chunk); rebuilds the library on the box per variant
mkdir -p gpurun_out
for V in "512 2048 4096" "256 1024 2048" "256 2048 2048" "384 1536 3072"; do
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
