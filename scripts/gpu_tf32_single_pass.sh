for TC in 2 1; do
  timeout 600 python bench.py --clips-per-gpu 8 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --configs cfg4 --tune simgemm_tc=$TC > gpurun_out/cfg4_tc$TC.log 2>&1
  python - <<PY
import json
line=[l for l in open('gpurun_out/cfg4_tc$TC.log') if l.startswith('{')][-1]
d=json.loads(line)['configs']['cfg4']
print('simgemm_tc=$TC', round(d['ms_per_step'],3), {k: round(v['ms_per_step'],3) for k,v in d['kernels'].items()})
PY
done
