#!/bin/bash
# final-ish N = 1 bench (full line with configs) + new stream tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sim.py tests/test_gpu_general.py tests/test_gpu_extended_adaptive.py -m gpu -q --maxfail=20 --tb=short > gpurun_out/pytest_r2g.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r2g.log
tail -12 gpurun_out/pytest_r2g.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2g.log 2> gpurun_out/bench_r2g.err; echo "bench exit $?"
python - <<PY
import json
line=[l for l in open('gpurun_out/bench_r2g.log') if l.startswith('{')][-1]
d=json.loads(line)
print({k: d[k] for k in ('value','ms_per_step','clocks')})
for k in ('e2e','e2e_pcm16','e2e_numpy_f64','host_link','cpu_baseline'): print(k, d.get(k))
print({k:(round(v['ms_total']/d['steps'],3), round(v.get('frac',0),3)) for k,v in d['roofline']['kernels'].items()}, d['roofline']['whole_path_frac'])
for name,c in d['configs'].items():
    print(name, round(c.get('ms_per_step',0),3), round(c.get('x_realtime',0)), {k: round(v['ms_per_step'],3) for k,v in c.get('kernels',{}).items()}, c.get('whole_path_frac_of_hbm_peak'), c.get('latency_ms_p50'), c.get('latency_ms_p99'))
PY
