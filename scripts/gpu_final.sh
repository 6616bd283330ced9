#!/bin/bash
# final pass of the round: full GPU test suite, smoke, the general-path rates, the full bench line, the launch list
mkdir -p gpurun_out
TAG=${TAG:-r2z}
timeout 2400 python -m pytest tests -m gpu -q --maxfail=60 --tb=short --durations=8 > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_$TAG.log
tail -16 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 900 python scripts/general_rate.py > gpurun_out/general_rate_$TAG.jsonl 2>&1; tail -8 gpurun_out/general_rate_$TAG.jsonl
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_$TAG.log 2>&1; echo "reference arm exit $?"; tail -c 600 gpurun_out/bench_reference_$TAG.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --clips-per-gpu 128 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --configs none > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list exit $?"
python - <<PY
import json
line=[l for l in open('gpurun_out/bench_$TAG.log') if l.startswith('{')][-1]
d=json.loads(line)
print({k: d[k] for k in ('value','ms_per_step','clocks')})
for k in ('e2e','e2e_pcm16','e2e_numpy_f64','host_link','cpu_baseline'): print(k, d.get(k))
print({k:(round(v['ms_total']/d['steps'],3), round(v.get('frac',0),3)) for k,v in d['roofline']['kernels'].items()}, d['roofline']['whole_path_frac'])
for name,c in d['configs'].items():
    print(name, round(c.get('ms_per_step',0),3), round(c.get('x_realtime',0)), {k: round(v['ms_per_step'],3) for k,v in c.get('kernels',{}).items()}, c.get('whole_path_frac_of_hbm_peak'), c.get('latency_ms_p50'), c.get('latency_ms_p99'), c.get('gemm',{}).get('issued_frac_of_tf32_peak'))
PY
