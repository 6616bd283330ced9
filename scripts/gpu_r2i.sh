#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_formats.py tests/test_gpu_sim.py "tests/test_gpu_long.py::test_sim_long_tracks" "tests/test_gpu_long.py::test_simonline_long_stream" tests/test_gpu_edges.py tests/test_gpu_helpers.py -m gpu -q --maxfail=20 --tb=short > gpurun_out/pytest_r2i.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r2i.log
tail -25 gpurun_out/pytest_r2i.log
timeout 600 python bench.py --clips-per-gpu 8 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --configs cfg4,cfg5 > gpurun_out/cfg45_r2i.log 2>&1
python - <<PY
import json
line=[l for l in open('gpurun_out/cfg45_r2i.log') if l.startswith('{')][-1]
for name,c in json.loads(line)['configs'].items():
    print(name, round(c.get('ms_per_step',0),3), {k: round(v['ms_per_step'],3) for k,v in c.get('kernels',{}).items()}, c.get('latency_ms_p50'))
PY
