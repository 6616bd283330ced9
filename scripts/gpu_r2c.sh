#!/bin/bash
# sim iteration: parity tests of the sim path, then cfg4 with both GEMM tile shapes, then an ncu capture of the GEMM / topk
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sim.py tests/test_gpu_helpers.py "tests/test_gpu_long.py::test_sim_long_tracks" tests/test_gpu_edges.py -m gpu -q --maxfail=20 --tb=short > gpurun_out/pytest_r2c.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r2c.log
tail -25 gpurun_out/pytest_r2c.log
for BN in 128 256; do
  timeout 600 python bench.py --clips-per-gpu 8 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --configs cfg4 --tune simgemm_bn=$BN > gpurun_out/cfg4_bn$BN.log 2>&1
  python - <<PY
import json
line=[l for l in open('gpurun_out/cfg4_bn$BN.log') if l.startswith('{')][-1]
d=json.loads(line)['configs']['cfg4']
print('BN=$BN', round(d['ms_per_step'],3), {k: round(v['ms_per_step'],3) for k,v in d['kernels'].items()}, d.get('gemm',{}).get('issued_frac_of_tf32_peak'))
PY
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_simgemm|k_topk' -c 2 -f -o gpurun_out/prof_sim_r2c \
    python bench.py --clips-per-gpu 8 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --configs cfg4 > gpurun_out/ncu_sim_r2c.log 2>&1
echo "sim capture exit $?"
