#!/bin/bash
# Tuning sweep on the GPU box: prints one compact line per configuration.
mkdir -p gpurun_out
CLIPS=${SWEEP_CLIPS:-256}
run() {
  python bench.py --clips-per-gpu $CLIPS --steps 5 --warmup 3 --no-e2e --no-cpu-baseline "$@" 2>gpurun_out/sweep_err.log | python -c "
import json,sys
for line in sys.stdin:
    try: d=json.loads(line)
    except Exception: continue
    k=d['roofline']['kernels']
    print('%-46s value %9.0f ms/step %7.3f | '%(' '.join(sys.argv[1:]), d['value'], d['ms_per_step']) + ' '.join('%s %.3f'%(n[2:],k[n]['ms_total']/d['steps']) for n in k), 'clk', d['clocks']['sm_mhz'])
" "$@"
}
for cfg in "$@"; do run $cfg; done
