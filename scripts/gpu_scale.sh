#!/bin/bash
# multi-GPU: the copy-only host-link probe and the bench line at N ranks (N = number of GPUs gpurun gave us)
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --probe-host-link --gpus $N --steps 5 > gpurun_out/host_link_n$N.json 2> gpurun_out/host_link_n$N.err
cat gpurun_out/host_link_n$N.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
echo "bench exit $?"
python - <<PY
import json
line=[l for l in open('gpurun_out/bench_n$N.log') if l.startswith('{')][-1]
d=json.loads(line)
print({k: d[k] for k in ('n_gpus','value','ms_per_step')})
for k in ('host_link','e2e','e2e_pcm16'): print(k, d.get(k))
PY
tail -3 gpurun_out/bench_n$N.err
