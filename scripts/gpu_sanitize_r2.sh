#!/bin/bash
# compute-sanitizer on the kernels added or changed in round 2 (small inputs: the tools slow execution 10-50x)
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import sys, os
sys.path.insert(0, "repet-python_b200"); sys.path.insert(0, "oracle")
import numpy as np, repet, repet_synth
FS = 44100
x = repet_synth.make_clip(3, 12 * FS + 321).T.astype(np.float64)
for name in ("original", "extended", "adaptive", "sim", "simonline"):
    y = getattr(repet, name)(x, FS); assert np.all(np.isfinite(y)), name
audio = repet_synth.make_batch(900, 2, 7 * FS)
pcm = np.clip(np.rint(np.transpose(audio, (0, 2, 1)) * 32768.0), -32768, 32767).astype(np.int16)
for m in ("original", "adaptive", "sim"):
    q, ints = repet.separate_batch(pcm, FS, m, in_format="pcm16", out_format="pcm16")
repet._host.set_tuning(topk_force_exact=1); repet.sim(x, FS); repet._host.set_tuning(topk_force_exact=0)
x3 = repet_synth.make_clip(5, 6 * FS, 3).T.astype(np.float64)
for name in ("original", "adaptive", "sim"):
    y = getattr(repet, name)(x3, FS)
x96 = repet_synth.make_clip(6, 4 * 96000, 2, 96000, 2048).T.astype(np.float64)
repet.original(x96, 96000)
import scipy.signal.windows as W
rng = np.random.default_rng(1); s = rng.standard_normal(9000)
X = repet._stft(s, W.hamming(1000, sym=False), 300); repet._istft(X, W.hamming(1000, sym=False), 300)
repet._beatspectrum(np.abs(rng.standard_normal((60, 1500)))); repet._acorr(np.abs(rng.standard_normal((1500, 9))))
saved = repet.buffer_length; repet.buffer_length = 20
repet.simonline(repet_synth.make_clip(7, 23 * FS).T.astype(np.float64), FS); repet.buffer_length = saved
print("sanitizer case done")
PY
for TOOL in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $TOOL --print-limit 20 python /tmp/san_case.py > gpurun_out/sanitizer_r2_$TOOL.log 2>&1
  echo "$TOOL exit $?"; tail -4 gpurun_out/sanitizer_r2_$TOOL.log
done
