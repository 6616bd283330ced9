#!/usr/bin/env python
"""Throughput of the general float64 path on the shapes it exists for (not a BASELINE config): wall time of the drop-in
calls on a 10-minute stereo track at 96 kHz and on a 10-minute 5-channel track at 44.1 kHz."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "repet-python_b200"))
import numpy as np  # noqa: E402

import repet  # noqa: E402
import repet_synth  # noqa: E402

for label, fs, channels, seconds, methods in (("96 kHz stereo", 96000, 2, 600, ("original", "extended", "adaptive")),
                                              ("44.1 kHz 5 channels", 44100, 5, 600, ("original", "adaptive")),
                                              ("96 kHz stereo", 96000, 2, 120, ("sim", "simonline"))):
    x = repet_synth.make_clip(77, seconds * fs, channels, fs, 2048 if fs > 51200 else 1024, redraw_seconds=(60, 120)).T.astype(np.float64)
    for method in methods:
        getattr(repet, method)(x[: 20 * fs], fs)
        t0 = time.perf_counter()
        y = getattr(repet, method)(x, fs)
        dt = time.perf_counter() - t0
        print(json.dumps({"general_path": label, "method": method, "seconds_of_audio": seconds, "wall_s": dt,
                          "x_realtime": seconds / dt, "finite": bool(np.all(np.isfinite(y)))}), flush=True)
