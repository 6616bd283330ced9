#!/usr/bin/env python
"""Throughput of the other BASELINE.json configs on one GPU (device-resident, CUDA events):
  cfg3  repet.adaptive on 10-min stereo tracks
  cfg4  repet.sim on a 10-min stereo track (T = 25 841: 1.37 TFLOP similarity GEMM)
  cfg5  repet.simonline and repet.extended on 1 hour of audio
Prints one JSON line per config with the per-kernel time split.  Not the driver's bench.py."""
import argparse
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "repet-python_b200"))
import numpy as np  # noqa: E402

FS = 44100


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="cfg3,cfg4,cfg5")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--tracks", type=int, default=4)
    ap.add_argument("--tune", default="")
    ap.add_argument("--stream-seconds", type=int, default=120, help="length of the streamed simonline latency test")
    args = ap.parse_args()
    import repet_synth

    specs = {
        "cfg3": ("adaptive", 600 * FS, args.tracks),
        "cfg4": ("sim", 600 * FS, 1),
        "cfg5_simonline": ("simonline", 155038 * 1024 + 2048, 1),
        "cfg5_extended": ("extended", 3600 * FS, 1),
    }
    wanted = []
    for name in args.configs.split(","):
        wanted += [k for k in specs if k.startswith(name)]
    clips = {}
    for name in wanted:
        driver, S, B = specs[name]
        t0 = time.perf_counter()
        clips[name] = repet_synth.make_batch(7000, B, S, processes=True, redraw_seconds=(60, 120))
        print("# synthesised %s: %d x %.0f s in %.1f s" % (name, B, S / FS, time.perf_counter() - t0), file=sys.stderr)

    import torch

    import repet

    device = torch.device("cuda", 0)
    torch.cuda.set_device(device)
    handle = repet._host.Handle(0)
    if args.tune:
        repet._host.set_tuning(**{k: int(v) for k, v in (kv.split("=") for kv in args.tune.split(","))})
    stream = torch.cuda.Stream(device)
    torch.cuda.set_stream(stream)
    handle.set_stream(stream.cuda_stream)
    tun = repet._tunables()
    for name in wanted:
        driver, S, B = specs[name]
        params, _ = repet._host.derive_params(FS, tun, driver)
        handle.ensure_window(params.window_length)
        audio = torch.from_numpy(clips[name]).to(device)
        out = torch.empty_like(audio)
        fn = getattr(handle.lib, "repet_%s_batch_dev" % driver)

        def step():
            handle.check(fn(handle.h, ctypes.c_void_p(audio.data_ptr()), B, 2, S, ctypes.byref(params),
                            ctypes.c_void_p(out.data_ptr()), None, None))

        step()
        torch.cuda.synchronize()
        handle.profile_read(reset=True)
        handle.set_profiling(True)
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record(stream)
        for _ in range(args.steps):
            step()
        end.record(stream)
        torch.cuda.synchronize()
        ms = start.elapsed_time(end) / args.steps
        prof = handle.profile_read(reset=True)
        handle.set_profiling(False)
        audio_seconds = B * S / FS
        line = {"config": name, "driver": driver, "tracks": B, "seconds_each": S / FS, "ms_per_step": ms,
                "x_realtime": audio_seconds / (ms / 1e3),
                "kernels_ms": {k: v[0] / args.steps for k, v in prof.items()}}
        if driver == "sim":
            T = int(np.ceil(S / 1024)) + 1
            gemm_ms = prof.get("k_simgemm", (0, 0))[0] / args.steps
            if gemm_ms:
                # only the triangle of S = A A^T is formed (square tiles), so the useful work is
                # T (T+1) F flop; the full-product figure says what a caller of the whole matrix gets per second
                line["gemm_useful_tflops"] = float(T) * (T + 1) * 1025 / (gemm_ms / 1e3) / 1e12
                line["gemm_full_product_equivalent_tflops"] = 2.0 * T * T * 1025 / (gemm_ms / 1e3) / 1e12
                line["gemm_tf32_mma_tflops_issued"] = 3 * line["gemm_useful_tflops"] * 1056 / 1025
        print(json.dumps(line), flush=True)
        del audio, out
        torch.cuda.empty_cache()
    if any(name.startswith("cfg5") for name in wanted) and args.stream_seconds > 0:
        # cfg5: latency per 1 s block of the streaming front end (host NumPy float64 in and out)
        x = repet_synth.make_clip(7100, args.stream_seconds * FS, redraw_seconds=(60, 120)).T.astype(np.float64)
        handle.set_stream(None)
        saved = repet._host._handles.get(0)
        repet._host._handles[0] = handle
        stream_obj = repet.SimOnline(FS, 2)
        latencies = []
        for k in range(0, len(x), FS):
            t0 = time.perf_counter()
            stream_obj.process(x[k : k + FS])
            latencies.append(1e3 * (time.perf_counter() - t0))
        stream_obj.flush()
        steady = np.array(latencies[12:])  # after the 10 s warm-up of the algorithm
        print(json.dumps({"config": "cfg5_simonline_stream", "block_seconds": 1.0, "blocks": len(steady),
                          "latency_ms_p50": float(np.percentile(steady, 50)), "latency_ms_p99": float(np.percentile(steady, 99)),
                          "latency_ms_max": float(steady.max()), "x_realtime_per_stream": 1e3 / float(np.mean(steady))}), flush=True)
        if saved is not None:
            repet._host._handles[0] = saved


if __name__ == "__main__":
    main()
