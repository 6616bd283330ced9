#!/usr/bin/env python
"""
bench.py -- REPET separation throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...     # the reference algorithm on the host cores

Workload (BASELINE.json configs[1]): `repet.original` on a batch of synthetic 30 s stereo
44.1 kHz clips, 512 clips per GPU (4096 over 8 GPUs; weak scaling, clips are independent, no
collective on the data path).  One "step" = one pass of the whole hot path (STFT, beat
spectrum, period, median model, mask, ISTFT) over the rank's 512 clips.

Printed JSON line (rank 0):
  value     audio-seconds separated per second, all ranks, inputs/outputs resident in HBM,
            timed with CUDA events on the launching stream, max over ranks
  e2e       same metric through the C ABI with HOST (pinned) buffers: H2D of the inputs and
            D2H of the results inside the timed region
  roofline  the dominant kernel's algorithmic bytes / its event-timed duration vs the measured
            HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline  the oracle port of the reference timed on this box's host cores on a bounded
            sample of the same clips (N=1, rank 0 only)
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "repet-python_b200"))

import numpy as np  # noqa: E402

FS = 44100
CLIP_SECONDS = 30
CLIP_SAMPLES = CLIP_SECONDS * FS  # 1 323 000
CHANNELS = 2
METRIC = "audio-seconds separated/sec (x realtime), repet.original, 30 s stereo 44.1 kHz clips"
UNIT = "audio-s/s"
TUNABLES = dict(
    cutoff_frequency=100, period_range=[1, 10], segment_length=10, segment_step=5, filter_order=5,
    similarity_threshold=0, similarity_distance=1, similarity_number=100, buffer_length=10,
)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--clips-per-gpu", type=int, default=512)
    ap.add_argument("--cpu-sample-clips", type=int, default=0, help="clips of the CPU baseline sample (0 = 2 per core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not bind the rank to the GPU's NUMA-local CPUs")
    ap.add_argument("--workspace-mb", type=int, default=0, help="per-chunk workspace cap of the library (0 = default)")
    ap.add_argument("--tune", default="", help="comma separated knob=value pairs for repet_set_tuning (experiments)")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------
# CPU side: the oracle port of the reference on the host cores
# --------------------------------------------------------------------------------------------
def _cpu_worker_init():
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["MKL_NUM_THREADS"] = "1"
    sys.path.insert(0, os.path.join(ROOT, "oracle"))


def _cpu_worker(index):
    import repet_oracle
    import repet_synth

    clip = repet_synth.make_clip(index, CLIP_SAMPLES, CHANNELS, FS)
    x = clip.T.astype(np.float64)
    t0 = time.perf_counter()
    y = repet_oracle.original(x, FS)
    return time.perf_counter() - t0, float(y[0, 0])


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_pool_size():
    cores = host_cores()
    try:
        import psutil

        by_memory = int(psutil.virtual_memory().available / (1.5 * (1 << 30)))  # ~1 GB peak per worker
        return max(1, min(cores, by_memory))
    except Exception:
        return max(1, min(cores, 64))


class CpuArm:
    """The reference algorithm (oracle port: NumPy restatement of repet.py, see oracle/) on a
    process pool, one clip per task, BLAS/OpenMP pinned to one thread per worker."""

    def __init__(self):
        import multiprocessing

        self.workers = cpu_pool_size()
        self.pool = multiprocessing.get_context("fork").Pool(self.workers, initializer=_cpu_worker_init)

    def run(self, first_index, number_clips):
        """Returns (wall seconds of the separation only, audio seconds)."""
        t0 = time.perf_counter()
        results = self.pool.map(_cpu_worker, range(first_index, first_index + number_clips), chunksize=1)
        wall = time.perf_counter() - t0
        # clip synthesis runs inside the workers too; subtract nothing -- report compute time as the
        # sum of per-clip separation times divided by the worker count actually busy
        compute = sum(r[0] for r in results)
        busy = min(self.workers, number_clips)
        return compute / busy, wall, number_clips * CLIP_SECONDS

    def close(self):
        self.pool.close()
        self.pool.join()


def run_reference_arm(args, rank, world):
    """--impl reference: rank 0 alone times the CPU implementation; other ranks exit."""
    if rank != 0:
        return
    arm = CpuArm()
    sample = args.cpu_sample_clips or arm.workers
    for w in range(args.warmup):
        arm.run(10_000 + w * sample, min(sample, arm.workers))
    total_time = 0.0
    total_audio = 0.0
    for k in range(args.steps):
        compute, wall, audio = arm.run(20_000 + k * sample, sample)
        total_time += compute
        total_audio += audio
    arm.close()
    value = total_audio / total_time
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_time / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, sample_clips=sample),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.workers, "kind": "port",
                         "sample": "%d clips of 30 s per step (the full step is %d clips per GPU)" % (sample, args.clips_per_gpu)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def config_dict(args, sample_clips=None):
    cfg = {
        "workload": "BASELINE configs[1]: repet.original, synthetic 30 s stereo 44.1 kHz clips, %d clips per GPU "
                    "(4096 over 8 GPUs), 2048-pt STFT" % args.clips_per_gpu,
        "clips_per_gpu": args.clips_per_gpu, "clip_seconds": CLIP_SECONDS, "channels": CHANNELS,
        "sampling_frequency": FS, "parallelism": "clip-sharded, no collective",
        "l2": "inputs larger than L2 (%.1f GB of audio per step per GPU)" % (args.clips_per_gpu * CHANNELS * CLIP_SAMPLES * 4 / 1e9),
    }
    if sample_clips is not None:
        cfg["reference_sample_clips_per_step"] = sample_clips
    return cfg


# --------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled through NVML every few ms DURING the timed region."""

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.samples = []
        self.reasons = set()
        self.running = False
        self.thread = None
        self.max_mhz = None
        self.error = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nvml = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(self.gpu_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM))
        except Exception as exc:  # pragma: no cover
            self.error = "nvml unavailable: %r" % (exc,)
            return
        self.running = True
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def _loop(self):
        n = self.nvml
        flags = {
            "hw_slowdown": n.nvmlClocksEventReasonHwSlowdown if hasattr(n, "nvmlClocksEventReasonHwSlowdown") else 0x8,
            "hw_thermal_slowdown": 0x40,
            "sw_thermal_slowdown": 0x20,
            "sw_power_cap": 0x4,
        }
        while self.running:
            try:
                self.samples.append(float(n.nvmlDeviceGetClockInfo(self.dev, n.NVML_CLOCK_SM)))
                try:
                    mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.dev)
                except AttributeError:
                    mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
                for name, bit in flags.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception as exc:  # pragma: no cover
                self.error = repr(exc)
                break
            time.sleep(0.001)

    def stop(self):
        self.running = False
        if self.thread:
            self.thread.join(timeout=2)
        out = {
            "sm_mhz": float(np.median(self.samples)) if self.samples else None,
            "sm_max_mhz": self.max_mhz,
            "samples": len(self.samples),
            "reasons": sorted(self.reasons),
        }
        if self.error:
            out["error"] = self.error
        return out


# --------------------------------------------------------------------------------------------
# the CUDA arm
# --------------------------------------------------------------------------------------------
def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel, clips_per_launch):
    """dram__bytes_read + dram__bytes_write per launch of `kernel`, scaled from the committed
    ncu --set full capture (profiles/traffic.json holds bytes per clip), or None."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            return json.load(f)[kernel]["dram_bytes_per_clip"] * clips_per_launch
    except Exception:
        return None


def bind_to_gpu_numa_node(gpu_index):
    """Pin this process (and the pinned host buffers it allocates next: first touch) to the CPUs NVML reports
    as local to the GPU.  With several ranks per box the host-buffer copies otherwise cross the socket
    interconnect.  Best effort: returns a description, or why it was skipped."""
    try:
        import pynvml

        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        before = len(os.sched_getaffinity(0))
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        after = len(os.sched_getaffinity(0))
        if after == 0:
            raise RuntimeError("empty affinity")
        return "bound to %d of %d cpus (NVML affinity of GPU %d)" % (after, before, gpu_index)
    except Exception as exc:  # not fatal: containers may forbid it
        return "not bound: %r" % (exc,)


def run_b200_arm(args, rank, local_rank, world):
    import repet_synth

    numa = "disabled" if args.no_numa_bind else bind_to_gpu_numa_node(local_rank)
    B = args.clips_per_gpu
    # 1) synthesise this rank's clips on the host BEFORE CUDA is initialised (fork pool)
    t_gen = time.perf_counter()
    workers = max(1, host_cores() // max(1, world))
    host_audio = repet_synth.make_batch(rank * B, B, CLIP_SAMPLES, CHANNELS, FS, processes=True, workers=workers)
    t_gen = time.perf_counter() - t_gen

    import torch
    import torch.distributed as dist

    import repet

    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()

    handle = repet._host.Handle(local_rank)
    if args.tune:
        repet._host.set_tuning(**{k: int(v) for k, v in (kv.split("=") for kv in args.tune.split(","))})
    if args.workspace_mb:
        handle.set_workspace_limit(args.workspace_mb << 20)
    stream = torch.cuda.Stream(device)  # everything below is enqueued on this one stream
    torch.cuda.set_stream(stream)
    handle.set_stream(stream.cuda_stream)

    pinned_in = torch.from_numpy(host_audio).pin_memory()
    del host_audio
    pinned_out = torch.empty_like(pinned_in).pin_memory()
    audio_dev = pinned_in.to(device, non_blocking=True)
    out_dev = torch.empty_like(audio_dev)
    periods_dev = torch.zeros(B, dtype=torch.int32, device=device)
    torch.cuda.synchronize(device)

    def step_device():
        repet._host.original_batch_device(
            audio_dev.data_ptr(), out_dev.data_ptr(), B, CHANNELS, CLIP_SAMPLES, FS, TUNABLES, handle=handle,
            periods_ptr=periods_dev.data_ptr())

    params, _ = repet._host.derive_params(FS, TUNABLES)
    handle.ensure_window(params.window_length)
    import ctypes

    periods_host = np.zeros(B, dtype=np.int32)

    def step_host():
        handle.check(handle.lib.repet_original_batch(
            handle.h, ctypes.c_void_p(pinned_in.data_ptr()), B, CHANNELS, CLIP_SAMPLES, ctypes.byref(params),
            ctypes.c_void_p(pinned_out.data_ptr()), periods_host.ctypes.data_as(ctypes.c_void_p)))

    # ---- device-resident timing ---------------------------------------------------------------
    for _ in range(max(3, args.warmup)):
        step_device()
    torch.cuda.synchronize(device)
    handle.profile_read(reset=True)
    handle.set_profiling(True)
    sampler = ClockSampler(local_rank)
    launches_before = handle.launch_count()
    barrier()
    torch.cuda.synchronize(device)
    sampler.start()
    start = torch.cuda.Event(enable_timing=True)
    end = torch.cuda.Event(enable_timing=True)
    start.record(stream)
    t_host = time.perf_counter()
    for _ in range(args.steps):
        step_device()
    t_host = time.perf_counter() - t_host  # host time to ENQUEUE the steps (no synchronisation inside)
    end.record(stream)
    torch.cuda.synchronize(device)
    barrier()
    clocks = sampler.stop()
    elapsed_ms = start.elapsed_time(end)
    launches = handle.launch_count() - launches_before
    profile = handle.profile_read(reset=True)
    handle.set_profiling(False)
    periods_first = periods_dev.cpu().numpy().copy()

    # ---- end to end through the host-buffer ABI --------------------------------------------------
    e2e_ms = None
    if not args.no_e2e:
        for _ in range(1):
            step_host()
        barrier()
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_host()
        torch.cuda.synchronize(device)
        e2e_ms = 1e3 * (time.perf_counter() - t0)
        barrier()
        assert np.array_equal(periods_host, periods_first), "host-buffer and device-resident paths disagree"

    # ---- the same with int16 PCM input (what a WAVE file holds): half the H2D bytes ----------------
    pcm_ms = None
    if not args.no_e2e:
        pcm_host = torch.empty((B, CLIP_SAMPLES, CHANNELS), dtype=torch.int16).pin_memory()
        chunk = 32
        for lo in range(0, B, chunk):  # quantise the same clips to 16 bits, WAV sample order
            block = pinned_in[lo : lo + chunk].transpose(1, 2)
            pcm_host[lo : lo + chunk] = torch.clamp(torch.round(block * 32767.0), -32768, 32767).to(torch.int16)

        def step_pcm():
            handle.check(handle.lib.repet_original_batch_pcm16(
                handle.h, ctypes.c_void_p(pcm_host.data_ptr()), B, CHANNELS, CLIP_SAMPLES, ctypes.byref(params),
                ctypes.c_void_p(pinned_out.data_ptr()), periods_host.ctypes.data_as(ctypes.c_void_p)))

        step_pcm()
        barrier()
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_pcm()
        torch.cuda.synchronize(device)
        pcm_ms = 1e3 * (time.perf_counter() - t0)
        barrier()

    # ---- reduce over ranks: max time -------------------------------------------------------------
    times = torch.tensor([elapsed_ms, e2e_ms if e2e_ms is not None else 0.0, pcm_ms if pcm_ms is not None else 0.0],
                         dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    elapsed_ms, e2e_max_ms, pcm_max_ms = float(times[0]), float(times[1]), float(times[2])

    if rank == 0:
        audio_seconds_per_step = world * B * CLIP_SECONDS
        value = audio_seconds_per_step * args.steps / (elapsed_ms / 1e3)
        T = int(np.ceil(CLIP_SAMPLES / 1024)) + 1
        audio_bytes = CLIP_SAMPLES * CHANNELS * 4
        x_bytes = T * CHANNELS * 1025 * 8
        p_bytes = T * 1025 * 4
        algorithmic = {  # bytes per clip, DESIGN.md section "algorithmic bytes"
            "k_stft": audio_bytes + x_bytes + p_bytes,
            "k_beat": p_bytes,
            "k_model": x_bytes,
            "k_mask_istft": x_bytes + audio_bytes,
        }
        peak, peak_source = measured_hbm_peak()
        kernels = {}
        for name, (ms, count) in profile.items():
            entry = {"ms_total": ms, "launches": count, "share_of_step": ms / elapsed_ms if elapsed_ms else None}
            if name in algorithmic and ms > 0:
                gbs = algorithmic[name] * B * args.steps / (ms / 1e3) / 1e9
                entry.update({"achieved_gbs": gbs, "frac": gbs / peak,
                              "algorithmic_bytes_per_launch": algorithmic[name] * B * args.steps / max(1, count)})
            kernels[name] = entry
        dominant = max((k for k in kernels if "achieved_gbs" in kernels[k]), key=lambda k: kernels[k]["ms_total"], default=None)
        roofline = None
        if dominant:
            d = kernels[dominant]
            roofline = {"bound": "hbm", "kernel": dominant, "achieved": d["achieved_gbs"], "peak": peak, "unit": "GB/s",
                        "frac": d["frac"],
                        "traffic": ncu_traffic(dominant, B * args.steps / max(1, d["launches"])), "peak_source": peak_source,
                        "algorithmic_bytes_per_launch": d["algorithmic_bytes_per_launch"],
                        "avg_launch_ms": d["ms_total"] / max(1, d["launches"]), "kernels": kernels,
                        "whole_path_algorithmic_gbs": (2 * audio_bytes + 2 * x_bytes + 2 * p_bytes) * B * args.steps / (elapsed_ms / 1e3) / 1e9}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config_dict(args), "clocks": clocks,
            "gpu_launches": int(launches), "roofline": roofline,
            "periods_sample": periods_first[:8].tolist(), "synthesis_seconds": t_gen,
            "host_enqueue_ms_per_step": 1e3 * t_host / args.steps, "numa": numa,
        }
        if e2e_ms is not None:
            line["e2e"] = {"value": audio_seconds_per_step * args.steps / (e2e_max_ms / 1e3), "unit": UNIT,
                           "h2d_bytes_per_step": B * audio_bytes, "d2h_bytes_per_step": B * audio_bytes + B * 4,
                           "ms_per_step": e2e_max_ms / args.steps,
                           "api": "repet_original_batch (C ABI, pinned host fp32 planar buffers in and out)"}
            line["e2e_pcm16"] = {"value": audio_seconds_per_step * args.steps / (pcm_max_ms / 1e3), "unit": UNIT,
                                 "h2d_bytes_per_step": B * audio_bytes // 2, "d2h_bytes_per_step": B * audio_bytes + B * 4,
                                 "ms_per_step": pcm_max_ms / args.steps,
                                 "api": "repet_original_batch_pcm16 (int16 PCM in WAV order in, fp32 planar out)"}
        if world == 1 and not args.no_cpu_baseline:
            # fork the CPU pool only now; workers never touch CUDA
            arm = CpuArm()
            sample = args.cpu_sample_clips or 2 * arm.workers
            compute, wall, audio = arm.run(0, sample)
            arm.close()
            line["cpu_baseline"] = {"value": audio / compute, "unit": UNIT, "cores": arm.workers, "kind": "port",
                                    "sample": "%d of the %d clips (30 s each), oracle port of repet.original, one clip per core"
                                              % (sample, B), "wall_seconds": wall}
        print(json.dumps(line), flush=True)
    handle.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1 and args.gpus > 1 and args.impl == "b200":
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29517"), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
    else:
        run_b200_arm(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
