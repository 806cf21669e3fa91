#!/usr/bin/env python
"""bench.py -- VQT frames/sec on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--workload chords60|hires60|streams4096] [--streams S]

A "step" is one pass of the hot path over one batch of synthetic audio.  At N=1 the workload is
BASELINE.json configs[1]: 60 s of synthetic polyphonic audio (random chords), default VqtParameters,
hop 368 -> 3507 frames.  With N ranks (torchrun, one rank per GPU) every rank transforms its own
60 s recording (weak scaling, independent streams, no collective on the data path).
`--workload streams4096` is BASELINE.json configs[2]: 4096 independent 10 s streams (511 frames each,
2,093,056 frames in all), contiguous blocks of 4096/N streams per rank (strong scaling, total work fixed).

  value        frames/s, device-timed (CUDA events per step), audio already resident in HBM
  e2e          frames/s through the host-buffer C-ABI entry (pvqt_calc_batch_db): H2D + kernels + D2H
  roofline     K-fft (the dominant kernel): algorithmic bytes / measured launch duration vs measured HBM peak
  cpu_baseline the CPU oracle (a C port of the reference algorithm, f32, OpenMP) on the host cores

`--impl reference` times that CPU port alone (the Rust reference cannot be built in this image).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "vqt_frames_per_sec"
UNIT = "frames/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def _stream(seed: int):
    from pitchvis_b200 import synth
    return synth.polyphonic_chords(10.0, 22050.0, seed=seed)


def workload(name: str, seed: int, rank: int = 0, world: int = 1, n_streams_total: int = 4096):
    """-> params, audio ([samples] or [streams][samples]), hop, n_streams (this rank), frames_per_stream"""
    from pitchvis_b200 import synth
    import pitchvis_b200 as pv
    if name == "chords60":
        params = pv.VqtParameters.default()
        audio = synth.polyphonic_chords(60.0, params.sr, seed=seed)
        hop = synth.HOP_DEFAULT
    elif name == "hires60":
        params = pv.VqtParameters.hires()
        audio = synth.polyphonic_chords(60.0, params.sr, seed=seed)
        hop = synth.HOP_HIRES
    elif name == "streams4096":
        # stream s = the config-2 generator with seed s, 10 s each; rank r owns the
        # contiguous block [r * S / world, (r + 1) * S / world)
        params = pv.VqtParameters.default()
        hop = synth.HOP_DEFAULT
        s0, s1 = rank * n_streams_total // world, (rank + 1) * n_streams_total // world
        import multiprocessing as mp
        procs = max(1, min(host_threads(), 64))
        t0 = time.perf_counter()
        with mp.get_context("fork").Pool(procs) as pool:
            rows = pool.map(_stream, range(s0, s1), chunksize=4)
        audio = np.stack(rows)
        log(f"generated {s1 - s0} streams in {time.perf_counter() - t0:.1f} s on {procs} processes")
    else:
        raise SystemExit(f"unknown workload {name}")
    if audio.ndim == 1:
        return params, audio, hop, 1, synth.frames_in(audio.shape[0], params.n_fft, hop)
    return params, audio, hop, audio.shape[0], synth.frames_in(audio.shape[1], params.n_fft, hop)


def workload_name(name: str, hop: int, n_frames: int, n_streams_total: int = 4096) -> str:
    if name == "streams4096":
        return (f"streams4096: {n_streams_total} independent 10 s synthetic streams (random chords, stream s = seed s), "
                f"hop {hop}, {n_frames // max(1, n_streams_total)} frames/stream, {n_frames} frames/step over all "
                f"GPUs, contiguous stream blocks per rank (BASELINE.json configs[2])")
    which = "BASELINE.json configs[1]" if name == "chords60" else "BASELINE.json configs[3]"
    return (f"{name}: 60 s synthetic polyphonic audio per GPU (random chords, seed = rank), hop {hop}, "
            f"{n_frames} frames/step/GPU ({which})")


def oracle_params(name: str):
    import orc
    return orc.hires_params() if name == "hires60" else orc.default_params()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower() == "active":
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def bind_near_gpu(local_rank: int):
    """Best effort: restrict this process to the CPUs NVML names as nearest to its GPU, so that the pinned host
    buffers of the end-to-end arm are allocated on the GPU's NUMA node (with several ranks per host the default
    placement sends half of the PCIe traffic across the socket interconnect).  Returns (previous affinity, cpus bound)."""
    prev = os.sched_getaffinity(0)
    try:
        import pynvml
        pynvml.nvmlInit()
        idx = local_rank
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        if vis and all(t.strip().isdigit() for t in vis.split(",")):
            idx = int(vis.split(",")[local_rank])
        words = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(idx), (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1} & prev
        if cpus and cpus != prev:
            os.sched_setaffinity(0, cpus)
            return prev, len(cpus)
    except Exception:
        pass
    return prev, 0


def host_threads() -> int:
    """Host cores this process may use (torchrun exports OMP_NUM_THREADS=1, so do not ask OpenMP)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def run_reference(args, rank: int, world: int):
    """CPU arm: the oracle's reference-faithful f32 path, all host threads, bounded sample per step."""
    if rank != 0:
        return
    import orc
    threads = host_threads()
    if args.workload == "streams4096":
        # bounded sample of config 3: the first 2 x threads streams of the 4096 (same frames per stream, same
        # parameters); one host thread per stream, as pitchvis_train/src/train.rs:146-154 runs one Vqt per worker
        sample = max(16, 2 * threads)
        params, audio, hop, n_streams, fps_ = workload(args.workload, seed=0, rank=0, world=max(1, args.streams // sample),
                                                       n_streams_total=args.streams)
    else:
        params, audio, hop, n_streams, fps_ = workload(args.workload, seed=0)
    n_frames = n_streams * fps_
    rows = audio.reshape(n_streams, -1)
    if n_streams == 1:
        v = orc.OracleVqt(oracle_params(args.workload))

        def one_pass():
            v.calculate_batch_db(rows[0], hop, fps_, mode=1, n_threads=threads)
    else:
        from concurrent.futures import ThreadPoolExecutor
        workers = [orc.OracleVqt(oracle_params(args.workload)) for _ in range(threads)]
        pool = ThreadPoolExecutor(threads)

        def work(w):   # the C call releases the GIL
            for i in range(w, n_streams, threads):
                workers[w].calculate_batch_db(rows[i], hop, fps_, mode=1, n_threads=1)

        def one_pass():
            list(pool.map(work, range(threads)))

    for _ in range(args.warmup):
        one_pass()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_pass()
    dt = time.perf_counter() - t0
    fps = args.steps * n_frames / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "strong" if args.workload == "streams4096" else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.workload, hop, n_frames if args.workload != "streams4096"
                                             else args.streams * fps_, args.streams),
                   "sample": ("every step transforms all frames of the workload on the host cores (oracle f32 path)"
                              if args.workload != "streams4096" else
                              f"every step transforms the first {n_streams} streams ({n_frames} frames) of the workload")},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": (f"{args.steps} passes over all {n_frames} frames of the workload, " if n_streams == 1
                                    else f"{args.steps} passes over the first {n_streams} streams ({n_frames} frames), "
                                         "one host thread per stream, ")
                                   + "oracle f32 path (C port of vqt.rs:866-954; the Rust crate cannot be built here)"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(args, audio, hop, n_frames):
    import orc
    v = orc.OracleVqt(oracle_params(args.workload))
    threads = host_threads()
    if audio.ndim == 2:   # streams: the sample is the first stream (same parameters, same frames per stream)
        audio = audio[0]
        n_frames = n_frames // max(1, args.local_streams)
    v.calculate_batch_db(audio, hop, min(n_frames, 256), mode=1, n_threads=threads)  # warm caches / plans
    t0 = time.perf_counter()
    v.calculate_batch_db(audio, hop, n_frames, mode=1, n_threads=threads)
    one = time.perf_counter() - t0
    reps = max(1, min(200, int(12.0 * threads / max(one * threads, 1e-3))))  # ~12 s of CPU work in total
    t0 = time.perf_counter()
    for _ in range(reps):
        v.calculate_batch_db(audio, hop, n_frames, mode=1, n_threads=threads)
    dt = time.perf_counter() - t0
    t1 = time.perf_counter()
    n1 = min(n_frames, 1024)
    v.calculate_batch_db(audio, hop, n1, mode=1, n_threads=1)
    dt1 = time.perf_counter() - t1
    return {"value": reps * n_frames / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{reps} passes over all {n_frames} frames (oracle f32 path, OpenMP, {threads} threads)",
            "single_thread_value": n1 / dt1, "single_thread_ms_per_frame": 1e3 * dt1 / n1,
            "reference_published_ms_per_frame": 0.091}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="chords60", choices=["chords60", "hires60", "streams4096"])
    ap.add_argument("--streams", type=int, default=4096, help="streams4096 only: total streams over all ranks")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--flush", default="write+read", choices=["write", "write+read"],
                    help="L2 flush between timed steps: write a 512 MiB buffer, or write it and read it back "
                         "(no dirty lines left for the first timed kernel to write back)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # Only the JSON line goes to stdout: libraries that print there (NCCL's version banner under NCCL_DEBUG=VERSION)
    # are sent to stderr by pointing fd 1 at fd 2 for the rest of the run; `emit` writes to the real stdout.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    def emit(obj):
        real_stdout.write(json.dumps(obj) + "\n")
        real_stdout.flush()

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod

    import pitchvis_b200 as pv
    from pitchvis_b200 import _ffi
    lib = _ffi.load()

    params, audio, hop, n_streams, frames_per_stream = workload(args.workload, seed=rank, rank=rank, world=world,
                                                                n_streams_total=args.streams)
    n_frames = n_streams * frames_per_stream     # frames per step on this rank
    args.local_streams = n_streams
    stream_stride = audio.shape[1] if audio.ndim == 2 else 0
    strong = args.workload == "streams4096"
    if strong and args.steps == 50:
        args.steps = 10                            # one step is ~50 ms of kernels here
    vqt = pv.Vqt(params, device=local_rank)
    nb = vqt.n_buckets
    h = vqt.handle

    def chk(rc):
        if rc != 0:
            raise RuntimeError(_ffi.last_error())

    # ---- device-resident arm ----------------------------------------------------------------
    d_audio = pv.DeviceBuffer(vqt, audio.nbytes)
    d_out = pv.DeviceBuffer(vqt, n_frames * nb * 4)
    d_audio.upload(audio)
    flush_bytes = 512 << 20   # > 126 MB L2: evicts audio, spectra scratch and output between steps
    d_flush = pv.DeviceBuffer(vqt, flush_bytes)
    ev = [C.c_void_p() for _ in range(2 * args.steps)]
    for e in ev:
        chk(lib.pvqt_event_create(h, C.byref(e)))

    def step():
        pv.calc_db_device(vqt, d_audio, n_streams, stream_stride, hop, frames_per_stream, d_out)

    def flush():
        if args.flush == "write":
            chk(lib.pvqt_dev_memset(h, d_flush.ptr, 0, flush_bytes))
        else:
            chk(lib.pvqt_dev_flush_l2(h, d_flush.ptr, flush_bytes))

    for _ in range(args.warmup):
        flush(); step()
    pv.synchronize(vqt)
    launches0 = vqt.launch_count

    def barrier():
        pv.synchronize(vqt)
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        flush()
        chk(lib.pvqt_event_record(h, ev[2 * i]))
        step()
        chk(lib.pvqt_event_record(h, ev[2 * i + 1]))
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    launches = vqt.launch_count - launches0
    step_ms = []
    for i in range(args.steps):
        ms = C.c_float()
        chk(lib.pvqt_event_elapsed_ms(h, ev[2 * i], ev[2 * i + 1], C.byref(ms)))
        step_ms.append(ms.value)
    dev_ms_total = float(sum(step_ms))

    # ---- per-kernel durations (same steps, event pairs around each launch) ---------------------
    chk(lib.pvqt_set_profiling(h, 1))
    for _ in range(args.steps):
        flush(); step()
    k_ms = (C.c_double * _ffi.PROFILE_KINDS)()
    k_n = (C.c_uint64 * _ffi.PROFILE_KINDS)()
    chk(lib.pvqt_get_profile(h, 1, k_ms, k_n))
    chk(lib.pvqt_set_profiling(h, 0))

    # ---- end-to-end arm: host buffers through the C ABI ---------------------------------------
    prev_affinity, near_cpus = bind_near_gpu(local_rank)
    pin_in, pin_out = C.c_void_p(), C.c_void_p()
    chk(lib.pvqt_host_alloc_pinned(audio.nbytes, C.byref(pin_in)))
    chk(lib.pvqt_host_alloc_pinned(n_frames * nb * 4, C.byref(pin_out)))
    C.memmove(pin_in, audio.ctypes.data, audio.nbytes)
    fp = C.POINTER(C.c_float)
    e2e_steps = max(3, min(args.steps, 20 if not strong else 3))

    def e2e_call():
        if audio.ndim == 1:
            chk(lib.pvqt_calc_batch_db(h, C.cast(pin_in, fp), audio.shape[0], hop, n_frames, C.cast(pin_out, fp)))
        else:
            chk(lib.pvqt_calc_streams_db(h, C.cast(pin_in, fp), n_streams, stream_stride, audio.shape[1], hop,
                                         frames_per_stream, C.cast(pin_out, fp)))

    for _ in range(2 if not strong else 1):
        e2e_call()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        # no L2 flush here: every step's input arrives from pinned host memory through H2D copies
        e2e_call()
    pv.synchronize(vqt)
    e2e_s = time.perf_counter() - t0
    result_checksum = float(np.ctypeslib.as_array(C.cast(pin_out, fp), shape=(n_frames * nb,)).sum())
    os.sched_setaffinity(0, prev_affinity)   # the CPU baseline below uses every core

    # ---- reduce over ranks: max time, total frames -------------------------------------------
    dev_ms_max, e2e_s_max = dev_ms_total, e2e_s
    if dist is not None:
        import torch
        t = torch.tensor([dev_ms_total, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms_max, e2e_s_max = float(t[0]), float(t[1])

    total_frames = world * n_frames
    if strong and dist is not None:
        import torch
        t = torch.tensor([n_frames], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        total_frames = int(t[0])
    value = total_frames * args.steps / (dev_ms_max * 1e-3)
    e2e_value = total_frames * e2e_steps / e2e_s_max

    if rank == 0:
        first = int(lib.pvqt_first_sample_used(h))
        union = params.n_fft - first
        bytes_per_frame = 4 * union + 4 * nb                      # 35,120 B at the defaults
        # roofline of the dominant kernel (largest share of the step), live CUDA-event durations
        kinds = [i for i in range(_ffi.PROFILE_KINDS) if k_n[i] > 0]
        top = max(kinds, key=lambda i: k_ms[i])
        top_avg_ms = k_ms[top] / k_n[top]
        frames_per_launch = n_frames * args.steps / k_n[top]
        peak, peak_src = measured_peak_gbs()
        achieved = bytes_per_frame * frames_per_launch / (top_avg_ms * 1e-3) / 1e9
        whole = bytes_per_frame * n_frames / (statistics.median(step_ms) * 1e-3) / 1e9
        traffic = None
        try:  # dram bytes of that kernel from the committed ncu --set full capture, per launch
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
                traffic = json.load(fh).get(args.workload, {}).get(_ffi.KERNEL_KIND_NAMES[top])
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": workload_name(args.workload, hop, total_frames if strong else n_frames, args.streams),
                "n_fft": params.n_fft, "n_buckets": nb, "hop": hop, "frames_per_step_per_gpu": n_frames,
                "l2": f"flushed between timed steps ({flush_bytes >> 20} MiB memset"
                      + (" followed by a read sweep of the same buffer, so the flush leaves no dirty lines)"
                         if args.flush != "write" else ")"),
                "timing": "CUDA events per step on the launching stream, summed; max over ranks",
            },
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(audio.nbytes),
                    "d2h_bytes_per_step": int(n_frames * nb * 4), "steps": e2e_steps,
                    "api": ("pvqt_calc_batch_db" if audio.ndim == 1 else "pvqt_calc_streams_db")
                           + " (pinned host buffers in and out)", "checksum": result_checksum,
                    "host_cpus_near_gpu": near_cpus},
            "gpu_launches": int(launches),
            "plan": vqt.plan_info(),
            "clocks": clocks,
            "roofline": {
                "bound": "hbm", "kernel": _ffi.KERNEL_KIND_NAMES[top], "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_frame": bytes_per_frame, "frames_per_launch": frames_per_launch,
                "kernel_avg_ms": top_avg_ms, "kernel_share_of_step": k_ms[top] / max(1e-9, sum(k_ms)),
                "kernels_avg_ms": {_ffi.KERNEL_KIND_NAMES[i]: k_ms[i] / k_n[i] for i in kinds},
                "whole_step_achieved": whole, "whole_step_frac": whole / peak,
                "note": "the path is FP32/shared-memory bound (SURVEY.md 8d); HBM fraction reported as BASELINE asks",
            },
            "step_ms": {"median": statistics.median(step_ms), "min": min(step_ms), "max": max(step_ms)},
            "wall_s_timed_region": t_wall,
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(args, audio, hop, n_frames)
        elif not args.no_cpu_baseline:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port",
                                    "sample": "measured at N=1 only"}
        emit(line)

    for e in ev:
        lib.pvqt_event_destroy(h, e)
    lib.pvqt_host_free_pinned(pin_in)
    lib.pvqt_host_free_pinned(pin_out)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
