#!/usr/bin/env python
"""bench.py -- VQT frames/sec on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--workload chords60|hires60|streams4096] [--configs all|none|a,b,...] [--streams S]

A "step" is one pass of the hot path over one batch of synthetic audio.  The headline workload (`value`, `e2e`,
`roofline`) is BASELINE.json configs[1]: 60 s of synthetic polyphonic audio (random chords), default VqtParameters,
hop 368 -> 3507 frames; with N ranks (torchrun, one rank per GPU) every rank transforms its own recording (weak
scaling, independent recordings, no collective on the data path).  The same line carries the other configurations
BASELINE.json names as sub-records under "configs" (each measured the same way: device-timed with CUDA events, L2
flushed between steps, max over ranks):

  configs.streams4096  configs[2]: 4096 independent 10 s streams (2,093,056 frames per step), contiguous blocks of
                       4096/N streams per rank -- STRONG scaling under --gpus N; value + e2e
  configs.hires60      configs[3]: hi-res parameters (1344 bins, 5 FFTs up to 16384, hop 735), 60 s per GPU
  configs.instant      configs[0]: pvqt_calc_instant_db, one frame per call, p50 / p99 latency, the CPU port beside it
  configs.pipeline5    configs[4]: VQT + AnalysisState in one call (pvqt_calc_*_analysis), results only over PCIe

  value        frames/s, device-timed (CUDA events per step), audio already resident in HBM
  e2e          frames/s through the host-buffer C-ABI entry: H2D + kernels + D2H, pinned host buffers; at N > 1 rank 0
               drives all N devices from one process through pvqt_multi_calc_streams_db (the north-star design: results
               gathered into one pinned host buffer), next to the box's measured PCIe ceiling for the same bytes
  roofline     frac = WHOLE-STEP algorithmic bytes (4 U + 4 n_buckets per frame) / step time / measured
               HBM peak; per kernel, that kernel's own compulsory bytes; fp32_frac against SMs x 128 x 2 x clock
  cpu_baseline the CPU oracle (a C port of the reference algorithm, f32, OpenMP) on the host cores

`--impl reference` times that CPU port alone (the Rust reference cannot be built in this image).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
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
FRAME_NS = 16_689_342   # 368 / 22050 s: the viewer's 60 FPS frame time
FP = C.POINTER(C.c_float)


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
        procs = max(1, min(host_threads() // max(1, world), 64))
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
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
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
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower() == "active":
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(pw) if pw else None}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def host_threads() -> int:
    """Host cores this process may use (torchrun exports OMP_NUM_THREADS=1, so do not ask OpenMP)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def numa_report(lib, device: int) -> dict:
    """Where the GPU hangs and where this process may run: the end-to-end arm is PCIe / host-memory bound, so the
    placement of the pinned buffers is part of the number."""
    node = C.c_int32(-1)
    lib.pvqt_device_attributes(device, None, None, None, C.byref(node))
    rep = {"gpu_numa_node": int(node.value), "host_nodes": None, "bound_cpus": 0}
    try:
        nodes = sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
        rep["host_nodes"] = len(nodes)
        if node.value >= 0 and len(nodes) > 1:
            with open(f"/sys/devices/system/node/node{node.value}/cpulist") as fh:
                cpus = set()
                for part in fh.read().strip().split(","):
                    a, _, b = part.partition("-")
                    cpus.update(range(int(a), int(b or a) + 1))
            cpus &= os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)   # pinned buffers allocated (first-touched) from here land on the GPU's node
                rep["bound_cpus"] = len(cpus)
    except Exception as e:  # containers often hide /sys/devices/system/node
        rep["note"] = f"{type(e).__name__}: {e}"
    return rep


# ------------------------------------------------------------------------------------------------------------
# reference arm: the CPU port on the host cores
# ------------------------------------------------------------------------------------------------------------
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


def cpu_baseline(workload_key, audio, hop, n_frames, local_streams):
    import orc
    v = orc.OracleVqt(oracle_params(workload_key))
    threads = host_threads()
    if audio.ndim == 2:   # streams: the sample is the first stream (same parameters, same frames per stream)
        audio = audio[0]
        n_frames = n_frames // max(1, local_streams)
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


# ------------------------------------------------------------------------------------------------------------
# the B200 arm
# ------------------------------------------------------------------------------------------------------------
class Ctx:
    """Everything a measurement needs: library, rank layout, torch.distributed (barrier / max over ranks only)."""

    def __init__(self, args, rank, world, local_rank, dist):
        from pitchvis_b200 import _ffi
        self.args, self.rank, self.world, self.local_rank, self.dist = args, rank, world, local_rank, dist
        self.cpu_group = dist.new_group(backend="gloo") if dist is not None else None
        self.lib = _ffi.load()
        self.ffi = _ffi
        self.peak, self.peak_src = measured_peak_gbs()
        sm, khz = C.c_int32(), C.c_int32()
        self.lib.pvqt_device_attributes(local_rank, C.byref(sm), C.byref(khz), None, None)
        self.sm_count, self.sm_clock_khz = int(sm.value), int(khz.value)

    def chk(self, rc):
        if rc != 0:
            raise RuntimeError(self.ffi.last_error())

    def barrier(self, vqt=None):
        if vqt is not None:
            import pitchvis_b200 as pv
            pv.synchronize(vqt)
        if self.dist is not None:
            import torch
            # a host-side (gloo) barrier: an NCCL barrier is a kernel that spins on the waiting ranks' GPUs, and in the
            # end-to-end arm rank 0 drives those very GPUs through pvqt_multi_* while the other ranks wait
            self.dist.barrier(group=self.cpu_group)
            torch.cuda.synchronize()

    def reduce_max(self, *xs):
        if self.dist is None:
            return [float(x) for x in xs]
        import torch
        t = torch.tensor(list(xs), dtype=torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def reduce_sum(self, x):
        if self.dist is None:
            return int(x)
        import torch
        t = torch.tensor([int(x)], dtype=torch.int64, device="cuda")
        self.dist.all_reduce(t)
        return int(t[0])


def plan_numbers(vqt, params, hop):
    """Algorithmic bytes / flops per frame, for the whole step and per kernel."""
    k = vqt.kernel()
    groups = k.window_groups
    nb = vqt.n_buckets
    mask = vqt.plan_info()["sdft_group_mask"]
    cols = [vqt.group_columns(g)[1] for g in range(len(groups))]
    first = min(g.window[0] for g in groups)
    union = params.n_fft - first
    fft_groups = [i for i in range(len(groups)) if not (mask >> i) & 1]
    sdft_groups = [i for i in range(len(groups)) if (mask >> i) & 1]
    u_fft = (max(groups[i].window[1] for i in fft_groups) - min(groups[i].window[0] for i in fft_groups)) if fft_groups else 0
    cols_fft = sum(cols[i] for i in fft_groups)
    nk_sdft = sum(cols[i] for i in sdft_groups)
    flops = 10 * nb
    for g in groups:
        n = g.window_size()
        flops += 2.5 * n * math.log2(n) + 8 * (g.filter_bank.nnz() + (g.negative_filter_bank.nnz() if g.negative_filter_bank else 0))
    return {
        "whole": 4 * union + 4 * nb,                          # 35,120 B at the defaults
        "streaming": 4 * hop + 4 * nb,
        "flops": flops,                                       # nominal: 621,608 at the defaults
        # per kernel: the bytes that kernel cannot avoid moving per frame
        "fft_groups_kernel": 4 * u_fft + 8 * cols_fft,        # its windows' union in, its consumed bins out
        "sdft_partial_kernel": (4 * hop + 16 * nk_sdft) if sdft_groups else 0,   # one hop of samples in, one C and one R row out
        "spmm_db_fused_kernel": 8 * cols_fft + 16 * nk_sdft + 4 * nb,            # spectra + one new chunk row pair in, dB out
        "sdft_groups": sdft_groups,
    }


def measure(ctx: Ctx, name: str, steps: int, warmup: int, *, e2e_steps: int, profile: bool, sustain_s: float = 0.0,
            want_cpu: bool = False):
    """Device-timed frames/s of one workload on this rank's GPU (max over ranks), per-kernel durations, the end-to-end
    arm through host buffers, and optionally a sustained run of the same step for the clock record."""
    import pitchvis_b200 as pv
    args, lib, chk = ctx.args, ctx.lib, ctx.chk
    strong = name == "streams4096"
    params, audio, hop, n_streams, fps_ = workload(name, seed=ctx.rank, rank=ctx.rank, world=ctx.world,
                                                   n_streams_total=args.streams)
    n_frames = n_streams * fps_
    stream_stride = audio.shape[1] if audio.ndim == 2 else 0
    vqt = pv.Vqt(params, device=ctx.local_rank)
    nb, h = vqt.n_buckets, vqt.handle
    d_audio = pv.DeviceBuffer(vqt, audio.nbytes)
    d_out = pv.DeviceBuffer(vqt, n_frames * nb * 4)
    d_audio.upload(audio)
    flush_bytes = 512 << 20   # > 126 MB L2: evicts audio, spectra scratch and output between steps
    d_flush = pv.DeviceBuffer(vqt, flush_bytes)

    def step():
        pv.calc_db_device(vqt, d_audio, n_streams, stream_stride, hop, fps_, d_out)

    def flush():
        if args.flush == "write":
            chk(lib.pvqt_dev_memset(h, d_flush.ptr, 0, flush_bytes))
        else:
            chk(lib.pvqt_dev_flush_l2(h, d_flush.ptr, flush_bytes))

    def timed(n):
        ev = [C.c_void_p() for _ in range(2 * n)]
        for e in ev:
            chk(lib.pvqt_event_create(h, C.byref(e)))
        for i in range(n):
            flush()
            chk(lib.pvqt_event_record(h, ev[2 * i]))
            step()
            chk(lib.pvqt_event_record(h, ev[2 * i + 1]))
        pv.synchronize(vqt)
        out = []
        for i in range(n):
            ms = C.c_float()
            chk(lib.pvqt_event_elapsed_ms(h, ev[2 * i], ev[2 * i + 1], C.byref(ms)))
            out.append(ms.value)
        for e in ev:
            lib.pvqt_event_destroy(h, e)
        return out

    for _ in range(warmup):
        flush(); step()
    pv.synchronize(vqt)
    launches0 = vqt.launch_count
    sampler = ClockSampler(ctx.local_rank)
    ctx.barrier(vqt)
    sampler.start()
    t_wall0 = time.perf_counter()
    step_ms = timed(steps)
    ctx.barrier(vqt)
    t_wall = time.perf_counter() - t_wall0
    launches = vqt.launch_count - launches0
    # the same step, back to back, long enough for nvidia-smi to see it (clocks / throttle reasons under load)
    sustained = None
    if sustain_s > 0:
        t0 = time.perf_counter()
        ms_all = []
        while time.perf_counter() - t0 < sustain_s:
            ms_all += timed(50)
        sustained = {"value": n_frames * len(ms_all) / (sum(ms_all) * 1e-3), "steps": len(ms_all),
                     "seconds": time.perf_counter() - t0, "step_ms_median": statistics.median(ms_all)}
    clocks = sampler.stop()

    rec = {"n_frames_rank": n_frames, "hop": hop, "n_buckets": nb, "n_fft": params.n_fft}
    k_avg = {}
    if profile:
        chk(lib.pvqt_set_profiling(h, 1))
        for _ in range(min(steps, 20)):
            flush(); step()
        k_ms = (C.c_double * ctx.ffi.PROFILE_KINDS)()
        k_n = (C.c_uint64 * ctx.ffi.PROFILE_KINDS)()
        chk(lib.pvqt_get_profile(h, 1, k_ms, k_n))
        chk(lib.pvqt_set_profiling(h, 0))
        k_avg = {ctx.ffi.KERNEL_KIND_NAMES[i]: (k_ms[i] / k_n[i], int(k_n[i]) // min(steps, 20))
                 for i in range(ctx.ffi.PROFILE_KINDS) if k_n[i] > 0}
    nums = plan_numbers(vqt, params, hop)

    # ---- end-to-end arm: host buffers through the C ABI ---------------------------------------------------------
    e2e = None
    if e2e_steps > 0 and ctx.world == 1:
        pin_in, pin_out = C.c_void_p(), C.c_void_p()
        chk(lib.pvqt_host_alloc_pinned(audio.nbytes, C.byref(pin_in)))
        chk(lib.pvqt_host_alloc_pinned(n_frames * nb * 4, C.byref(pin_out)))
        C.memmove(pin_in, audio.ctypes.data, audio.nbytes)

        def e2e_call():
            if audio.ndim == 1:
                chk(lib.pvqt_calc_batch_db(h, C.cast(pin_in, FP), audio.shape[0], hop, n_frames, C.cast(pin_out, FP)))
            else:
                chk(lib.pvqt_calc_streams_db(h, C.cast(pin_in, FP), n_streams, stream_stride, audio.shape[1], hop, fps_,
                                             C.cast(pin_out, FP)))

        for _ in range(2 if not strong else 1):
            e2e_call()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_call()    # no L2 flush here: every step's input arrives from pinned host memory through H2D copies
        e2e_s = time.perf_counter() - t0
        checksum = float(np.ctypeslib.as_array(C.cast(pin_out, FP), shape=(n_frames * nb,)).sum(dtype=np.float64))
        lib.pvqt_host_free_pinned(pin_in)
        lib.pvqt_host_free_pinned(pin_out)
        e2e = {"value": n_frames * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(audio.nbytes),
               "d2h_bytes_per_step": int(n_frames * nb * 4), "steps": e2e_steps,
               "api": ("pvqt_calc_batch_db" if audio.ndim == 1 else "pvqt_calc_streams_db") + " (pinned host buffers in and out)",
               "checksum": checksum}

    cpu = cpu_baseline(name, audio, hop, n_frames, n_streams) if want_cpu else None
    plan = vqt.plan_info()
    for b in (d_audio, d_out, d_flush):
        b.free()
    vqt.close()

    dev_ms_total = float(sum(step_ms))
    (dev_ms_max,) = ctx.reduce_max(dev_ms_total)
    total_frames = ctx.reduce_sum(n_frames) if strong else ctx.world * n_frames
    rec.update({
        "value": total_frames * steps / (dev_ms_max * 1e-3), "ms_per_step": dev_ms_max / steps, "steps": steps,
        "total_frames_per_step": total_frames, "launches": int(launches), "clocks": clocks, "sustained": sustained,
        "step_ms": {"median": statistics.median(step_ms), "min": min(step_ms), "max": max(step_ms)},
        "wall_s_timed_region": t_wall, "kernels": k_avg, "numbers": nums, "e2e": e2e, "cpu": cpu, "plan": plan,
        "audio": audio if (ctx.world > 1 or strong) else None, "stream_stride": stream_stride, "n_streams": n_streams, "fps": fps_,
        "params": params,
    })
    return rec


def roofline_record(ctx: Ctx, rec, workload_key: str):
    """roofline object of one measurement: frac = whole-step fraction; per kernel its own compulsory bytes."""
    nums, n_frames = rec["numbers"], rec["n_frames_rank"]
    step_s = rec["step_ms"]["median"] * 1e-3
    whole = nums["whole"] * n_frames / step_s / 1e9
    fp32_peak = ctx.sm_count * 128 * 2 * ctx.sm_clock_khz * 1e3            # FLOP/s
    kernels = {}
    top = None
    total = sum(ms * n for ms, n in rec["kernels"].values()) or 1.0
    for name, (avg_ms, per_step) in rec["kernels"].items():
        b = nums.get(name, 0)
        frames_per_launch = n_frames / max(1, per_step)
        ach = b * frames_per_launch / (avg_ms * 1e-3) / 1e9
        kernels[name] = {"avg_ms": avg_ms, "launches_per_step": per_step, "bytes_per_frame": b, "achieved": ach,
                         "frac": ach / ctx.peak, "share_of_kernel_time": avg_ms * per_step / total}
        if top is None or avg_ms * per_step > rec["kernels"][top][0] * rec["kernels"][top][1]:
            top = name
    traffic = None
    try:  # dram bytes of the dominant kernel from the committed ncu --set full capture, per launch
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            traffic = json.load(fh).get(workload_key, {}).get(top)
    except Exception:
        pass
    return {
        "bound": "hbm", "achieved": whole, "peak": ctx.peak, "unit": "GB/s", "frac": whole / ctx.peak,
        "traffic": traffic, "traffic_note": "dram read + write bytes of the dominant kernel per launch, ncu --set full "
                                            "(cold cache, serialised replays): profiles/traffic.json",
        "peak_source": ctx.peak_src, "what": "WHOLE STEP: algorithmic bytes per frame x frames / median step time",
        "algorithmic_bytes_per_frame": nums["whole"], "streaming_bytes_per_frame": nums["streaming"],
        "dominant_kernel": top, "kernels": kernels,
        "fp32_frac": rec["value"] / max(1, ctx.world) * nums["flops"] / fp32_peak, "nominal_flop_per_frame": nums["flops"],
        "fp32_peak_tflops": fp32_peak / 1e12, "sm_count": ctx.sm_count, "sm_clock_mhz": ctx.sm_clock_khz / 1e3,
        "note": "the path is FP32 / shared-memory bound (SURVEY.md 8d); the HBM fraction is reported as BASELINE asks",
    }


def multi_e2e(ctx: Ctx, name: str, rec, steps: int):
    """N > 1: rank 0 drives all N devices from one process through pvqt_multi_calc_streams_db -- the north-star design,
    results gathered into one pinned host buffer -- while the other ranks wait; beside it the box's PCIe ceiling for the
    same bytes (all devices copying at once, no kernel)."""
    import pitchvis_b200 as pv
    lib, chk = ctx.lib, ctx.chk
    audio = rec["audio"]
    rows = audio.reshape(rec["n_streams"], -1)
    shm = f"/dev/shm/pvqt_bench_{os.environ.get('MASTER_PORT', '0')}_{name}"
    total_streams = ctx.reduce_sum(rows.shape[0])
    # every rank writes its block of streams into one shared file (rank order = stream order)
    import torch
    t = torch.zeros(ctx.world, dtype=torch.int64, device="cuda")
    t[ctx.rank] = rows.shape[0]
    ctx.dist.all_reduce(t)
    counts = [int(x) for x in t]
    offs = [sum(counts[:r]) for r in range(ctx.world)]
    if ctx.rank == 0:
        mm = np.lib.format.open_memmap(shm, mode="w+", dtype=np.float32, shape=(total_streams, rows.shape[1]))
        del mm
    ctx.barrier()
    mm = np.load(shm, mmap_mode="r+")
    mm[offs[ctx.rank]:offs[ctx.rank] + rows.shape[0]] = rows
    mm.flush()
    del mm
    ctx.barrier()
    out = None
    if ctx.rank == 0:
        prev_affinity = os.sched_getaffinity(0)
        numa = numa_report(lib, 0)
        mm = np.load(shm, mmap_mode="r")
        n_samples, fps_, hop, nb = rows.shape[1], rec["fps"], rec["hop"], rec["n_buckets"]
        in_bytes, out_bytes = total_streams * n_samples * 4, total_streams * fps_ * nb * 4
        pin_in, pin_out = C.c_void_p(), C.c_void_p()
        chk(lib.pvqt_host_alloc_pinned(in_bytes, C.byref(pin_in)))
        chk(lib.pvqt_host_alloc_pinned(out_bytes, C.byref(pin_out)))
        np.ctypeslib.as_array(C.cast(pin_in, FP), shape=(total_streams, n_samples))[:] = mm
        del mm
        m = pv.MultiVqt(rec["params"], list(range(ctx.world)))

        def call():
            chk(lib.pvqt_multi_calc_streams_db(m.handle, C.cast(pin_in, FP), total_streams, n_samples, n_samples, hop, fps_,
                                               C.cast(pin_out, FP)))
        call()
        t0 = time.perf_counter()
        for _ in range(steps):
            call()
        dt = time.perf_counter() - t0
        checksum = float(np.ctypeslib.as_array(C.cast(pin_out, FP), shape=(total_streams * fps_ * nb,)).sum(dtype=np.float64))
        probe_s = m.pcie_probe(in_bytes // ctx.world, out_bytes // ctx.world, 3)
        frames = total_streams * fps_
        out = {"value": frames * steps / dt, "unit": UNIT, "h2d_bytes_per_step": int(in_bytes), "d2h_bytes_per_step": int(out_bytes),
               "steps": steps, "api": f"pvqt_multi_calc_streams_db over {ctx.world} devices from one process (rank 0), "
                                      "one pinned host buffer in, one out", "checksum": checksum,
               "pcie_ceiling_frames_per_s": frames * 3 / probe_s,
               "pcie_ceiling_note": "all devices copying the same bytes at once (pinned, both directions in flight, no kernel): "
                                    "pvqt_multi_pcie_probe", "numa": numa}
        out["frac_of_pcie_ceiling"] = out["value"] / out["pcie_ceiling_frames_per_s"]
        m.close()
        lib.pvqt_host_free_pinned(pin_in)
        lib.pvqt_host_free_pinned(pin_out)
        os.sched_setaffinity(0, prev_affinity)
        try:
            os.unlink(shm)
        except OSError:
            pass
    ctx.barrier()
    return out


def single_e2e_ceiling(ctx: Ctx, rec):
    """N = 1: the PCIe ceiling of the same bytes on this box (one device)."""
    import pitchvis_b200 as pv
    m = pv.MultiVqt(rec["params"], [ctx.local_rank])
    try:
        s = m.pcie_probe(rec["e2e"]["h2d_bytes_per_step"], rec["e2e"]["d2h_bytes_per_step"], 5)
    finally:
        m.close()
    return rec["n_frames_rank"] * 5 / s


def bench_instant(ctx: Ctx):
    """configs[0]: pvqt_calc_instant_db, the reference's real-time call (vqt.rs:866, 60 FPS in the viewer)."""
    import orc
    import pitchvis_b200 as pv
    v = pv.Vqt(pv.VqtParameters.default(), device=ctx.local_rank)
    o = orc.OracleVqt()
    x = orc.test_create_sines(o.params, [440.0, 880.0, 1320.0, 1760.0, 2200.0])
    out = np.empty(v.n_buckets, np.float32)
    xp, op = x.ctypes.data_as(FP), out.ctypes.data_as(FP)
    for _ in range(30):
        ctx.chk(ctx.lib.pvqt_calc_instant_db(v.handle, xp, x.shape[0], op))
    ts = []
    for _ in range(1000):
        t0 = time.perf_counter()
        ctx.lib.pvqt_calc_instant_db(v.handle, xp, x.shape[0], op)
        ts.append(time.perf_counter() - t0)
    ts = np.sort(np.array(ts)) * 1e6
    ref = o.calculate_vqt_instant_in_db(x, 1)
    err = float(np.abs(out - ref).max())
    tc = []
    for _ in range(20):
        o.calculate_vqt_instant_in_db(x, 1)
    for _ in range(300):
        t0 = time.perf_counter()
        o.calculate_vqt_instant_in_db(x, 1)
        tc.append(time.perf_counter() - t0)
    tc = np.sort(np.array(tc)) * 1e6
    # steady state of the port on one core (scratch allocated once, as the reference's `Vqt` owns its scratch): 256 frames
    rep = np.tile(x, 4)[: x.shape[0] + 255 * 64]
    o.calculate_batch_db(rep, 64, 256, mode=1, n_threads=1)
    t0 = time.perf_counter()
    o.calculate_batch_db(rep, 64, 256, mode=1, n_threads=1)
    steady_us = (time.perf_counter() - t0) / 256 * 1e6
    v.close()
    return {"workload": "single-frame VQT, default VqtParameters, 440 Hz tone + 4 harmonics (BASELINE.json configs[0]); "
                        "host x[n_fft] in, host dB[588] out, one call per frame",
            "api": "pvqt_calc_instant_db (pinned staging inside the handle, only the 8192 samples the windows read are "
                   "uploaded, one captured graph: H2D, K-fft, K-spmm-db, D2H)",
            "p50_us": float(ts[500]), "p99_us": float(ts[989]), "min_us": float(ts[0]), "calls": 1000,
            "max_abs_db_vs_oracle_f32": err,
            "cpu_port": {"us_per_frame": steady_us, "cores": 1, "kind": "port",
                         "per_call_p50_us": float(tc[150]), "per_call_p99_us": float(tc[296]),
                         "note": "us_per_frame: oracle f32 path on one core with its scratch set up once (256 frames in "
                                 "one call) -- the comparable of the reference's published 0.091 ms/frame; per_call_*: "
                                 "one oracle call per frame, which sets its scratch up on every call"}}


def bench_pipeline(ctx: Ctx, streams_audio=None):
    """configs[4]: VQT + AnalysisState in one call; the spectra never visit the host."""
    import orc
    import pitchvis_b200 as pv
    from pitchvis_b200 import synth
    v = pv.Vqt(pv.VqtParameters.default(), device=ctx.local_rank)
    audio = synth.polyphonic_chords(60.0, 22050.0, seed=0)
    hop = synth.HOP_DEFAULT
    T = v.frames_in(audio.shape[0], hop)
    rec = {}
    # one 60 s stream: the epilogue is a recurrence in time, 3507 dependent steps on one CTA -- latency, not throughput
    a = pv.AnalysisState(pv.VqtRange(), device=ctx.local_rank)
    a.calculate_and_preprocess(v, audio, hop, FRAME_NS, max_peaks=32)
    ts = []
    for _ in range(3):
        a2 = pv.AnalysisState(pv.VqtRange(), device=ctx.local_rank)
        t0 = time.perf_counter()
        res = a2.calculate_and_preprocess(v, audio, hop, FRAME_NS, max_peaks=32)
        ts.append(time.perf_counter() - t0)
        a2.close()
    a.close()
    # CPU port of the same chain: oracle VQT on all cores, then the (sequential) oracle epilogue
    o = orc.OracleVqt()
    threads = host_threads()
    t0 = time.perf_counter()
    db = o.calculate_batch_db(audio, hop, T, mode=1, n_threads=threads)
    t_vqt = time.perf_counter() - t0
    oa = orc.OracleAnalysisState()
    t0 = time.perf_counter()
    n_same = 0
    for t in range(T):
        oa.preprocess(db[t], FRAME_NS)
        if t % 97 == 0:
            n = int(res["peak_count"][0, t])
            n_same += int(np.array_equal(res["peak_indices"][0, t, :n], oa.peaks))
    t_ana = time.perf_counter() - t0
    rec["single_stream"] = {
        "workload": f"chords60 (one 60 s stream, {T} frames) through pvqt_calc_batch_analysis: peak_count, peak_indices, "
                    "peaks_continuous (max 32), scene calmness, tuning inaccuracy back; spectra stay in HBM",
        "seconds": min(ts), "value": T / min(ts), "unit": UNIT, "us_per_frame": 1e6 * min(ts) / T,
        "h2d_bytes_per_step": int(audio.nbytes), "d2h_bytes_per_step": int(res["d2h_bytes"]),
        "d2h_bytes_if_spectra_were_returned": int(T * v.n_buckets * 4),
        "note": "the epilogue is strictly sequential in time per stream (EMA horizon <- scene calmness <- previous frame): "
                "one CTA walks the frames; this is a latency number",
        "sampled_frames_with_same_peaks_as_cpu_chain": f"{n_same}/{len(range(0, T, 97))} (different dB bits: near-ties may flip)",
        "cpu_port": {"seconds": t_vqt + t_ana, "value": T / (t_vqt + t_ana), "vqt_seconds": t_vqt, "analysis_seconds": t_ana,
                     "cores": threads, "kind": "port",
                     "note": "oracle VQT (f32, all host threads) + oracle AnalysisState epilogue (sequential, one thread)"}}
    # the viewer's loop (vqt_system.rs:40-68 + analysis_system.rs:10-20): one frame of audio per call, 60 times a second
    n_fft = 32768
    live = pv.AnalysisState(pv.VqtRange(), device=ctx.local_rank)
    outs = ctx.ffi.PvqtAnalysisOutputs()
    outs.max_peaks = 32
    keep = {"peak_count": np.zeros(1, np.uint32), "peak_indices": np.zeros(32, np.uint32),
            "peaks_continuous": np.zeros((32, 2), np.float32), "smoothed_scene_calmness": np.zeros(1, np.float32),
            "smoothed_tuning_grid_inaccuracy": np.zeros(1, np.float32)}
    for name, arr in keep.items():
        setattr(outs, name, arr.ctypes.data_as(C.c_void_p))
    moved = C.c_uint64(0)
    lat = []
    for t in range(450):
        x = np.ascontiguousarray(audio[t * hop:t * hop + n_fft])
        xp = x.ctypes.data_as(C.POINTER(C.c_float))
        t0 = time.perf_counter()
        rc = ctx.lib.pvqt_calc_batch_analysis(v.handle, live._h, xp, n_fft, hop, 1, FRAME_NS, C.byref(outs), None, C.byref(moved))
        lat.append(time.perf_counter() - t0)
        ctx.chk(rc)
    live.close()
    lat = np.array(lat[50:]) * 1e6
    # the CPU port frame by frame on one core: the VQT with its scratch set up once, then the epilogue
    t0 = time.perf_counter()
    o.calculate_batch_db(audio[:255 * hop + n_fft], hop, 256, mode=1, n_threads=1)
    cpu_vqt_us = (time.perf_counter() - t0) / 256 * 1e6
    rec["frame_by_frame"] = {
        "workload": "one frame of audio per call through pvqt_calc_batch_analysis (n_frames = 1): host x[n_fft] in, peaks and "
                    "scalars out, the AnalysisState advancing from call to call",
        "api": "pinned staging both ways, one captured graph per call: H2D, K-fft, K-spmm-db, K-analysis, one D2H",
        "p50_us": float(np.percentile(lat, 50)), "p99_us": float(np.percentile(lat, 99)), "min_us": float(lat.min()),
        "calls": int(lat.size),
        "cpu_port": {"us_per_frame": cpu_vqt_us + 1e6 * t_ana / T, "vqt_us": cpu_vqt_us, "analysis_us": 1e6 * t_ana / T,
                     "cores": 1, "kind": "port"}}
    # many streams: stream-parallel epilogue, one CTA per stream
    if streams_audio is not None and streams_audio.shape[0] >= 64:
        S = min(streams_audio.shape[0], 1024)
        src = np.ascontiguousarray(streams_audio[:S])
        pin = C.c_void_p()      # pinned host audio, like the e2e arm of the headline (pageable input is copied at ~6 GB/s)
        ctx.chk(ctx.lib.pvqt_host_alloc_pinned(src.nbytes, C.byref(pin)))
        sa = np.ctypeslib.as_array(C.cast(pin, C.POINTER(C.c_float)), shape=src.shape)
        sa[...] = src
        fps_ = v.frames_in(sa.shape[1], hop)
        a = pv.AnalysisState(pv.VqtRange(), n_streams=S, device=ctx.local_rank)
        # the C ABI with result arrays the caller keeps (allocated and touched once, as a consumer in a loop would): a fresh
        # numpy array per call costs more in page faults during the copy back than the copy itself
        so = ctx.ffi.PvqtAnalysisOutputs()
        so.max_peaks = 32
        held = {"peak_count": np.empty((S, fps_), np.uint32), "peak_indices": np.empty((S, fps_, 32), np.uint32),
                "peaks_continuous": np.empty((S, fps_, 32, 2), np.float32),
                "smoothed_scene_calmness": np.empty((S, fps_), np.float32),
                "smoothed_tuning_grid_inaccuracy": np.empty((S, fps_), np.float32)}
        for name, arr in held.items():
            arr.fill(0)
            setattr(so, name, arr.ctypes.data_as(C.c_void_p))
        got = C.c_uint64(0)
        sap = sa.ctypes.data_as(C.POINTER(C.c_float))

        def streams_call():
            ctx.chk(ctx.lib.pvqt_calc_streams_analysis(v.handle, a._h, sap, S, sa.shape[1], sa.shape[1], hop, fps_, FRAME_NS,
                                                       C.byref(so), None, C.byref(got)))
        streams_call()
        t0 = time.perf_counter()
        streams_call()
        dt = time.perf_counter() - t0
        res = {"d2h_bytes": int(got.value)}
        a.close()
        rec["streams"] = {
            "workload": f"the first {S} streams of streams4096 ({S * fps_} frames) through pvqt_calc_streams_analysis "
                        "(pinned host audio in; peaks and scalars back into pageable arrays the caller keeps)",
            "seconds": dt, "value": S * fps_ / dt, "unit": UNIT, "h2d_bytes_per_step": int(sa.nbytes),
            "d2h_bytes_per_step": int(res["d2h_bytes"]), "d2h_bytes_if_spectra_were_returned": int(S * fps_ * v.n_buckets * 4)}
        del sa
        ctx.lib.pvqt_host_free_pinned(pin)
    v.close()
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="chords60", choices=["chords60", "hires60", "streams4096"])
    ap.add_argument("--configs", default="all",
                    help="sub-records beside the headline workload: all | none | comma list of streams4096, hires60, "
                         "instant, pipeline5")
    ap.add_argument("--streams", type=int, default=4096, help="streams4096: total streams over all ranks")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sustain", type=float, default=0.6, help="seconds of back-to-back steps for the clock record")
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

    ctx = Ctx(args, rank, world, local_rank, dist)
    want = {"streams4096", "hires60", "instant", "pipeline5"} if args.configs == "all" else \
        (set() if args.configs == "none" else {c.strip() for c in args.configs.split(",") if c.strip()})
    want.discard(args.workload)
    strong = args.workload == "streams4096"
    steps = 10 if strong and args.steps == 50 else args.steps

    t_all = time.perf_counter()
    main_rec = measure(ctx, args.workload, steps, args.warmup, e2e_steps=max(3, min(steps, 20 if not strong else 3)),
                       profile=True, sustain_s=args.sustain if not strong else 0.0,
                       want_cpu=(not args.no_cpu_baseline and world == 1 and rank == 0))
    e2e = main_rec["e2e"]
    if world > 1:
        e2e = multi_e2e(ctx, args.workload, main_rec, 5 if not strong else 2)
    elif e2e is not None:
        e2e["pcie_ceiling_frames_per_s"] = single_e2e_ceiling(ctx, main_rec)
        e2e["frac_of_pcie_ceiling"] = e2e["value"] / e2e["pcie_ceiling_frames_per_s"]
    main_rec["audio"] = None

    configs = {}
    streams_audio = None
    if "hires60" in want:
        r = measure(ctx, "hires60", min(steps, 20), args.warmup, e2e_steps=3, profile=True)
        configs["hires60"] = sub_record(ctx, r, "hires60", args)
    if "streams4096" in want:
        r = measure(ctx, "streams4096", min(steps, 5), 3, e2e_steps=2, profile=True)
        if world > 1:
            r["e2e"] = multi_e2e(ctx, "streams4096", r, 2)
        elif r["e2e"] is not None:
            r["e2e"]["pcie_ceiling_frames_per_s"] = single_e2e_ceiling(ctx, r)
            r["e2e"]["frac_of_pcie_ceiling"] = r["e2e"]["value"] / r["e2e"]["pcie_ceiling_frames_per_s"]
        streams_audio = r["audio"] if world == 1 else None
        r["audio"] = None
        configs["streams4096"] = sub_record(ctx, r, "streams4096", args)
    if rank == 0 and "instant" in want:
        configs["instant"] = bench_instant(ctx)
    if rank == 0 and "pipeline5" in want:
        configs["pipeline5"] = bench_pipeline(ctx, streams_audio)
    ctx.barrier()

    if rank == 0:
        rec = main_rec
        n_frames = rec["n_frames_rank"]
        line = {
            "metric": METRIC, "value": rec["value"], "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": args.warmup, "ms_per_step": rec["ms_per_step"], "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": workload_name(args.workload, rec["hop"], rec["total_frames_per_step"] if strong else n_frames, args.streams),
                "n_fft": rec["n_fft"], "n_buckets": rec["n_buckets"], "hop": rec["hop"], "frames_per_step_per_gpu": n_frames,
                "l2": "flushed between timed steps (512 MiB memset"
                      + (" followed by a read sweep of the same buffer, so the flush leaves no dirty lines)"
                         if args.flush != "write" else ")"),
                "timing": "CUDA events per step on the launching stream, summed; max over ranks",
            },
            "e2e": e2e, "gpu_launches": rec["launches"], "plan": rec["plan"], "clocks": rec["clocks"],
            "clocks_note": "sampled every 50 ms over the K timed steps AND the sustained run of the same step that follows "
                           "(`sustained`), so that the record has more than a couple of samples under load",
            "sustained": rec["sustained"],
            "roofline": roofline_record(ctx, rec, args.workload),
            "step_ms": rec["step_ms"], "wall_s_timed_region": rec["wall_s_timed_region"],
            "configs": configs, "bench_wall_s": time.perf_counter() - t_all,
        }
        if rec["cpu"] is not None:
            line["cpu_baseline"] = rec["cpu"]
        elif not args.no_cpu_baseline:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "measured at N=1 only"}
        emit(line)

    if dist is not None:
        dist.barrier(group=ctx.cpu_group)
        dist.destroy_process_group()


def sub_record(ctx: Ctx, r, key: str, args):
    strong = key == "streams4096"
    out = {"workload": workload_name(key, r["hop"], r["total_frames_per_step"] if strong else r["n_frames_rank"], args.streams),
           "metric": METRIC, "value": r["value"], "unit": UNIT, "ms_per_step": r["ms_per_step"], "steps": r["steps"],
           "scaling": "strong" if strong else "weak", "frames_per_step_per_gpu": r["n_frames_rank"], "step_ms": r["step_ms"],
           "gpu_launches": r["launches"], "e2e": r["e2e"], "clocks": r["clocks"], "roofline": roofline_record(ctx, r, key)}
    return out


if __name__ == "__main__":
    main()
