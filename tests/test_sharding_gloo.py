"""The N>1 path on CPU: world_size-2 `gloo` processes shard one recording by frame range (with the
n_fft - hop halo) and a set of streams by stream, exactly as bench.py / pvqt_multi_* do, using the
library's own shard arithmetic (pvqt_shard_range, pvqt_frame_range_samples: pure host code).  The
compute inside each rank is the CPU oracle (no GPU here); the gathered result must be bit-identical
to the unsharded run -- there is no collective on the data path, only the final gather."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOP = 368


def _worker(rank: int, world: int, port: int, tmpdir: str):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import orc
    from pitchvis_b200 import _ffi, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = _ffi.load()
    v = orc.OracleVqt()
    nb, n_fft = v.n_buckets, v.n_fft

    # ---- one recording, frame-range sharded ------------------------------------------------------
    audio = synth.polyphonic_chords(3.0, 22050.0, seed=11)
    n_frames = synth.frames_in(audio.shape[0], n_fft, HOP)
    f0, f1, s0, s1 = C.c_size_t(), C.c_size_t(), C.c_size_t(), C.c_size_t()
    assert lib.pvqt_shard_range(n_frames, world, rank, C.byref(f0), C.byref(f1)) == 0
    assert lib.pvqt_frame_range_samples(n_fft, HOP, f0.value, f1.value, C.byref(s0), C.byref(s1)) == 0
    mine = v.calculate_batch_db(audio[s0.value:s1.value], HOP, f1.value - f0.value, mode=1, n_threads=1)
    parts = [None] * world if rank == 0 else None      # ragged shards (46 + 45 frames): gather as objects
    dist.gather_object((f0.value, mine), parts, dst=0)
    if rank == 0:
        np.save(os.path.join(tmpdir, "frames.npy"), np.concatenate([p for _, p in sorted(parts, key=lambda x: x[0])]))

    # ---- independent streams, stream sharded -------------------------------------------------------
    n_streams, n_samples = 5, n_fft + 3 * HOP
    streams = np.stack([synth.polyphonic_chords(2.0, 22050.0, seed=20 + s)[:n_samples] for s in range(n_streams)])
    assert lib.pvqt_shard_range(n_streams, world, rank, C.byref(f0), C.byref(f1)) == 0
    mine = np.stack([v.calculate_batch_db(streams[s], HOP, mode=1, n_threads=1) for s in range(f0.value, f1.value)])
    parts = [None] * world if rank == 0 else None
    dist.gather_object((f0.value, mine), parts, dst=0)
    if rank == 0:
        np.save(os.path.join(tmpdir, "streams.npy"), np.concatenate([p for _, p in sorted(parts, key=lambda x: x[0])]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_unsharded(tmp_path, built_lib):
    import orc
    from pitchvis_b200 import synth
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    v = orc.OracleVqt()
    audio = synth.polyphonic_chords(3.0, 22050.0, seed=11)
    np.testing.assert_array_equal(np.load(tmp_path / "frames.npy"), v.calculate_batch_db(audio, HOP, mode=1, n_threads=1))
    n_samples = v.n_fft + 3 * HOP
    ref = np.stack([v.calculate_batch_db(synth.polyphonic_chords(2.0, 22050.0, seed=20 + s)[:n_samples], HOP, mode=1,
                                         n_threads=1) for s in range(5)])
    np.testing.assert_array_equal(np.load(tmp_path / "streams.npy"), ref)
