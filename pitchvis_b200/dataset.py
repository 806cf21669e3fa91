"""Offline dataset caller: the sliding loop of pitchvis_train/src/train.rs:252-351 as ONE
batched call, and its `.npy` writer (train.rs:192-208).

The reference renders audio in chunks of `vqt_delay_in_samples` (the VQT delay in whole milliseconds, rounded down to
a multiple of 64 samples, train.rs:128-129), pushes every chunk through the AGC into a zero-initialised ring buffer and,
every STEP_SIZE_IN_CHUNKS = 3 chunks (train.rs:43, :312), transforms the last N_FFT samples of the ring buffer
(train.rs:341).  That is a sliding window with hop 3 * chunk over the AGC'd stream, left-padded with N_FFT zeros:
frame j = padded[(j + 1) * hop : (j + 1) * hop + N_FFT].  MIDI synthesis itself is out of scope; the caller passes
rendered audio and (optionally) the 128 key gains per frame.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from .agc import MonoAgc
from .vqt import Vqt, VqtParameters, VqtRange

# train.rs:30-41
TRAIN_SR = 22050
TRAIN_PARAMS = VqtParameters(sr=float(TRAIN_SR), n_fft=2 * 16384, range=VqtRange(55.0, 7, 36), sparsity_quantile=0.999,
                             quality=10.0, gamma=5.3 * 10.0)
STEP_SIZE_IN_CHUNKS = 3   # train.rs:43
N_KEYS = 128              # train.rs:177


def train_chunk_samples(vqt: Vqt, sr: int = TRAIN_SR) -> int:
    """vqt_delay_in_samples, train.rs:128-129: delay.as_millis() * SR / 1000, rounded down to a multiple of 64."""
    delay_ms = int(vqt.delay * 1000.0)
    return (delay_ms * sr // 1000) // 64 * 64


def annotated_vqt(vqt: Vqt, audio: np.ndarray, agc: Optional[MonoAgc] = None,
                  step_size_in_chunks: int = STEP_SIZE_IN_CHUNKS, chunk: Optional[int] = None) -> np.ndarray:
    """dB spectra of the reference's dataset loop for one rendered recording: [frames][n_buckets].

    `agc` (e.g. MonoAgc(0.07, 0.001), train.rs:271) is applied chunk by chunk first (silent chunks frozen,
    train.rs:296-298)."""
    audio = np.ascontiguousarray(audio, np.float32)
    chunk = chunk or train_chunk_samples(vqt, int(vqt.params().sr))
    n_chunks = audio.shape[0] // chunk                 # the reference renders whole chunks only
    audio = audio[:n_chunks * chunk]
    if agc is not None:
        audio = agc.process_chunks(audio, chunk, 1e-6)
    hop = step_size_in_chunks * chunk
    n_frames = n_chunks // step_size_in_chunks
    if n_frames == 0:
        return np.zeros((0, vqt.n_buckets), np.float32)
    padded = np.concatenate([np.zeros(vqt.n_fft, np.float32), audio])
    return vqt.calculate_vqt_batch_in_db(padded[hop:], hop, n_frames=n_frames)


def write_dataset_npy(path: str, x_vqt: np.ndarray, targets: Optional[np.ndarray] = None) -> int:
    """train.rs:156-208: one flat little-endian f32 array, per data point n_buckets dB values then 128 targets.
    Returns the number of floats written."""
    x_vqt = np.ascontiguousarray(x_vqt, np.float32)
    if targets is None:
        targets = np.zeros((x_vqt.shape[0], N_KEYS), np.float32)
    targets = np.ascontiguousarray(targets, np.float32)
    if targets.shape != (x_vqt.shape[0], N_KEYS):
        raise ValueError("targets must be [frames][128]")   # assert!(target.len() == 128), train.rs:177
    data = np.concatenate([x_vqt, targets], axis=1).reshape(-1).astype("<f4")
    np.save(path, data)
    return int(data.shape[0])
