"""Python mirror of dagc_fork::MonoAgc (dagc_fork/src/lib.rs:19-87) over libpvqt.so (include/pvqt_agc.h).

`MonoAgc(desired_output_rms, distortion_factor)` keeps the reference's names (`process`, `freeze_gain`,
`is_gain_frozen`, `gain`); `n_streams > 1` holds that many independent states and `process_chunks` runs the callers'
chunk loop (freeze on silent chunks, audio_desktop.rs:101-117 / train.rs:296-310) for all streams in one launch.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _ffi
from .vqt import _check, _fptr


class AgcError(ValueError):
    """Error::InvalidDesiredOutputRms / Error::InvalidDistortionFactor (lib.rs:8-16)."""


class MonoAgc:
    def __init__(self, desired_output_rms: float, distortion_factor: float, n_streams: int = 1, device: int = 0):
        self._lib = _ffi.load()
        self._h = C.c_void_p()
        rc = self._lib.pvqt_agc_create(desired_output_rms, distortion_factor, n_streams, device, C.byref(self._h))
        if rc == _ffi.PVQT_INVALID_ARGUMENT:
            raise AgcError(_ffi.last_error())
        _check(rc)
        self.n_streams = n_streams
        self._frozen = False

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.pvqt_agc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def freeze_gain(self, freeze: bool):
        self._frozen = bool(freeze)
        _check(self._lib.pvqt_agc_freeze_gain(self._h, 1 if freeze else 0))

    def is_gain_frozen(self) -> bool:
        return self._frozen

    @property
    def gains(self) -> np.ndarray:
        out = np.empty(self.n_streams, np.float32)
        _check(self._lib.pvqt_agc_gains(self._h, _fptr(out)))
        return out

    def gain(self) -> float:
        return float(self.gains[0])

    def _run(self, samples, chunk: int, threshold: float) -> np.ndarray:
        x = np.array(samples, np.float32, copy=True, order="C")
        flat = x.reshape(self.n_streams, -1) if x.ndim > 1 or self.n_streams == 1 else x
        if flat.shape[0] != self.n_streams:
            raise ValueError("audio must be [n_streams][n_samples]")
        _check(self._lib.pvqt_agc_process(self._h, _fptr(flat), flat.shape[1], flat.shape[1], chunk, threshold))
        return flat.reshape(x.shape)

    def process(self, samples) -> np.ndarray:
        """MonoAgc::process on one chunk (per stream), honouring freeze_gain(); returns the processed samples."""
        return self._run(samples, 0, math.nan)

    def process_chunks(self, audio, chunk: int, silence_threshold: float = 1e-6) -> np.ndarray:
        """The callers' loop: per chunk, freeze when sum(x^2) < silence_threshold, then process."""
        return self._run(audio, chunk, silence_threshold)
