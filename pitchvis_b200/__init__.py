"""pitchvis_b200 -- B200-native (sm_100a) implementation of pitchvis_analysis's VQT hot path.

The product is the C-ABI library `pitchvis_b200/lib/libpvqt.so` (include/pvqt.h), built from
`pitchvis_b200/csrc/` by `python -m pitchvis_b200.build`.  This package is the thin Python host
used by the tests and bench; the Rust shim crates (rust/) bind the same ABI.  There is no CPU fallback.
"""
from .vqt import (  # noqa: F401
    AboveNyquist, CsMat, DeviceBuffer, HostKernel, MultiVqt, PvqtRuntimeError, Vqt, VqtError, VqtKernel,
    VqtParameters, VqtRange, WindowExceedsNFft, WindowGroup, calc_db_device, fft_device, filter_bank_params,
    set_log_callback, synchronize,
)
from .analysis import (  # noqa: F401,E402
    AnalysisParameters, AnalysisState, PeakDetectionParameters, chroma, ml_input_windows, spectrogram_peaks, spectrogram_vqt,
)
from .agc import AgcError, MonoAgc  # noqa: F401,E402
