//! `pitchvis_analysis` on a B200: the module tree of the reference crate (lib.rs:1-4 upstream), with
//! `vqt` and `analysis` implemented over libpvqt.so.  `analysis_modules` keeps only the plain-data types
//! callers name (`ContinuousPeak`, `PeakDetectionParameters`); the arithmetic of those modules runs in
//! the CUDA epilogue kernel.
pub mod analysis;
pub mod analysis_modules {
    pub use crate::analysis::{ContinuousPeak, PeakDetectionParameters};
}
pub mod util;
pub mod vqt;
