//! Host-side helpers of the reference's `util.rs` that callers use directly.  `EmaMeasurement` here is a
//! read-only snapshot of the device-side state (the recurrence util.rs:106-125 runs in the epilogue kernel).
use std::time::Duration;

pub fn arg_max(sl: &[f32]) -> usize {
    // util.rs:31-43: index of the largest element
    let mut best = 0;
    for (i, v) in sl.iter().enumerate() {
        if *v > sl[best] {
            best = i;
        }
    }
    best
}

pub fn max(sl: &[f32]) -> f32 {
    sl.iter().fold(f32::MIN, |a, b| a.max(*b))
}

pub fn min(sl: &[f32]) -> f32 {
    sl.iter().fold(f32::MAX, |a, b| a.min(*b))
}

/// Snapshot of one exponentially smoothed value (util.rs:91-137).  `get()` as in the reference.
#[derive(Debug, Clone, Copy, Default)]
pub struct EmaMeasurement {
    pub(crate) y: f32,
    pub(crate) time_horizon: Option<Duration>,
}

impl EmaMeasurement {
    pub fn get(&self) -> f32 {
        self.y
    }
    pub fn time_horizon(&self) -> Option<Duration> {
        self.time_horizon
    }
}

/// `test_create_sines` (util.rs:62-79): the reference's test-signal generator, kept bit-compatible
/// (f32 phase arithmetic) so its `#[cfg(test)]` modules run unchanged over this crate.
pub fn test_create_sines(params: &crate::vqt::VqtParameters, freqs: &[f32], t_diff: f32) -> Vec<f32> {
    let mut wave = vec![0.0f32; params.n_fft];
    for f in freqs {
        for (i, w) in wave.iter_mut().enumerate() {
            let amp = (((i as f32 + t_diff * params.sr) * 2.0 * std::f32::consts::PI / params.sr) * f).sin() / 12.0;
            *w += amp;
        }
    }
    wave
}
