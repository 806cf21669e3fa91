//! `dagc::MonoAgc` (dagc_fork/src/lib.rs:19-87 upstream) over the CUDA AGC stage of libpvqt.so.
//! Same constructor checks, `freeze_gain`, `is_gain_frozen`, `gain` and `process(&mut [f32])`; new:
//! `process_chunks` runs the callers' per-chunk loop (freeze on silent chunks) in one launch.
use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

#[derive(Debug, thiserror::Error)]
#[allow(missing_docs)]
pub enum Error {
    #[error("`desired_output_rms` must be a finite positive number, but got {value}")]
    InvalidDesiredOutputRms { value: f32 },
    #[error("`distortion_factor` must be a number within `0.0 ..= 1.0`, but got {value}")]
    InvalidDistortionFactor { value: f32 },
    #[error("pvqt backend error (status {status}): {message}")]
    Backend { status: i32, message: String },
}

extern "C" {
    fn pvqt_last_error_string() -> *const c_char;
    fn pvqt_agc_create(rms: f32, distortion: f32, n_streams: usize, device: c_int, out: *mut *mut c_void) -> c_int;
    fn pvqt_agc_destroy(a: *mut c_void);
    fn pvqt_agc_gains(a: *mut c_void, out: *mut f32) -> c_int;
    fn pvqt_agc_freeze_gain(a: *mut c_void, freeze: c_int) -> c_int;
    fn pvqt_agc_process(a: *mut c_void, audio: *mut f32, stream_stride: usize, n_samples: usize, chunk: usize,
                        silence_threshold: f32) -> c_int;
}

#[derive(Debug)]
pub struct MonoAgc {
    handle: *mut c_void,
    frozen: bool,
}

unsafe impl Send for MonoAgc {}

impl Drop for MonoAgc {
    fn drop(&mut self) {
        unsafe { pvqt_agc_destroy(self.handle) }
    }
}

impl MonoAgc {
    pub fn new(desired_output_rms: f32, distortion_factor: f32) -> Result<Self, Error> {
        if !(desired_output_rms > 0.0 && desired_output_rms.is_finite()) {
            return Err(Error::InvalidDesiredOutputRms { value: desired_output_rms });
        }
        if !(0.0..=1.0).contains(&distortion_factor) {
            return Err(Error::InvalidDistortionFactor { value: distortion_factor });
        }
        let device = std::env::var("PVQT_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
        let mut handle = std::ptr::null_mut();
        let status = unsafe { pvqt_agc_create(desired_output_rms, distortion_factor, 1, device, &mut handle) };
        if status != 0 {
            let message = unsafe { CStr::from_ptr(pvqt_last_error_string()).to_string_lossy().into_owned() };
            return Err(Error::Backend { status, message });
        }
        Ok(Self { handle, frozen: false })
    }

    pub fn freeze_gain(&mut self, freeze: bool) {
        self.frozen = freeze;
        unsafe { pvqt_agc_freeze_gain(self.handle, freeze as c_int) };
    }

    pub const fn is_gain_frozen(&self) -> bool {
        self.frozen
    }

    pub fn gain(&self) -> f32 {
        let mut g = 1.0f32;
        unsafe { pvqt_agc_gains(self.handle, &mut g) };
        g
    }

    /// One chunk, honouring `freeze_gain` (silence_threshold = NaN selects the flag).
    pub fn process(&mut self, samples: &mut [f32]) {
        let rc = unsafe { pvqt_agc_process(self.handle, samples.as_mut_ptr(), samples.len(), samples.len(), 0, f32::NAN) };
        assert!(rc == 0, "pvqt_agc_process failed");
    }

    /// New: the callers' loop (audio_desktop.rs:101-117, train.rs:296-310) over a whole recording.
    pub fn process_chunks(&mut self, samples: &mut [f32], chunk: usize, silence_threshold: f32) {
        let rc = unsafe {
            pvqt_agc_process(self.handle, samples.as_mut_ptr(), samples.len(), samples.len(), chunk, silence_threshold)
        };
        assert!(rc == 0, "pvqt_agc_process failed");
    }
}
