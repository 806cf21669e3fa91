//! `pitchvis_analysis::vqt` backed by libpvqt.so (B200 / sm_100a).
//!
//! Drop-in for the reference module of the same path: the public items keep their names, field
//! layout and error behaviour, so `pitchvis_viewer/src/vqt_system.rs`, `pitchvis_serial`,
//! `pitchvis_train` etc. compile unchanged.  All arithmetic happens behind the C ABI declared in
//! `include/pvqt.h`; this file only converts types.  (No Rust toolchain exists in the build image:
//! this shim is source-only and kept small enough to review by eye against the header.)

use num_complex::Complex32;
use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};
use std::time::Duration;

pub const DEFAULT_SR: u32 = 22050;
pub const DEFAULT_N_FFT: usize = 2 * 16384;
pub const DEFAULT_MIN_FREQ: f32 = 55.0;
pub const DEFAULT_UPSCALE_FACTOR: u16 = 1;
pub const DEFAULT_BUCKETS_PER_SEMITONE: u16 = 7 * DEFAULT_UPSCALE_FACTOR;
pub const DEFAULT_BUCKETS_PER_OCTAVE: u16 = 12 * DEFAULT_BUCKETS_PER_SEMITONE;
pub const DEFAULT_OCTAVES: u8 = 7;
pub const DEFAULT_SPARSITY_QUANTILE: f32 = 0.999;
pub const DEFAULT_Q: f32 = 1.6 / DEFAULT_UPSCALE_FACTOR as f32;
pub const DEFAULT_GAMMA: f32 = 4.8 * DEFAULT_Q;

#[derive(Debug, Clone)]
pub struct VqtRange {
    pub min_freq: f32,
    pub octaves: u8,
    pub buckets_per_octave: u16,
}

impl VqtRange {
    pub fn n_buckets(&self) -> usize {
        self.buckets_per_octave as usize * self.octaves as usize
    }
}

#[derive(Debug, Clone)]
pub struct VqtParameters {
    pub sr: f32,
    pub n_fft: usize,
    pub range: VqtRange,
    pub sparsity_quantile: f32,
    pub quality: f32,
    pub gamma: f32,
}

impl Default for VqtParameters {
    fn default() -> Self {
        Self {
            sr: DEFAULT_SR as f32,
            n_fft: DEFAULT_N_FFT,
            range: VqtRange {
                min_freq: DEFAULT_MIN_FREQ,
                octaves: DEFAULT_OCTAVES,
                buckets_per_octave: DEFAULT_BUCKETS_PER_OCTAVE,
            },
            sparsity_quantile: DEFAULT_SPARSITY_QUANTILE,
            quality: DEFAULT_Q,
            gamma: DEFAULT_GAMMA,
        }
    }
}

#[derive(Debug, thiserror::Error)]
pub enum VqtError {
    #[error(
        "the highest VQT bin frequency ({highest_frequency} Hz) exceeds the Nyquist \
         frequency ({nyquist_frequency} Hz); reduce octaves or increase the sample rate"
    )]
    AboveNyquist { highest_frequency: f32, nyquist_frequency: f32 },
    #[error(
        "the longest filter window ({window_length} samples) exceeds n_fft ({n_fft} \
         samples); increase n_fft or gamma, or decrease quality"
    )]
    WindowExceedsNFft { window_length: f32, n_fft: usize },
    /// New: the GPU backend could not be brought up (there is no CPU fallback).
    #[error("pvqt backend error (status {status}): {message}")]
    Backend { status: i32, message: String },
}

pub struct WindowGroup {
    pub window: (usize, usize),
    pub filter_bank: sprs::CsMat<Complex32>,
    pub negative_filter_bank: Option<sprs::CsMat<Complex32>>,
}

impl WindowGroup {
    pub fn window_size(&self) -> usize {
        self.window.1 - self.window.0
    }
}

pub struct VqtKernel {
    pub window_groups: Vec<WindowGroup>,
}

// ---- include/pvqt.h -------------------------------------------------------------------------------
#[repr(C)]
struct PvqtParams {
    sr: f32,
    n_fft: u64,
    min_freq: f32,
    octaves: u32,
    buckets_per_octave: u32,
    sparsity_quantile: f32,
    quality: f32,
    gamma: f32,
}

#[repr(C)]
#[derive(Default)]
struct PvqtError {
    status: i32,
    highest_frequency: f32,
    nyquist_frequency: f32,
    window_length: f32,
    n_fft: u64,
    cuda_error: i32,
}

#[repr(C)]
struct PvqtCsrView {
    rows: i32,
    cols: i32,
    nnz: i64,
    indptr: *const i32,
    indices: *const i32,
    data: *const f32,
}

const PVQT_OK: c_int = 0;
const PVQT_ABOVE_NYQUIST: c_int = 1;
const PVQT_WINDOW_EXCEEDS_NFFT: c_int = 2;
const PVQT_BAD_LENGTH: c_int = 4;

#[link(name = "pvqt")]
extern "C" {
    fn pvqt_set_log_callback(f: Option<extern "C" fn(c_int, *const c_char, *mut c_void)>, user: *mut c_void, max_level: c_int) -> c_int;
    fn pvqt_last_error_string() -> *const c_char;
    fn pvqt_create(params: *const PvqtParams, device: c_int, out: *mut *mut c_void, err: *mut PvqtError) -> c_int;
    fn pvqt_destroy(v: *mut c_void);
    fn pvqt_delay_seconds(v: *const c_void) -> f64;
    fn pvqt_num_window_groups(v: *const c_void) -> usize;
    fn pvqt_group_window(v: *const c_void, group: usize, begin: *mut u64, end: *mut u64) -> c_int;
    fn pvqt_group_csr(v: *const c_void, group: usize, negative: c_int, out: *mut PvqtCsrView) -> c_int;
    fn pvqt_frames_in(v: *const c_void, n_samples: usize, hop: usize) -> usize;
    fn pvqt_calc_instant_db(v: *mut c_void, x: *const f32, n: usize, out: *mut f32) -> c_int;
    fn pvqt_calc_batch_db(v: *mut c_void, audio: *const f32, n_samples: usize, hop: usize, n_frames: usize,
                          out: *mut f32) -> c_int;
    fn pvqt_calc_streams_db(v: *mut c_void, audio: *const f32, n_streams: usize, stream_stride: usize,
                            n_samples: usize, hop: usize, frames_per_stream: usize, out: *mut f32) -> c_int;
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(pvqt_last_error_string()).to_string_lossy().into_owned() }
}

pub struct Vqt {
    params: VqtParameters,
    vqt_kernel: VqtKernel,
    pub delay: Duration,
    handle: *mut c_void,
}

// `&mut self` on every compute method: a handle is never used from two threads at once.
unsafe impl Send for Vqt {}
unsafe impl Sync for Vqt {}

impl Drop for Vqt {
    fn drop(&mut self) {
        unsafe { pvqt_destroy(self.handle) }
    }
}

unsafe fn csmat(v: &PvqtCsrView) -> sprs::CsMat<Complex32> {
    let nnz = v.nnz as usize;
    let indptr = std::slice::from_raw_parts(v.indptr, v.rows as usize + 1).iter().map(|&i| i as usize).collect();
    let indices = std::slice::from_raw_parts(v.indices, nnz).iter().map(|&i| i as usize).collect();
    let raw = std::slice::from_raw_parts(v.data, 2 * nnz);
    let data = raw.chunks_exact(2).map(|c| Complex32::new(c[0], c[1])).collect();
    sprs::CsMat::new((v.rows as usize, v.cols as usize), indptr, indices, data)
}

impl Vqt {
    /// `Vqt::new` on CUDA device 0 (`PVQT_DEVICE` selects another one).
    pub fn new(params: &VqtParameters) -> Result<Self, VqtError> {
        // the reference logs through the `log` crate (vqt.rs:468, :661-667, :688-710, :741-746): forward the library's lines
        extern "C" fn sink(level: c_int, message: *const c_char, _user: *mut c_void) {
            let m = unsafe { CStr::from_ptr(message) }.to_string_lossy();
            match level {
                1 => log::warn!("{m}"),
                2 => log::info!("{m}"),
                _ => log::debug!("{m}"),
            }
        }
        unsafe { pvqt_set_log_callback(Some(sink), std::ptr::null_mut(), 3) };
        let device = std::env::var("PVQT_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
        Self::new_on_device(params, device)
    }

    pub fn new_on_device(params: &VqtParameters, device: i32) -> Result<Self, VqtError> {
        let p = PvqtParams {
            sr: params.sr,
            n_fft: params.n_fft as u64,
            min_freq: params.range.min_freq,
            octaves: params.range.octaves as u32,
            buckets_per_octave: params.range.buckets_per_octave as u32,
            sparsity_quantile: params.sparsity_quantile,
            quality: params.quality,
            gamma: params.gamma,
        };
        let mut handle = std::ptr::null_mut();
        let mut err = PvqtError::default();
        match unsafe { pvqt_create(&p, device, &mut handle, &mut err) } {
            PVQT_OK => {}
            PVQT_ABOVE_NYQUIST => {
                return Err(VqtError::AboveNyquist {
                    highest_frequency: err.highest_frequency,
                    nyquist_frequency: err.nyquist_frequency,
                })
            }
            PVQT_WINDOW_EXCEEDS_NFFT => {
                return Err(VqtError::WindowExceedsNFft { window_length: err.window_length, n_fft: err.n_fft as usize })
            }
            status => return Err(VqtError::Backend { status, message: last_error() }),
        }
        let mut window_groups = Vec::new();
        unsafe {
            for g in 0..pvqt_num_window_groups(handle) {
                let (mut b, mut e) = (0u64, 0u64);
                pvqt_group_window(handle, g, &mut b, &mut e);
                let mut pos = std::mem::zeroed::<PvqtCsrView>();
                let mut neg = std::mem::zeroed::<PvqtCsrView>();
                pvqt_group_csr(handle, g, 0, &mut pos);
                pvqt_group_csr(handle, g, 1, &mut neg);
                window_groups.push(WindowGroup {
                    window: (b as usize, e as usize),
                    filter_bank: csmat(&pos),
                    negative_filter_bank: (neg.nnz > 0).then(|| csmat(&neg)),
                });
            }
        }
        let delay = Duration::from_secs_f32(unsafe { pvqt_delay_seconds(handle) } as f32);
        log::info!("VQT analysis delay: {} ms.", delay.as_millis());
        Ok(Self { params: params.clone(), vqt_kernel: VqtKernel { window_groups }, delay, handle })
    }

    /// The C handle, for entries that take a `Vqt` and an `AnalysisState` (analysis::calculate_and_preprocess).
    pub(crate) fn raw_handle(&mut self) -> *mut c_void {
        self.handle
    }

    pub fn frames_in(&self, n_samples: usize, hop: usize) -> usize {
        unsafe { pvqt_frames_in(self.handle, n_samples, hop) }
    }

    pub fn params(&self) -> &VqtParameters {
        &self.params
    }

    pub fn kernel(&self) -> &VqtKernel {
        &self.vqt_kernel
    }

    /// Same contract as the reference: panics unless `x.len() == n_fft`.
    pub fn calculate_vqt_instant_in_db(&mut self, x: &[f32]) -> Vec<f32> {
        assert_eq!(x.len(), self.params.n_fft, "input must be exactly n_fft samples");
        let mut out = vec![0.0f32; self.params.range.n_buckets()];
        let rc = unsafe { pvqt_calc_instant_db(self.handle, x.as_ptr(), x.len(), out.as_mut_ptr()) };
        assert!(rc != PVQT_BAD_LENGTH, "input must be exactly n_fft samples");
        assert!(rc == PVQT_OK, "pvqt_calc_instant_db failed: {}", last_error());
        out
    }

    /// New: frame `t` is the transform of `audio[t * hop .. t * hop + n_fft]`; returns
    /// `frames * n_buckets` values, frame-major.
    pub fn calculate_vqt_batch_in_db(&mut self, audio: &[f32], hop: usize) -> Vec<f32> {
        let n_frames = unsafe { pvqt_frames_in(self.handle, audio.len(), hop) };
        let mut out = vec![0.0f32; n_frames * self.params.range.n_buckets()];
        let rc = unsafe { pvqt_calc_batch_db(self.handle, audio.as_ptr(), audio.len(), hop, n_frames, out.as_mut_ptr()) };
        assert!(rc == PVQT_OK, "pvqt_calc_batch_db failed: {}", last_error());
        out
    }

    /// New: `n_streams` recordings of `n_samples` each, stored back to back.
    pub fn calculate_vqt_streams_in_db(&mut self, audio: &[f32], n_streams: usize, hop: usize) -> Vec<f32> {
        assert!(n_streams > 0 && audio.len() % n_streams == 0);
        let n_samples = audio.len() / n_streams;
        let frames = unsafe { pvqt_frames_in(self.handle, n_samples, hop) };
        let mut out = vec![0.0f32; n_streams * frames * self.params.range.n_buckets()];
        let rc = unsafe {
            pvqt_calc_streams_db(self.handle, audio.as_ptr(), n_streams, n_samples, n_samples, hop, frames, out.as_mut_ptr())
        };
        assert!(rc == PVQT_OK, "pvqt_calc_streams_db failed: {}", last_error());
        out
    }
}
