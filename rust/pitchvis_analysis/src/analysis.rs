//! `pitchvis_analysis::analysis` backed by the CUDA epilogue of libpvqt.so (include/pvqt_analysis.h).
//!
//! Same public surface as the reference (analysis.rs:36-98, :119-177, :192-404): `AnalysisParameters`
//! with the same defaults, `AnalysisState::new`, `preprocess(&[f32], Duration)` (panics on a wrong length
//! like the reference's assert at :289), `update_vqt_smoothing_duration`, and the public result fields,
//! refreshed after every call.  New: `preprocess_batch` for T consecutive frames in one launch.
use std::collections::HashSet;
use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};
use std::time::Duration;

use crate::util::EmaMeasurement;
use crate::vqt::VqtRange;

#[derive(Debug, Clone, Copy)]
#[repr(C)]
pub struct PeakDetectionParameters {
    pub min_prominence: f32,
    pub min_height: f32,
}

#[derive(Debug, Clone, Copy, Default, PartialEq)]
#[repr(C)]
pub struct ContinuousPeak {
    pub center: f32,
    pub size: f32,
}

#[derive(Debug, Clone)]
pub struct AnalysisParameters {
    pub spectrogram_length: usize,
    pub peak_config: PeakDetectionParameters,
    pub bassline_peak_config: PeakDetectionParameters,
    pub highest_bassnote: usize,
    pub vqt_smoothing_duration_base: Duration,
    pub vqt_smoothing_calmness_min: f32,
    pub vqt_smoothing_calmness_max: f32,
    pub note_calmness_smoothing_duration: Duration,
    pub scene_calmness_smoothing_duration: Duration,
    pub tuning_inaccuracy_smoothing_duration: Duration,
    pub harmonic_threshold: f32,
}

impl Default for AnalysisParameters {
    fn default() -> Self {
        // analysis.rs:72-98
        Self {
            spectrogram_length: 400,
            peak_config: PeakDetectionParameters { min_prominence: 10.0, min_height: 4.0 },
            bassline_peak_config: PeakDetectionParameters { min_prominence: 5.0, min_height: 3.5 },
            highest_bassnote: 12 * 2 + 4,
            vqt_smoothing_duration_base: Duration::from_millis(70),
            vqt_smoothing_calmness_min: 0.6,
            vqt_smoothing_calmness_max: 2.0,
            note_calmness_smoothing_duration: Duration::from_millis(3_500),
            scene_calmness_smoothing_duration: Duration::from_millis(800),
            tuning_inaccuracy_smoothing_duration: Duration::from_millis(4_000),
            harmonic_threshold: 0.3,
        }
    }
}

// ---- include/pvqt_analysis.h ----------------------------------------------------------------------
#[repr(C)]
struct PvqtAnalysisParams {
    spectrogram_length: u64,
    peak_config: PeakDetectionParameters,
    bassline_peak_config: PeakDetectionParameters,
    highest_bassnote: u64,
    vqt_smoothing_duration_base_ns: u64,
    vqt_smoothing_calmness_min: f32,
    vqt_smoothing_calmness_max: f32,
    note_calmness_smoothing_duration_ns: u64,
    scene_calmness_smoothing_duration_ns: u64,
    tuning_inaccuracy_smoothing_duration_ns: u64,
    harmonic_threshold: f32,
}

#[repr(C)]
struct PvqtRange {
    min_freq: f32,
    octaves: u32,
    buckets_per_octave: u32,
}

#[repr(C)]
struct PvqtAnalysisOutputs {
    max_peaks: u32,
    peak_count: *mut u32,
    peak_indices: *mut u32,
    peaks_continuous: *mut ContinuousPeak,
    x_vqt_smoothed: *mut f32,
    x_vqt_peakfiltered: *mut f32,
    x_vqt_afterglow: *mut f32,
    calmness: *mut f32,
    pitch_accuracy: *mut f32,
    pitch_deviation: *mut f32,
    smoothed_scene_calmness: *mut f32,
    smoothed_tuning_grid_inaccuracy: *mut f32,
}

const PVQT_OK: c_int = 0;
const PVQT_BAD_LENGTH: c_int = 4;
const MAX_PEAKS: usize = 128;

extern "C" {
    fn pvqt_last_error_string() -> *const c_char;
    fn pvqt_analysis_create(range: *const PvqtRange, params: *const PvqtAnalysisParams, n_streams: usize, device: c_int,
                            out: *mut *mut c_void) -> c_int;
    fn pvqt_analysis_destroy(a: *mut c_void);
    fn pvqt_analysis_update_vqt_smoothing_duration(a: *mut c_void, has_duration: c_int, duration_ns: u64) -> c_int;
    fn pvqt_analysis_preprocess_batch(a: *mut c_void, db: *const f32, n_buckets: usize, n_frames: usize,
                                      frame_time_ns: u64, out: *const PvqtAnalysisOutputs) -> c_int;
    // VQT + AnalysisState in one call: the spectra stay in HBM between the transform and the epilogue
    fn pvqt_calc_batch_analysis(v: *mut c_void, a: *mut c_void, audio: *const f32, n_samples: usize, hop: usize,
                                n_frames: usize, frame_time_ns: u64, out: *const PvqtAnalysisOutputs, out_db: *mut f32,
                                d2h_bytes: *mut u64) -> c_int;
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(pvqt_last_error_string()).to_string_lossy().into_owned() }
}

/// Results of one frame of a batch call (what the public fields of `AnalysisState` held after it).
#[derive(Debug, Clone, Default)]
pub struct FrameAnalysis {
    pub peaks: Vec<usize>,
    pub peaks_continuous: Vec<ContinuousPeak>,
    pub smoothed_scene_calmness: f32,
    pub smoothed_tuning_grid_inaccuracy: f32,
}

pub struct AnalysisState {
    pub params: AnalysisParameters,
    pub range: VqtRange,
    pub x_vqt_smoothed: Vec<EmaMeasurement>,
    pub x_vqt_peakfiltered: Vec<f32>,
    pub x_vqt_afterglow: Vec<f32>,
    pub peaks: HashSet<usize>,
    pub peaks_continuous: Vec<ContinuousPeak>,
    pub ml_midi_base_pitches: Vec<f32>,
    pub calmness: Vec<EmaMeasurement>,
    pub pitch_accuracy: Vec<f32>,
    pub pitch_deviation: Vec<f32>,
    pub smoothed_scene_calmness: EmaMeasurement,
    pub smoothed_tuning_grid_inaccuracy: EmaMeasurement,
    handle: *mut c_void,
}

unsafe impl Send for AnalysisState {}
unsafe impl Sync for AnalysisState {}

impl Drop for AnalysisState {
    fn drop(&mut self) {
        unsafe { pvqt_analysis_destroy(self.handle) }
    }
}

impl AnalysisState {
    pub fn new(range: VqtRange, params: AnalysisParameters) -> Self {
        let n = range.n_buckets();
        let r = PvqtRange { min_freq: range.min_freq, octaves: range.octaves as u32, buckets_per_octave: range.buckets_per_octave as u32 };
        let p = PvqtAnalysisParams {
            spectrogram_length: params.spectrogram_length as u64,
            peak_config: params.peak_config,
            bassline_peak_config: params.bassline_peak_config,
            highest_bassnote: params.highest_bassnote as u64,
            vqt_smoothing_duration_base_ns: params.vqt_smoothing_duration_base.as_nanos() as u64,
            vqt_smoothing_calmness_min: params.vqt_smoothing_calmness_min,
            vqt_smoothing_calmness_max: params.vqt_smoothing_calmness_max,
            note_calmness_smoothing_duration_ns: params.note_calmness_smoothing_duration.as_nanos() as u64,
            scene_calmness_smoothing_duration_ns: params.scene_calmness_smoothing_duration.as_nanos() as u64,
            tuning_inaccuracy_smoothing_duration_ns: params.tuning_inaccuracy_smoothing_duration.as_nanos() as u64,
            harmonic_threshold: params.harmonic_threshold,
        };
        let device = std::env::var("PVQT_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
        let mut handle = std::ptr::null_mut();
        let rc = unsafe { pvqt_analysis_create(&r, &p, 1, device, &mut handle) };
        assert!(rc == PVQT_OK, "pvqt_analysis_create failed: {}", last_error());
        Self {
            params,
            range,
            x_vqt_smoothed: vec![EmaMeasurement::default(); n],
            x_vqt_peakfiltered: vec![0.0; n],
            x_vqt_afterglow: vec![0.0; n],
            peaks: HashSet::new(),
            peaks_continuous: Vec::new(),
            ml_midi_base_pitches: vec![0.0; 128],
            calmness: vec![EmaMeasurement::default(); n],
            pitch_accuracy: vec![0.0; n],
            pitch_deviation: vec![0.0; n],
            smoothed_scene_calmness: EmaMeasurement::default(),
            smoothed_tuning_grid_inaccuracy: EmaMeasurement::default(),
            handle,
        }
    }

    /// analysis.rs:251-270
    pub fn update_vqt_smoothing_duration(&mut self, vqt_smoothing_duration: Option<Duration>) {
        let (has, ns) = match vqt_smoothing_duration {
            Some(d) => (1, d.as_nanos() as u64),
            None => (0, 0),
        };
        let rc = unsafe { pvqt_analysis_update_vqt_smoothing_duration(self.handle, has, ns) };
        assert!(rc == PVQT_OK, "pvqt_analysis_update_vqt_smoothing_duration failed: {}", last_error());
    }

    /// analysis.rs:288-404.  Panics if `x_vqt.len() != range.n_buckets()` (analysis.rs:289).
    pub fn preprocess(&mut self, x_vqt: &[f32], frame_time: Duration) {
        assert_eq!(x_vqt.len(), self.range.n_buckets());
        self.preprocess_batch(x_vqt, frame_time);
    }

    /// New: what every caller of the reference does per frame -- `vqt.calculate_vqt_instant_in_db(..)` then
    /// `analysis.preprocess(..)` (pitchvis_viewer/src/vqt_system.rs:40-68 + analysis_system.rs:10-20) -- for a whole
    /// recording in one library call: frame t = audio[t*hop .. t*hop + n_fft].  The dB spectra never leave the GPU; the
    /// per-frame peak results come back (the public per-bin fields are not refreshed by this entry: call
    /// `preprocess_batch` on returned spectra when they are needed).
    pub fn calculate_and_preprocess(&mut self, vqt: &mut crate::vqt::Vqt, audio: &[f32], hop: usize,
                                    frame_time: Duration) -> Vec<FrameAnalysis> {
        let t = vqt.frames_in(audio.len(), hop);
        if t == 0 {
            return Vec::new();
        }
        let mut count = vec![0u32; t];
        let mut indices = vec![0u32; t * MAX_PEAKS];
        let mut cont = vec![ContinuousPeak::default(); t * MAX_PEAKS];
        let mut scene = vec![0.0f32; t];
        let mut tuning = vec![0.0f32; t];
        let null = std::ptr::null_mut();
        let out = PvqtAnalysisOutputs {
            max_peaks: MAX_PEAKS as u32,
            peak_count: count.as_mut_ptr(),
            peak_indices: indices.as_mut_ptr(),
            peaks_continuous: cont.as_mut_ptr(),
            x_vqt_smoothed: null, x_vqt_peakfiltered: null, x_vqt_afterglow: null, calmness: null,
            pitch_accuracy: null, pitch_deviation: null,
            smoothed_scene_calmness: scene.as_mut_ptr(),
            smoothed_tuning_grid_inaccuracy: tuning.as_mut_ptr(),
        };
        let rc = unsafe {
            pvqt_calc_batch_analysis(vqt.raw_handle(), self.handle, audio.as_ptr(), audio.len(), hop, t,
                                     frame_time.as_nanos() as u64, &out, std::ptr::null_mut(), std::ptr::null_mut())
        };
        assert!(rc == PVQT_OK, "pvqt_calc_batch_analysis failed: {}", last_error());
        let frames: Vec<FrameAnalysis> = (0..t).map(|f| {
            let c = (count[f] as usize).min(MAX_PEAKS);
            FrameAnalysis {
                peaks: indices[f * MAX_PEAKS..f * MAX_PEAKS + c].iter().map(|&i| i as usize).collect(),
                peaks_continuous: cont[f * MAX_PEAKS..f * MAX_PEAKS + c].to_vec(),
                smoothed_scene_calmness: scene[f],
                smoothed_tuning_grid_inaccuracy: tuning[f],
            }
        }).collect();
        let last = &frames[t - 1];
        self.peaks = last.peaks.iter().copied().collect();
        self.peaks_continuous = last.peaks_continuous.clone();
        self.smoothed_scene_calmness.y = last.smoothed_scene_calmness;
        self.smoothed_tuning_grid_inaccuracy.y = last.smoothed_tuning_grid_inaccuracy;
        frames
    }

    /// New: `db` holds T consecutive frames (frame-major).  The public fields hold the state after the last
    /// frame; the per-frame peak results of all T frames are returned.
    pub fn preprocess_batch(&mut self, db: &[f32], frame_time: Duration) -> Vec<FrameAnalysis> {
        let n = self.range.n_buckets();
        assert!(n > 0 && db.len() % n == 0, "db must hold whole frames of n_buckets values");
        let t = db.len() / n;
        if t == 0 {
            return Vec::new();
        }
        let mut count = vec![0u32; t];
        let mut indices = vec![0u32; t * MAX_PEAKS];
        let mut cont = vec![ContinuousPeak::default(); t * MAX_PEAKS];
        let mut vecs: Vec<Vec<f32>> = (0..6).map(|_| vec![0.0f32; t * n]).collect();
        let mut scene = vec![0.0f32; t];
        let mut tuning = vec![0.0f32; t];
        let out = PvqtAnalysisOutputs {
            max_peaks: MAX_PEAKS as u32,
            peak_count: count.as_mut_ptr(),
            peak_indices: indices.as_mut_ptr(),
            peaks_continuous: cont.as_mut_ptr(),
            x_vqt_smoothed: vecs[0].as_mut_ptr(),
            x_vqt_peakfiltered: vecs[1].as_mut_ptr(),
            x_vqt_afterglow: vecs[2].as_mut_ptr(),
            calmness: vecs[3].as_mut_ptr(),
            pitch_accuracy: vecs[4].as_mut_ptr(),
            pitch_deviation: vecs[5].as_mut_ptr(),
            smoothed_scene_calmness: scene.as_mut_ptr(),
            smoothed_tuning_grid_inaccuracy: tuning.as_mut_ptr(),
        };
        let rc = unsafe {
            pvqt_analysis_preprocess_batch(self.handle, db.as_ptr(), n, t, frame_time.as_nanos() as u64, &out)
        };
        assert!(rc != PVQT_BAD_LENGTH, "x_vqt.len() must equal range.n_buckets()");
        assert!(rc == PVQT_OK, "pvqt_analysis_preprocess_batch failed: {}", last_error());

        let mut frames = Vec::with_capacity(t);
        for f in 0..t {
            let c = (count[f] as usize).min(MAX_PEAKS);
            frames.push(FrameAnalysis {
                peaks: indices[f * MAX_PEAKS..f * MAX_PEAKS + c].iter().map(|&i| i as usize).collect(),
                peaks_continuous: cont[f * MAX_PEAKS..f * MAX_PEAKS + c].to_vec(),
                smoothed_scene_calmness: scene[f],
                smoothed_tuning_grid_inaccuracy: tuning[f],
            });
        }
        // public fields = state after the last frame
        let last = &frames[t - 1];
        let row = (t - 1) * n;
        let horizon = Some(self.params.vqt_smoothing_duration_base);
        for i in 0..n {
            self.x_vqt_smoothed[i] = EmaMeasurement { y: vecs[0][row + i], time_horizon: horizon };
            self.calmness[i] = EmaMeasurement { y: vecs[3][row + i], time_horizon: Some(self.params.note_calmness_smoothing_duration) };
        }
        self.x_vqt_peakfiltered.copy_from_slice(&vecs[1][row..row + n]);
        self.x_vqt_afterglow.copy_from_slice(&vecs[2][row..row + n]);
        self.pitch_accuracy.copy_from_slice(&vecs[4][row..row + n]);
        self.pitch_deviation.copy_from_slice(&vecs[5][row..row + n]);
        self.peaks = last.peaks.iter().copied().collect();
        self.peaks_continuous = last.peaks_continuous.clone();
        self.smoothed_scene_calmness = EmaMeasurement {
            y: last.smoothed_scene_calmness,
            time_horizon: Some(self.params.scene_calmness_smoothing_duration),
        };
        self.smoothed_tuning_grid_inaccuracy = EmaMeasurement {
            y: last.smoothed_tuning_grid_inaccuracy,
            time_horizon: Some(self.params.tuning_inaccuracy_smoothing_duration),
        };
        frames
    }
}
