// agc_kernels.cu -- K-agc: dagc_fork::MonoAgc for many streams, and its C ABI (include/pvqt_agc.h).
//
// The gain update gain *= max(1 + d (1 - (x gain)^2 / rms), d) is a nonlinear recurrence per sample: there is no
// scan formulation, so a stream is one thread walking its samples in order; parallelism is across streams (config 3:
// 4096 of them).  Every f32 operation is an explicit round-to-nearest intrinsic in the order of the Rust source
// (dagc_fork/src/lib.rs:76-86), so the result is bit-identical to the CPU oracle.
#include <cmath>
#include <string>

#include <cuda_runtime.h>

#include "last_error.hpp"
#include "pvqt_agc.h"

namespace {

struct AgcParams {
    float *audio;
    unsigned long long stream_stride, n_samples, chunk;
    unsigned n_streams;
    float desired_output_rms, distortion_factor, silence_threshold;
    int use_flag, frozen_flag;
    float *gain;
};

__global__ void __launch_bounds__(32) agc_kernel(const AgcParams P)
{
    const unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= P.n_streams) return;
    float *x = P.audio + (size_t)s * P.stream_stride;
    float gain = P.gain[s];
    const float rms = P.desired_output_rms, d = P.distortion_factor;
    for (unsigned long long b = 0; b < P.n_samples; b += P.chunk) {
        const unsigned long long m = min(P.chunk, P.n_samples - b);
        bool frozen = P.frozen_flag != 0;
        if (!P.use_flag) {
            float sq = 0.0f;  // data.iter().map(|x| x.powi(2)).sum::<f32>()
            for (unsigned long long i = 0; i < m; ++i) {
                const float v = x[b + i];
                sq = __fadd_rn(sq, __fmul_rn(v, v));
            }
            frozen = sq < P.silence_threshold;
        }
        for (unsigned long long i = 0; i < m; ++i) {
            const float v = __fmul_rn(x[b + i], gain);                  // *x *= self.gain
            x[b + i] = v;
            if (!frozen) {
                const float y = __fdiv_rn(__fmul_rn(v, v), rms);        // x.powi(2) / desired_output_rms
                float g = __fadd_rn(1.0f, __fmul_rn(d, __fsub_rn(1.0f, y)));
                g = fmaxf(g, d);                                        // g.max(distortion_factor)
                gain = __fmul_rn(gain, g);
            }
        }
    }
    P.gain[s] = gain;
}

int fail(int st, const std::string &m)
{
    pvqt_detail::set_last_error(m);
    return st;
}

int cuda_fail(cudaError_t e, const char *what)
{
    pvqt_detail::set_last_error(std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
    return PVQT_CUDA_ERROR;
}

#define AGC_CUDA(call)                                          \
    do {                                                        \
        cudaError_t _e = (call);                                \
        if (_e != cudaSuccess) return cuda_fail(_e, #call);     \
    } while (0)

}  // namespace

struct pvqt_agc {
    float desired_output_rms = 0, distortion_factor = 0;
    size_t n_streams = 0;
    int device = 0;
    int frozen = 0;
    cudaStream_t stream = nullptr;
    float *d_gain = nullptr;
    float *d_audio = nullptr;
    size_t d_audio_bytes = 0;
};

extern "C" {

int pvqt_agc_create(float desired_output_rms, float distortion_factor, size_t n_streams, int device, pvqt_agc **out)
{
    if (!out || n_streams == 0) return fail(PVQT_INVALID_ARGUMENT, "null output or no streams");
    *out = nullptr;
    if (!(desired_output_rms > 0.0f && std::isfinite(desired_output_rms)))   // Error::InvalidDesiredOutputRms
        return fail(PVQT_INVALID_ARGUMENT, "`desired_output_rms` must be a finite positive number, but got " +
                                               std::to_string(desired_output_rms));
    if (!(distortion_factor >= 0.0f && distortion_factor <= 1.0f))          // Error::InvalidDistortionFactor
        return fail(PVQT_INVALID_ARGUMENT, "`distortion_factor` must be a number within `0.0 ..= 1.0`, but got " +
                                               std::to_string(distortion_factor));
    int n_dev = 0;
    AGC_CUDA(cudaGetDeviceCount(&n_dev));
    if (device < 0 || device >= n_dev) return fail(PVQT_INVALID_ARGUMENT, "device out of range");
    AGC_CUDA(cudaSetDevice(device));
    pvqt_agc *a = new pvqt_agc();
    a->desired_output_rms = desired_output_rms;
    a->distortion_factor = distortion_factor;
    a->n_streams = n_streams;
    a->device = device;
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&a->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaMalloc(&a->d_gain, n_streams * sizeof(float))) != cudaSuccess) {
        pvqt_agc_destroy(a);
        return cuda_fail(e, "create AGC state");
    }
    std::string ones(n_streams * sizeof(float), '\0');
    float *h = reinterpret_cast<float *>(&ones[0]);
    for (size_t i = 0; i < n_streams; ++i) h[i] = 1.0f;  // gain: 1.0 (lib.rs:50)
    if ((e = cudaMemcpy(a->d_gain, h, n_streams * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess) {
        pvqt_agc_destroy(a);
        return cuda_fail(e, "initialise AGC gains");
    }
    *out = a;
    return PVQT_OK;
}

void pvqt_agc_destroy(pvqt_agc *a)
{
    if (!a) return;
    cudaSetDevice(a->device);
    if (a->stream) { cudaStreamSynchronize(a->stream); cudaStreamDestroy(a->stream); }
    if (a->d_gain) cudaFree(a->d_gain);
    if (a->d_audio) cudaFree(a->d_audio);
    delete a;
}

size_t pvqt_agc_n_streams(const pvqt_agc *a) { return a ? a->n_streams : 0; }

int pvqt_agc_gains(pvqt_agc *a, float *out)
{
    if (!a || !out) return fail(PVQT_INVALID_ARGUMENT, "null argument");
    AGC_CUDA(cudaSetDevice(a->device));
    AGC_CUDA(cudaMemcpyAsync(out, a->d_gain, a->n_streams * sizeof(float), cudaMemcpyDeviceToHost, a->stream));
    AGC_CUDA(cudaStreamSynchronize(a->stream));
    return PVQT_OK;
}

int pvqt_agc_freeze_gain(pvqt_agc *a, int freeze)
{
    if (!a) return fail(PVQT_INVALID_ARGUMENT, "null handle");
    a->frozen = freeze != 0;
    return PVQT_OK;
}

int pvqt_agc_process_device(pvqt_agc *a, float *d_audio, size_t stream_stride, size_t n_samples, size_t chunk,
                            float silence_threshold, void *cuda_stream)
{
    if (!a || !d_audio) return fail(PVQT_INVALID_ARGUMENT, "null argument");
    if (n_samples == 0) return PVQT_OK;
    if (a->n_streams > 1 && stream_stride < n_samples) return fail(PVQT_INVALID_ARGUMENT, "stream_stride must be >= n_samples");
    AGC_CUDA(cudaSetDevice(a->device));
    AgcParams p{};
    p.audio = d_audio;
    p.stream_stride = stream_stride;
    p.n_samples = n_samples;
    p.chunk = chunk == 0 ? n_samples : chunk;
    p.n_streams = (unsigned)a->n_streams;
    p.desired_output_rms = a->desired_output_rms;
    p.distortion_factor = a->distortion_factor;
    p.silence_threshold = silence_threshold;
    p.use_flag = std::isnan(silence_threshold) ? 1 : 0;
    p.frozen_flag = a->frozen;
    p.gain = a->d_gain;
    cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : a->stream;
    agc_kernel<<<(unsigned)((a->n_streams + 31) / 32), 32, 0, st>>>(p);
    AGC_CUDA(cudaGetLastError());
    return PVQT_OK;
}

int pvqt_agc_process(pvqt_agc *a, float *audio, size_t stream_stride, size_t n_samples, size_t chunk,
                     float silence_threshold)
{
    if (!a || !audio) return fail(PVQT_INVALID_ARGUMENT, "null argument");
    if (n_samples == 0) return PVQT_OK;
    if (a->n_streams > 1 && stream_stride < n_samples) return fail(PVQT_INVALID_ARGUMENT, "stream_stride must be >= n_samples");
    AGC_CUDA(cudaSetDevice(a->device));
    const size_t bytes = a->n_streams * n_samples * sizeof(float);
    if (bytes > a->d_audio_bytes) {
        if (a->d_audio) cudaFree(a->d_audio);
        a->d_audio = nullptr;
        a->d_audio_bytes = 0;
        AGC_CUDA(cudaMalloc(&a->d_audio, bytes));
        a->d_audio_bytes = bytes;
    }
    AGC_CUDA(cudaMemcpy2DAsync(a->d_audio, n_samples * sizeof(float), audio, stream_stride * sizeof(float),
                               n_samples * sizeof(float), a->n_streams, cudaMemcpyHostToDevice, a->stream));
    int rc = pvqt_agc_process_device(a, a->d_audio, n_samples, n_samples, chunk, silence_threshold, a->stream);
    if (rc != PVQT_OK) return rc;
    AGC_CUDA(cudaMemcpy2DAsync(audio, stream_stride * sizeof(float), a->d_audio, n_samples * sizeof(float),
                               n_samples * sizeof(float), a->n_streams, cudaMemcpyDeviceToHost, a->stream));
    AGC_CUDA(cudaStreamSynchronize(a->stream));
    return PVQT_OK;
}

int pvqt_agc_synchronize(pvqt_agc *a)
{
    if (!a) return fail(PVQT_INVALID_ARGUMENT, "null handle");
    AGC_CUDA(cudaSetDevice(a->device));
    AGC_CUDA(cudaStreamSynchronize(a->stream));
    return PVQT_OK;
}

}  // extern "C"
