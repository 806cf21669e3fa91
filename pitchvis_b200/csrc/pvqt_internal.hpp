// pvqt_internal.hpp -- calls between the translation units of libpvqt.so that are not part of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "pvqt_analysis.h"

namespace pvqt_detail {

// K-analysis on streams [first_stream, first_stream + n_streams) of `a`, n_frames frames each.  d_db points at the
// first of those streams' frames ([n_streams][n_frames][NB], device); the launch writes frame slots
// out_frame_offset, out_frame_offset + 1, ... of every non-NULL member of d_out.  Asynchronous on `stream`.
int analysis_run_device(pvqt_analysis *a, const float *d_db, size_t first_stream, size_t n_streams, size_t n_frames,
                        uint64_t frame_time_ns, const pvqt_analysis_outputs *d_out, size_t out_frame_offset,
                        cudaStream_t stream);
int analysis_device(const pvqt_analysis *a);

// Device mirrors of the host result buffers `host` names (NULL members stay NULL), S * T frames; the buffers are
// owned by `a` and reused across calls.  download() copies every mirrored member back (asynchronous on `stream`).
int analysis_outputs_reserve(pvqt_analysis *a, const pvqt_analysis_outputs *host, size_t frames, pvqt_analysis_outputs *dev,
                             cudaStream_t stream);
int analysis_outputs_download(pvqt_analysis *a, const pvqt_analysis_outputs *host, const pvqt_analysis_outputs *dev,
                              size_t frames, cudaStream_t stream, size_t *bytes);

// The result arrays of pvqt_analysis_outputs by index (declaration order): the member, and its bytes per frame.
constexpr int kAnalysisOutputs = 11;
void *&analysis_output_member(pvqt_analysis_outputs &o, int i);
size_t analysis_output_bytes_per_frame(const pvqt_analysis *a, int i, size_t max_peaks);
// Changes whenever something a captured launch of K-analysis has baked in does: the parameters, the result mirrors.
uint64_t analysis_generation(const pvqt_analysis *a);

}  // namespace pvqt_detail
