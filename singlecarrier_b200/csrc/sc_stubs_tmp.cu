#include "sc_kernels.h"
namespace sc {
cudaError_t launch_fir_batch(bool, long, float2*, float2*, long, int, cudaStream_t){return cudaErrorNotSupported;}
cudaError_t launch_search_batch(long, const float2*, long, int*, float*, cudaStream_t){return cudaErrorNotSupported;}
cudaError_t launch_track_window_batch(long, const float2*, long, const int*, const float*, int*, uint32_t, unsigned long long, sc_frame_result*, float*, cudaStream_t){return cudaErrorNotSupported;}
cudaError_t launch_fft_batch(long, int, int, const float2*, float2*, cudaStream_t){return cudaErrorNotSupported;}
}
