// oracle/ref_shims/cudaAsserts.h — TEST INFRASTRUCTURE ONLY (see core/xmipp_error.h in this directory).
// Stand-in for cuFFTAdvisor's cudaAsserts.h (a build-time download of the reference, cmake/fetch_cufftadvisor.cmake):
// reconstruction_cuda/cuda_asserts.h:31-32 wraps cuFFTAdvisor::gpuErrchk(code, file, line), which prints and aborts.
#pragma once
#include <cuda_runtime_api.h>
#include <cstdio>
#include <cstdlib>
namespace cuFFTAdvisor {
inline void gpuErrchk(cudaError_t code, const char* file, int line) {
    if (code != cudaSuccess) {
        fprintf(stderr, "GPUassert: %s %s %d\n", cudaGetErrorString(code), file, line);
        abort();
    }
}
template <typename T>
inline void gpuErrchkFFT(T code, const char* file, int line) {
    if ((int)code != 0) {
        fprintf(stderr, "GPUassert (cuFFT): %d %s %d\n", (int)code, file, line);
        abort();
    }
}
}  // namespace cuFFTAdvisor
