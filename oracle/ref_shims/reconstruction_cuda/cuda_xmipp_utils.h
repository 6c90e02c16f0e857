// oracle/ref_shims/reconstruction_cuda/cuda_xmipp_utils.h — TEST INFRASTRUCTURE ONLY.
// Shadows the reference's reconstruction_cuda/cuda_xmipp_utils.h (whose implementation needs cuFFTAdvisor and
// xmippCore).  cuda_gpu_reconstruct_fourier.cpp uses it in convertImages() only (the --fftOnGPU input path,
// :1008-1040): GpuMultidimArrayAtGpu<T>::fft + mycufftHandle.  The harness always hands over ready transforms
// (hasFFTs = true), so these stand-ins only have to compile; calling fft() aborts.
#pragma once
#include <complex>
#include <cstdio>
#include <cstdlib>
struct mycufftHandle {
    void* ptr = nullptr;
    void clear() {}
};
template <typename T>
struct GpuMultidimArrayAtGpu {
    size_t Xdim = 0, Ydim = 0, Zdim = 0, Ndim = 0;
    T* d_data = nullptr;
    GpuMultidimArrayAtGpu() = default;
    GpuMultidimArrayAtGpu(size_t x, size_t y, size_t z, size_t n, T* d) : Xdim(x), Ydim(y), Zdim(z), Ndim(n), d_data(d) {}
    template <typename U>
    void fft(GpuMultidimArrayAtGpu<U>&, mycufftHandle&) {
        fprintf(stderr, "ref harness: the --fftOnGPU input path is not available (needs cuFFTAdvisor)\n");
        abort();
    }
};
