// oracle/ref_shims/core/xmipp_error.h — TEST INFRASTRUCTURE ONLY.
// Stand-in for xmippCore's core/xmipp_error.h (xmippCore is not in the reference tree) so that the reference's
// own CUDA translation unit (libraries/reconstruction_cuda/cuda_gpu_reconstruct_fourier.cpp) compiles where it lies.
// Only what that file and reconstruction_cuda/gpu.h use: REPORT_ERROR with an error code and a message.
#pragma once
#include <stdexcept>
#include <string>
enum ErrorType { ERR_VALUE_INCORRECT = 1, ERR_LOGIC_ERROR, ERR_GPU_MEMORY, ERR_MEM_NOTENOUGH, ERR_ARG_INCORRECT, ERR_NOT_IMPLEMENTED, ERR_UNCLASSIFIED };
struct XmippError : public std::runtime_error {
    ErrorType code;
    XmippError(ErrorType c, const std::string& msg) : std::runtime_error(msg), code(c) {}
};
#define REPORT_ERROR(code, msg) throw XmippError((code), std::string(msg))
