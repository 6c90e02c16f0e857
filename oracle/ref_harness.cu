// oracle/ref_harness.cu — TEST / BENCH INFRASTRUCTURE ONLY.  Never linked into the product.
//
// Compiles the REFERENCE's own CUDA translation unit for this path,
//     /root/reference/src/xmipp/libraries/reconstruction_cuda/cuda_gpu_reconstruct_fourier.cpp
// from where it lies (it is #included below by path; no reference source is copied into this repository), with the four
// stand-in headers of oracle/ref_shims/ replacing xmippCore / cuFFTAdvisor, and puts a small C ABI around its public
// functions (cuda_gpu_reconstruct_fourier.h:67-156: createStreams, allocateWrapper, allocateTempVolumeGPU,
// copyConstants, copyBlobTable, processBufferGPU, ...).  Built by oracle/build_ref.py into oracle/_ref/librefkernel.so.
//
// Used for (a) the same-box GPU baseline: the reference's processBufferKernel timed on the B200 next to ours
// (bench.py `ref_gpu_kernel`), and (b) pinning the restatements to reference code: the temporary volume / weights this
// kernel produces from the same buffers are compared with oracle/recfourier_fast_oracle.cpp (--fast) and, after the
// restated host post-processing, with the exact path (tests/test_gpu_refkernel.py).
#include "reconstruction_cuda/cuda_gpu_reconstruct_fourier.cpp"      // the reference translation unit (via -I <reference>/libraries)

#include <cstring>
#include <new>
#include <vector>

// ---- pieces of reconstruction_cuda/gpu.cpp the translation unit links against (that file needs NVML and xmippCore)
void GPU::pinMemory(const void* h_mem, size_t bytes, unsigned int flags) {
    gpuErrchk(cudaHostRegister(const_cast<void*>(h_mem), bytes, flags));
}
void GPU::unpinMemory(const void* h_mem) { gpuErrchk(cudaHostUnregister(const_cast<void*>(h_mem))); }

namespace {

struct RefState {
    int S = 0, sx = 0, sy = 0, maxImages = 0, nSym = 1, blobOrder = 0;
    bool hasCTF = false;
    float blobRadius = 0, blobAlpha = 0;
    float* tempVolumeGPU = nullptr;
    float* tempWeightsGPU = nullptr;
    RecFourierBufferData* buffer = nullptr;
    bool created = false;
};
RefState g;

}  // namespace

// plain description of one traverse space (reconstruct_fourier_projection_traverse_space.h:37-59); converted to the
// reference's struct here so that its memory layout is the reference header's
struct refk_space {
    int minX, minY, minZ, maxX, maxY, maxZ;
    int dir;                       // 0 XY, 1 XZ, 2 YZ
    int projectionIndex;
    float maxDistanceSqr;
    float unitNormal[3], topOrigin[3], bottomOrigin[3];
    float transformInv[9];
    float weight;
};

extern "C" {

// One instance per process (the reference keeps its streams, wrappers and blob table in module globals).
int refk_create(int S, int fftSizeX, int fftSizeY, int maxImages, int nSym, int hasCTF, float blobRadius, float blobAlpha,
                int blobOrder, float iDeltaSqrt, float iw0, float oneOverBessiOrderAlpha, const float* blobTableSqrt) {
    if (g.created) return -1;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return -2;
    g.S = S; g.sx = fftSizeX; g.sy = fftSizeY; g.maxImages = maxImages; g.nSym = nSym; g.hasCTF = hasCTF != 0;
    g.blobRadius = blobRadius; g.blobAlpha = blobAlpha; g.blobOrder = blobOrder;
    // ProgRecFourierGPU::processImages, reconstruct_fourier_gpu.cpp:832-849
    allocateTempVolumeGPU(g.tempVolumeGPU, S + 1, sizeof(std::complex<float>));
    allocateTempVolumeGPU(g.tempWeightsGPU, S + 1, sizeof(float));
    createStreams(1);
    copyConstants(S, S, blobRadius, blobAlpha, iDeltaSqrt, iw0, oneOverBessiOrderAlpha);
    copyBlobTable(const_cast<float*>(blobTableSqrt), BLOB_TABLE_SIZE_SQRT);
    // threadRoutine, :432-438
    g.buffer = new RecFourierBufferData(true, hasCTF != 0, fftSizeX, fftSizeY, 0, maxImages, nSym);
    pinMemory(g.buffer);
    allocateWrapper(g.buffer, 0);
    g.created = true;
    return 0;
}

static int fill_buffer(const float* FFTs, const float* CTFs, const float* mods, const refk_space* spaces, int nImages) {
    if (!g.created || nImages < 0 || nImages > g.maxImages) return -1;
    RecFourierBufferData* b = g.buffer;
    const size_t px = (size_t)g.sx * g.sy;
    b->noOfImages = nImages;
    b->hasFFTs = true;
    std::memcpy(b->FFTs, FFTs, sizeof(float) * 2 * px * nImages);
    if (g.hasCTF) {
        std::memcpy(b->CTFs, CTFs, sizeof(float) * px * nImages);
        std::memcpy(b->modulators, mods, sizeof(float) * px * nImages);
    }
    for (int i = 0; i < nImages * g.nSym; ++i) {
        const refk_space& s = spaces[i];
        RecFourierProjectionTraverseSpace& d = b->spaces[i];
        d.minX = s.minX; d.minY = s.minY; d.minZ = s.minZ; d.maxX = s.maxX; d.maxY = s.maxY; d.maxZ = s.maxZ;
        d.maxDistanceSqr = s.maxDistanceSqr;
        d.dir = s.dir == 0 ? RecFourierProjectionTraverseSpace::XY : (s.dir == 1 ? RecFourierProjectionTraverseSpace::XZ : RecFourierProjectionTraverseSpace::YZ);
        d.unitNormal.x = s.unitNormal[0]; d.unitNormal.y = s.unitNormal[1]; d.unitNormal.z = s.unitNormal[2];
        d.topOrigin.x = s.topOrigin[0]; d.topOrigin.y = s.topOrigin[1]; d.topOrigin.z = s.topOrigin[2];
        d.bottomOrigin.x = s.bottomOrigin[0]; d.bottomOrigin.y = s.bottomOrigin[1]; d.bottomOrigin.z = s.bottomOrigin[2];
        d.projectionIndex = s.projectionIndex;
        for (int a = 0; a < 3; ++a)
            for (int c = 0; c < 3; ++c) d.transformInv[a][c] = s.transformInv[3 * a + c];
        d.weight = s.weight;
    }
    return 0;
}

// One processBufferGPU call of the reference (H2D of the buffer + processBufferKernel), then wait.
int refk_process(const float* FFTs, const float* CTFs, const float* mods, const refk_space* spaces, int nImages, int useFast,
                 float maxResolutionSqr) {
    if (fill_buffer(FFTs, CTFs, mods, spaces, nImages)) return -1;
    processBufferGPU(g.tempVolumeGPU, g.tempWeightsGPU, g.buffer, g.blobRadius, g.S, useFast != 0, maxResolutionSqr, 0, g.blobOrder, g.blobAlpha);
    waitForGPU();
    return 0;
}

// Time `reps` repetitions on the buffer last handed to refk_process (already resident on the device):
//   call_ms   = whole processBufferGPU calls (host-side copy of the buffer, H2D, kernel), wall clock around waitForGPU
//   kernel_ms = processBufferKernel alone, launched exactly as processBufferGPU_ does (:1135-1194), CUDA events on its stream
int refk_time(int useFast, float maxResolutionSqr, int reps, float* call_ms, float* kernel_ms) {
    if (!g.created || reps < 1) return -1;
    cudaStream_t stream = streams[0];
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    waitForGPU();
    cudaEventRecord(a, stream);
    for (int r = 0; r < reps; ++r)
        processBufferGPU(g.tempVolumeGPU, g.tempWeightsGPU, g.buffer, g.blobRadius, g.S, useFast != 0, maxResolutionSqr, 0, g.blobOrder, g.blobAlpha);
    cudaEventRecord(b, stream);
    cudaEventSynchronize(b);
    float t = 0;
    cudaEventElapsedTime(&t, a, b);
    *call_ms = t / reps;
    // kernel alone: same launch configuration as processBufferGPU_
    const int size2D = g.S + 1;
    const int imgCacheDim = (int)ceil(sqrt(2.f) * sqrt(3.f) * (BLOCK_DIM + 2 * g.blobRadius));
    dim3 dimBlock(BLOCK_DIM, BLOCK_DIM);
    dim3 dimGrid((unsigned)ceil(size2D / (float)dimBlock.x), (unsigned)ceil(size2D / (float)dimBlock.y), GRID_DIM_Z);
    const int sharedMemSize = SHARED_IMG ? (int)(imgCacheDim * imgCacheDim * sizeof(float2)) : 0;
    FRecBufferDataGPUWrapper* w = wrappers[0];
    const bool fastKaiser = g.blobOrder == 0 && g.blobAlpha <= 15.0f;
    if (g.blobOrder != 0) return -3;       // the timing entry covers the default blob (order 0), like the bench configuration
    cudaEventRecord(a, stream);
    for (int r = 0; r < reps; ++r) {
#define REFK_LAUNCH(FAST, CTF, FK) processBufferKernel<FAST, CTF, 0, FK><<<dimGrid, dimBlock, (FAST) ? 0 : sharedMemSize, stream>>>( \
        g.tempVolumeGPU, g.tempWeightsGPU, w->gpuCopy, devBlobTableSqrt, imgCacheDim)
        if (useFast) {
            if (g.hasCTF) { if (fastKaiser) REFK_LAUNCH(true, true, true); else REFK_LAUNCH(true, true, false); }
            else { if (fastKaiser) REFK_LAUNCH(true, false, true); else REFK_LAUNCH(true, false, false); }
        } else {
            if (g.hasCTF) { if (fastKaiser) REFK_LAUNCH(false, true, true); else REFK_LAUNCH(false, true, false); }
            else { if (fastKaiser) REFK_LAUNCH(false, false, true); else REFK_LAUNCH(false, false, false); }
        }
#undef REFK_LAUNCH
    }
    cudaEventRecord(b, stream);
    cudaEventSynchronize(b);
    cudaEventElapsedTime(&t, a, b);
    *kernel_ms = t / reps;
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// temporary volume (interleaved re, im) and weights, (S+1)^3 each, [z][y][x] (copyTempVolumes, :356-367, in one copy)
int refk_download(float* Vri, float* W) {
    if (!g.created) return -1;
    const size_t n = (size_t)(g.S + 1) * (g.S + 1) * (g.S + 1);
    waitForGPU();
    cudaMemcpy(Vri, g.tempVolumeGPU, sizeof(float) * 2 * n, cudaMemcpyDeviceToHost);
    cudaMemcpy(W, g.tempWeightsGPU, sizeof(float) * n, cudaMemcpyDeviceToHost);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

int refk_clear() {
    if (!g.created) return -1;
    const size_t n = (size_t)(g.S + 1) * (g.S + 1) * (g.S + 1);
    cudaMemset(g.tempVolumeGPU, 0, sizeof(float) * 2 * n);
    cudaMemset(g.tempWeightsGPU, 0, sizeof(float) * n);
    return 0;
}

void refk_destroy() {
    if (!g.created) return;
    waitForGPU();
    releaseWrapper(0);
    unpinMemory(g.buffer);
    delete g.buffer;
    g.buffer = nullptr;
    releaseBlobTable();
    deleteStreams(1);
    releaseTempVolumeGPU(g.tempVolumeGPU);
    releaseTempVolumeGPU(g.tempWeightsGPU);
    g = RefState();
}

// the tuning profile the reference was compiled with (reconstruct_fourier_defines.h:33-63)
void refk_profile(int* blockDim, int* sharedBlobTable, int* sharedImg, int* precomputeBlobVal, int* tile, int* gridDimZ) {
    *blockDim = BLOCK_DIM; *sharedBlobTable = SHARED_BLOB_TABLE; *sharedImg = SHARED_IMG;
    *precomputeBlobVal = PRECOMPUTE_BLOB_VAL; *tile = TILE; *gridDimZ = GRID_DIM_Z;
}

}  // extern "C"
