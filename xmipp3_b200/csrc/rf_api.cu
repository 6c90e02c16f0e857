// rf_api.cu — handle management and the extern "C" boundary declared in
// include/recfourier_b200.h.  Host orchestration only; the arithmetic lives in
// rf_kernels.cuh (device) and rf_host.hpp (double-precision precompute).
//
// Stream structure per handle: `copy` stream (H2D of raw particles, double buffered) and
// `compute` stream (K1a -> cuFFT R2C -> K1b' -> K2' x3 classes -> K2e' -> K2r per chunk), linked by
// events, so the PCIe transfer of chunk c+1 overlaps the kernels of chunk c.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cufft.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/recfourier_b200.h"
#include "rf_host.hpp"
#include "rf_kernels.cuh"
#include "rf_sticks.cuh"
#include "rf_fft.cuh"
#include "rf_fast.cuh"

#if __has_include(<nccl.h>)
#include <nccl.h>
#define RFB200_HAVE_NCCL_H 1
#else
#define RFB200_HAVE_NCCL_H 0
#endif

using namespace rfb200;

namespace {

std::mutex g_errMutex;
std::string g_createError;

struct Stage { enum { H2D, PAD, FFT2D, SLICE, GATHER, EDGE, FINALIZE, REDUCE, COUNT }; };

struct EvPair { int stage; cudaEvent_t a, b; };

struct ParamSlot {      // pinned host staging for one chunk's parameters
    ImgParams* img = nullptr;
    CtfConsts* ctf = nullptr;
    PlaneD* planesD = nullptr;
    int* planeImg = nullptr;
    // stick gather: class-sorted, (a,b,d)-permuted copies, one region per launch sub-range
    PlaneD* planesDp = nullptr;
    PlaneS* planesS = nullptr;
    float* soaP = nullptr;
    int* imgPlane0 = nullptr;
    FastSpace* fast = nullptr;     // --fast: traverse spaces of the chunk
    cudaEvent_t done = nullptr;
    bool used = false;
    struct Launch { int cls, start, count; };
    std::vector<Launch> launches;  // gather launches of the chunk: planesS / planesDp [start, start + count) are of class cls
};

#if RFB200_HAVE_NCCL_H
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi g_nccl;
std::once_flag g_ncclOnce;
void load_nccl() {
    const char* names[] = {"libnccl.so.2", "libnccl.so", "/usr/lib/x86_64-linux-gnu/libnccl.so.2"};
    for (const char* n : names) {
        g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) return;
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(g_nccl.lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(g_nccl.lib, "ncclCommInitRank");
    g_nccl.Reduce = (decltype(g_nccl.Reduce))dlsym(g_nccl.lib, "ncclReduce");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(g_nccl.lib, "ncclAllReduce");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(g_nccl.lib, "ncclCommDestroy");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(g_nccl.lib, "ncclGetErrorString");
    g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.Reduce && g_nccl.CommDestroy;
}
#endif

}  // namespace

struct rfb200_handle_s {
    rfb200_config cfg{};
    std::vector<double> sym;        // (n_sym+1) x 9, identity first (R_repository, RF.cpp:272-286)
    int nSymTot = 1;
    Geometry geo{};
    host::Tables tables;
    std::vector<int> jmax;
    int iLo = 0, iHi = 0;
    int chunkImages = 0;            // images per preprocessing chunk
    std::string err;

    cudaStream_t compute = nullptr, copy = nullptr;
    // static device data
    float* dBlobTable = nullptr;
    int* dJmax = nullptr;
    EdgeItem* dEdge = nullptr;
    int32_t* dEdgeGroups = nullptr;
    int nEdge = 0, nEdgeGroups = 0;
    float* dG = nullptr;
    // accumulators (blocked layout)
    float2* dVb = nullptr;
    float* dWb = nullptr;
    float* dWb2 = nullptr;          // only with use_ctf && n_iter_weight > 1
    float2* dVsaved = nullptr;      // half-set snapshot (--prepare_fsc)
    float* dWsaved = nullptr;
    float* dW2saved = nullptr;
    int64_t nBlocked = 0;
    // per-chunk buffers
    float* dRaw[2] = {nullptr, nullptr};
    cudaEvent_t evH2D[2] = {nullptr, nullptr}, evRawFree[2] = {nullptr, nullptr};
    bool rawBusy[2] = {false, false};
    float* dPad = nullptr;
    float* dCoef = nullptr;          // B-spline coefficients of the chunk (allocated on first fractional shift)
    float2* dFft = nullptr;
    ImgParams* dImg = nullptr;
    CtfConsts* dCtf = nullptr;
    PlaneD* dPlanesD = nullptr;
    int* dPlaneImg = nullptr;
    ParamSlot slots[2];
    int slotIdx = 0;
    // Device copies of a chunk's parameter tables, one set per slot.  dImg, dCtf, dPlanesD, dPlaneImg, dPlanesSoAp, dImgPlane0 and
    // dFastSpaces above are the CURRENT set (the launches of a chunk capture the pointers); the next chunk's set is filled
    // by DMA on the copy stream while this chunk computes, so no parameter traffic sits between two chunks on the compute
    // stream (it used to: six small kernels reading pinned memory through a PCIe link busy with the next image batch).
    struct DevParams {
        ImgParams* img = nullptr; CtfConsts* ctf = nullptr; PlaneD* planesD = nullptr; int* planeImg = nullptr;
        float* soaP = nullptr; int* imgPlane0 = nullptr; FastSpace* fast = nullptr;
        size_t nImg = 0, nCtf = 0, nPlanesD = 0, nPlaneImg = 0, nSoaP = 0, nImgPlane0 = 0, nFast = 0;   // bytes
        cudaEvent_t up = nullptr;
    } dpar[2];
    std::map<int, cufftHandle> plans2d;
    // finalize
    cufftHandle plan3d = 0;
    bool havePlan3d = false;
    float2* dNorm = nullptr;
    float* dVol = nullptr;
    float* dOut = nullptr;
    // timings
    std::vector<EvPair> pending;
    std::vector<cudaEvent_t> evPool;
    double ms[Stage::COUNT] = {0};
    int64_t nImages = 0, nPlanes = 0, nGatherLaunches = 0, nKernelLaunches = 0;
    int lastChunkImages = 0;
#if RFB200_HAVE_NCCL_H
    ncclComm_t comm = nullptr;
#endif
    int nRanks = 1, rank = 0;
    // peer-memory reduce (rfb200_reduce_p2p): accumulators of the other ranks' handles, opened through CUDA IPC
    struct Peer { float2* V = nullptr; float* W = nullptr; float* W2 = nullptr; bool open = false; };
    std::vector<Peer> peers;
    float* dBarrier = nullptr;      // 4 bytes for the stream-ordered rank barrier (a tiny all-reduce)
    // ---- stick gather
    float4* dSlices2 = nullptr;     // per image side x pitch entries (pixel(i,j), pixel(i,j+1))
    float2* dCol02 = nullptr;
    float* dDamped = nullptr;       // per image (2R+1) x (R+1) weights of the CTF-damped (flagged) pixels (use_ctf only)
    float* dDamped2 = nullptr;      // their un-modulated weights (use_ctf && n_iter_weight > 1)
    uint32_t* dDampedMask = nullptr;   // per image (2R+1) x ceil((R+1)/32) words: which pixels are flagged
    unsigned long long* dD = nullptr;   // blocked volume of 2^32 fixed-point damped-pixel weights (use_ctf only)
    unsigned long long* dD2 = nullptr;  // same for the un-modulated weights
    bool dampedDirty = false;
    int32_t* dRimTab = nullptr;
    StickUnit* dUnits[3] = {nullptr, nullptr, nullptr};
    int nUnits[3] = {0, 0, 0};
    int* dStickCounters = nullptr;  // 3 ints
    PlaneD* dPlanesDp = nullptr;
    float* dPlanesSoAp = nullptr;
    PlaneS* dPlanesSStage = nullptr;
    int* dImgPlane0 = nullptr;
    int stickGrid = 0;
    CUtensorMap sliceMap;           // TMA descriptor of dSlices2 (L2 prefetch of slice patches); zeroed when unavailable
    bool haveSliceMap = false;
    // ---- fused FFT chain (power-of-two padded sizes, integer shifts); cuFFT is the general path
    bool fusedFft = false;
    float2* dTwiddle = nullptr;
    cudaEvent_t swStart = nullptr, swStop = nullptr;
    bool swStarted = false;
    // ---- --fast (rf_fast.cuh): dVb / dWb are then the (S+1)^3 temporary volume and weights
    bool fast = false;
    FastGeo fgeo{};
    float4* dFastPix = nullptr;     // per image sy x sx folded pixels
    float4* dFastAcc = nullptr;     // interleaved scratch accumulators of the insertion, flushed into dVb / dWb on demand
    bool fastDirty = false;
    FastSpace* dFastSpaces = nullptr;
    float2* dFastVh = nullptr;      // finalisation: mirrored half space, then its blob convolution
    float2* dFastVc = nullptr;
    float* dFastWh = nullptr;
    float* dFastWc = nullptr;
    double* dSum = nullptr;         // 1024 partials + 1 result
    double* hSum = nullptr;         // pinned
    cudaEvent_t evSum = nullptr;    // completion of the last rfb200_weight_sum_begin
    bool sumPending = false;
};

namespace {

#define RF_CUDA(h, call)                                                                      \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            char buf_[512];                                                                   \
            snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            (h)->err = buf_;                                                                  \
            return RFB200_ERR_CUDA;                                                           \
        }                                                                                     \
    } while (0)

#define RF_CUFFT(h, call)                                                                     \
    do {                                                                                      \
        cufftResult r_ = (call);                                                              \
        if (r_ != CUFFT_SUCCESS) {                                                            \
            char buf_[512];                                                                   \
            snprintf(buf_, sizeof buf_, "%s failed: cufft error %d (%s:%d)", #call, (int)r_, __FILE__, __LINE__); \
            (h)->err = buf_;                                                                  \
            return RFB200_ERR_CUDA;                                                           \
        }                                                                                     \
    } while (0)

int fail(rfb200_handle h, int code, const std::string& msg) {
    h->err = msg;
    return code;
}

cudaEvent_t get_event(rfb200_handle h) {
    if (!h->evPool.empty()) {
        cudaEvent_t e = h->evPool.back();
        h->evPool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;            // stage timing is best effort: a pair with a missing event is skipped
    }
    return e;
}
struct StageTimer {   // records an event pair around a stage on a stream
    rfb200_handle h; int stage; cudaStream_t s; cudaEvent_t a, b;
    StageTimer(rfb200_handle h_, int stage_, cudaStream_t s_) : h(h_), stage(stage_), s(s_) {
        a = get_event(h); b = get_event(h);
        if (a && b) cudaEventRecord(a, s);
    }
    ~StageTimer() {
        if (a && b) {
            cudaEventRecord(b, s);
            h->pending.push_back({stage, a, b});
        } else {
            if (a) h->evPool.push_back(a);
            if (b) h->evPool.push_back(b);
        }
    }
};
// Recycle the event pairs whose stage has completed (no waiting): called at the start of every insert so that a long
// run keeps a bounded number of live events instead of ~14 per chunk until the next sync.
void drain_timings(rfb200_handle h) {
    size_t keep = 0;
    for (size_t i = 0; i < h->pending.size(); ++i) {
        EvPair& p = h->pending[i];
        if (cudaEventQuery(p.b) == cudaSuccess) {
            float t = 0;
            if (cudaEventElapsedTime(&t, p.a, p.b) == cudaSuccess) h->ms[p.stage] += t;
            h->evPool.push_back(p.a);
            h->evPool.push_back(p.b);
        } else {
            cudaGetLastError();     // cudaErrorNotReady is not an error
            h->pending[keep++] = p;
        }
    }
    h->pending.resize(keep);
}
void resolve_timings(rfb200_handle h) {
    static const bool trace = getenv("RFB200_TRACE") != nullptr;     // developer aid: stage timeline on stderr
    static const char* names[] = {"h2d", "pad", "fft2d", "slice", "gather", "edge", "finalize", "reduce"};
    for (auto& p : h->pending) {
        float t = 0;
        if (cudaEventSynchronize(p.b) == cudaSuccess && cudaEventElapsedTime(&t, p.a, p.b) == cudaSuccess) h->ms[p.stage] += t;
        if (trace && h->swStarted) {
            float t0 = 0, t1 = 0;
            if (cudaEventElapsedTime(&t0, h->swStart, p.a) == cudaSuccess && cudaEventElapsedTime(&t1, h->swStart, p.b) == cudaSuccess)
                fprintf(stderr, "[rfb200 trace] %-8s %9.3f -> %9.3f ms\n", names[p.stage], t0, t1);
            else
                cudaGetLastError();   // events recorded before the stopwatch started
        }
        h->evPool.push_back(p.a);
        h->evPool.push_back(p.b);
    }
    h->pending.clear();
}

// Parameter tables of a chunk: DMA from the slot's pinned staging arrays into the slot's device set on the COPY stream (the
// slot's previous user has finished: upload_chunk_params waited for its `done` event), one event, and the compute stream
// waits for that event before the chunk's first kernel.
int fetch_params(rfb200_handle h, void* dst, const void* srcPinned, size_t bytes) {
    if (!bytes) return RFB200_OK;
    RF_CUDA(h, cudaMemcpyAsync(dst, srcPinned, bytes, cudaMemcpyHostToDevice, h->copy));
    return RFB200_OK;
}
void use_param_set(rfb200_handle h, int si) {
    const rfb200_handle_s::DevParams& d = h->dpar[si];
    h->dImg = d.img; h->dCtf = d.ctf; h->dPlanesD = d.planesD; h->dPlaneImg = d.planeImg;
    h->dPlanesSoAp = d.soaP; h->dImgPlane0 = d.imgPlane0; h->dFastSpaces = d.fast;
}
int params_uploaded(rfb200_handle h, int si) {
    RF_CUDA(h, cudaEventRecord(h->dpar[si].up, h->copy));
    RF_CUDA(h, cudaStreamWaitEvent(h->compute, h->dpar[si].up, 0));
    return RFB200_OK;
}
// second device set with the sizes of the first (the cudaMalloc calls of do_create fill set 0 through the h->d* members)
int make_param_sets(rfb200_handle h, size_t nImg, size_t nCtf, size_t nPlanesD, size_t nPlaneImg, size_t nSoaP, size_t nImgPlane0, size_t nFast) {
    rfb200_handle_s::DevParams& a = h->dpar[0];
    a.img = h->dImg; a.ctf = h->dCtf; a.planesD = h->dPlanesD; a.planeImg = h->dPlaneImg; a.soaP = h->dPlanesSoAp;
    a.imgPlane0 = h->dImgPlane0; a.fast = h->dFastSpaces;
    rfb200_handle_s::DevParams& b = h->dpar[1];
    if (nImg) RF_CUDA(h, cudaMalloc(&b.img, nImg));
    if (nCtf) RF_CUDA(h, cudaMalloc(&b.ctf, nCtf));
    if (nPlanesD) RF_CUDA(h, cudaMalloc(&b.planesD, nPlanesD));
    if (nPlaneImg) RF_CUDA(h, cudaMalloc(&b.planeImg, nPlaneImg));
    if (nSoaP) RF_CUDA(h, cudaMalloc(&b.soaP, nSoaP));
    if (nImgPlane0) RF_CUDA(h, cudaMalloc(&b.imgPlane0, nImgPlane0));
    if (nFast) RF_CUDA(h, cudaMalloc(&b.fast, nFast));
    for (int i = 0; i < 2; ++i) RF_CUDA(h, cudaEventCreateWithFlags(&h->dpar[i].up, cudaEventDisableTiming));
    return RFB200_OK;
}

template <int K, int CLS>
int launch_sticks_kc(rfb200_handle h, const StickLaunch& a, int grid) {
    // only a CTF can damp (flag) a pixel: without it the gather runs the variant without the per-candidate flag test
    if (h->cfg.use_ctf) {
        RF_CUDA(h, cudaFuncSetAttribute(k_gather_sticks<K, CLS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStickSmem));
        k_gather_sticks<K, CLS, true><<<grid, kStickThreads, kStickSmem, h->compute>>>(a);
    } else {
        RF_CUDA(h, cudaFuncSetAttribute(k_gather_sticks<K, CLS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStickSmem));
        k_gather_sticks<K, CLS, false><<<grid, kStickThreads, kStickSmem, h->compute>>>(a);
    }
    RF_CUDA(h, cudaGetLastError());
    return RFB200_OK;
}
template <int K>
int launch_sticks_k(rfb200_handle h, const StickLaunch& a, int grid) {
    switch (a.a.cls) {
        case 0: return launch_sticks_kc<K, 0>(h, a, grid);
        case 1: return launch_sticks_kc<K, 1>(h, a, grid);
        default: return launch_sticks_kc<K, 2>(h, a, grid);
    }
}
int launch_sticks(rfb200_handle h, const StickLaunch& a, int grid) {
    switch (h->geo.K) {
        case 1: return launch_sticks_k<1>(h, a, grid);
        case 2: return launch_sticks_k<2>(h, a, grid);
        case 3: return launch_sticks_k<3>(h, a, grid);
        case 4: return launch_sticks_k<4>(h, a, grid);
        case 5: return launch_sticks_k<5>(h, a, grid);
        case 6: return launch_sticks_k<6>(h, a, grid);
        case 7: return launch_sticks_k<7>(h, a, grid);
        case 8: return launch_sticks_k<8>(h, a, grid);
    }
    return fail(h, RFB200_ERR_ARG, "blob radius / padding ratio gives an unsupported interpolation window");
}

template <int P, int kCtf>
int launch_k1c(rfb200_handle h, const FftColsArgs& ca, int n, int threadsC, size_t smemC) {
    constexpr int NC = kColsPerCta<P>;
    RF_CUDA(h, cudaFuncSetAttribute(k_fft_cols_slices<P, kCtf>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemC));
    RF_CUDA(h, cudaFuncSetAttribute(k_fft_cols_slices<P, kCtf>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    k_fft_cols_slices<P, kCtf><<<dim3((h->geo.R + 2 + NC - 1) / NC, n), threadsC, smemC, h->compute>>>(ca);
    RF_CUDA(h, cudaGetLastError());
    return RFB200_OK;
}
// ctfMode: 0 = no CTF, 1 = every image of the chunk takes the fixed-point phase path, 2 = general (envelope / phase plate)
template <int P>
int launch_fused_fft_p(rfb200_handle h, const FftRowsArgs& ra, const FftColsArgs& ca, int n, int ctfMode) {
    const Geometry& g = h->geo;
    const int threads = kFftSeqs * P / 8;
    const size_t smem = sizeof(float2) * (P + kFftSeqs * kFftBuf<P>);
    constexpr int NC = kColsPerCta<P>;
    const int threadsC = (NC + kK1cHalo) * P / 8;
    const size_t smemC = sizeof(float2) * (P + (NC + kK1cHalo) * kFftBuf<P> + (P + 2)) + sizeof(int) * (P + 2);   // twiddles, the sequences, halo values, cut-off table
    RF_CUDA(h, cudaFuncSetAttribute(k_fft_rows<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RF_CUDA(h, cudaFuncSetAttribute(k_fft_rows<P>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    {
        StageTimer t(h, Stage::FFT2D, h->compute);
        k_fft_rows<P><<<dim3((g.N + 2 * kFftSeqs - 1) / (2 * kFftSeqs), n), threads, smem, h->compute>>>(ra);
        RF_CUDA(h, cudaGetLastError());
    }
    {
        StageTimer t(h, Stage::SLICE, h->compute);
        if (h->dDampedMask)
            RF_CUDA(h, cudaMemsetAsync(h->dDampedMask, 0, sizeof(uint32_t) * (size_t)n * (2 * g.R + 1) * ((g.R + 1 + 31) / 32), h->compute));
        const int rck = ctfMode == 0 ? launch_k1c<P, 0>(h, ca, n, threadsC, smemC)
                      : ctfMode == 1 ? launch_k1c<P, 1>(h, ca, n, threadsC, smemC) : launch_k1c<P, 2>(h, ca, n, threadsC, smemC);
        if (rck) return rck;
    }
    h->nKernelLaunches += 2;
    return RFB200_OK;
}
int launch_fused_fft(rfb200_handle h, const FftRowsArgs& ra, const FftColsArgs& ca, int n, int ctfMode) {
    switch (h->geo.P) {
        case 64: return launch_fused_fft_p<64>(h, ra, ca, n, ctfMode);
        case 128: return launch_fused_fft_p<128>(h, ra, ca, n, ctfMode);
        case 256: return launch_fused_fft_p<256>(h, ra, ca, n, ctfMode);
        case 512: return launch_fused_fft_p<512>(h, ra, ca, n, ctfMode);
        case 1024: return launch_fused_fft_p<1024>(h, ra, ca, n, ctfMode);
    }
    return fail(h, RFB200_ERR_STATE, "fused FFT chain selected for an unsupported padded size");
}

// W += weights of the CTF-damped pixels; must run before W is read, reduced or copied
int flush_deficit(rfb200_handle h) {
    if (h->fast && h->fastDirty) {      // --fast: fold the scratch accumulators into V and W
        k_fast_flush<<<2048, 256, 0, h->compute>>>(h->dFastAcc, h->dVb, h->dWb, h->nBlocked);
        RF_CUDA(h, cudaGetLastError());
        h->nKernelLaunches += 1;
        h->fastDirty = false;
    }
    if (!h->dampedDirty || !h->dD) return RFB200_OK;
    k_fold_damped<<<2048, 256, 0, h->compute>>>(h->dWb, h->dD, h->nBlocked);
    if (h->dD2) k_fold_damped<<<2048, 256, 0, h->compute>>>(h->dWb2, h->dD2, h->nBlocked);
    RF_CUDA(h, cudaGetLastError());
    h->nKernelLaunches += h->dD2 ? 2 : 1;
    h->dampedDirty = false;
    return RFB200_OK;
}

int get_plan2d(rfb200_handle h, int batch, cufftHandle* out) {
    auto it = h->plans2d.find(batch);
    if (it != h->plans2d.end()) { *out = it->second; return RFB200_OK; }
    cufftHandle p;
    int n[2] = {h->geo.P, h->geo.P};
    RF_CUFFT(h, cufftPlanMany(&p, 2, n, nullptr, 1, 0, nullptr, 1, 0, CUFFT_R2C, batch));
    RF_CUFFT(h, cufftSetStream(p, h->compute));
    h->plans2d[batch] = p;
    *out = p;
    return RFB200_OK;
}

// weight and shift parameters of one metadata row (RF.cpp:362-381)
ImgParams make_img_params(const rfb200_config& cfg, const rfb200_particle& p, bool& anySpline) {
    ImgParams q{};
    double w = cfg.use_weights ? p.weight : 1.0;     // RF.cpp:374-381
    q.weight = (float)w;
    q.skip = (w == 0.0) ? 1 : 0;                        // RF.cpp:483-484
    // readApplyGeo(only_apply_shifts): out(x) = in(x - shift).  Integer shifts are an exact circular
    // shift; if either component is fractional the image goes through cubic B-spline interpolation.
    double rx = std::nearbyint(p.shift_x), ry = std::nearbyint(p.shift_y);
    bool integer = std::fabs(p.shift_x - rx) < 1e-9 && std::fabs(p.shift_y - ry) < 1e-9;
    if (integer) {
        q.mx = (int)(-rx);
        q.my = (int)(-ry);
        q.ux = q.uy = 0.f;
        q.spline = 0;
    } else {
        double fx = std::floor(-p.shift_x), fy = std::floor(-p.shift_y);
        q.mx = (int)fx;
        q.my = (int)fy;
        q.ux = (float)(-p.shift_x - fx);
        q.uy = (float)(-p.shift_y - fy);
        q.spline = 1;
        anySpline = true;
    }
    return q;
}

// fill one chunk's parameter slot on the host (double precision) and upload it
int upload_chunk_params(rfb200_handle h, const rfb200_particle* meta, int n, ParamSlot** slotOut, int* nPlanesOut, bool* anySplineOut) {
    bool anySpline = false;
    ParamSlot& s = h->slots[h->slotIdx];
    h->slotIdx ^= 1;
    if (s.used) RF_CUDA(h, cudaEventSynchronize(s.done));
    const Geometry& g = h->geo;
    const double pixPerVox = (double)g.P / (double)g.Z;
    int np = 0;
    for (int i = 0; i < n; ++i) {
        const rfb200_particle& p = meta[i];
        ImgParams q = make_img_params(h->cfg, p, anySpline);
        s.img[i] = q;
        if (s.imgPlane0) s.imgPlane0[i] = q.skip ? -1 : np;
        if (h->cfg.use_ctf)
            s.ctf[i] = host::make_ctf(p.kV, p.defocusU, p.defocusV, p.defocus_angle, p.Cs, p.Ca, p.espr, p.ispr, p.alpha, p.DeltaF,
                                      p.DeltaR, p.Q0, p.K, p.envR0, p.envR1, p.envR2, p.phase_shift, p.vpp_radius);
        if (q.skip) continue;
        for (int sIdx = 0; sIdx < h->nSymTot; ++sIdx) {
            host::make_plane(&h->sym[9 * sIdx], p.rot, p.tilt, p.psi, pixPerVox, s.planesD[np]);
            s.planeImg[np] = i;
            ++np;
        }
    }
    use_param_set(h, &s == &h->slots[0] ? 0 : 1);
    int rcf = fetch_params(h, h->dImg, s.img, sizeof(ImgParams) * n);
    if (!rcf && h->cfg.use_ctf) rcf = fetch_params(h, h->dCtf, s.ctf, sizeof(CtfConsts) * n);
    if (!rcf && np) {
        rcf = fetch_params(h, h->dPlanesD, s.planesD, sizeof(PlaneD) * np);
        if (!rcf) rcf = fetch_params(h, h->dPlaneImg, s.planeImg, sizeof(int) * np);
    }
    if (!rcf) {
        // stable sort of the chunk's planes by class (axis dominating the normal), components permuted to the (a,b,d)
        // order of the class; every class is then cut into launches of <= kLaunchPlanes planes
        int start[4] = {0, 0, 0, 0};
        for (int k = 0; k < np; ++k) start[host::plane_class(s.planesD[k]) + 1]++;
        for (int c = 0; c < 3; ++c) start[c + 1] += start[c];
        int fill[3] = {start[0], start[1], start[2]};
        for (int k = 0; k < np; ++k) {
            const int cls = host::plane_class(s.planesD[k]);
            const int pos = fill[cls]++;
            const int img = s.planeImg[k];
            host::permute_plane(s.planesD[k], cls, img, s.img[img].weight, s.planesDp[pos], s.planesS[pos]);
        }
        s.launches.clear();
        for (int c = 0; c < 3; ++c) {
            const int nc = start[c + 1] - start[c];
            if (!nc) continue;
            // full launches first, the remainder last: every launch reads and writes all the sticks its planes touch, so an
            // overflow of a few planes is cheap as a small launch (few sticks) and expensive as two half-full ones
            int at = start[c];
            for (int left = nc; left > 0;) {
                const int cnt = std::min(left, kLaunchPlanes);
                s.launches.push_back({c, at, cnt});
                at += cnt;
                left -= cnt;
            }
        }
        for (size_t g = 0; g < s.launches.size(); ++g) {
            float* soa = s.soaP + g * 9 * kLaunchPlanes;
            for (int k = 0; k < s.launches[g].count; ++k) {
                const PlaneS& f = s.planesS[s.launches[g].start + k];
                const float comp[9] = {f.e1a, f.e1b, f.e1d, f.e2a, f.e2b, f.e2d, f.na, f.nb, f.nd};
                for (int c = 0; c < 9; ++c) soa[c * kLaunchPlanes + k] = comp[c];
            }
        }
        if (!s.launches.empty()) rcf = fetch_params(h, h->dPlanesSoAp, s.soaP, sizeof(float) * 9 * kLaunchPlanes * s.launches.size());
        if (!rcf) rcf = fetch_params(h, h->dImgPlane0, s.imgPlane0, sizeof(int) * n);
    }
    if (!rcf) rcf = params_uploaded(h, &s == &h->slots[0] ? 0 : 1);
    if (rcf) return rcf;
    s.used = true;
    *anySplineOut = anySpline;
    *slotOut = &s;
    *nPlanesOut = np;
    return RFB200_OK;
}

SliceParams make_slice_params(rfb200_handle h) {
    const Geometry& g = h->geo;
    SliceParams sp{};
    sp.P = g.P; sp.Xh = g.P / 2 + 1; sp.iLo = h->iLo; sp.iHi = h->iHi;
    sp.R = g.R; sp.Rp = g.Rp; sp.side = g.side;
    sp.useCtf = h->cfg.use_ctf; sp.phaseFlipped = h->cfg.phase_flipped;
    const double aStep = (h->cfg.use_ctf ? 1.0 / h->cfg.sampling : 1.0) / (double)g.P;
    sp.a2 = aStep * aStep;
    sp.a = (float)aStep;
    sp.minCtfF = (float)h->cfg.min_ctf;
    sp.minCtf = h->cfg.min_ctf;
    sp.invP2 = (float)(1.0 / ((double)g.P * (double)g.P));
    return sp;
}

// K1b' -> K2' (one launch per plane class) -> K2e' -> K2r for the n images whose half-plane FFTs sit in dFft
Slice2Args make_slice_args(rfb200_handle h) {
    const Geometry& g = h->geo;
    Slice2Args a{};
    a.sp = make_slice_params(h);
    a.pitch = g.pitch; a.planeStride = g.planeStride; a.colOff = g.colOff;
    a.fft = h->dFft; a.slices = h->dSlices2; a.col0 = h->dCol02; a.damped = h->dDamped; a.damped2 = h->dDamped2; a.dampedMask = h->dDampedMask;
    a.ip = h->dImg; a.ctfs = h->dCtf; a.jmax = h->dJmax;
    return a;
}

int insert_planes_sticks(rfb200_handle h, ParamSlot* slot, int n, int nPlanes) {
    const Geometry& g = h->geo;
    int rc = RFB200_OK;
    if (nPlanes) {
        RF_CUDA(h, cudaMemsetAsync(h->dStickCounters, 0, sizeof(int) * slot->launches.size(), h->compute));
        StageTimer t(h, Stage::GATHER, h->compute);
        // launches that touch the same voxels are serialised on the stream; each owns its sticks
        static thread_local StickLaunch L;          // 31 KB parameter block, copied by the launch
        for (size_t gi = 0; gi < slot->launches.size(); ++gi) {
            const ParamSlot::Launch& la = slot->launches[gi];
            if (h->nUnits[la.cls] == 0) continue;
            StickArgs& a = L.a;
            a = StickArgs{};
            a.geo = g;
            a.units = h->dUnits[la.cls]; a.nUnits = h->nUnits[la.cls]; a.counter = h->dStickCounters + gi;
            a.cls = la.cls; a.nPlanes = la.count;
            a.blobTable = h->dBlobTable;
            a.planesSoA = h->dPlanesSoAp + gi * 9 * kLaunchPlanes;
            a.slices = h->dSlices2; a.rimTab = h->dRimTab + g.Rp;
            a.Vb = h->dVb; a.Wb = h->dWb; a.Wb2 = h->dWb2;
#ifdef RF_L2_PREFETCH
            L.sliceMap = h->sliceMap;
#endif
            std::memcpy(L.ps, slot->planesS + la.start, sizeof(PlaneS) * la.count);
            std::memcpy(L.pd, slot->planesDp + la.start, sizeof(PlaneD) * la.count);
            rc = launch_sticks(h, L, std::min(h->stickGrid, (h->nUnits[la.cls] + kStickWarps - 1) / kStickWarps));
            if (rc) return rc;
            h->nKernelLaunches += 1;
            h->nGatherLaunches += 1;
        }
    }
    if (h->nEdge && nPlanes) {
        StageTimer t(h, Stage::EDGE, h->compute);
        Edge2Args e{};
        e.geo = g;
        e.items = h->dEdge; e.groupStart = h->dEdgeGroups; e.nGroups = h->nEdgeGroups;
        e.planesD = h->dPlanesD; e.planeImg = h->dPlaneImg; e.img = h->dImg; e.nPlanes = nPlanes;
        e.blobTable = h->dBlobTable; e.slices = h->dSlices2; e.col0 = h->dCol02; e.rimTab = h->dRimTab + g.Rp;
        e.Vb = h->dVb; e.Wb = h->dWb; e.Wb2 = h->dWb2;
        e.iDeltaD = h->tables.iDeltaSqrt;
        k_edge2<<<(h->nEdgeGroups + 3) / 4, 128, 0, h->compute>>>(e);      // one warp per target voxel
        RF_CUDA(h, cudaGetLastError());
        h->nKernelLaunches += 1;
    }
    // (Measured and dropped: running the damped-weight scatter on a side stream beside the gather launches — it writes its
    // own fixed-point volume — changed nothing: 44.80 vs 44.90 ms per 4096 particles; the persistent gather leaves no idle
    // SM time to fill.)
    if (h->dDamped && nPlanes) {
        StageTimer t(h, Stage::EDGE, h->compute);
        DampedArgs d{};
        d.geo = g;
        d.mask = h->dDampedMask; d.damped = h->dDamped; d.damped2 = h->dDamped2; d.nImg = n; d.imgPlane0 = h->dImgPlane0; d.nSym = h->nSymTot;
        d.planesD = h->dPlanesD; d.blobTable = h->dBlobTable; d.iDeltaD = h->tables.iDeltaSqrt;
        d.D = h->dD; d.D2 = h->dD2;
        const int nWords = ((g.R + 1 + 31) / 32) * (2 * g.R + 1);
        k_damped_scatter<<<dim3((nWords + 255) / 256, n), 256, 0, h->compute>>>(d);
        RF_CUDA(h, cudaGetLastError());
        h->nKernelLaunches += 1;
        h->dampedDirty = true;
    }
    return RFB200_OK;
}

// K1a -> cuFFT -> K1b -> K2 (+K2e) for n images whose raw data sit at dRaw (device)
// shift + pad (K1a) and the batched R2C of one chunk into dFft
int pad_and_fft(rfb200_handle h, const float* dRaw, int n, bool anySpline) {
    const Geometry& g = h->geo;
    {
        StageTimer t(h, Stage::PAD, h->compute);
        if (anySpline) {
            if (!h->dCoef) RF_CUDA(h, cudaMalloc(&h->dCoef, sizeof(float) * (size_t)h->chunkImages * g.N * g.N));
            dim3 pg((g.N + 127) / 128, n);
            k_bspline_prefilter<<<pg, 128, 0, h->compute>>>(dRaw, h->dCoef, h->dImg, g.N, 0);
            k_bspline_prefilter<<<pg, 128, 0, h->compute>>>(dRaw, h->dCoef, h->dImg, g.N, 1);
            RF_CUDA(h, cudaGetLastError());
            h->nKernelLaunches += 2;
        }
        dim3 grid((g.N * g.N + 255) / 256, n);
        k_pad_images<<<grid, 256, 0, h->compute>>>(dRaw, h->dCoef, h->dPad, h->dImg, g.N, g.P);
        RF_CUDA(h, cudaGetLastError());
        h->nKernelLaunches += 1;
    }
    {
        StageTimer t(h, Stage::FFT2D, h->compute);
        cufftHandle plan;
        int rc = get_plan2d(h, n, &plan);
        if (rc) return rc;
        RF_CUFFT(h, cufftExecR2C(plan, h->dPad, reinterpret_cast<cufftComplex*>(h->dFft)));
    }
    return RFB200_OK;
}

// --fast: K1a -> cuFFT R2C -> K1f (crop, re-centre, CTF, weights folded) -> K2f (nearest-pixel insertion)
int process_chunk_fast(rfb200_handle h, const float* dRaw, const rfb200_particle* meta, int n) {
    const FastGeo& fg = h->fgeo;
    ParamSlot& s = h->slots[h->slotIdx];
    h->slotIdx ^= 1;
    if (s.used) RF_CUDA(h, cudaEventSynchronize(s.done));
    bool anySpline = false;
    int np = 0;
    for (int i = 0; i < n; ++i) {
        const rfb200_particle& p = meta[i];
        ImgParams q = make_img_params(h->cfg, p, anySpline);
        if (h->cfg.use_weights && (float)p.weight == 0.f) q.skip = 1;     // G:351-353 compares the float weight
        s.img[i] = q;
        if (h->cfg.use_ctf)
            s.ctf[i] = host::make_ctf(p.kV, p.defocusU, p.defocusV, p.defocus_angle, p.Cs, p.Ca, p.espr, p.ispr, p.alpha, p.DeltaF,
                                      p.DeltaR, p.Q0, p.K, p.envR0, p.envR1, p.envR2, p.phase_shift, p.vpp_radius);
        if (q.skip) continue;
        for (int sIdx = 0; sIdx < h->nSymTot; ++sIdx)
            host::make_fast_space(fg, &h->sym[9 * sIdx], p.rot, p.tilt, p.psi, i, q.weight, s.fast[np++]);
    }
    use_param_set(h, &s == &h->slots[0] ? 0 : 1);
    int rc = fetch_params(h, h->dImg, s.img, sizeof(ImgParams) * n);
    if (!rc && h->cfg.use_ctf) rc = fetch_params(h, h->dCtf, s.ctf, sizeof(CtfConsts) * n);
    if (!rc && np) rc = fetch_params(h, h->dFastSpaces, s.fast, sizeof(FastSpace) * np);
    if (!rc) rc = params_uploaded(h, &s == &h->slots[0] ? 0 : 1);
    if (rc) return rc;
    s.used = true;
    rc = pad_and_fft(h, dRaw, n, anySpline);
    if (rc) return rc;
    {
        StageTimer t(h, Stage::SLICE, h->compute);
        FastPrepArgs a{};
        a.g = fg; a.fft = h->dFft; a.pix = h->dFastPix; a.ip = h->dImg; a.ctfs = h->dCtf;
        a.useCtf = h->cfg.use_ctf; a.phaseFlipped = h->cfg.phase_flipped;
        a.iTs = h->cfg.use_ctf ? 1.0 / h->cfg.sampling : 1.0;
        a.minCtf = h->cfg.min_ctf;
        a.maxRes2 = (float)(h->cfg.max_resolution * h->cfg.max_resolution);
        a.ctfInt = (fg.Pv & (fg.Pv - 1)) == 0 ? 1 : 0;
        a.sp = make_slice_params(h);          // frequency step 1/(Pv * sampling), --minCTF, --phaseFlipped
        dim3 grid((fg.sx * fg.sy + 255) / 256, n);
        k_fast_prepare<<<grid, 256, 0, h->compute>>>(a);
        RF_CUDA(h, cudaGetLastError());
        h->nKernelLaunches += 1;
    }
    if (np) {
        StageTimer t(h, Stage::GATHER, h->compute);
        for (int p0 = 0; p0 < np; p0 += 65535) {
            const int cnt = std::min(65535, np - p0);
            FastInsertArgs a{};
            a.g = fg; a.spaces = h->dFastSpaces + p0; a.pix = h->dFastPix; a.A = h->dFastAcc;
            h->fastDirty = true;
            dim3 grid((fg.S + 1 + 31) / 32, (fg.S + 1 + 7) / 8, cnt);
            k_fast_insert<<<grid, dim3(32, 8), 0, h->compute>>>(a);
            RF_CUDA(h, cudaGetLastError());
            h->nKernelLaunches += 1;
            h->nGatherLaunches += 1;
        }
    }
    RF_CUDA(h, cudaEventRecord(s.done, h->compute));
    h->nImages += n;
    h->nPlanes += np;
    h->lastChunkImages = n;
    return RFB200_OK;
}

int process_chunk(rfb200_handle h, const float* dRaw, const rfb200_particle* meta, int n) {
    if (h->fast) return process_chunk_fast(h, dRaw, meta, n);
    const Geometry& g = h->geo;
    ParamSlot* slot = nullptr;
    int nPlanes = 0;
    bool anySpline = false;
    int rc = upload_chunk_params(h, meta, n, &slot, &nPlanes, &anySpline);
    if (rc) return rc;
    if (h->fusedFft && !anySpline) {
        // K1r -> K1c: shift + pad + FFT + CTF + slices without the padded image and the half-plane transform in HBM
        FftRowsArgs ra{};
        ra.raw = dRaw; ra.ip = h->dImg; ra.twiddle = h->dTwiddle; ra.T = h->dFft; ra.N = g.N;
        FftColsArgs ca{};
        ca.s = make_slice_args(h); ca.twiddle = h->dTwiddle; ca.T = h->dFft; ca.N = g.N;
        // which CTF code the slice pass needs: none, the fixed-point phase path only, or the general one
        int ctfMode = 0;
        if (h->cfg.use_ctf) {
            ctfMode = ca.s.sp.a >= 1e-6f ? 1 : 2;
            for (int i = 0; i < n && ctfMode == 1; ++i)
                if (slot->ctf[i].has_envelope || slot->ctf[i].has_vpp) ctfMode = 2;
        }
        rc = launch_fused_fft(h, ra, ca, n, ctfMode);
        if (rc) return rc;
    } else {
    rc = pad_and_fft(h, dRaw, n, anySpline);
    if (rc) return rc;
    {
        StageTimer t(h, Stage::SLICE, h->compute);
        Slice2Args a = make_slice_args(h);
        dim3 grid((g.R + 1 + 31) / 32, (2 * g.R + 1 + 8 * kSliceRowsPerThread - 1) / (8 * kSliceRowsPerThread), n);
        k_make_slices2<<<grid, dim3(32, 8), 0, h->compute>>>(a);
        RF_CUDA(h, cudaGetLastError());
        h->nKernelLaunches += 1;
    }
    }
    rc = insert_planes_sticks(h, slot, n, nPlanes);
    if (rc) return rc;
    RF_CUDA(h, cudaEventRecord(slot->done, h->compute));
    h->nImages += n;
    h->nPlanes += nPlanes;
    h->lastChunkImages = n;
    return RFB200_OK;
}

int validate(const rfb200_config* c, std::string& why) {
    if (!c) { why = "null config"; return RFB200_ERR_ARG; }
    if (c->abi_version != RFB200_ABI_VERSION) { why = "abi_version mismatch"; return RFB200_ERR_ARG; }
    if (c->img_size < 4 || c->img_size > 4096) { why = "img_size out of range"; return RFB200_ERR_ARG; }
    if (c->pad_proj < 1.0 || c->pad_vol < 1.0) { why = "padding factors must be >= 1"; return RFB200_ERR_ARG; }
    if (!(c->max_resolution > 0.0) || c->max_resolution > 0.5) { why = "max_resolution must be in (0, 0.5]"; return RFB200_ERR_ARG; }
    if (!(c->blob_radius > 0.0)) { why = "blob radius must be positive"; return RFB200_ERR_ARG; }
    if (c->blob_order != 0 && c->blob_order != 2) { why = "blob order must be 0 or 2 (kaiser_Fourier_value, blobs.cpp:146)"; return RFB200_ERR_ARG; }
    if (c->n_sym < 0 || (c->n_sym > 0 && !c->sym_matrices)) { why = "symmetry matrices missing"; return RFB200_ERR_ARG; }
    if (c->n_iter_weight < 0) { why = "n_iter_weight must be >= 0"; return RFB200_ERR_ARG; }
    if (c->use_ctf && !(c->sampling > 0.0)) { why = "sampling must be positive with use_ctf"; return RFB200_ERR_ARG; }
    return RFB200_OK;
}

void free_all(rfb200_handle h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    if (h->compute) cudaStreamSynchronize(h->compute);
    if (h->copy) cudaStreamSynchronize(h->copy);
    resolve_timings(h);
    for (auto e : h->evPool) cudaEventDestroy(e);
    if (h->swStart) cudaEventDestroy(h->swStart);
    if (h->swStop) cudaEventDestroy(h->swStop);
    if (h->dSum) cudaFree(h->dSum);
    if (h->hSum) cudaFreeHost(h->hSum);
    if (h->evSum) cudaEventDestroy(h->evSum);
    for (auto& kv : h->plans2d) cufftDestroy(kv.second);
    if (h->havePlan3d) cufftDestroy(h->plan3d);
    void* dev[] = {h->dBlobTable, h->dJmax, h->dEdge, h->dEdgeGroups, h->dG, h->dVb, h->dWb, h->dWb2, h->dVsaved, h->dWsaved, h->dW2saved, h->dRaw[0], h->dRaw[1],
                   h->dPad, h->dCoef, h->dFft, h->dNorm,
                   h->dVol, h->dOut, h->dSlices2, h->dCol02, h->dDamped, h->dDamped2, h->dDampedMask, h->dD, h->dD2, h->dRimTab, h->dUnits[0], h->dUnits[1], h->dUnits[2],
                   h->dStickCounters, h->dTwiddle, h->dPlanesDp, h->dPlanesSStage,
                   h->dFastPix, h->dFastVh, h->dFastVc, h->dFastWh, h->dFastWc, h->dFastAcc};
    for (void* p : dev) if (p) cudaFree(p);
    if (!h->dpar[0].up) {        // creation failed before the parameter sets existed: the members still own set 0
        void* own[] = {h->dImg, h->dCtf, h->dPlanesD, h->dPlaneImg, h->dPlanesSoAp, h->dImgPlane0, h->dFastSpaces};
        for (void* p : own) if (p) cudaFree(p);
    }
    for (auto& d : h->dpar) {
        void* dp[] = {d.img, d.ctf, d.planesD, d.planeImg, d.soaP, d.imgPlane0, d.fast};
        for (void* p : dp) if (p) cudaFree(p);
        if (d.up) cudaEventDestroy(d.up);
    }
    for (auto& s : h->slots) {
        void* hp[] = {s.img, s.ctf, s.planesD, s.planeImg, s.planesDp, s.planesS, s.soaP, s.imgPlane0, s.fast};
        for (void* p : hp) if (p) cudaFreeHost(p);
        if (s.done) cudaEventDestroy(s.done);
    }
    for (int i = 0; i < 2; ++i) {
        if (h->evH2D[i]) cudaEventDestroy(h->evH2D[i]);
        if (h->evRawFree[i]) cudaEventDestroy(h->evRawFree[i]);
    }
    for (auto& p : h->peers) {
        if (!p.open) continue;
        cudaIpcCloseMemHandle(p.V);
        cudaIpcCloseMemHandle(p.W);
        if (p.W2) cudaIpcCloseMemHandle(p.W2);
    }
    if (h->dBarrier) cudaFree(h->dBarrier);
#if RFB200_HAVE_NCCL_H
    if (h->comm && g_nccl.ok) g_nccl.CommDestroy(h->comm);
#endif
    if (h->compute) cudaStreamDestroy(h->compute);
    if (h->copy) cudaStreamDestroy(h->copy);
    delete h;
}

// --fast handle: images are padded to N*pad_vol (both --fast programs do, G:229, 384-394), the accumulators are the
// (S+1)^3 temporary volume and weights (G:838-843), finalisation buffers are allocated on first use
int do_create_fast(rfb200_handle h) {
    const rfb200_config& c = h->cfg;
    h->fast = true;
    h->fgeo = host::make_fast_geo(c.img_size, c.pad_vol, c.max_resolution);
    const FastGeo& fg = h->fgeo;
    if (fg.S < 2 || fg.S > fg.Pv) return fail(h, RFB200_ERR_ARG, "--fast: resolution sphere does not fit the padded volume");
    h->tables = host::build_tables(c.img_size, c.pad_proj, c.pad_vol, c.blob_radius, c.blob_order, c.blob_alpha);
    Geometry& g = h->geo;
    g = Geometry{};
    g.N = c.img_size; g.P = fg.Pv; g.Z = fg.Pv; g.X = fg.Pv / 2 + 1;
    double meanF2 = 0;
    std::vector<float> G = host::build_gridding_table(c.img_size, c.pad_proj, c.pad_vol, h->tables, 1, &meanF2);
    std::vector<float> blobF(kBlobTable);
    for (int i = 0; i < kBlobTable; ++i) blobF[i] = (float)h->tables.blobSqrt[i];
    int maxBatch = c.max_batch > 0 ? c.max_batch : 1024;
    h->chunkImages = std::min(maxBatch, 1024);
    RF_CUDA(h, cudaStreamCreateWithFlags(&h->compute, cudaStreamNonBlocking));
    RF_CUDA(h, cudaStreamCreateWithFlags(&h->copy, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        RF_CUDA(h, cudaEventCreateWithFlags(&h->evH2D[i], cudaEventDisableTiming));
        RF_CUDA(h, cudaEventCreateWithFlags(&h->evRawFree[i], cudaEventDisableTiming));
    }
    RF_CUDA(h, cudaMalloc(&h->dBlobTable, sizeof(float) * kBlobTable));
    RF_CUDA(h, cudaMemcpy(h->dBlobTable, blobF.data(), sizeof(float) * kBlobTable, cudaMemcpyHostToDevice));
    RF_CUDA(h, cudaMalloc(&h->dSum, sizeof(double) * 1025));
    RF_CUDA(h, cudaMallocHost(&h->hSum, sizeof(double)));
    RF_CUDA(h, cudaEventCreateWithFlags(&h->evSum, cudaEventDisableTiming));
    RF_CUDA(h, cudaEventCreate(&h->swStart));
    RF_CUDA(h, cudaEventCreate(&h->swStop));
    RF_CUDA(h, cudaMalloc(&h->dG, sizeof(float) * G.size()));
    RF_CUDA(h, cudaMemcpy(h->dG, G.data(), sizeof(float) * G.size(), cudaMemcpyHostToDevice));
    h->nBlocked = (int64_t)(fg.S + 1) * (fg.S + 1) * (fg.S + 1);
    RF_CUDA(h, cudaMalloc(&h->dVb, sizeof(float2) * h->nBlocked));
    RF_CUDA(h, cudaMalloc(&h->dWb, sizeof(float) * h->nBlocked));
    RF_CUDA(h, cudaMemset(h->dVb, 0, sizeof(float2) * h->nBlocked));
    RF_CUDA(h, cudaMemset(h->dWb, 0, sizeof(float) * h->nBlocked));
    RF_CUDA(h, cudaMalloc(&h->dFastAcc, sizeof(float4) * h->nBlocked));
    RF_CUDA(h, cudaMemset(h->dFastAcc, 0, sizeof(float4) * h->nBlocked));
    const size_t CH = h->chunkImages;
    const size_t nRaw = CH * g.N * g.N, nPad = CH * (size_t)g.P * g.P, nFft = CH * (size_t)g.P * (g.P / 2 + 1);
    for (int i = 0; i < 2; ++i) RF_CUDA(h, cudaMalloc(&h->dRaw[i], sizeof(float) * nRaw));
    RF_CUDA(h, cudaMalloc(&h->dPad, sizeof(float) * nPad));
    RF_CUDA(h, cudaMemset(h->dPad, 0, sizeof(float) * nPad));
    RF_CUDA(h, cudaMalloc(&h->dFft, sizeof(float2) * nFft));
    RF_CUDA(h, cudaMalloc(&h->dFastPix, sizeof(float4) * CH * fg.sx * fg.sy));
    const size_t maxSpaces = CH * h->nSymTot;
    RF_CUDA(h, cudaMalloc(&h->dFastSpaces, sizeof(FastSpace) * maxSpaces + 16));
    RF_CUDA(h, cudaMalloc(&h->dImg, sizeof(ImgParams) * CH));
    RF_CUDA(h, cudaMalloc(&h->dCtf, sizeof(CtfConsts) * CH));
    if (int rcp = make_param_sets(h, sizeof(ImgParams) * CH, sizeof(CtfConsts) * CH, 0, 0, 0, 0, sizeof(FastSpace) * maxSpaces + 16)) return rcp;
    for (auto& s : h->slots) {
        RF_CUDA(h, cudaMallocHost(&s.img, sizeof(ImgParams) * CH));
        RF_CUDA(h, cudaMallocHost(&s.ctf, sizeof(CtfConsts) * CH));
        RF_CUDA(h, cudaMallocHost(&s.fast, sizeof(FastSpace) * maxSpaces + 16));
        RF_CUDA(h, cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    }
    RF_CUDA(h, cudaDeviceSynchronize());
    return RFB200_OK;
}

// --fast: mirrorAndCrop -> applyBlob -> forceHermitianSymmetry + processWeights + convertToExpectedSpace (G:879-893),
// then the common inverse FFT, crop and gridding correction
int finalize_fast(rfb200_handle h, float* out, float* fourierOut = nullptr) {
    const FastGeo& fg = h->fgeo;
    const Geometry& g = h->geo;
    const rfb200_config& c = h->cfg;
    const size_t nH = (size_t)(fg.S + 1) * (fg.S + 1) * (fg.X + 1);
    const size_t nHalf = (size_t)g.Z * g.Z * g.X, nVol = (size_t)g.Z * g.Z * g.Z, nOut = (size_t)g.N * g.N * g.N;
    if (!h->havePlan3d) {
        RF_CUFFT(h, cufftPlan3d(&h->plan3d, g.Z, g.Z, g.Z, CUFFT_C2R));
        RF_CUFFT(h, cufftSetStream(h->plan3d, h->compute));
        h->havePlan3d = true;
    }
    if (!h->dFastVh) {
        RF_CUDA(h, cudaMalloc(&h->dFastVh, sizeof(float2) * nH));
        RF_CUDA(h, cudaMalloc(&h->dFastVc, sizeof(float2) * nH));
        RF_CUDA(h, cudaMalloc(&h->dFastWh, sizeof(float) * nH));
        RF_CUDA(h, cudaMalloc(&h->dFastWc, sizeof(float) * nH));
    }
    if (!h->dNorm) RF_CUDA(h, cudaMalloc(&h->dNorm, sizeof(float2) * nHalf));
    if (!h->dVol) RF_CUDA(h, cudaMalloc(&h->dVol, sizeof(float) * nVol));
    if (!h->dOut) RF_CUDA(h, cudaMalloc(&h->dOut, sizeof(float) * nOut));
    if (int rcf = flush_deficit(h)) return rcf;
    {
        StageTimer t(h, Stage::FINALIZE, h->compute);
        const unsigned gh = (unsigned)((nH + 255) / 256);
        k_fast_mirror_crop<<<gh, 256, 0, h->compute>>>(fg, h->dVb, h->dWb, h->dFastVh, h->dFastWh);
        k_fast_blob<<<gh, 256, 0, h->compute>>>(fg, h->dBlobTable, (float)c.blob_radius, (float)h->tables.iDeltaSqrt, h->dFastVh, h->dFastWh,
                                                 h->dFastVc, h->dFastWc);
        const float corr = (float)(std::pow(c.pad_proj, 2.0) / (c.img_size * std::pow(c.pad_vol, 3.0)));   // G:753-754
        k_fast_to_fourier<<<(unsigned)((nHalf + 255) / 256), 256, 0, h->compute>>>(fg, h->dFastVc, h->dFastWc, corr, h->dNorm);
        RF_CUDA(h, cudaGetLastError());
        if (fourierOut) {       // diagnostics: the transform handed to the inverse FFT
            RF_CUDA(h, cudaMemcpyAsync(fourierOut, h->dNorm, sizeof(float2) * nHalf, cudaMemcpyDeviceToHost, h->compute));
            RF_CUDA(h, cudaStreamSynchronize(h->compute));
            return RFB200_OK;
        }
        RF_CUFFT(h, cufftExecC2R(h->plan3d, reinterpret_cast<cufftComplex*>(h->dNorm), h->dVol));
        k_crop_correct<<<(unsigned)((nOut + 255) / 256), 256, 0, h->compute>>>(h->dVol, h->dG, h->dOut, g.N, g.Z);
        RF_CUDA(h, cudaGetLastError());
        RF_CUDA(h, cudaMemcpyAsync(out, h->dOut, sizeof(float) * nOut, cudaMemcpyDeviceToHost, h->compute));
    }
    h->nKernelLaunches += 4;
    RF_CUDA(h, cudaStreamSynchronize(h->compute));
    resolve_timings(h);
    return RFB200_OK;
}

int do_create(rfb200_handle h) {
    const rfb200_config& c = h->cfg;
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(h, RFB200_ERR_CUDA, std::string("no CUDA device available (") + cudaGetErrorString(ce) + "); this library has no CPU fallback");
    if (c.device < 0 || c.device >= ndev) return fail(h, RFB200_ERR_ARG, "device ordinal out of range");
    RF_CUDA(h, cudaSetDevice(c.device));
    cudaDeviceProp prop;
    RF_CUDA(h, cudaGetDeviceProperties(&prop, c.device));

    // ---- host precompute (double)
    h->nSymTot = c.n_sym + 1;
    h->sym.assign((size_t)9 * h->nSymTot, 0.0);
    h->sym[0] = h->sym[4] = h->sym[8] = 1.0;
    if (c.n_sym) std::memcpy(&h->sym[9], c.sym_matrices, sizeof(double) * 9 * c.n_sym);
    if (c.fast) return do_create_fast(h);
    h->tables = host::build_tables(c.img_size, c.pad_proj, c.pad_vol, c.blob_radius, c.blob_order, c.blob_alpha);
    int P = (int)(c.img_size * c.pad_proj);
    int R = 0;
    host::build_cutoff(P, c.max_resolution, h->jmax, h->iLo, h->iHi, R);
    h->geo = host::make_geometry(c.img_size, c.pad_proj, c.pad_vol, c.max_resolution, c.blob_radius, R);
    Geometry& g = h->geo;
    if (g.K > kMaxWin) return fail(h, RFB200_ERR_ARG, "blob radius too large for the interpolation window (max 8 pixels)");
    std::vector<int32_t> rimTab = host::build_rim_table(g, h->jmax, h->iLo, h->iHi);
    std::vector<EdgeItem> edge = host::build_edge_items(g);
    h->nEdge = (int)edge.size();
    double meanF2 = 0;
    std::vector<float> G = host::build_gridding_table(c.img_size, c.pad_proj, c.pad_vol, h->tables, c.n_iter_weight, &meanF2);
    std::vector<float> blobF(kBlobTable);
    for (int i = 0; i < kBlobTable; ++i) blobF[i] = (float)h->tables.blobSqrt[i];

    int maxBatch = c.max_batch > 0 ? c.max_batch : 1024;
    // Images per preprocessing chunk.  A gather launch carries at most kLaunchPlanes planes of one class; without symmetry a
    // chunk of 720 images gives 240 +- 13 planes per class = one launch each, 3 % of the classes overflowing into a second
    // one (1024 images would give 341 = two launches of 171, i.e. the per-launch work of a stick amortised over fewer planes:
    // measured 44.07 vs 44.59 ms per 4096 particles).  With symmetry every class needs many launches anyway and the split
    // is even.
    h->chunkImages = std::min(maxBatch, h->nSymTot == 1 ? (3 * kLaunchPlanes * 10) / 11 : 1024);

    // ---- streams / events
    RF_CUDA(h, cudaStreamCreateWithFlags(&h->compute, cudaStreamNonBlocking));
    RF_CUDA(h, cudaStreamCreateWithFlags(&h->copy, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        RF_CUDA(h, cudaEventCreateWithFlags(&h->evH2D[i], cudaEventDisableTiming));
        RF_CUDA(h, cudaEventCreateWithFlags(&h->evRawFree[i], cudaEventDisableTiming));
    }

    // ---- static device data
    RF_CUDA(h, cudaMalloc(&h->dBlobTable, sizeof(float) * kBlobTable));
    RF_CUDA(h, cudaMemcpy(h->dBlobTable, blobF.data(), sizeof(float) * kBlobTable, cudaMemcpyHostToDevice));
    RF_CUDA(h, cudaMalloc(&h->dJmax, sizeof(int) * h->jmax.size()));
    RF_CUDA(h, cudaMemcpy(h->dJmax, h->jmax.data(), sizeof(int) * h->jmax.size(), cudaMemcpyHostToDevice));
    if (!edge.empty()) {
        RF_CUDA(h, cudaMalloc(&h->dEdge, sizeof(EdgeItem) * edge.size()));
        RF_CUDA(h, cudaMemcpy(h->dEdge, edge.data(), sizeof(EdgeItem) * edge.size(), cudaMemcpyHostToDevice));
        std::vector<int32_t> starts = host::edge_group_starts(edge);
        h->nEdgeGroups = (int)starts.size() - 1;
        RF_CUDA(h, cudaMalloc(&h->dEdgeGroups, sizeof(int32_t) * starts.size()));
        RF_CUDA(h, cudaMemcpy(h->dEdgeGroups, starts.data(), sizeof(int32_t) * starts.size(), cudaMemcpyHostToDevice));
    }
    {
        const int Pp = g.P;
        const char* e = getenv("RFB200_FFT");          // developer switch: "cufft" forces the general path
        h->fusedFft = (Pp == 64 || Pp == 128 || Pp == 256 || Pp == 512 || Pp == 1024) && g.N <= Pp && !(e && std::string(e) == "cufft");
        if (h->fusedFft) {
            std::vector<float2> tw(Pp);
            for (int k = 0; k < Pp; ++k) {
                const double ang = -2.0 * host::kPi * (double)k / (double)Pp;
                tw[k] = make_float2((float)std::cos(ang), (float)std::sin(ang));
            }
            RF_CUDA(h, cudaMalloc(&h->dTwiddle, sizeof(float2) * Pp));
            RF_CUDA(h, cudaMemcpy(h->dTwiddle, tw.data(), sizeof(float2) * Pp, cudaMemcpyHostToDevice));
        }
    }
    RF_CUDA(h, cudaMalloc(&h->dSum, sizeof(double) * 1025));
    RF_CUDA(h, cudaMallocHost(&h->hSum, sizeof(double)));
    RF_CUDA(h, cudaEventCreateWithFlags(&h->evSum, cudaEventDisableTiming));
    RF_CUDA(h, cudaEventCreate(&h->swStart));
    RF_CUDA(h, cudaEventCreate(&h->swStop));
    RF_CUDA(h, cudaMalloc(&h->dG, sizeof(float) * G.size()));
    RF_CUDA(h, cudaMemcpy(h->dG, G.data(), sizeof(float) * G.size(), cudaMemcpyHostToDevice));

    // ---- accumulators
    h->nBlocked = (int64_t)g.tx * g.ty * g.tz * kTileVox;
    RF_CUDA(h, cudaMalloc(&h->dVb, sizeof(float2) * h->nBlocked));
    RF_CUDA(h, cudaMalloc(&h->dWb, sizeof(float) * h->nBlocked));
    RF_CUDA(h, cudaMemset(h->dVb, 0, sizeof(float2) * h->nBlocked));
    RF_CUDA(h, cudaMemset(h->dWb, 0, sizeof(float) * h->nBlocked));
    if (c.use_ctf && c.n_iter_weight > 1) {
        RF_CUDA(h, cudaMalloc(&h->dWb2, sizeof(float) * h->nBlocked));
        RF_CUDA(h, cudaMemset(h->dWb2, 0, sizeof(float) * h->nBlocked));
    }

    // ---- per-chunk buffers
    const size_t CH = h->chunkImages;
    const size_t nRaw = CH * g.N * g.N, nPad = CH * (size_t)g.P * g.P, nFft = CH * (size_t)g.P * (g.P / 2 + 1);
    const size_t nC0 = CH * (size_t)g.side;
    for (int i = 0; i < 2; ++i) RF_CUDA(h, cudaMalloc(&h->dRaw[i], sizeof(float) * nRaw));
    RF_CUDA(h, cudaMalloc(&h->dPad, sizeof(float) * nPad));
    RF_CUDA(h, cudaMemset(h->dPad, 0, sizeof(float) * nPad));
    RF_CUDA(h, cudaMalloc(&h->dFft, sizeof(float2) * nFft));
    {
        const size_t nSl2 = CH * (size_t)g.planeStride;
        RF_CUDA(h, cudaMalloc(&h->dSlices2, sizeof(float4) * nSl2 + 64));
        RF_CUDA(h, cudaMemset(h->dSlices2, 0, sizeof(float4) * nSl2 + 64));
        RF_CUDA(h, cudaMalloc(&h->dCol02, sizeof(float2) * nC0));
        RF_CUDA(h, cudaMemset(h->dCol02, 0, sizeof(float2) * nC0));
        RF_CUDA(h, cudaMalloc(&h->dRimTab, sizeof(int32_t) * rimTab.size()));
        RF_CUDA(h, cudaMemcpy(h->dRimTab, rimTab.data(), sizeof(int32_t) * rimTab.size(), cudaMemcpyHostToDevice));
        if (c.use_ctf) {
            RF_CUDA(h, cudaMalloc(&h->dDamped, sizeof(float) * CH * (size_t)(2 * g.R + 1) * (g.R + 1)));
            RF_CUDA(h, cudaMalloc(&h->dDampedMask, sizeof(uint32_t) * CH * (size_t)(2 * g.R + 1) * ((g.R + 1 + 31) / 32)));
            RF_CUDA(h, cudaMalloc(&h->dD, sizeof(unsigned long long) * h->nBlocked));
            RF_CUDA(h, cudaMemset(h->dD, 0, sizeof(unsigned long long) * h->nBlocked));
            if (c.n_iter_weight > 1) {
                RF_CUDA(h, cudaMalloc(&h->dDamped2, sizeof(float) * CH * (size_t)(2 * g.R + 1) * (g.R + 1)));
                RF_CUDA(h, cudaMalloc(&h->dD2, sizeof(unsigned long long) * h->nBlocked));
                RF_CUDA(h, cudaMemset(h->dD2, 0, sizeof(unsigned long long) * h->nBlocked));
            }
        }
        for (int cls = 0; cls < 3; ++cls) {
            std::vector<StickUnit> units = host::build_stick_units(g, cls);
            h->nUnits[cls] = (int)units.size();
            if (units.empty()) continue;
            RF_CUDA(h, cudaMalloc(&h->dUnits[cls], sizeof(StickUnit) * units.size()));
            RF_CUDA(h, cudaMemcpy(h->dUnits[cls], units.data(), sizeof(StickUnit) * units.size(), cudaMemcpyHostToDevice));
        }
        RF_CUDA(h, cudaMalloc(&h->dStickCounters, sizeof(int) * (((size_t)h->chunkImages * h->nSymTot + kLaunchPlanes - 1) / kLaunchPlanes + 3)));
    }
    RF_CUDA(h, cudaMalloc(&h->dImg, sizeof(ImgParams) * CH));
    RF_CUDA(h, cudaMalloc(&h->dCtf, sizeof(CtfConsts) * CH));
    const size_t maxPlanes = CH * h->nSymTot;
    const size_t maxLaunches = (maxPlanes + kLaunchPlanes - 1) / kLaunchPlanes + 3;    // every class rounds up
    RF_CUDA(h, cudaMalloc(&h->dPlanesD, sizeof(PlaneD) * maxPlanes + 16));
    RF_CUDA(h, cudaMalloc(&h->dPlaneImg, sizeof(int) * maxPlanes + 16));
    {
        RF_CUDA(h, cudaMalloc(&h->dPlanesSoAp, sizeof(float) * 9 * kLaunchPlanes * maxLaunches));
        RF_CUDA(h, cudaMalloc(&h->dImgPlane0, sizeof(int) * CH + 16));
    }
    if (int rcp = make_param_sets(h, sizeof(ImgParams) * CH, sizeof(CtfConsts) * CH, sizeof(PlaneD) * maxPlanes + 16, sizeof(int) * maxPlanes + 16,
                                  sizeof(float) * 9 * kLaunchPlanes * maxLaunches, sizeof(int) * CH + 16, 0)) return rcp;
    for (auto& s : h->slots) {
        RF_CUDA(h, cudaMallocHost(&s.img, sizeof(ImgParams) * CH));
        RF_CUDA(h, cudaMallocHost(&s.ctf, sizeof(CtfConsts) * CH));
        RF_CUDA(h, cudaMallocHost(&s.planesD, sizeof(PlaneD) * maxPlanes + 16));
        RF_CUDA(h, cudaMallocHost(&s.planeImg, sizeof(int) * maxPlanes + 16));
        {
            RF_CUDA(h, cudaMallocHost(&s.planesDp, sizeof(PlaneD) * maxPlanes + 16));
            RF_CUDA(h, cudaMallocHost(&s.planesS, sizeof(PlaneS) * maxPlanes + 16));
            RF_CUDA(h, cudaMallocHost(&s.soaP, sizeof(float) * 9 * kLaunchPlanes * maxLaunches));
            RF_CUDA(h, cudaMallocHost(&s.imgPlane0, sizeof(int) * CH + 16));
        }
        RF_CUDA(h, cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    }
    {
        // TMA descriptor of the slices: floats of a row (pitch entries x 4), rows, images; box = 16 entries x 16 rows
        std::memset(&h->sliceMap, 0, sizeof h->sliceMap);
        typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && fn && qres == cudaDriverEntryPointSuccess) {
            const cuuint64_t dims[3] = {(cuuint64_t)g.pitch * 4, (cuuint64_t)g.side, (cuuint64_t)CH};
            const cuuint64_t strides[2] = {(cuuint64_t)g.pitch * 16, (cuuint64_t)g.planeStride * 16};
            const cuuint32_t box[3] = {(cuuint32_t)kPfBoxCols * 4, (cuuint32_t)kPfBoxRows, 1}, es[3] = {1, 1, 1};
            h->haveSliceMap = ((EncodeFn)fn)(&h->sliceMap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, h->dSlices2, dims, strides, box, es,
                                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
        } else
            cudaGetLastError();
#ifdef RF_L2_PREFETCH
        if (!h->haveSliceMap) return fail(h, RFB200_ERR_CUDA, "cuTensorMapEncodeTiled failed for the slice buffer");
#endif
    }
    // persistent grid: one CTA per SM
    h->stickGrid = prop.multiProcessorCount;
    RF_CUDA(h, cudaDeviceSynchronize());
    return RFB200_OK;
}

}  // namespace

// =============================================================================================
extern "C" {

int rfb200_create(const rfb200_config* cfg, rfb200_handle* out) {
    if (!out) return RFB200_ERR_ARG;
    *out = nullptr;
    std::string why;
    int rc = validate(cfg, why);
    if (rc) {
        std::lock_guard<std::mutex> g(g_errMutex);
        g_createError = why;
        return rc;
    }
    rfb200_handle h = new rfb200_handle_s();
    h->cfg = *cfg;               // the caller's sym_matrices pointer is only read inside do_create
    rc = do_create(h);
    h->cfg.sym_matrices = nullptr;
    if (rc) {
        {
            std::lock_guard<std::mutex> g(g_errMutex);
            g_createError = h->err;
        }
        free_all(h);
        return rc;
    }
    *out = h;
    return RFB200_OK;
}

void rfb200_destroy(rfb200_handle h) { free_all(h); }

const char* rfb200_last_error(rfb200_handle h) {
    if (h) return h->err.c_str();
    std::lock_guard<std::mutex> g(g_errMutex);
    static thread_local std::string copy;
    copy = g_createError;
    return copy.c_str();
}

int rfb200_get_info(rfb200_handle h, rfb200_info* info) {
    if (!h || !info) return RFB200_ERR_ARG;
    const Geometry& g = h->geo;
    info->N = g.N; info->P = g.P; info->Z = g.Z; info->X = g.X;
    info->tiles_x = g.tx; info->tiles_y = g.ty; info->tiles_z = g.tz; info->tile = kTileX;
    info->n_blocked = h->nBlocked;
    info->chunk_images = h->chunkImages;
    info->n_tiles_active = h->nUnits[0] + h->nUnits[1] + h->nUnits[2];   // work units (sticks) of the gather
    info->n_edge_items = h->nEdge;
    if (h->fast) {          // --fast: Z = padded size, tile = S + 1 (edge of the temporary volume), no blocked layout
        info->tiles_x = info->tiles_y = info->tiles_z = 1;
        info->tile = h->fgeo.S + 1;
    }
    return RFB200_OK;
}

int rfb200_insert_batch(rfb200_handle h, const float* images, const rfb200_particle* meta, int32_t n) {
    if (!h || (n > 0 && (!images || !meta)) || n < 0) return RFB200_ERR_ARG;
    RF_CUDA(h, cudaSetDevice(h->cfg.device));
    static const bool trace = getenv("RFB200_TRACE") != nullptr;
    if (!trace) drain_timings(h);
    const Geometry& g = h->geo;
    const size_t imgElems = (size_t)g.N * g.N;
    int buf = 0;
    for (int i0 = 0; i0 < n; i0 += h->chunkImages, buf ^= 1) {
        int cnt = std::min(h->chunkImages, n - i0);
        if (h->rawBusy[buf]) RF_CUDA(h, cudaStreamWaitEvent(h->copy, h->evRawFree[buf], 0));
        {
            StageTimer t(h, Stage::H2D, h->copy);
            RF_CUDA(h, cudaMemcpyAsync(h->dRaw[buf], images + (size_t)i0 * imgElems, sizeof(float) * imgElems * cnt,
                                       cudaMemcpyHostToDevice, h->copy));
        }
        RF_CUDA(h, cudaEventRecord(h->evH2D[buf], h->copy));
        RF_CUDA(h, cudaStreamWaitEvent(h->compute, h->evH2D[buf], 0));
        int rc = process_chunk(h, h->dRaw[buf], meta + i0, cnt);
        if (rc) return rc;
        RF_CUDA(h, cudaEventRecord(h->evRawFree[buf], h->compute));
        h->rawBusy[buf] = true;
    }
    // the caller may reuse `images` once every H2D copy has completed
    RF_CUDA(h, cudaStreamSynchronize(h->copy));
    return RFB200_OK;
}

int rfb200_insert_batch_device(rfb200_handle h, const float* d_images, const rfb200_particle* meta, int32_t n) {
    if (!h || (n > 0 && (!d_images || !meta)) || n < 0) return RFB200_ERR_ARG;
    RF_CUDA(h, cudaSetDevice(h->cfg.device));
    if (getenv("RFB200_TRACE") == nullptr) drain_timings(h);
    const size_t imgElems = (size_t)h->geo.N * h->geo.N;
    for (int i0 = 0; i0 < n; i0 += h->chunkImages) {
        int cnt = std::min(h->chunkImages, n - i0);
        int rc = process_chunk(h, d_images + (size_t)i0 * imgElems, meta + i0, cnt);
        if (rc) return rc;
    }
    return RFB200_OK;
}

int rfb200_sync(rfb200_handle h) {
    if (!h) return RFB200_ERR_ARG;
    RF_CUDA(h, cudaSetDevice(h->cfg.device));
    RF_CUDA(h, cudaStreamSynchronize(h->copy));
    RF_CUDA(h, cudaStreamSynchronize(h->compute));
    resolve_timings(h);
    return RFB200_OK;
}

int rfb200_reset(rfb200_handle h) {
    if (!h) return RFB200_ERR_ARG;
    int rc = rfb200_sync(h);
    if (rc) return rc;
    RF_CUDA(h, cudaMemsetAsync(h->dVb, 0, sizeof(float2) * h->nBlocked, h->compute));
    RF_CUDA(h, cudaMemsetAsync(h->dWb, 0, sizeof(float) * h->nBlocked, h->compute));
    if (h->dWb2) RF_CUDA(h, cudaMemsetAsync(h->dWb2, 0, sizeof(float) * h->nBlocked, h->compute));
    if (h->dD) RF_CUDA(h, cudaMemsetAsync(h->dD, 0, sizeof(unsigned long long) * h->nBlocked, h->compute));
    if (h->dD2) RF_CUDA(h, cudaMemsetAsync(h->dD2, 0, sizeof(unsigned long long) * h->nBlocked, h->compute));
    if (h->dFastAcc) RF_CUDA(h, cudaMemsetAsync(h->dFastAcc, 0, sizeof(float4) * h->nBlocked, h->compute));
    h->fastDirty = false;
    h->dampedDirty = false;
    RF_CUDA(h, cudaStreamSynchronize(h->compute));
    for (double& m : h->ms) m = 0;
    h->nImages = h->nPlanes = h->nGatherLaunches = h->nKernelLaunches = 0;
    return RFB200_OK;
}

int rfb200_nccl_unique_id(void* id128) {
#if RFB200_HAVE_NCCL_H
    if (!id128) return RFB200_ERR_ARG;
    std::call_once(g_ncclOnce, load_nccl);
    if (!g_nccl.ok) return RFB200_ERR_NCCL;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return RFB200_ERR_NCCL;
    std::memcpy(id128, &id, 128);
    return RFB200_OK;
#else
    (void)id128;
    return RFB200_ERR_NCCL;
#endif
}

int rfb200_nccl_init(rfb200_handle h, const void* id128, int32_t n_ranks, int32_t rank) {
#if RFB200_HAVE_NCCL_H
    if (!h || !id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) return RFB200_ERR_ARG;
    std::call_once(g_ncclOnce, load_nccl);
    if (!g_nccl.ok) return fail(h, RFB200_ERR_NCCL, "libnccl.so.2 could not be loaded");
    RF_CUDA(h, cudaSetDevice(h->cfg.device));
    ncclUniqueId id;
    std::memcpy(&id, id128, 128);
    ncclResult_t r = g_nccl.CommInitRank(&h->comm, n_ranks, id, rank);
    if (r != ncclSuccess) return fail(h, RFB200_ERR_NCCL, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error"));
    h->nRanks = n_ranks;
    h->rank = rank;
    return RFB200_OK;
#else
    (void)id128; (void)n_ranks; (void)rank;
    return h ? fail(h, RFB200_ERR_NCCL, "built without nccl.h") : RFB200_ERR_NCCL;
#endif
}

int rfb200_reduce_nccl(rfb200_handle h, int32_t root) {
#if RFB200_HAVE_NCCL_H
    if (!h) return RFB200_ERR_ARG;
    if (!h->comm) return fail(h, RFB200_ERR_STATE, "rfb200_nccl_init has not been called");
    RF_CUDA(h, cudaSetDevice(h->cfg.device));
    if (int rcf = flush_deficit(h)) return rcf;
    {
        StageTimer t(h, Stage::REDUCE, h->compute);
        ncclResult_t r1 = g_nccl.Reduce(h->dVb, h->dVb, (size_t)h->nBlocked * 2, ncclFloat, ncclSum, root, h->comm, h->compute);
        ncclResult_t r2 = g_nccl.Reduce(h->dWb, h->dWb, (size_t)h->nBlocked, ncclFloat, ncclSum, root, h->comm, h->compute);
        if (h->dWb2 && r2 == ncclSuccess) r2 = g_nccl.Reduce(h->dWb2, h->dWb2, (size_t)h->nBlocked, ncclFloat, ncclSum, root, h->comm, h->compute);
        if (r1 != ncclSuccess || r2 != ncclSuccess) return fail(h, RFB200_ERR_NCCL, "ncclReduce failed");
    }
    return RFB200_OK;
#else
    (void)root;
    return h ? fail(h, RFB200_ERR_NCCL, "built without nccl.h") : RFB200_ERR_NCCL;
#endif
}

// ---- peer-memory reduce
namespace {
struct IpcBlob {                      // RFB200_IPC_BYTES = 256
    cudaIpcMemHandle_t V, W, W2;      // 3 x 64 bytes
    int64_t nBlocked;
    int32_t hasW2, device;
    char pad[256 - 3 * 64 - 16];
};
static_assert(sizeof(IpcBlob) == 256, "IPC blob layout");

int launch_p2p(rfb200_handle h, const P2PArgs& a) {
    const long long n = a.hi - a.lo;
    if (n <= 0) return RFB200_OK;
    const int grid = (int)std::min<long long>((n + 255) / 256, 148 * 8);
    switch (h->nRanks) {
        case 2: k_reduce_p2p<2><<<grid, 256, 0, h->compute>>>(a, 2); break;
        case 4: k_reduce_p2p<4><<<grid, 256, 0, h->compute>>>(a, 4); break;
        case 8: k_reduce_p2p<8><<<grid, 256, 0, h->compute>>>(a, 8); break;
        default: k_reduce_p2p<0><<<grid, 256, 0, h->compute>>>(a, h->nRanks); break;
    }
    RF_CUDA(h, cudaGetLastError());
    h->nKernelLaunches += 1;
    return RFB200_OK;
}
// stream-ordered barrier over the ranks of the communicator: a 4-byte all-reduce on the compute stream
int rank_barrier(rfb200_handle h) {
#if RFB200_HAVE_NCCL_H
    if (!h->dBarrier) {
        RF_CUDA(h, cudaMalloc(&h->dBarrier, 16));
        RF_CUDA(h, cudaMemsetAsync(h->dBarrier, 0, 16, h->compute));
    }
    if (g_nccl.AllReduce(h->dBarrier, h->dBarrier, 1, ncclFloat, ncclSum, h->comm, h->compute) != ncclSuccess)
        return fail(h, RFB200_ERR_NCCL, "ncclAllReduce (rank barrier) failed");
    return RFB200_OK;
#else
    return fail(h, RFB200_ERR_NCCL, "built without nccl.h");
#endif
}
}  // namespace

int rfb200_ipc_export(rfb200_handle h, void* out256) {
    if (!h || !out256) return RFB200_ERR_ARG;
    // the kernel moves 16-byte vectors: the blocked layout of the exact path is a multiple of 2048 voxels, the (S+1)^3
    // temporary volume of --fast need not be; callers then stay with ncclReduce (the CLI's auto mode does by itself)
    if (h->nBlocked % 4 != 0) return fail(h, RFB200_ERR_UNSUPPORTED, "peer-memory reduce: the accumulators of this handle are not a multiple of 4 elements (--fast); use rfb200_reduce_nccl");
    RF_CUDA(h, cudaSetDevice(h->cfg.device));
    IpcBlob b;
    std::memset(&b, 0, sizeof b);
    RF_CUDA(h, cudaIpcGetMemHandle(&b.V, h->dVb));
    RF_CUDA(h, cudaIpcGetMemHandle(&b.W, h->dWb));
    if (h->dWb2) RF_CUDA(h, cudaIpcGetMemHandle(&b.W2, h->dWb2));
    b.nBlocked = (int64_t)h->nBlocked;
    b.hasW2 = h->dWb2 ? 1 : 0;
    b.device = h->cfg.device;
    std::memcpy(out256, &b, sizeof b);
    return RFB200_OK;
}

int rfb200_ipc_import(rfb200_handle h, int32_t rank, const void* in256) {
    if (!h || !in256 || rank < 0 || rank >= kMaxP2PRanks) return RFB200_ERR_ARG;
    if (h->nBlocked % 4 != 0) return fail(h, RFB200_ERR_UNSUPPORTED, "peer-memory reduce: not available for --fast handles");
    RF_CUDA(h, cudaSetDevice(h->cfg.device));
    IpcBlob b;
    std::memcpy(&b, in256, sizeof b);
    if (b.nBlocked != (int64_t)h->nBlocked || (b.hasW2 != 0) != (h->dWb2 != nullptr))
        return fail(h, RFB200_ERR_ARG, "rfb200_ipc_import: the peer's handle has a different geometry");
    if ((int)h->peers.size() <= rank) h->peers.resize(rank + 1);
    rfb200_handle_s::Peer& p = h->peers[rank];
    if (p.open) return fail(h, RFB200_ERR_STATE, "rfb200_ipc_import: this rank has been imported already");
    void *v = nullptr, *w = nullptr, *w2 = nullptr;
    RF_CUDA(h, cudaIpcOpenMemHandle(&v, b.V, cudaIpcMemLazyEnablePeerAccess));
    RF_CUDA(h, cudaIpcOpenMemHandle(&w, b.W, cudaIpcMemLazyEnablePeerAccess));
    if (b.hasW2) RF_CUDA(h, cudaIpcOpenMemHandle(&w2, b.W2, cudaIpcMemLazyEnablePeerAccess));
    p.V = (float2*)v; p.W = (float*)w; p.W2 = (float*)w2; p.open = true;
    return RFB200_OK;
}

int rfb200_ipc_release(rfb200_handle h) {
    if (!h) return RFB200_ERR_ARG;
    RF_CUDA(h, cudaSetDevice(h->cfg.device));
    RF_CUDA(h, cudaStreamSynchronize(h->compute));
    for (auto& p : h->peers) {
        if (!p.open) continue;
        cudaIpcCloseMemHandle(p.V);
        cudaIpcCloseMemHandle(p.W);
        if (p.W2) cudaIpcCloseMemHandle(p.W2);
        p = rfb200_handle_s::Peer{};
    }
    return RFB200_OK;
}

namespace {
int check_peers(rfb200_handle h, int32_t root) {
    if (root < 0 || root >= h->nRanks || h->nRanks > kMaxP2PRanks) return RFB200_ERR_ARG;
    for (int k = 0; k < h->nRanks; ++k)
        if (k != h->rank && ((int)h->peers.size() <= k || !h->peers[k].open))
            return fail(h, RFB200_ERR_STATE, "peer-memory reduce: rfb200_ipc_import has not been called for every other rank");
    return RFB200_OK;
}
// the kernels of the peer-memory reduce: this rank's slice of V, W (and the un-modulated weights) out of every rank's memory
int p2p_kernels(rfb200_handle h, int32_t root) {
    auto one = [&](auto member, void* mine, size_t nFloat4) -> int {
        P2PArgs a{};
        for (int k = 0; k < h->nRanks; ++k) a.src[k] = reinterpret_cast<const float4*>(k == h->rank ? mine : (void*)(h->peers[k].*member));
        a.dst = reinterpret_cast<float4*>(root == h->rank ? mine : (void*)(h->peers[root].*member));
        a.lo = (long long)(nFloat4 * (size_t)h->rank / (size_t)h->nRanks);
        a.hi = (long long)(nFloat4 * (size_t)(h->rank + 1) / (size_t)h->nRanks);
        return launch_p2p(h, a);
    };
    int rc = one(&rfb200_handle_s::Peer::V, h->dVb, h->nBlocked / 2);
    if (!rc) rc = one(&rfb200_handle_s::Peer::W, h->dWb, h->nBlocked / 4);
    if (!rc && h->dWb2) rc = one(&rfb200_handle_s::Peer::W2, h->dWb2, h->nBlocked / 4);
    return rc;
}
}  // namespace

int rfb200_reduce_p2p(rfb200_handle h, int32_t root) {
#if RFB200_HAVE_NCCL_H
    if (!h) return RFB200_ERR_ARG;
    if (!h->comm) return fail(h, RFB200_ERR_STATE, "rfb200_nccl_init has not been called (its communicator orders the ranks)");
    if (!g_nccl.AllReduce) return fail(h, RFB200_ERR_NCCL, "ncclAllReduce not found in libnccl");
    if (int rcp = check_peers(h, root)) return rcp;
    RF_CUDA(h, cudaSetDevice(h->cfg.device));
    if (int rcf = flush_deficit(h)) return rcf;
    StageTimer t(h, Stage::REDUCE, h->compute);
    // every rank has finished inserting (and folding its damped weights) before anybody reads its accumulators
    if (int rcb = rank_barrier(h)) return rcb;
    if (int rc = p2p_kernels(h, root)) return rc;
    // the root may use the sums, and nobody may change its accumulators, once every rank's slice has been written
    return rank_barrier(h);
#else
    (void)root;
    return h ? fail(h, RFB200_ERR_NCCL, "built without nccl.h") : RFB200_ERR_NCCL;
#endif
}

// ---- the same reduce for host programs that order their ranks themselves (MPI_Barrier, files): no NCCL involved
int rfb200_set_ranks(rfb200_handle h, int32_t n_ranks, int32_t rank) {
    if (!h || n_ranks < 1 || n_ranks > kMaxP2PRanks || rank < 0 || rank >= n_ranks) return RFB200_ERR_ARG;
#if RFB200_HAVE_NCCL_H
    if (h->comm) return fail(h, RFB200_ERR_STATE, "rfb200_set_ranks: the handle already belongs to an NCCL communicator");
#endif
    h->nRanks = n_ranks;
    h->rank = rank;
    return RFB200_OK;
}

int rfb200_reduce_p2p_prepare(rfb200_handle h) {
    if (!h) return RFB200_ERR_ARG;
    RF_CUDA(h, cudaSetDevice(h->cfg.device));
    if (int rcf = flush_deficit(h)) return rcf;
    RF_CUDA(h, cudaStreamSynchronize(h->compute));       // this rank's accumulators are final
    return RFB200_OK;
}

int rfb200_reduce_p2p_run(rfb200_handle h, int32_t root) {
    if (!h) return RFB200_ERR_ARG;
    if (int rcp = check_peers(h, root)) return rcp;
    RF_CUDA(h, cudaSetDevice(h->cfg.device));
    {
        StageTimer t(h, Stage::REDUCE, h->compute);
        if (int rc = p2p_kernels(h, root)) return rc;
    }
    RF_CUDA(h, cudaStreamSynchronize(h->compute));       // this rank's slice has been written into the root's memory
    return RFB200_OK;
}

int rfb200_accumulator_ptrs(rfb200_handle h, void** d_V, void** d_W, int64_t* n_blocked) {
    if (!h) return RFB200_ERR_ARG;
    if (int rcf = flush_deficit(h)) return rcf;
    if (d_V) *d_V = h->dVb;
    if (d_W) *d_W = h->dWb;
    if (n_blocked) *n_blocked = h->nBlocked;
    return RFB200_OK;
}

int rfb200_export_accumulators(rfb200_handle h, float* V, float* W) {
    if (!h || !V || !W) return RFB200_ERR_ARG;
    RF_CUDA(h, cudaSetDevice(h->cfg.device));
    if (h->fast) {
        // --fast: the (S+1)^3 temporary volume and weights [z][y][x], as the reference copies them back (copyTempVolumes)
        if (int rcf = flush_deficit(h)) return rcf;
        RF_CUDA(h, cudaMemcpyAsync(V, h->dVb, sizeof(float2) * h->nBlocked, cudaMemcpyDeviceToHost, h->compute));
        RF_CUDA(h, cudaMemcpyAsync(W, h->dWb, sizeof(float) * h->nBlocked, cudaMemcpyDeviceToHost, h->compute));
        RF_CUDA(h, cudaStreamSynchronize(h->compute));
        return RFB200_OK;
    }
    const Geometry& g = h->geo;
    size_t total = (size_t)g.Z * g.Z * g.X;
    if (int rcf = flush_deficit(h)) return rcf;
    float2* dV = nullptr;
    float* dW = nullptr;
    RF_CUDA(h, cudaMalloc(&dV, sizeof(float2) * total));
    RF_CUDA(h, cudaMalloc(&dW, sizeof(float) * total));
    k_export<<<(unsigned)((total + 255) / 256), 256, 0, h->compute>>>(g, h->dVb, h->dWb, dV, dW);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(V, dV, sizeof(float2) * total, cudaMemcpyDeviceToHost, h->compute);
    if (e == cudaSuccess) e = cudaMemcpyAsync(W, dW, sizeof(float) * total, cudaMemcpyDeviceToHost, h->compute);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->compute);
    cudaFree(dV);
    cudaFree(dW);
    if (e != cudaSuccess) return fail(h, RFB200_ERR_CUDA, std::string("export failed: ") + cudaGetErrorString(e));
    return RFB200_OK;
}

// plan and buffers of the finalisation (created on first use, or ahead of time by rfb200_warmup)
static int prepare_finalize(rfb200_handle h) {
    const Geometry& g = h->geo;
    size_t nHalf = (size_t)g.Z * g.Z * g.X, nVol = (size_t)g.Z * g.Z * g.Z, nOut = (size_t)g.N * g.N * g.N;
    if (!h->havePlan3d) {
        RF_CUFFT(h, cufftPlan3d(&h->plan3d, g.Z, g.Z, g.Z, CUFFT_C2R));
        RF_CUFFT(h, cufftSetStream(h->plan3d, h->compute));
        h->havePlan3d = true;
    }
    if (!h->dNorm) RF_CUDA(h, cudaMalloc(&h->dNorm, sizeof(float2) * nHalf));
    if (!h->dVol) RF_CUDA(h, cudaMalloc(&h->dVol, sizeof(float) * nVol));
    if (!h->dOut) RF_CUDA(h, cudaMalloc(&h->dOut, sizeof(float) * nOut));
    return RFB200_OK;
}

int rfb200_warmup(rfb200_handle h) {
    if (!h) return RFB200_ERR_ARG;
    RF_CUDA(h, cudaSetDevice(h->cfg.device));
    if (!h->fast)
        if (int rc = prepare_finalize(h)) return rc;
#if RFB200_HAVE_NCCL_H
    if (h->comm && g_nccl.ok) {
        // a 4-byte reduce on a private stream: every rank calls this once, before its first real reduce
        cudaStream_t s = nullptr;
        float* d = nullptr;
        RF_CUDA(h, cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        cudaError_t e = cudaMalloc(&d, 16);
        if (e == cudaSuccess) e = cudaMemsetAsync(d, 0, 16, s);
        ncclResult_t r = e == cudaSuccess ? g_nccl.Reduce(d, d, 1, ncclFloat, ncclSum, 0, h->comm, s) : ncclSuccess;
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (d) cudaFree(d);
        cudaStreamDestroy(s);
        if (e != cudaSuccess) return fail(h, RFB200_ERR_CUDA, std::string("warm-up failed: ") + cudaGetErrorString(e));
        if (r != ncclSuccess) return fail(h, RFB200_ERR_NCCL, "warm-up ncclReduce failed");
    }
#endif
    return RFB200_OK;
}

int rfb200_finalize(rfb200_handle h, float* out) {
    if (!h || !out) return RFB200_ERR_ARG;
    RF_CUDA(h, cudaSetDevice(h->cfg.device));
    if (h->fast) return finalize_fast(h, out);
    const Geometry& g = h->geo;
    const rfb200_config& c = h->cfg;
    size_t nHalf = (size_t)g.Z * g.Z * g.X, nOut = (size_t)g.N * g.N * g.N;
    if (int rcp = prepare_finalize(h)) return rcp;
    if (int rcf = flush_deficit(h)) return rcf;
    {
        StageTimer t(h, Stage::FINALIZE, h->compute);
        NormArgs a{};
        a.geo = g;
        a.Vb = h->dVb; a.Wb = h->dWb; a.Wb2 = h->dWb2 ? h->dWb2 : h->dWb; a.out = h->dNorm;
        a.corr = (float)(std::pow(c.pad_proj, 2.0) / (c.img_size * std::pow(c.pad_vol, 3.0)));   // RF.cpp:457-458
        a.nIterWeight = c.n_iter_weight;
        k_normalize<<<(unsigned)((nHalf + 255) / 256), 256, 0, h->compute>>>(a);
        RF_CUDA(h, cudaGetLastError());
        RF_CUFFT(h, cufftExecC2R(h->plan3d, reinterpret_cast<cufftComplex*>(h->dNorm), h->dVol));
        k_crop_correct<<<(unsigned)((nOut + 255) / 256), 256, 0, h->compute>>>(h->dVol, h->dG, h->dOut, g.N, g.Z);
        RF_CUDA(h, cudaGetLastError());
        RF_CUDA(h, cudaMemcpyAsync(out, h->dOut, sizeof(float) * nOut, cudaMemcpyDeviceToHost, h->compute));
    }
    h->nKernelLaunches += 2;
    RF_CUDA(h, cudaStreamSynchronize(h->compute));
    resolve_timings(h);
    return RFB200_OK;
}

int rfb200_get_timings(rfb200_handle h, rfb200_timings* t) {
    if (!h || !t) return RFB200_ERR_ARG;
    int rc = rfb200_sync(h);
    if (rc) return rc;
    t->h2d_ms = h->ms[Stage::H2D];
    t->preprocess_ms = h->ms[Stage::PAD];
    t->fft2d_ms = h->ms[Stage::FFT2D];
    t->slice_ms = h->ms[Stage::SLICE];
    t->gather_ms = h->ms[Stage::GATHER];
    t->edge_ms = h->ms[Stage::EDGE];
    t->finalize_ms = h->ms[Stage::FINALIZE];
    t->reduce_ms = h->ms[Stage::REDUCE];
    t->images = h->nImages;
    t->planes = h->nPlanes;
    t->gather_launches = h->nGatherLaunches;
    t->kernel_launches = h->nKernelLaunches;
    return RFB200_OK;
}

int rfb200_halfset_push(rfb200_handle h) {
    if (!h) return RFB200_ERR_ARG;
    RF_CUDA(h, cudaSetDevice(h->cfg.device));
    if (int rcf = flush_deficit(h)) return rcf;
    if (!h->dVsaved) {
        RF_CUDA(h, cudaMalloc(&h->dVsaved, sizeof(float2) * h->nBlocked));
        RF_CUDA(h, cudaMalloc(&h->dWsaved, sizeof(float) * h->nBlocked));
        if (h->dWb2) RF_CUDA(h, cudaMalloc(&h->dW2saved, sizeof(float) * h->nBlocked));
    }
    RF_CUDA(h, cudaMemcpyAsync(h->dVsaved, h->dVb, sizeof(float2) * h->nBlocked, cudaMemcpyDeviceToDevice, h->compute));
    RF_CUDA(h, cudaMemcpyAsync(h->dWsaved, h->dWb, sizeof(float) * h->nBlocked, cudaMemcpyDeviceToDevice, h->compute));
    RF_CUDA(h, cudaMemsetAsync(h->dVb, 0, sizeof(float2) * h->nBlocked, h->compute));
    RF_CUDA(h, cudaMemsetAsync(h->dWb, 0, sizeof(float) * h->nBlocked, h->compute));
    if (h->dWb2) {
        RF_CUDA(h, cudaMemcpyAsync(h->dW2saved, h->dWb2, sizeof(float) * h->nBlocked, cudaMemcpyDeviceToDevice, h->compute));
        RF_CUDA(h, cudaMemsetAsync(h->dWb2, 0, sizeof(float) * h->nBlocked, h->compute));
    }
    return RFB200_OK;
}

int rfb200_halfset_merge(rfb200_handle h) {
    if (!h) return RFB200_ERR_ARG;
    if (!h->dVsaved) return fail(h, RFB200_ERR_STATE, "rfb200_halfset_push has not been called");
    RF_CUDA(h, cudaSetDevice(h->cfg.device));
    if (int rcf = flush_deficit(h)) return rcf;
    const int64_t n = h->nBlocked;
    k_axpy<<<2048, 256, 0, h->compute>>>(reinterpret_cast<float*>(h->dVb), reinterpret_cast<const float*>(h->dVsaved), 2 * n);
    k_axpy<<<2048, 256, 0, h->compute>>>(h->dWb, h->dWsaved, n);
    if (h->dWb2) k_axpy<<<2048, 256, 0, h->compute>>>(h->dWb2, h->dW2saved, n);
    RF_CUDA(h, cudaGetLastError());
    h->nKernelLaunches += h->dWb2 ? 3 : 2;
    return RFB200_OK;
}

int rfb200_timer_start(rfb200_handle h) {
    if (!h) return RFB200_ERR_ARG;
    RF_CUDA(h, cudaSetDevice(h->cfg.device));
    // make the compute stream wait for any copy still in flight so that the stopwatch starts "behind everything"
    RF_CUDA(h, cudaStreamSynchronize(h->copy));
    RF_CUDA(h, cudaEventRecord(h->swStart, h->compute));
    h->swStarted = true;
    return RFB200_OK;
}

int rfb200_timer_stop(rfb200_handle h, double* elapsed_ms) {
    if (!h || !elapsed_ms) return RFB200_ERR_ARG;
    RF_CUDA(h, cudaSetDevice(h->cfg.device));
    RF_CUDA(h, cudaStreamSynchronize(h->copy));
    RF_CUDA(h, cudaEventRecord(h->swStop, h->compute));
    RF_CUDA(h, cudaEventSynchronize(h->swStop));
    float ms = 0;
    RF_CUDA(h, cudaEventElapsedTime(&ms, h->swStart, h->swStop));
    *elapsed_ms = ms;
    return RFB200_OK;
}

int rfb200_weight_sum_begin(rfb200_handle h) {
    if (!h) return RFB200_ERR_ARG;
    RF_CUDA(h, cudaSetDevice(h->cfg.device));
    const bool scratch = h->fast && h->fastDirty;       // --fast: add the W lane of the scratch instead of flushing it per step
    if (!scratch)
        if (int rcf = flush_deficit(h)) return rcf;
    k_weight_sum_partial<<<1024, 256, 0, h->compute>>>(h->dWb, h->nBlocked, h->dSum,
                                                        scratch ? reinterpret_cast<const float*>(h->dFastAcc) : nullptr);
    RF_CUDA(h, cudaGetLastError());
    k_weight_sum_final<<<1, 256, 0, h->compute>>>(h->dSum, 1024, h->dSum + 1024);
    RF_CUDA(h, cudaGetLastError());
    h->nKernelLaunches += 2;
    RF_CUDA(h, cudaMemcpyAsync(h->hSum, h->dSum + 1024, sizeof(double), cudaMemcpyDeviceToHost, h->compute));
    RF_CUDA(h, cudaEventRecord(h->evSum, h->compute));
    h->sumPending = true;
    return RFB200_OK;
}

int rfb200_weight_sum_end(rfb200_handle h, double* sum) {
    if (!h || !sum) return RFB200_ERR_ARG;
    if (!h->sumPending) return fail(h, RFB200_ERR_STATE, "rfb200_weight_sum_begin has not been called");
    RF_CUDA(h, cudaSetDevice(h->cfg.device));
    RF_CUDA(h, cudaEventSynchronize(h->evSum));
    h->sumPending = false;
    *sum = *h->hSum;
    return RFB200_OK;
}

int rfb200_weight_sum(rfb200_handle h, double* sum) {
    int rc = rfb200_weight_sum_begin(h);
    if (rc) return rc;
    return rfb200_weight_sum_end(h, sum);
}

int rfb200_device_count(int32_t* n) {
    if (!n) return RFB200_ERR_ARG;
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) {
        cudaGetLastError();
        *n = 0;
        return RFB200_ERR_CUDA;
    }
    *n = c;
    return RFB200_OK;
}

namespace {
__global__ void __launch_bounds__(256) k_fma_peak(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    const float s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 12345.678f) out[0] = s;          // keeps the chains alive without a store per thread
}
}  // namespace

int rfb200_measure_fp32_peak(int32_t device, double* tflops) {
    if (!tflops) return RFB200_ERR_ARG;
    *tflops = 0;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return RFB200_ERR_CUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return RFB200_ERR_CUDA;
    float* d = nullptr;
    if (cudaMalloc(&d, 256) != cudaSuccess) return RFB200_ERR_CUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int blocks = prop.multiProcessorCount * 8, iters = 4096;      // 8 x 256 threads per SM
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k_fma_peak<<<blocks, 256>>>(d, iters, 0.999f, 1e-3f);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) break;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 8 * 16 * (double)iters * 256.0 * blocks;
        if (rep > 0 && ms > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    if (cudaGetLastError() != cudaSuccess || best <= 0) return RFB200_ERR_CUDA;
    *tflops = best;
    return RFB200_OK;
}

int rfb200_host_alloc(void** ptr, size_t bytes) {
    if (!ptr) return RFB200_ERR_ARG;
    *ptr = nullptr;
    cudaError_t e = cudaHostAlloc(ptr, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *ptr = nullptr;
        return RFB200_ERR_CUDA;
    }
    return RFB200_OK;
}

int rfb200_host_free(void* ptr) {
    if (!ptr) return RFB200_OK;
    return cudaFreeHost(ptr) == cudaSuccess ? RFB200_OK : RFB200_ERR_CUDA;
}

int rfb200_get_streams(rfb200_handle h, void** compute_stream, void** copy_stream) {
    if (!h) return RFB200_ERR_ARG;
    if (compute_stream) *compute_stream = (void*)h->compute;
    if (copy_stream) *copy_stream = (void*)h->copy;
    return RFB200_OK;
}

int rfb200_debug_fast_fourier(rfb200_handle h, float* out) {
    if (!h || !out) return RFB200_ERR_ARG;
    if (!h->fast) return fail(h, RFB200_ERR_STATE, "handle was not created with cfg.fast");
    RF_CUDA(h, cudaSetDevice(h->cfg.device));
    std::vector<float> dummy(1);
    return finalize_fast(h, dummy.data(), out);
}

int rfb200_debug_slice_dims(rfb200_handle h, int32_t* side, int32_t* apron_radius) {
    if (!h) return RFB200_ERR_ARG;
    if (h->fast) return fail(h, RFB200_ERR_UNSUPPORTED, "no slices in --fast mode");
    if (side) *side = h->geo.side;
    if (apron_radius) *apron_radius = h->geo.Rp;
    return RFB200_OK;
}

int rfb200_debug_get_slice(rfb200_handle h, int32_t idx, float* out4) {
    if (!h || !out4 || idx < 0 || idx >= h->lastChunkImages) return RFB200_ERR_ARG;
    if (h->fast) return fail(h, RFB200_ERR_UNSUPPORTED, "no slices in --fast mode");
    RF_CUDA(h, cudaSetDevice(h->cfg.device));
    RF_CUDA(h, cudaStreamSynchronize(h->compute));
    // The full-plane view of the stored half-plane slice: (re, im) of pixel (i, j), columns j < -colOff from the Hermitian
    // mate; the third channel is rebuilt from the validity table (multiplicity of the pixel: 1 inside the resolution
    // disc, 2 on column 0, 0 outside)
    const Geometry& g = h->geo;
    std::vector<float4> A((size_t)g.planeStride);
    RF_CUDA(h, cudaMemcpy(A.data(), h->dSlices2 + (size_t)idx * g.planeStride, sizeof(float4) * A.size(), cudaMemcpyDeviceToHost));
    std::vector<int32_t> rim = host::build_rim_table(h->geo, h->jmax, h->iLo, h->iHi);
    for (int i = 0; i < g.side; ++i)
        for (int j = 0; j < g.side; ++j) {
            float4 v;
            if (j - g.Rp >= -g.colOff) v = A[(size_t)i * g.pitch + (j - g.Rp + g.colOff)];
            else {
                v = A[(size_t)(g.side - 1 - i) * g.pitch + (g.Rp - j + g.colOff)];
                v.y = -v.y;
            }
            const int rt = rim[i], jc = j - g.Rp;
            const int jPos = (rt & 0x3fff) - 1, jNeg = ((rt >> 14) & 0x3fff) - 1;
            float m = jc > 0 ? (jc <= jPos ? 1.f : 0.f) : (jc < 0 ? (-jc <= jNeg ? 1.f : 0.f) : (float)(rt >> 28));
            float* o = out4 + ((size_t)i * g.side + j) * 4;
            o[0] = v.x; o[1] = v.y; o[2] = m; o[3] = 0.f;
        }
    return RFB200_OK;
}

}  // extern "C"

#include "rf_projector.cuh"
