// rf_projector.cuh — central-slice (Fourier) projector on the GPU: the step BEFORE the reconstruction path in every
// workload.  It computes what FourierProjector computes (libraries/data/fourier_projection.cpp = "FP";
// used by xmipp_phantom_project, reconstruction/project.cpp:992, and by the continuous-alignment programs):
//   produceSideInfo FP:265-308        zero-pad the volume about its centre to P = padding*N, forward 3-D transform (1/size),
//                                     ShiftFFT by -P/2, CenterFFT, scale by P^3/N^2, cubic B-spline coefficients
//                                     (mirror-off-bounds) windowed to |index| <= maxFreq*P + 10 when degree = BSPLINE3
//   produceSideInfoProjection :310-333 phase ramp that moves the image origin back to the corner
//   project FP:91-262                 per half-plane pixel inside maxFrequency: volume frequency E^T (fx, fy, 0), NEAREST /
//                                     LINEAR / BSPLINE3 interpolation of the centred 3-D transform, phase ramp, optional CTF
//                                     image, inverse 2-D transform (c2r, unnormalised)
// The reference's own GPU twin is reconstruction_cuda/cuda_fourier_projection.cu:17-135 (one projection per call, texture
// fetches); here a batch of orientations is one gather kernel + one batched cuFFT C2R, the coefficient volume stays
// resident in HBM (P = 512: 1.07 GB) and is served from the 126 MB L2 (a slice touches a thin slab of it).
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>

#include <cmath>
#include <map>
#include <string>
#include <vector>

#include "../../include/recfourier_b200.h"
#include "rf_host.hpp"
#include "rf_kernels.cuh"

namespace rfb200 {

struct ProjGeo {
    int32_t N, P, hP, Xh;        // volume / image size, padded size, P/2, N/2+1
    int32_t degree;              // 0, 1, 3
    int32_t wStart, wDim;        // BSPLINE3: logical start and edge of the coefficient window (mirror bounds)
    double maxFreq2;
};

// ---- produceSideInfo
__global__ void __launch_bounds__(256) k_proj_pad(const float* __restrict__ vol, float* __restrict__ pad, int N, int P) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)N * N * N) return;
    const int j = (int)(idx % N);
    const size_t t = idx / N;
    const int i = (int)(t % N), k = (int)(t / N);
    const int off = P / 2 - N / 2;                                  // logical -N/2 sits at physical P/2 - N/2 (FP:270-272)
    pad[((size_t)(k + off) * P + (i + off)) * P + (j + off)] = vol[idx];
}

// C[lz][ly][lx] over the full logical cube [-hP, P-1-hP]^3: element k of the transform sits at logical l with k = l mod P
// (CenterFFT + setXmippOrigin, FP:279-280), times the ShiftFFT phase exp(2 pi i hP k / P) per axis (FP:278) and
// (1/P^3) * K = 1/N^2 (FP:283-285)
__global__ void __launch_bounds__(256) k_proj_center(const float2* __restrict__ F, float2* __restrict__ C, int P, int N) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)P * P * P) return;
    const int hP = P / 2, Xf = P / 2 + 1;
    const int x = (int)(idx % P);
    const size_t t = idx / P;
    const int y = (int)(t % P), z = (int)(t / P);
    const int kx = d_wrap(x - hP, P), ky = d_wrap(y - hP, P), kz = d_wrap(z - hP, P);
    float2 v;
    if (kx <= P / 2) v = __ldg(F + ((size_t)kz * P + ky) * Xf + kx);
    else {
        v = __ldg(F + ((size_t)((P - kz) % P) * P + (P - ky) % P) * Xf + (P - kx));
        v.y = -v.y;
    }
    const long long m = ((long long)hP * kx + (long long)hP * ky + (long long)hP * kz) % P;     // phase = 2 pi m / P
    double s, c;
    sincospi(2.0 * (double)m / (double)P, &s, &c);
    const double sc = 1.0 / ((double)N * (double)N);
    C[idx] = make_float2((float)(((double)v.x * c - (double)v.y * s) * sc), (float)(((double)v.x * s + (double)v.y * c) * sc));
}

// cubic B-spline direct transform of every line along one axis (bilib ChangeBasisVolume, mirror-off-bounds): one thread
// per line, FP64 recursion.  axis 0: x lines, 1: y lines, 2: z lines
__global__ void __launch_bounds__(128) k_proj_prefilter(float2* __restrict__ C, int P, int axis) {
    const size_t line = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (line >= (size_t)P * P) return;
    const size_t a = line / P, b = line % P;
    size_t base, stride;
    if (axis == 0) { base = (a * P + b) * P; stride = 1; }
    else if (axis == 1) { base = a * (size_t)P * P + b; stride = P; }
    else { base = a * P + b; stride = (size_t)P * P; }
    const double z = -0.26794919243112270647, lambda = 6.0;         // sqrt(3) - 2, (1 - z)(1 - 1/z)
    // causal initialisation over the half-sample mirrored signal: c+[0] = s[0] + z * sum_{m>=0} z^m s~[m]
    double sr = 0, si = 0, zm = 1.0;
    const int horizon = 80 > 2 * P ? 80 : 2 * P;
    for (int m = 0; m < horizon; ++m) {
        const int q = m % (2 * P);
        const int idx = q < P ? q : 2 * P - 1 - q;
        const float2 v = C[base + idx * stride];
        sr += zm * (double)v.x * lambda;
        si += zm * (double)v.y * lambda;
        zm *= z;
        if (m >= 80) break;
    }
    float2 v0 = C[base];
    double pr = (double)v0.x * lambda + z * sr, pi = (double)v0.y * lambda + z * si;
    C[base] = make_float2((float)pr, (float)pi);
    // the causal pass keeps FP64 state; the stored FP32 values are re-read by the anticausal pass
    for (int k = 1; k < P; ++k) {
        const float2 v = C[base + k * stride];
        pr = (double)v.x * lambda + z * pr;
        pi = (double)v.y * lambda + z * pi;
        C[base + k * stride] = make_float2((float)pr, (float)pi);
    }
    const double e = z / (z - 1.0);
    pr *= e;
    pi *= e;
    C[base + (size_t)(P - 1) * stride] = make_float2((float)pr, (float)pi);
    for (int k = P - 2; k >= 0; --k) {
        const float2 v = C[base + k * stride];
        pr = z * (pr - (double)v.x);
        pi = z * (pi - (double)v.y);
        C[base + k * stride] = make_float2((float)pr, (float)pi);
    }
}

// ---- project
struct ProjArgs {
    ProjGeo g;
    const float2* C;             // P^3 coefficients, logical origin at (hP, hP, hP)
    const double* E;             // per image: rows 0 and 1 of the Euler matrix (6 doubles)
    const float* ctf;            // per image N x Xh or nullptr
    float2* PF;                  // per image N x Xh
};

__device__ __forceinline__ float2 d_proj_fetch(const ProjArgs& a, int k, int i, int j) {      // logical indices, 0 outside
    const int P = a.g.P, hP = a.g.hP;
    const int z = k + hP, y = i + hP, x = j + hP;
    if ((unsigned)z >= (unsigned)P || (unsigned)y >= (unsigned)P || (unsigned)x >= (unsigned)P) return make_float2(0.f, 0.f);
    return __ldg(a.C + ((size_t)z * P + y) * P + x);
}
__device__ __forceinline__ float d_bspline03(float x) {
    float a = fabsf(x);
    if (a < 1.0f) return a * a * (a - 2.0f) * 0.5f + (2.0f / 3.0f);
    if (a < 2.0f) { a -= 2.0f; return a * a * a * (-1.0f / 6.0f); }
    return 0.f;
}

// value of half-plane pixel (i, j) before the Hermitian fix of the self-conjugate columns
__device__ __forceinline__ float2 d_proj_pixel(const ProjArgs& a, int img, int i, int j) {
    const ProjGeo& g = a.g;
    const double freqy = (i <= g.N / 2) ? (double)i / g.N : (double)(i - g.N) / g.N;          // FFT_IDX2DIGFREQ
    const double freqx = (double)j / g.N;
    if (freqy * freqy + freqx * freqx > g.maxFreq2) return make_float2(0.f, 0.f);             // FP:121-122
    const double* E = a.E + 6 * img;
    const double X = (E[3] * freqy + E[0] * freqx) * g.P, Y = (E[4] * freqy + E[1] * freqx) * g.P, Z = (E[5] * freqy + E[2] * freqx) * g.P;
    float2 v;
    if (g.degree == 0) {
        v = d_proj_fetch(a, (int)round(Z), (int)round(Y), (int)round(X));
    } else if (g.degree == 1) {
        const double x0d = floor(X), y0d = floor(Y), z0d = floor(Z);
        const int x0 = (int)x0d, y0 = (int)y0d, z0 = (int)z0d;
        const float fx = (float)(X - x0d), fy = (float)(Y - y0d), fz = (float)(Z - z0d);
        const float2 d000 = d_proj_fetch(a, z0, y0, x0), d001 = d_proj_fetch(a, z0, y0, x0 + 1);
        const float2 d010 = d_proj_fetch(a, z0, y0 + 1, x0), d011 = d_proj_fetch(a, z0, y0 + 1, x0 + 1);
        const float2 d100 = d_proj_fetch(a, z0 + 1, y0, x0), d101 = d_proj_fetch(a, z0 + 1, y0, x0 + 1);
        const float2 d110 = d_proj_fetch(a, z0 + 1, y0 + 1, x0), d111 = d_proj_fetch(a, z0 + 1, y0 + 1, x0 + 1);
#define RF_LIN(t, l, h) make_float2(fmaf((t), (h).x - (l).x, (l).x), fmaf((t), (h).y - (l).y, (l).y))
        const float2 dx00 = RF_LIN(fx, d000, d001), dx01 = RF_LIN(fx, d100, d101), dx10 = RF_LIN(fx, d010, d011), dx11 = RF_LIN(fx, d110, d111);
        const float2 dxy0 = RF_LIN(fy, dx00, dx10), dxy1 = RF_LIN(fy, dx01, dx11);
        v = RF_LIN(fz, dxy0, dxy1);
#undef RF_LIN
    } else {
        // coordinates relative to the coefficient window, mirror at its ends (FP:181-237)
        const double x = X - g.wStart, y = Y - g.wStart, z = Z - g.wStart;
        const int l1 = (int)ceil(x - 2), m1 = (int)ceil(y - 2), n1 = (int)ceil(z - 2);
        float wx[4], wy[4], wz[4];
        int ex[4], ey[4], ez[4];
        const int dim = g.wDim, off = g.wStart + g.hP;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int l = l1 + q, m = m1 + q, n = n1 + q;
            wx[q] = d_bspline03((float)(x - (double)l));
            wy[q] = d_bspline03((float)(y - (double)m));
            wz[q] = d_bspline03((float)(z - (double)n));
            ex[q] = (l < 0 ? -l - 1 : (l >= dim ? 2 * dim - l - 1 : l)) + off;
            ey[q] = (m < 0 ? -m - 1 : (m >= dim ? 2 * dim - m - 1 : m)) + off;
            ez[q] = (n < 0 ? -n - 1 : (n >= dim ? 2 * dim - n - 1 : n)) + off;
        }
        float accR = 0.f, accI = 0.f;
        const int P = g.P;
#pragma unroll
        for (int qz = 0; qz < 4; ++qz) {
            float yr = 0.f, yi = 0.f;
#pragma unroll
            for (int qy = 0; qy < 4; ++qy) {
                const float2* row = a.C + ((size_t)ez[qz] * P + ey[qy]) * P;
                float xr = 0.f, xi = 0.f;
#pragma unroll
                for (int qx = 0; qx < 4; ++qx) {
                    const float2 c = __ldg(row + ex[qx]);
                    xr = fmaf(c.x, wx[qx], xr);
                    xi = fmaf(c.y, wx[qx], xi);
                }
                yr = fmaf(xr, wy[qy], yr);
                yi = fmaf(xi, wy[qy], yi);
            }
            accR = fmaf(yr, wz[qz], accR);
            accI = fmaf(yi, wz[qz], accI);
        }
        v = make_float2(accR, accI);
    }
    // phase ramp exp(-2 pi i (N/2)(i + j)/N) (FP:317-331) and the optional CTF image (FP:243-248)
    const long long m = ((long long)(g.N / 2) * (i + j)) % g.N;
    float s, c;
    sincospif(-2.0f * (float)m / (float)g.N, &s, &c);
    if (a.ctf) {
        const float w = __ldg(a.ctf + ((size_t)img * g.N + i) * g.Xh + j);
        c *= w;
        s *= w;
    }
    return make_float2(c * v.x - s * v.y, c * v.y + s * v.x);
}

// grid (ceil(Xh/32), ceil(N/8), nImg), block (32, 8).  Columns j = 0 and j = N/2 (even N) are stored as their Hermitian
// part, the only part a c2r transform sees (FFTW drops the rest implicitly; cuFFT's behaviour is unspecified).
__global__ void __launch_bounds__(256) k_proj_slices(const __grid_constant__ ProjArgs a) {
    const ProjGeo& g = a.g;
    const int j = blockIdx.x * 32 + threadIdx.x, i = blockIdx.y * 8 + threadIdx.y, img = blockIdx.z;
    if (j >= g.Xh || i >= g.N) return;
    float2 v = d_proj_pixel(a, img, i, j);
    if (j == 0 || 2 * j == g.N) {
        const float2 m = d_proj_pixel(a, img, (g.N - i) % g.N, j);
        v.x = 0.5f * (v.x + m.x);
        v.y = 0.5f * (v.y - m.y);
    }
    a.PF[((size_t)img * g.N + i) * g.Xh + j] = v;
}

}  // namespace rfb200

// ------------------------------------------------------------------------------------------------ handle + C ABI
struct rfb200_projector_s {
    rfb200::ProjGeo g{};
    int device = 0;
    int chunk = 256;
    cudaStream_t stream = nullptr;
    float2* dC = nullptr;
    float2* dPF = nullptr;
    double* dE = nullptr;
    double* hE = nullptr;        // pinned
    float* dCtf = nullptr;
    float* dImg = nullptr;
    std::map<int, cufftHandle> plans;
    std::string err;
};

namespace rfb200 {
inline int proj_fail(rfb200_projector p, int code, const std::string& msg) {
    p->err = msg;
    return code;
}
#define RFP_CUDA(p, call)                                                                                       \
    do {                                                                                                        \
        cudaError_t e_ = (call);                                                                                \
        if (e_ != cudaSuccess) return proj_fail(p, RFB200_ERR_CUDA, std::string(#call " failed: ") + cudaGetErrorString(e_)); \
    } while (0)
#define RFP_CUFFT(p, call)                                                                                      \
    do {                                                                                                        \
        cufftResult r_ = (call);                                                                                \
        if (r_ != CUFFT_SUCCESS) return proj_fail(p, RFB200_ERR_CUDA, std::string(#call " failed: cufft error ") + std::to_string((int)r_)); \
    } while (0)

inline int proj_build(rfb200_projector p, const float* volume) {
    const ProjGeo& g = p->g;
    const int N = g.N, P = g.P;
    const size_t nVol = (size_t)N * N * N, nPad = (size_t)P * P * P, nHalf = (size_t)P * P * (P / 2 + 1);
    float *dVol = nullptr, *dPad = nullptr;
    float2* dF = nullptr;
    cufftHandle plan = 0;
    bool havePlan = false;
    auto body = [&]() -> int {
        RFP_CUDA(p, cudaMalloc(&dVol, sizeof(float) * nVol));
        RFP_CUDA(p, cudaMalloc(&dPad, sizeof(float) * nPad));
        RFP_CUDA(p, cudaMalloc(&dF, sizeof(float2) * nHalf));
        RFP_CUDA(p, cudaMalloc(&p->dC, sizeof(float2) * nPad));
        RFP_CUDA(p, cudaMemcpyAsync(dVol, volume, sizeof(float) * nVol, cudaMemcpyHostToDevice, p->stream));
        RFP_CUDA(p, cudaMemsetAsync(dPad, 0, sizeof(float) * nPad, p->stream));
        k_proj_pad<<<(unsigned)((nVol + 255) / 256), 256, 0, p->stream>>>(dVol, dPad, N, P);
        RFP_CUFFT(p, cufftPlan3d(&plan, P, P, P, CUFFT_R2C));
        havePlan = true;
        RFP_CUFFT(p, cufftSetStream(plan, p->stream));
        RFP_CUFFT(p, cufftExecR2C(plan, dPad, reinterpret_cast<cufftComplex*>(dF)));
        k_proj_center<<<(unsigned)((nPad + 255) / 256), 256, 0, p->stream>>>(dF, p->dC, P, N);
        if (g.degree == 3)
            for (int axis = 0; axis < 3; ++axis) k_proj_prefilter<<<(unsigned)(((size_t)P * P + 127) / 128), 128, 0, p->stream>>>(p->dC, P, axis);
        RFP_CUDA(p, cudaGetLastError());
        RFP_CUDA(p, cudaStreamSynchronize(p->stream));
        return RFB200_OK;
    };
    const int rc = body();          // the temporaries are released on every path
    if (havePlan) cufftDestroy(plan);
    cudaFree(dVol);
    cudaFree(dPad);
    cudaFree(dF);
    return rc;
}

inline int proj_run(rfb200_projector p, const double* angles, const float* ctf, bool ctfOnDevice, int n, float* images, bool imagesOnDevice) {
    const ProjGeo& g = p->g;
    const size_t nPix = (size_t)g.N * g.N, nHalf = (size_t)g.N * g.Xh;
    for (int i0 = 0; i0 < n; i0 += p->chunk) {
        const int cnt = std::min(p->chunk, n - i0);
        RFP_CUDA(p, cudaStreamSynchronize(p->stream));          // hE is reused
        for (int k = 0; k < cnt; ++k) {
            double A[9];
            host::euler_matrix(angles[3 * (i0 + k)], angles[3 * (i0 + k) + 1], angles[3 * (i0 + k) + 2], A);
            for (int q = 0; q < 6; ++q) p->hE[6 * k + q] = A[q];
        }
        RFP_CUDA(p, cudaMemcpyAsync(p->dE, p->hE, sizeof(double) * 6 * cnt, cudaMemcpyHostToDevice, p->stream));
        const float* dCtf = nullptr;
        if (ctf) {
            if (ctfOnDevice) dCtf = ctf + (size_t)i0 * nHalf;
            else {
                if (!p->dCtf) RFP_CUDA(p, cudaMalloc(&p->dCtf, sizeof(float) * nHalf * p->chunk));
                RFP_CUDA(p, cudaMemcpyAsync(p->dCtf, ctf + (size_t)i0 * nHalf, sizeof(float) * nHalf * cnt, cudaMemcpyHostToDevice, p->stream));
                dCtf = p->dCtf;
            }
        }
        ProjArgs a{};
        a.g = g; a.C = p->dC; a.E = p->dE; a.ctf = dCtf; a.PF = p->dPF;
        dim3 grid((g.Xh + 31) / 32, (g.N + 7) / 8, cnt);
        k_proj_slices<<<grid, dim3(32, 8), 0, p->stream>>>(a);
        RFP_CUDA(p, cudaGetLastError());
        auto it = p->plans.find(cnt);
        cufftHandle plan;
        if (it == p->plans.end()) {
            int dims[2] = {g.N, g.N};
            RFP_CUFFT(p, cufftPlanMany(&plan, 2, dims, nullptr, 1, 0, nullptr, 1, 0, CUFFT_C2R, cnt));
            RFP_CUFFT(p, cufftSetStream(plan, p->stream));
            p->plans[cnt] = plan;
        } else
            plan = it->second;
        float* dst = imagesOnDevice ? images + (size_t)i0 * nPix : p->dImg;
        RFP_CUFFT(p, cufftExecC2R(plan, reinterpret_cast<cufftComplex*>(p->dPF), dst));
        if (!imagesOnDevice)
            RFP_CUDA(p, cudaMemcpyAsync(images + (size_t)i0 * nPix, p->dImg, sizeof(float) * nPix * cnt, cudaMemcpyDeviceToHost, p->stream));
    }
    RFP_CUDA(p, cudaStreamSynchronize(p->stream));
    return RFB200_OK;
}
}  // namespace rfb200

extern "C" {

int rfb200_projector_create(const float* volume, int32_t N, double padding, double max_freq, int32_t degree, int32_t device,
                            rfb200_projector* out) {
    if (!out) return RFB200_ERR_ARG;
    *out = nullptr;
    if (!volume || N < 4 || N > 2048 || padding < 1.0 || !(max_freq > 0.0) || (degree != 0 && degree != 1 && degree != 3)) return RFB200_ERR_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return RFB200_ERR_CUDA;              // no CPU fallback
    }
    if (device < 0 || device >= ndev) return RFB200_ERR_ARG;
    rfb200_projector p = new rfb200_projector_s();
    p->device = device;
    rfb200::ProjGeo& g = p->g;
    g.N = N;
    g.P = (int)(padding * N);                // FP:268
    g.hP = g.P / 2;
    g.Xh = N / 2 + 1;
    g.degree = degree;
    g.maxFreq2 = max_freq * max_freq;
    g.wStart = -g.hP;
    g.wDim = g.P;
    if (degree == 3) {                       // FP:297-301
        int idxMax = (int)(max_freq * g.P + 10);
        idxMax = std::min(g.P - 1 - g.hP, idxMax);
        const int idxMin = std::max(-idxMax, -g.hP);
        g.wStart = idxMin;
        g.wDim = idxMax - idxMin + 1;
    }
    int rc = RFB200_OK;
    auto init = [&]() -> int {
        RFP_CUDA(p, cudaSetDevice(device));
        RFP_CUDA(p, cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
        const size_t nHalf = (size_t)N * g.Xh, nPix = (size_t)N * N;
        p->chunk = (int)std::max<size_t>(1, std::min<size_t>(256, ((size_t)256 << 20) / (nHalf * sizeof(float2))));
        RFP_CUDA(p, cudaMalloc(&p->dPF, sizeof(float2) * nHalf * p->chunk));
        RFP_CUDA(p, cudaMalloc(&p->dImg, sizeof(float) * nPix * p->chunk));
        RFP_CUDA(p, cudaMalloc(&p->dE, sizeof(double) * 6 * p->chunk));
        RFP_CUDA(p, cudaMallocHost(&p->hE, sizeof(double) * 6 * p->chunk));
        return rfb200::proj_build(p, volume);
    };
    rc = init();
    if (rc != RFB200_OK) {
        rfb200_projector_destroy(p);
        return rc;
    }
    *out = p;
    return RFB200_OK;
}

int rfb200_projector_project(rfb200_projector p, const double* angles, const float* ctf, int32_t n, float* images) {
    if (!p || n < 0 || (n > 0 && (!angles || !images))) return RFB200_ERR_ARG;
    RFP_CUDA(p, cudaSetDevice(p->device));
    return rfb200::proj_run(p, angles, ctf, false, n, images, false);
}

int rfb200_projector_project_device(rfb200_projector p, const double* angles, const float* d_ctf, int32_t n, float* d_images) {
    if (!p || n < 0 || (n > 0 && (!angles || !d_images))) return RFB200_ERR_ARG;
    RFP_CUDA(p, cudaSetDevice(p->device));
    return rfb200::proj_run(p, angles, d_ctf, true, n, d_images, true);
}

const char* rfb200_projector_last_error(rfb200_projector p) { return p ? p->err.c_str() : ""; }

void rfb200_projector_destroy(rfb200_projector p) {
    if (!p) return;
    cudaSetDevice(p->device);
    if (p->stream) cudaStreamSynchronize(p->stream);
    for (auto& kv : p->plans) cufftDestroy(kv.second);
    void* dev[] = {p->dC, p->dPF, p->dE, p->dCtf, p->dImg};
    for (void* d : dev)
        if (d) cudaFree(d);
    if (p->hE) cudaFreeHost(p->hE);
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
}

}  // extern "C"
