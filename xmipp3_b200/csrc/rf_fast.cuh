// rf_fast.cuh — the `--fast` arithmetic: nearest-pixel insertion + one final 3-D blob convolution.
//
// What it computes is the reference's "Do the blobing at the end of the computation" mode
// (reconstruction_adapt_cuda/reconstruct_fourier_gpu.cpp:71-72 = "G"; device functions of
// reconstruction_cuda/cuda_gpu_reconstruct_fourier.cpp = "D"; CPU twin reconstruct_fourier_accel.cpp):
//   * images are padded to N*pad_vol, transformed, cropped to the resolution sphere and re-centred in y
//     (cropAndShift G:292-321) -> sx x sy pixels, sx = S/2, sy = S, S = maxVolumeIndexYZ (G:229-232);
//   * per (image, symmetry) a traverse space (G:769-814): the voxel lattice is walked along the two axes that do
//     not dominate the plane normal, the third coordinate is the rounded plane hit (D:391-413, 655-735), and the
//     voxel takes the NEAREST pixel with its CTF weights (processVoxel D:455-503) into the (S+1)^3 temporary
//     volume and weights;
//   * at the end: mirrorAndCrop G:704-730, applyBlob G:623-661 on both, forceHermitianSymmetry G:732-749,
//     processWeights G:751-767, convertToExpectedSpace G:664-681, inverse FFT and gridding correction G:894-931.
// How it is computed here: the single-precision plane/voxel/pixel decisions use explicitly rounded operations
// (__fmul_rn / __fadd_rn / __fdiv_rn: no a*b+c contraction) in the reference's operation order, so they are
// reproducible and equal to the CPU oracle's; the CTF tables of the reference (two floats per pixel, computed on the
// host) are evaluated on the device in FP64 and folded into one float4 per pixel; the last five steps are three
// gather kernels (no atomics, no host round trip of the volume).  The insertion itself is a scatter with FP32
// atomics like the reference's (one voxel per lattice column and plane: there is nothing to gather over).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstring>
#include <limits>

#include "rf_kernels.cuh"
#include "rf_types.h"

namespace rfb200 {

struct FastGeo {
    int32_t N, Pv, S, sx, sy, X;      // image, padded size, maxVolumeIndexYZ, fftSizeX, fftSizeY, S/2
};

struct FastSpace {                    // RecFourierProjectionTraverseSpace (reconstruct_fourier_projection_traverse_space.h:37-59)
    int32_t minX, minY, minZ, maxX, maxY, maxZ;
    int32_t dir;                      // 0 = XY, 1 = XZ, 2 = YZ
    int32_t img;                      // image index inside the chunk
    float nx, ny, nz;                 // unitNormal
    float ox, oy, oz;                 // bottomOrigin
    float maxDist2;
    float tinv[9];
    float weight;
    float pad0, pad1, pad2;           // 112 B
};

namespace host {

struct F3 { float x, y, z; };

// D:417-424 / G:475-482, left to right
inline void fast_multiply(const float t[9], F3& p) {
    float a0 = t[0] * p.x, a1 = t[1] * p.y, a2 = t[2] * p.z;
    float b0 = t[3] * p.x, b1 = t[4] * p.y, b2 = t[5] * p.z;
    float c0 = t[6] * p.x, c1 = t[7] * p.y, c2 = t[8] * p.z;
    float s0 = a0 + a1, s1 = b0 + b1, s2 = c0 + c1;
    p.x = s0 + a2;
    p.y = s1 + b2;
    p.z = s2 + c2;
}

inline FastGeo make_fast_geo(int N, double padVol, double maxRes) {
    FastGeo g{};
    g.N = N;
    g.Pv = (int)(N * padVol);                                                   // G:229
    size_t conserveRows = (size_t)std::ceil((double)g.Pv * maxRes * 2.0);       // G:230-232
    conserveRows = (size_t)std::ceil((double)conserveRows / 2.0);
    g.S = 2 * (int)conserveRows;
    g.sx = g.S / 2;                                                             // G:434
    g.sy = g.S;
    g.X = g.S / 2;                                                              // G:698
    return g;
}

// A_SL = R * A^T and its inverse (adjugate over determinant, the 3x3 case of Matrix2D::inv) in double, then the
// traverse space in single precision exactly as computeTraverseSpace G:769-814 evaluates it
inline void make_fast_space(const FastGeo& g, const double R[9], double rot, double tilt, double psi, int img, float weight,
                            FastSpace& sp) {
    double A[9], M[9], I[9];
    euler_matrix(rot, tilt, psi, A);
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            double t = 0;
            for (int c = 0; c < 3; ++c) t += R[a * 3 + c] * A[b * 3 + c];        // (A^T)[c][b] = A[b][c]
            M[a * 3 + b] = t;
        }
    {
        const double* m = M;
        double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
        double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
        double id = 1.0 / det;
        I[0] = c00 * id; I[1] = (m[2] * m[7] - m[1] * m[8]) * id; I[2] = (m[1] * m[5] - m[2] * m[4]) * id;
        I[3] = c01 * id; I[4] = (m[0] * m[8] - m[2] * m[6]) * id; I[5] = (m[2] * m[3] - m[0] * m[5]) * id;
        I[6] = c02 * id; I[7] = (m[1] * m[6] - m[0] * m[7]) * id; I[8] = (m[0] * m[4] - m[1] * m[3]) * id;
    }
    float tr[9];
    for (int a = 0; a < 9; ++a) {
        tr[a] = (float)M[a];
        sp.tinv[a] = (float)I[a];
    }
    F3 c[8];
    {   // createProjectionCuboid G:484-495 with blobSize 0
        const float sizeX = (float)g.sx, sizeY = (float)g.sy, blobSize = 0.f;
        float halfY = sizeY / 2.0f;
        c[3].x = c[2].x = c[7].x = c[6].x = 0.f - blobSize;
        c[0].x = c[1].x = c[4].x = c[5].x = sizeX + blobSize;
        c[3].y = c[0].y = c[7].y = c[4].y = -(halfY + blobSize);
        c[1].y = c[2].y = c[5].y = c[6].y = halfY + blobSize;
        c[3].z = c[0].z = c[1].z = c[2].z = 0.f + blobSize;
        c[7].z = c[4].z = c[5].z = c[6].z = 0.f - blobSize;
    }
    const float org = g.S / 2.f;
    for (int i = 0; i < 8; ++i) fast_multiply(tr, c[i]);
    for (int i = 0; i < 8; ++i) {
        c[i].x += org;
        c[i].y += org;
        c[i].z += org;
    }
    F3 lo = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
    F3 hi = {std::numeric_limits<float>::min(), std::numeric_limits<float>::min(), std::numeric_limits<float>::min()};   // sic, G:509
    for (int i = 0; i < 8; ++i) {
        if (lo.x > c[i].x) lo.x = c[i].x;
        if (lo.y > c[i].y) lo.y = c[i].y;
        if (lo.z > c[i].z) lo.z = c[i].z;
        if (hi.x < c[i].x) hi.x = c[i].x;
        if (hi.y < c[i].y) hi.y = c[i].y;
        if (hi.z < c[i].z) hi.z = c[i].z;
    }
    const float mx = (float)g.S;
    if (lo.x < 0.f) lo.x = 0.f;
    if (lo.y < 0.f) lo.y = 0.f;
    if (lo.z < 0.f) lo.z = 0.f;
    if (hi.x > mx) hi.x = mx;
    if (hi.y > mx) hi.y = mx;
    if (hi.z > mx) hi.z = mx;
    sp.minZ = (int)std::floor(lo.z);
    sp.minY = (int)std::floor(lo.y);
    sp.minX = (int)std::floor(lo.x);
    sp.maxZ = (int)std::ceil(hi.z);
    sp.maxY = (int)std::ceil(hi.y);
    sp.maxX = (int)std::ceil(hi.x);
    sp.ox = c[0].x;
    sp.oy = c[0].y;
    sp.oz = c[0].z;
    const float e = (float)g.sx + 0.f;
    sp.maxDist2 = e * e;
    F3 n = {0.f, 0.f, 1.f};
    fast_multiply(tr, n);
    sp.nx = n.x;
    sp.ny = n.y;
    sp.nz = n.z;
    const float aX = std::fabs(n.x), aY = std::fabs(n.y), aZ = std::fabs(n.z);
    if (aX >= aY && aX >= aZ) sp.dir = 2;
    else if (aY >= aX && aY >= aZ) sp.dir = 1;
    else sp.dir = 0;
    sp.img = img;
    sp.weight = weight;
    sp.pad0 = sp.pad1 = sp.pad2 = 0.f;
}

}  // namespace host

// ------------------------------------------------------------------------------------------------ K1f
// getValuePureNoKAt at continuous frequencies (X, Y) in FP64 (ctf.h:452-502, 1002-1029): the reference evaluates its
// per-pixel CTF tables on the host in double (computeCTFCorrection G:552-593)
__device__ __forceinline__ double d_ctf_value_xy(const CtfConsts& c, double X, double Y) {
    const double u2 = X * X + Y * Y;
    double deltaf = 0.0;
    if (!(fabs(X) < 1e-6 && fabs(Y) < 1e-6)) {
        const double inv = 1.0 / u2;
        const double c2 = (X * X - Y * Y) * inv, s2 = (2.0 * X * Y) * inv;      // cos / sin of 2 atan2(Y, X)
        deltaf = c.defocus_average + c.defocus_deviation * (c2 * c.cos2az + s2 * c.sin2az);
    }
    double arg = c.K1 * deltaf * u2 + c.K2 * u2 * u2;
    if (c.has_vpp) arg += -c.phase_shift * (1.0 - exp(-u2 / (2.0 * c.vpp_radius * c.vpp_radius)));
    double sd, cd;
    sincos(arg, &sd, &cd);
    const double E = c.has_envelope ? d_ctf_envelope(c, u2, deltaf) : 1.0;
    return c.K * c.K * (c.Kcos * cd - c.Ksin * sd) * E;
}

struct FastPrepArgs {
    FastGeo g;
    const float2* fft;            // per image Pv x (Pv/2+1), unnormalised
    float4* pix;                  // per image sy x sx: (re w wCTF, im w wCTF, w, 0), w = wModulator * image weight
    const ImgParams* ip;
    const CtfConsts* ctfs;
    int useCtf, phaseFlipped;
    double iTs, minCtf;
    float maxRes2;                // float member maxResolutionSqr of the reference program
    // Power-of-two padded sizes: the reference's single-precision frequencies x/Pv and (y - Pv/2)/Pv are exact, i.e. integer
    // multiples of 1/Pv, so the CTF can go through the exact-path evaluator (exact fixed-point phase, FP32 sin/cos,
    // FP64 re-evaluation next to the --minCTF threshold) instead of FP64 per pixel
    int ctfInt;
    SliceParams sp;
};

// cropAndShift + computeCTFCorrection, the per-pixel products of processVoxel folded in.  grid (ceil(sx*sy/256), nImg)
__global__ void __launch_bounds__(256) k_fast_prepare(const __grid_constant__ FastPrepArgs a) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const FastGeo& g = a.g;
    const int img = blockIdx.y;
    __shared__ CtfFloat sCtfF;
    if (a.useCtf && a.ctfInt) {
        if (threadIdx.x == 0) d_ctf_prepare(a.ctfs[img], a.sp, sCtfF);
        __syncthreads();
    }
    if (idx >= g.sx * g.sy) return;
    const int y = idx / g.sx, x = idx - y * g.sx;
    const int i = (y >= g.sx) ? y - g.sx : y + g.Pv - g.sx;                    // inverse of myPadI (G:314)
    const int Xh = g.Pv / 2 + 1;
    float2 v = __ldg(a.fft + ((size_t)img * g.Pv + i) * Xh + x);
    const double inv = 1.0 / ((double)g.Pv * (double)g.Pv);                     // FourierTransform() scales by 1/size
    float re = (float)((double)v.x * inv), im = (float)((double)v.y * inv);
    {
        const double f0 = (double)x / (double)g.Pv;                             // FFT_IDX2DIGFREQ, x <= Pv/2
        const double f1 = (i <= g.Pv / 2) ? (double)i / (double)g.Pv : (double)(i - g.Pv) / (double)g.Pv;
        if (f0 * f0 + f1 * f1 > (double)a.maxRes2) re = im = 0.f;               // G:308-310
    }
    float wCTF = 1.f, wMod = 1.f;
    if (a.useCtf && a.ctfInt) {
        d_ctf_weights(a.ctfs[img], sCtfF, a.sp, x, y - g.Pv / 2, wCTF, wMod);
    } else if (a.useCtf) {
        const float freqY = __fdiv_rn(__fsub_rn((float)y, (float)g.Pv / 2.f), (float)g.Pv);   // G:562
        const float freqX = (float)((double)x / (double)g.Pv);                  // G:566
        float CTFVal = (float)d_ctf_value_xy(a.ctfs[img], (double)freqX * a.iTs, (double)freqY * a.iTs);
        float modulatorVal = 1.f;
        if (isnan(CTFVal)) {                                                    // G:569-576
            if (x == 0 && y == 0) modulatorVal = CTFVal = 1.0f;
            else modulatorVal = CTFVal = 0.0f;
        }
        if ((double)fabsf(CTFVal) < a.minCtf) {                                 // G:577-582
            modulatorVal = fabsf(CTFVal);
            CTFVal = (CTFVal >= 0.f) ? 1.f : -1.f;
        } else {
            CTFVal = (float)(1.0 / (double)CTFVal);
        }
        if (a.phaseFlipped) CTFVal = fabsf(CTFVal);
        wCTF = CTFVal;
        wMod = modulatorVal;
    }
    float weight = __fmul_rn(1.f, wMod);                                        // D:494
    weight = __fmul_rn(weight, a.ip[img].weight);
    float4 o;
    o.x = __fmul_rn(__fmul_rn(re, weight), wCTF);                               // D:497-498
    o.y = __fmul_rn(__fmul_rn(im, weight), wCTF);
    o.z = weight;
    o.w = 0.f;
    a.pix[(size_t)img * g.sx * g.sy + idx] = o;
}

// ------------------------------------------------------------------------------------------------ K2f
struct FastInsertArgs {
    FastGeo g;
    const FastSpace* spaces;
    const float4* pix;
    float4* A;                    // (S+1)^3 [z][y][x] interleaved scratch accumulators (V.re, V.im, W, unused)
};

__device__ __forceinline__ float d_fast_hit(float na, float nb, float nc, float a, float b, float oa, float ob, float oc) {
    // (-na*(a-oa) - nb*(b-ob)) / nc + oc, every operation rounded on its own (D:391-413)
    const float t0 = __fmul_rn(-na, __fsub_rn(a, oa));
    const float t1 = __fmul_rn(nb, __fsub_rn(b, ob));
    return __fadd_rn(__fdiv_rn(__fsub_rn(t0, t1), nc), oc);
}

// processProjection<useFast> + processVoxel.  grid (ceil((S+1)/32), ceil((S+1)/8), nSpaces), block (32, 8)
__global__ void __launch_bounds__(256) k_fast_insert(const __grid_constant__ FastInsertArgs a) {
    const FastGeo& g = a.g;
    const FastSpace& sp = a.spaces[blockIdx.z];
    const int idx = blockIdx.x * 32 + threadIdx.x, idy = blockIdx.y * 8 + threadIdx.y;
    int x, y, z;
    if (sp.dir == 0) {                                                          // iterate the XY plane
        if (!(idy >= sp.minY && idy <= sp.maxY && idx >= sp.minX && idx <= sp.maxX)) return;
        const float hit = d_fast_hit(sp.nx, sp.ny, sp.nz, (float)idx, (float)idy, sp.ox, sp.oy, sp.oz);
        x = idx; y = idy; z = (int)__fadd_rn(hit, 0.5f);
    } else if (sp.dir == 1) {                                                   // XZ
        if (!(idy >= sp.minZ && idy <= sp.maxZ && idx >= sp.minX && idx <= sp.maxX)) return;
        const float hit = d_fast_hit(sp.nx, sp.nz, sp.ny, (float)idx, (float)idy, sp.ox, sp.oz, sp.oy);
        x = idx; z = idy; y = (int)__fadd_rn(hit, 0.5f);
    } else {                                                                    // YZ
        if (!(idy >= sp.minZ && idy <= sp.maxZ && idx >= sp.minY && idx <= sp.maxY)) return;
        const float hit = d_fast_hit(sp.ny, sp.nz, sp.nx, (float)idx, (float)idy, sp.oy, sp.oz, sp.ox);
        y = idx; z = idy; x = (int)__fadd_rn(hit, 0.5f);
    }
    const int half = g.S / 2;
    const float px = (float)(x - half), py = (float)(y - half), pz = (float)(z - half);
    const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)), __fmul_rn(pz, pz));
    if (d2 > sp.maxDist2) return;                                               // D:477-479 (also keeps the voxel inside the volume)
    const float ix = __fadd_rn(__fadd_rn(__fmul_rn(sp.tinv[0], px), __fmul_rn(sp.tinv[1], py)), __fmul_rn(sp.tinv[2], pz));
    const float iy = __fadd_rn(__fadd_rn(__fmul_rn(sp.tinv[3], px), __fmul_rn(sp.tinv[4], py)), __fmul_rn(sp.tinv[5], pz));
    if (ix < 0.f) return;                                                       // D:482
    int imgX = (int)__fadd_rn(ix, 0.5f);
    int imgY = (int)__fadd_rn(__fadd_rn(iy, 0.5f), (float)half);
    imgX = min(max(imgX, 0), g.sx - 1);
    imgY = min(max(imgY, 0), g.sy - 1);
    const float4 p = __ldg(a.pix + (size_t)sp.img * g.sx * g.sy + (size_t)imgY * g.sx + imgX);
    const size_t i3 = ((size_t)z * (g.S + 1) + y) * (g.S + 1) + x;
    // one fire-and-forget 16-byte vector reduction per hit (the reference issues three atomicAdd, D:499-502): V and W of
    // a voxel share a sector, which halves the DRAM traffic of this scatter (every reduction reads and writes the
    // 32-byte sector it touches)
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a.A + i3), "f"(p.x), "f"(p.y), "f"(p.z), "f"(0.f) : "memory");
}

// scratch -> accumulators: V += A.xy, W += A.z, A = 0 (before anything reads V or W)
__global__ void __launch_bounds__(256) k_fast_flush(float4* __restrict__ A, float2* __restrict__ V, float* __restrict__ W, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 a = A[i];
        if (a.x != 0.f || a.y != 0.f || a.z != 0.f) {
            float2 v = V[i];
            v.x += a.x;
            v.y += a.y;
            V[i] = v;
            W[i] += a.z;
            A[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}

// ------------------------------------------------------------------------------------------------ K3f
// mirrorAndCrop G:704-730: half space [S+1][S+1][X+1]; out(z,y,x') = in(z,y,x'+X) + f(in(S-z,S-y,X-x')) for x' >= 1
__global__ void __launch_bounds__(256) k_fast_mirror_crop(FastGeo g, const float2* __restrict__ V, const float* __restrict__ W,
                                                          float2* __restrict__ Vh, float* __restrict__ Wh) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int n1 = g.S + 1, nx = g.X + 1;
    if (idx >= (size_t)n1 * n1 * nx) return;
    const int x = (int)(idx % nx);
    const size_t t = idx / nx;
    const int y = (int)(t % n1), z = (int)(t / n1);
    const size_t d = ((size_t)z * n1 + y) * n1 + (x + g.X);
    float2 v = V[d];
    float w = W[d];
    if (x >= 1) {
        const size_t m = ((size_t)(g.S - z) * n1 + (g.S - y)) * n1 + (g.X - x);
        const float2 mv = V[m];
        // whichever of the two the reference adds first lands on 0, and a + b is commutative
        v.x = __fadd_rn(v.x, mv.x);
        v.y = __fadd_rn(v.y, -mv.y);
        w = __fadd_rn(w, W[m]);
    }
    Vh[idx] = v;
    Wh[idx] = w;
}

// applyBlob G:623-661 on V and W at once (same neighbours, same weights), ascending z, y, x like the reference
__global__ void __launch_bounds__(256) k_fast_blob(FastGeo g, const float* __restrict__ table, float blobSize, float iDeltaSqrt,
                                                   const float2* __restrict__ Vh, const float* __restrict__ Wh,
                                                   float2* __restrict__ Vc, float* __restrict__ Wc) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int n1 = g.S + 1, nx = g.X + 1;
    if (idx >= (size_t)n1 * n1 * nx) return;
    const int k = (int)(idx % nx);
    const size_t t = idx / nx;
    const int j = (int)(t % n1), i = (int)(t / n1);
    const float blobSizeSqr = __fmul_rn(blobSize, blobSize);
    const int blob = (int)floorf(blobSize);
    float ar = 0.f, ai = 0.f, aw = 0.f;
    for (int z = max(0, i - blob); z <= min(g.S, i + blob); ++z) {
        const float dZ = (float)((i - z) * (i - z));
        for (int y = max(0, j - blob); y <= min(g.S, j + blob); ++y) {
            const float dY = (float)((j - y) * (j - y));
            for (int x = max(0, k - blob); x <= min(g.X, k + blob); ++x) {
                const float dX = (float)((k - x) * (k - x));
                const float d2 = __fadd_rn(__fadd_rn(dZ, dY), dX);
                if (d2 > blobSizeSqr) continue;
                const int aux = (int)__fadd_rn(__fmul_rn(d2, iDeltaSqrt), 0.5f);
                const float w = __ldg(table + aux);
                const size_t s = ((size_t)z * n1 + y) * nx + x;
                const float2 v = __ldg(Vh + s);
                ar = __fadd_rn(ar, __fmul_rn(w, v.x));
                ai = __fadd_rn(ai, __fmul_rn(w, v.y));
                aw = __fadd_rn(aw, __fmul_rn(w, __ldg(Wh + s)));
            }
        }
    }
    Vc[idx] = make_float2(ar, ai);
    Wc[idx] = aw;
}

// forceHermitianSymmetry G:732-749 + processWeights G:751-767 of one half-space voxel
__device__ __forceinline__ float2 d_fast_value(const FastGeo& g, const float2* __restrict__ Vc, const float* __restrict__ Wc, float corr,
                                               int z, int y, int x) {
    const int n1 = g.S + 1, nx = g.X + 1;
    const size_t self = ((size_t)z * n1 + y) * nx + x;
    float2 v = __ldg(Vc + self);
    float w = __ldg(Wc + self);
    if (x == 0) {
        // the sequential in-place loop visits (z, y <= S/2) as `a` with mirror `b` = (S-z, S-y); on the row y = S/2 both ends
        // are visited and the first visit (smaller z) decides, the second reproduces it
        int az = z, ay = y;
        bool selfIsA = true;
        if (y > g.X) { az = g.S - z; ay = g.S - y; selfIsA = false; }
        else if (y == g.X && z > g.S - z) { az = g.S - z; selfIsA = false; }
        const size_t ia = ((size_t)az * n1 + ay) * nx, ib = ((size_t)(g.S - az) * n1 + (g.S - ay)) * nx;
        const float2 va = __ldg(Vc + ia), vb = __ldg(Vc + ib);
        const float t1x = __fmul_rn(0.5f, __fadd_rn(vb.x, va.x));
        const float t1y = __fmul_rn(0.5f, __fadd_rn(vb.y, -va.y));
        v = selfIsA ? make_float2(t1x, -t1y) : make_float2(t1x, t1y);
        w = __fmul_rn(0.5f, __fadd_rn(__ldg(Wc + ib), __ldg(Wc + ia)));
    }
    if (w > 0.001f) {                                                           // ACCURACY
        const float s = __fdiv_rn(corr, w);
        return make_float2(__fmul_rn(v.x, s), __fmul_rn(v.y, s));
    }
    return make_float2(0.f, 0.f);
}

// convertToExpectedSpace G:664-681 as a gather: sum of the (up to four) half-space voxels that land on target (zt, yt, xt)
__device__ __forceinline__ float2 d_fast_target(const FastGeo& g, const float2* __restrict__ Vc, const float* __restrict__ Wc, float corr,
                                                int zt, int yt, int xt) {
    const int Z = g.Pv, half = g.S / 2;
    float2 acc = make_float2(0.f, 0.f);
    // sources of target index c: s < half with Z - half + s == c, and s >= half with s - half == c (ascending s)
    int ys[2], zs[2], ny = 0, nz = 0;
    { int s = yt - (Z - half); if (s >= 0 && s < half) ys[ny++] = s; s = yt + half; if (s <= g.S) ys[ny++] = s; }
    { int s = zt - (Z - half); if (s >= 0 && s < half) zs[nz++] = s; s = zt + half; if (s <= g.S) zs[nz++] = s; }
    for (int a = 0; a < nz; ++a)
        for (int b = 0; b < ny; ++b) {
            const float2 v = d_fast_value(g, Vc, Wc, corr, zs[a], ys[b], xt);
            acc.x += v.x;
            acc.y += v.y;
        }
    return acc;
}

// out is the Pv x Pv x (Pv/2+1) transform handed to the inverse FFT.  The x = 0 plane is Hermitian by construction
// (forceHermitianSymmetry); the x = Pv/2 plane (reached when S == Pv) is not, and FFTW's c2r only sees its Hermitian part,
// so that part is what is stored (cuFFT's result for a non-Hermitian plane is unspecified).
__global__ void __launch_bounds__(256) k_fast_to_fourier(FastGeo g, const float2* __restrict__ Vc, const float* __restrict__ Wc, float corr,
                                                         float2* __restrict__ out) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int Z = g.Pv, Xf = g.Pv / 2 + 1, half = g.S / 2;
    if (idx >= (size_t)Z * Z * Xf) return;
    const int xt = (int)(idx % Xf);
    const size_t t = idx / Xf;
    const int yt = (int)(t % Z), zt = (int)(t / Z);
    float2 acc = make_float2(0.f, 0.f);
    if (xt <= half) {
        acc = d_fast_target(g, Vc, Wc, corr, zt, yt, xt);
        if (2 * xt == Z) {
            const float2 m = d_fast_target(g, Vc, Wc, corr, (Z - zt) % Z, (Z - yt) % Z, xt);
            acc.x = 0.5f * (acc.x + m.x);
            acc.y = 0.5f * (acc.y - m.y);
        }
    }
    out[idx] = acc;
}

}  // namespace rfb200
