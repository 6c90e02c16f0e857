// rf_kernels.cuh — hand-written sm_100a kernels of the direct Fourier reconstruction.
//
//   K1a k_pad_images     raw N x N particle -> zero-padded, centred, shifted P x P image  (RF.cpp:388-402)
//   (cuFFT R2C batched)                                                                  (RF.cpp:405-407)
//   K1b k_make_slices    half-plane FFT -> resolution-cropped, CTF-weighted FULL-plane slice
//                        (originals + Hermitian mirrors), RF.cpp:594-625 hoisted out of the insertion
//   K2  k_gather         voxel-centric gather with Kaiser-Bessel blob interpolation; replaces the
//                        scatter loop RF.cpp:586-792.  One CTA owns an 8^3 tile of the half volume,
//                        each thread one voxel; accumulators live in registers; no atomics.
//   K2e k_edge           lattice points the tile gather does not own (orig-only planes, Nyquist wraps)
//   K3a k_normalize      weight symmetrisation + normalisation (RF.cpp:1056-1101, 453-479, 1188-1221)
//   (cuFFT C2R 3-D)                                                                      (RF.cpp:1145)
//   K3b k_crop_correct   CenterFFT + crop + gridding correction (RF.cpp:1146-1178)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rf_types.h"

namespace rfb200 {

__constant__ PlaneF c_planes[kMaxPlanes];   // per-(image,symmetry) rotation data, 48 KB

// ------------------------------------------------------------------ small helpers
__device__ __forceinline__ int d_wrap(int x, int n) {
    int r = x % n;
    return r < 0 ? r + n : r;
}
__device__ __forceinline__ int d_tile_slot(int vx, int vy, int vz) {
    int brick = (vx >> 2) | ((vy >> 2) << 2) | ((vz >> 1) << 4);
    int lane = (vx & 3) | ((vy & 3) << 2) | ((vz & 1) << 4);
    return brick * 32 + lane;
}
__device__ __forceinline__ int64_t d_blocked_index(const Geometry& geo, int ux, int uy, int uz) {
    int x = ux, y = uy - geo.lo, z = uz - geo.lo;
    int64_t tile = ((int64_t)(z / kTileZ) * geo.ty + (y / kTileY)) * geo.tx + (x / kTileX);
    return tile * kTileVox + d_tile_slot(x % kTileX, y % kTileY, z % kTileZ);
}
// natural lattice point owned by the tile gather? (host twin: host::main_owns)
__device__ __forceinline__ bool d_main_owns(const Geometry& geo, int ux, int uy) {
    const int Z = geo.Z;
    if (ux == 0) return uy <= geo.yHalf;
    if (ux > Z / 2) return false;
    return d_wrap(-ux, Z) > Z / 2;
}

// ================================================================== K1a
// Cubic B-spline prefilter (Unser's recursive filter, pole sqrt(3)-2, mirror boundaries) for the images
// whose shift is fractional: xmippCore's readApplyGeo interpolates with BSPLINE3.
// One thread filters one line; `stride` selects rows (1) or columns (N).  Recursion state in double.
__device__ __forceinline__ void d_bspline_line(float* c, int n, long stride) {
    if (n == 1) return;
    const double z1 = -0.26794919243112270647, lambda = 6.0;   // sqrt(3)-2, (1-z1)(1-1/z1)
    const int horizon = 28;                                     // ceil(log(DBL_EPSILON)/log|z1|)
    double sum;
    if (horizon < n) {
        double zn = z1;
        sum = lambda * c[0];
        for (int k = 1; k < horizon; ++k) { sum += zn * lambda * c[k * stride]; zn *= z1; }
    } else {
        double zn = z1, iz = 1.0 / z1, z2n = pow(z1, (double)(n - 1));
        sum = lambda * c[0] + z2n * lambda * c[(n - 1) * stride];
        z2n *= z2n * iz;
        for (int k = 1; k <= n - 2; ++k) { sum += (zn + z2n) * lambda * c[k * stride]; zn *= z1; z2n *= iz; }
        sum /= (1.0 - zn * zn);
    }
    // causal pass; the running value is kept in double, the line holds float
    double prev = sum;
    c[0] = (float)prev;
    double last2 = prev;
    for (int k = 1; k < n; ++k) {
        double v = lambda * c[k * stride] + z1 * prev;
        last2 = prev;
        prev = v;
        c[k * stride] = (float)v;
    }
    // anticausal pass
    double nxt = (z1 / (z1 * z1 - 1.0)) * (z1 * last2 + prev);
    c[(n - 1) * stride] = (float)nxt;
    for (int k = n - 2; k >= 0; --k) {
        nxt = z1 * (nxt - (double)c[k * stride]);
        c[k * stride] = (float)nxt;
    }
}
// grid (ceil(N/128), nImg): rows pass (axis 0) reads raw and writes coef, columns pass (axis 1) works in place
__global__ void __launch_bounds__(128) k_bspline_prefilter(const float* __restrict__ raw, float* __restrict__ coef,
                                                           const ImgParams* __restrict__ ip, int N, int axis) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int img = blockIdx.y;
    if (t >= N || !ip[img].spline) return;
    float* c = coef + (size_t)img * N * N;
    if (axis == 0) {
        const float* r = raw + (size_t)img * N * N + (size_t)t * N;
        float* line = c + (size_t)t * N;
        for (int k = 0; k < N; ++k) line[k] = r[k];
        d_bspline_line(line, N, 1);
    } else {
        d_bspline_line(c + t, N, N);
    }
}

__device__ __forceinline__ void d_bspline3_weights(float t, float w[4]) {
    float t2 = t * t, t3 = t2 * t;
    w[0] = (1.0f - 3.0f * t + 3.0f * t2 - t3) * (1.0f / 6.0f);
    w[1] = (4.0f - 6.0f * t2 + 3.0f * t3) * (1.0f / 6.0f);
    w[2] = (1.0f + 3.0f * t + 3.0f * t2 - 3.0f * t3) * (1.0f / 6.0f);
    w[3] = t3 * (1.0f / 6.0f);
}

// grid (ceil(N*N/256), nImg).  The P x P buffer is zeroed once at create; the set of written
// positions is the same for every image, so zeros never need rewriting.
__global__ void __launch_bounds__(256) k_pad_images(const float* __restrict__ raw, const float* __restrict__ coef,
                                                    float* __restrict__ pad, const ImgParams* __restrict__ ip, int N, int P) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * N) return;
    int img = blockIdx.y;
    int i = idx / N, j = idx - i * N;
    ImgParams q = ip[img];
    // content moves by +shift with wrap (readApplyGeo, only_apply_shifts; RF.cpp:313-314,362):
    // out(x) = in(x - shift); x - shift = (j + mx) + ux
    float v;
    if (!q.spline) {
        int si = d_wrap(i + q.my, N), sj = d_wrap(j + q.mx, N);
        v = __ldg(raw + (size_t)img * N * N + (size_t)si * N + sj);
    } else {
        float wx[4], wy[4];
        d_bspline3_weights(q.ux, wx);
        d_bspline3_weights(q.uy, wy);
        const float* c = coef + (size_t)img * N * N;
        v = 0.f;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const float* row = c + (size_t)d_wrap(i + q.my - 1 + a, N) * N;
            float r = 0.f;
#pragma unroll
            for (int b = 0; b < 4; ++b) r = fmaf(wx[b], __ldg(row + d_wrap(j + q.mx - 1 + b, N)), r);
            v = fmaf(wy[a], r, v);
        }
    }
    // logical index l = i - N/2 lands on physical (l mod P): pad + CenterFFT fused (RF.cpp:390-402)
    int pi = d_wrap(i - N / 2, P), pj = d_wrap(j - N / 2, P);
    pad[(size_t)img * P * P + (size_t)pi * P + pj] = v;
}

// ================================================================== K1b
struct SliceParams {
    int P, Xh;                 // padded size, P/2+1
    int iLo, iHi;              // valid signed row range of the half-plane FFT
    int R, Rp, side;
    int useCtf, phaseFlipped;
    double a2;                 // (1/(P*sampling))^2: integer |freq|^2 -> continuous u^2
    float a;                   // 1/(P*sampling)
    float minCtfF;
    double minCtf;
    float invP2;               // forward FFT normalisation 1/P^2 (RF.cpp:405-407)
};

__device__ __forceinline__ double d_bessj0(double x) {
    double ax = fabs(x);
    if (ax < 8.0) {
        double y = x * x;
        double a1 = 57568490574.0 + y * (-13362590354.0 + y * (651619640.7 + y * (-11214424.18 + y * (77392.33017 + y * (-184.9052456)))));
        double a2 = 57568490411.0 + y * (1029532985.0 + y * (9494680.718 + y * (59272.64853 + y * (267.8532712 + y * 1.0))));
        return a1 / a2;
    }
    double z = 8.0 / ax, y = z * z, xx = ax - 0.785398164;
    double a1 = 1.0 + y * (-0.1098628627e-2 + y * (0.2734510407e-4 + y * (-0.2073370639e-5 + y * 0.2093887211e-6)));
    double a2 = -0.1562499995e-1 + y * (0.1430488765e-3 + y * (-0.6911147651e-5 + y * (0.7621095161e-6 - y * 0.934935152e-7)));
    return sqrt(0.636619772 / ax) * (cos(xx) * a1 - z * sin(xx) * a2);
}

// damping envelope of getValuePureAt (ctf.h:474-484), double precision; only evaluated when a CTF row
// carries envelope parameters
__device__ __noinline__ double d_ctf_envelope(const CtfConsts& c, double u2, double deltaf) {
    double u = sqrt(u2), u4 = u2 * u2;
    double Eespr = exp(-c.K3 * u4);
    double EdeltaF = d_bessj0(c.K5 * u2);
    double xr = u * c.DeltaR;
    double EdeltaR = (fabs(xr) < 0.0001) ? 1.0 : sin(3.14159265358979323846 * xr) / (3.14159265358979323846 * xr);
    double aux = c.K7 * u2 * u + deltaf * u;
    double Ealpha = exp(-c.K6 * aux * aux);
    double Ed = Eespr * EdeltaF * EdeltaR * Ealpha + c.envR0 + c.envR1 * u + c.envR2 * u2;
    return Ed < 0 ? 0.0 : Ed;
}

// RF.cpp:609-624 applied to a CTF value: NaN handling, "invert or damp" by minCTF, phase flipping
template <typename T>
__device__ __forceinline__ void d_ctf_rules(T v, T minCtf, int phaseFlipped, int j, int ip, float& wCTF, float& wMod) {
    T wm = 1, wc = v;
    if (isnan(v)) {                                       // :609-615
        if (ip == 0 && j == 0) wm = wc = 1;
        else wm = wc = 0;
    }
    if (fabs(wc) < minCtf) {                              // :616-622
        wm = fabs(wc);
        wc = (wc >= 0) ? T(1) : T(-1);
    } else
        wc = T(1) / wc;
    if (phaseFlipped) wc = fabs(wc);                      // :623-624
    wCTF = (float)wc;
    wMod = (float)wm;
}

// Full double-precision evaluation of getValuePureNoKAt (ctf.h:452-502, 1002-1029).  Only used for the rare
// pixels whose |CTF| is within FP32 noise of the minCTF threshold, so that the invert-or-damp decision is the
// one a double-precision program takes.
__device__ __noinline__ void d_ctf_weights_exact(const CtfConsts& c, const SliceParams& sp, int j, int ip, float& wCTF, float& wMod) {
    const int r2i = j * j + ip * ip;
    const double u2 = (double)r2i * sp.a2;
    double deltaf = 0.0;
    if (!(fabs((double)j) * sp.a < 1e-6 && fabs((double)ip) * sp.a < 1e-6)) {
        const double inv = 1.0 / (double)r2i;
        const double c2 = (double)(j * j - ip * ip) * inv, s2 = (double)(2 * j * ip) * inv;
        deltaf = c.defocus_average + c.defocus_deviation * (c2 * c.cos2az + s2 * c.sin2az);
    }
    double arg = u2 * fma(c.K1, deltaf, c.K2 * u2);
    if (c.has_vpp) arg += -c.phase_shift * (1.0 - exp(-u2 / (2.0 * c.vpp_radius * c.vpp_radius)));
    double sd, cd;
    sincos(arg, &sd, &cd);
    const double E = c.has_envelope ? d_ctf_envelope(c, u2, deltaf) : 1.0;
    const double v = c.K * c.K * (c.Kcos * cd - c.Ksin * sd) * E;     // K * ( -K (Ksin sin - Kcos cos) E )
    d_ctf_rules<double>(v, sp.minCtf, sp.phaseFlipped, j, ip, wCTF, wMod);
}

// wCTF / wModulator of half-plane pixel (j, ip) — RF.cpp:600-625.  Only the phase argument needs double
// precision (it reaches hundreds of radians): it is formed from exact integer frequencies and range-reduced
// in double; the astigmatism angle term, sin/cos and the minCTF rules run in FP32.
// cos(2(atan2(Y,X) - az)) is expanded so that no atan2 is needed.
__device__ __forceinline__ void d_ctf_weights(const CtfConsts& c, const SliceParams& sp, int j, int ip, float& wCTF, float& wMod) {
    const int r2i = j * j + ip * ip;                      // exact: |freq|^2 in units of (1/(P*Ts))^2
    const double u2 = (double)r2i * sp.a2;
    double deltaf = 0.0;
    const float ax = fabsf((float)j) * sp.a, ay = fabsf((float)ip) * sp.a;
    if (!(ax < 1e-6f && ay < 1e-6f)) {                    // precomputeValues(X,Y): deltaf = 0 at the origin
        const float inv = 1.0f / (float)r2i;
        const float c2 = (float)(j * j - ip * ip) * inv, s2 = (float)(2 * j * ip) * inv;
        deltaf = c.defocus_average + c.defocus_deviation * (double)(c2 * (float)c.cos2az + s2 * (float)c.sin2az);
    }
    double arg = u2 * fma(c.K1, deltaf, c.K2 * u2);
    if (c.has_vpp) arg += -c.phase_shift * (1.0 - exp(-u2 / (2.0 * c.vpp_radius * c.vpp_radius)));
    const double inv2pi = 0.15915494309189535, twopi_hi = 6.283185307179586, twopi_lo = 2.4492935982947064e-16;
    const double kk = rint(arg * inv2pi);
    double red = fma(-kk, twopi_hi, arg);
    red = fma(-kk, twopi_lo, red);
    float sn, cs;
    sincosf((float)red, &sn, &cs);
    float E = 1.0f;
    if (c.has_envelope) E = (float)d_ctf_envelope(c, u2, deltaf);
    const float v = (float)(c.K * c.K) * ((float)c.Kcos * cs - (float)c.Ksin * sn) * E;
    if (fabsf(fabsf(v) - sp.minCtfF) < 4e-6f) {
        d_ctf_weights_exact(c, sp, j, ip, wCTF, wMod);
        return;
    }
    d_ctf_rules<float>(v, sp.minCtfF, sp.phaseFlipped, j, ip, wCTF, wMod);
}

// contribution of original half-plane pixel (j >= 0, ip): (re, im, m) = (wCTF*w*F, w), w = weight*wModulator
__device__ __forceinline__ float4 d_pixel_contrib(const float2* __restrict__ fft, const int* __restrict__ jmax,
                                                  const SliceParams& sp, const CtfConsts* ctf, float weight, int j, int ip) {
    float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ip < sp.iLo || ip > sp.iHi) return out;
    if (j > jmax[ip - sp.iLo]) return out;            // resolution cut-off, RF.cpp:597
    int row = ip < 0 ? ip + sp.P : ip;
    float2 F = __ldg(fft + (size_t)row * sp.Xh + j);
    float wc = 1.f, wm = 1.f;
    if (sp.useCtf) d_ctf_weights(*ctf, sp, j, ip, wc, wm);
    float w = weight * wm;
    float s = w * wc * sp.invP2;
    out.x = F.x * s;
    out.y = F.y * s;
    out.z = w;
    out.w = weight;      // weight without the CTF modulator: what a --iter > 1 re-insertion pass adds (RF.cpp:770-775)
    return out;
}

// grid (ceil((R+1)/32), ceil((2R+1)/32), nImg), block (32, 8): threadIdx.x runs along j (coalesced reads of
// the FFT rows and writes of the slice rows), every thread handles 4 rows.  A thread owns original half-plane
// pixels (j >= 0, ip) inside the bounding square; it writes its own entry and the Hermitian mirror entry of
// the full-plane slice.
constexpr int kSliceRowsPerThread = 4;
__global__ void __launch_bounds__(256, 4) k_make_slices(const float2* __restrict__ fft, float4* __restrict__ slices,
                                                     float4* __restrict__ col0, const ImgParams* __restrict__ ip,
                                                     const CtfConsts* __restrict__ ctfs, const int* __restrict__ jmax,
                                                     const __grid_constant__ SliceParams sp) {
    const int j = blockIdx.x * 32 + threadIdx.x;
    if (j > sp.R) return;
    const int img = blockIdx.z;
    const float2* f = fft + (size_t)img * sp.P * sp.Xh;
    const CtfConsts* ctf = sp.useCtf ? ctfs + img : nullptr;
    const float weight = ip[img].weight;
    float4* S = slices + (size_t)img * sp.side * sp.side;
    const int rowBase = blockIdx.y * (8 * kSliceRowsPerThread) + threadIdx.y;
#pragma unroll 2
    for (int q = 0; q < kSliceRowsPerThread; ++q) {
        const int r = rowBase + 8 * q;
        if (r > 2 * sp.R) break;
        const int ipx = r - sp.R;
        float4 c = d_pixel_contrib(f, jmax, sp, ctf, weight, j, ipx);
        if (j > 0) {
            S[(size_t)(ipx + sp.Rp) * sp.side + (j + sp.Rp)] = c;
            S[(size_t)(-ipx + sp.Rp) * sp.side + (-j + sp.Rp)] = make_float4(c.x, -c.y, c.z, c.w);
        } else {
            // column j = 0 holds original (0,ip) plus the mirror of original (0,-ip): the reference
            // inserts this column twice for x > 0 voxels
            float4 m = d_pixel_contrib(f, jmax, sp, ctf, weight, 0, -ipx);
            S[(size_t)(ipx + sp.Rp) * sp.side + sp.Rp] = make_float4(c.x + m.x, c.y - m.y, c.z + m.z, c.w + m.w);
            col0[(size_t)img * sp.side + (ipx + sp.Rp)] = c;
        }
    }
}

// ================================================================== K2
struct GatherArgs {
    Geometry geo;
    const int32_t* tileList;
    int nTiles;
    int* tileCounter;
    const float* blobTable;      // kBlobTable floats
    const PlaneD* planesD;       // nPlanes
    const float* planesSoA;      // 9 x kMaxPlanes floats: e1x,e1y,e1z,e2x,e2y,e2z,nx,ny,nz
    int nPlanes;
    const float4* slices;        // per image side*side float4
    size_t sliceStride;          // side*side
    float2* Vb;
    float* Wb;
    float* Wb2;                  // un-modulated weight sum, only for --iter > 1 with CTF (else nullptr)
};

#ifndef RF_GATHER_THREADS
#define RF_GATHER_THREADS 512
#endif
constexpr int kGatherThreads = RF_GATHER_THREADS;   // warps take bricks of the tile from a shared counter
constexpr size_t kGatherSmem = kMaxPlanes * sizeof(Hit) + 64 * sizeof(int);   // dynamic part (the blob table is static)
#ifndef RF_GATHER_CTAS
#define RF_GATHER_CTAS 2
#endif
static_assert(kMaxPlanes <= kGatherThreads, "phase A maps one thread to one plane");

template <int K, bool kTwoW>
__global__ void __launch_bounds__(kGatherThreads, RF_GATHER_CTAS) k_gather(const __grid_constant__ GatherArgs a) {
    const Geometry& c_geo = a.geo;
    __shared__ float tbl[kBlobTable];          // static: its shared address is a compile-time constant
    extern __shared__ __align__(16) unsigned char smem[];
    Hit* hits = reinterpret_cast<Hit*>(smem);
    int* sInt = reinterpret_cast<int*>(smem + kMaxPlanes * sizeof(Hit));
    // sInt[0..31] warp counts, sInt[32] tile, sInt[33] hit count, sInt[34] next brick, sInt[40] table address

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < kBlobTable; i += kGatherThreads) tbl[i] = __ldg(a.blobTable + i);

    const int Z = c_geo.Z, lo = c_geo.lo, hi = c_geo.hi;
    const float r2 = c_geo.r2, rho = c_geo.rho, s2 = c_geo.s2, iDelta = c_geo.iDelta, rr = c_geo.r;
    const float kI = s2 * iDelta, sMax = c_geo.sMax;
    // Shared byte address of tbl[0] minus (bits(2^23) << 2), modulo 2^32 (see the lookup below).  Routed through
    // shared memory so that the compiler treats it as an opaque value: the lookup address is then one LEA.
    if (threadIdx.x == 0) sInt[40] = (int)((uint32_t)__cvta_generic_to_shared(tbl) - (0x4B000000u << 2));
    __syncthreads();
    const uint32_t tblAdj = (uint32_t)((volatile int*)sInt)[40];
    const int Rp = c_geo.Rp, side = c_geo.side;
    // lane -> voxel inside a 4x4x2 brick
    const int lx = lane & 3, ly = (lane >> 2) & 3, lz = lane >> 4;
    // tile-level culling constants: half extents of the 16x16x8 lattice box around its centre
    const float hx = 0.5f * (kTileX - 1), hy = 0.5f * (kTileY - 1), hz = 0.5f * (kTileZ - 1);
    const float inplaneLim = c_geo.inplane_reach + sqrtf(hx * hx + hy * hy + hz * hz) * sqrtf(1.0f / s2) + 1.0f;
    const float reach2 = c_geo.reach * c_geo.reach + 1.0f;

    for (;;) {
        __syncthreads();
        if (tid == 0) { sInt[32] = atomicAdd(a.tileCounter, 1); sInt[33] = 0; sInt[34] = 0; }
        __syncthreads();
        const int t = sInt[32];
        if (t >= a.nTiles) break;
        const int tileId = __ldg(a.tileList + t);
        const int ttx = tileId % c_geo.tx, tty = (tileId / c_geo.tx) % c_geo.ty, ttz = tileId / (c_geo.tx * c_geo.ty);
        const int ox = ttx * kTileX, oy = lo + tty * kTileY, oz = lo + ttz * kTileZ;

        // ---------------- phase A: which planes of the chunk come near this tile? (thread <-> plane)
        {
            const float cx = ox + hx, cy = oy + hy, cz = oz + hz;
            const int k = tid;
            bool hit = false;
            Hit e;
            if (k < a.nPlanes) {
                const float* s = a.planesSoA + k;
                const float nx = __ldg(s + 6 * kMaxPlanes), ny = __ldg(s + 7 * kMaxPlanes), nz = __ldg(s + 8 * kMaxPlanes);
                const float hc = cx * nx + cy * ny + cz * nz;
                const float supp = hx * fabsf(nx) + hy * fabsf(ny) + hz * fabsf(nz);
                if (fabsf(hc) <= rr + supp + 1e-2f) {
                    float ac = cx * __ldg(s) + cy * __ldg(s + kMaxPlanes) + cz * __ldg(s + 2 * kMaxPlanes);
                    float bc = cx * __ldg(s + 3 * kMaxPlanes) + cy * __ldg(s + 4 * kMaxPlanes) + cz * __ldg(s + 5 * kMaxPlanes);
                    if (fabsf(ac) <= inplaneLim && fabsf(bc) <= inplaneLim) {
                        const PlaneD pd = a.planesD[k];
                        double a0 = ox * pd.e1[0] + oy * pd.e1[1] + oz * pd.e1[2];
                        double b0 = ox * pd.e2[0] + oy * pd.e2[1] + oz * pd.e2[2];
                        double h0 = ox * pd.n[0] + oy * pd.n[1] + oz * pd.n[2];
                        double ja = rint(a0), jb = rint(b0);
                        e.k = k;
                        e.ja0 = (int)ja;
                        e.jb0 = (int)jb;
                        e.fa = (float)(a0 - ja);
                        e.fb = (float)(b0 - jb);
                        e.h0 = (float)h0;
                        // brick mask: bricks whose 4x4x2 box comes within the blob radius of the plane
                        const float bs = 1.5f * fabsf(nx) + 1.5f * fabsf(ny) + 0.5f * fabsf(nz) + rr + 1e-2f;
                        const float hb0 = e.h0 + 1.5f * nx + 1.5f * ny + 0.5f * nz;   // centre of brick (0,0,0)
                        uint32_t mlo = 0, mhi = 0;
#pragma unroll
                        for (int bk = 0; bk < 4; ++bk)
#pragma unroll
                            for (int bj = 0; bj < 4; ++bj)
#pragma unroll
                                for (int bi = 0; bi < 4; ++bi) {
                                    const float hb = hb0 + 4.0f * bi * nx + 4.0f * bj * ny + 2.0f * bk * nz;
                                    const int b = bi | (bj << 2) | (bk << 4);
                                    if (fabsf(hb) <= bs) {
                                        if (b < 32) mlo |= 1u << b;
                                        else mhi |= 1u << (b - 32);
                                    }
                                }
                        e.maskLo = mlo;
                        e.maskHi = mhi;
                        hit = (mlo | mhi) != 0;
                    }
                }
            }
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (lane == 0) sInt[warp] = __popc(m);
            __syncthreads();
            int off = 0, total = 0;
#pragma unroll
            for (int w = 0; w < kGatherThreads / 32; ++w) {
                int c = sInt[w];
                if (w < warp) off += c;
                total += c;
            }
            if (hit) hits[off + __popc(m & ((1u << lane) - 1u))] = e;
            if (tid == 0) sInt[33] = total;
        }
        __syncthreads();
        const int nHits = sInt[33];
        if (nHits == 0) continue;

        // ---------------- phase B: warps take bricks from a shared counter; every lane gathers for its own voxel
        for (;;) {
            int brick = 0;
            if (lane == 0) brick = atomicAdd(&sInt[34], 1);
            brick = __shfl_sync(0xffffffffu, brick, 0);
            if (brick >= kBricks) break;
            const int vx = ((brick & 3) << 2) | lx, vy = (((brick >> 2) & 3) << 2) | ly, vz = ((brick >> 4) << 1) | lz;
            const int ux = ox + vx, uy = oy + vy, uz = oz + vz;
            bool owned = (ux <= Z / 2) && (uy <= hi) && (uz <= hi) && d_main_owns(c_geo, ux, uy);
            owned = owned && ((float)ux * ux + (float)uy * uy + (float)uz * uz <= reach2);
            if (!__any_sync(0xffffffffu, owned)) continue;
            const float vxf = (float)vx, vyf = (float)vy, vzf = (float)vz;
            const uint32_t bitLo = brick < 32 ? (1u << brick) : 0u, bitHi = brick < 32 ? 0u : (1u << (brick - 32));
            float accRe = 0.f, accIm = 0.f, accW = 0.f, accW2 = 0.f;
            for (int eI = 0; eI < nHits; ++eI) {
                const uint2 mk = *reinterpret_cast<const uint2*>(&hits[eI].maskLo);
                if (((mk.x & bitLo) | (mk.y & bitHi)) == 0) continue;
                const Hit H = hits[eI];
                const PlaneF& pl = c_planes[H.k];
                const float h = fmaf(vzf, pl.n[2], fmaf(vyf, pl.n[1], fmaf(vxf, pl.n[0], H.h0)));
                const float h2 = h * h;
                const bool in = owned && (h2 <= r2);
                if (!__any_sync(0xffffffffu, in)) continue;
                if (in) {
                    const float ar = fmaf(vzf, pl.e1[2], fmaf(vyf, pl.e1[1], fmaf(vxf, pl.e1[0], H.fa)));
                    const float br = fmaf(vzf, pl.e2[2], fmaf(vyf, pl.e2[1], fmaf(vxf, pl.e2[0], H.fb)));
                    const int jw = __float2int_ru(ar - rho);
                    const int iw = __float2int_ru(br - rho);
                    const int jAbs = H.ja0 + jw + Rp, iAbs = H.jb0 + iw + Rp;   // slice coordinates of the window origin
                    if ((unsigned)jAbs <= (unsigned)(side - K) && (unsigned)iAbs <= (unsigned)(side - K)) {
                        // squared distances pre-scaled by iDelta: S = d^2 * iDelta is directly the table coordinate
                        float dxs[K], dys[K];
                        const float da0 = ar - __int2float_rn(jw), db0 = br - __int2float_rn(iw);
                        const float h2s = h2 * iDelta;
#pragma unroll
                        for (int q = 0; q < K; ++q) {
                            float da = da0 - (float)q, db = db0 - (float)q;
                            dxs[q] = kI * da * da;
                            dys[q] = fmaf(kI * db, db, h2s);
                        }
                        const float4* p = a.slices + ((size_t)pl.img * a.sliceStride + (size_t)(iAbs * side + jAbs));
#pragma unroll
                        for (int ti = 0; ti < K; ++ti) {
#pragma unroll
                            for (int tj = 0; tj < K; ++tj) {
                                const float S = dys[ti] + dxs[tj];
                                if (S <= sMax) {
                                    // (int)(d2*iDelta + 0.5) of RF.cpp:725: adding 2^23 rounds S to the nearest integer in
                                    // the mantissa; (bits << 2) + tblAdj is then the shared-memory byte address of the entry
                                    const uint32_t addr = (__float_as_uint(S + 8388608.0f) << 2) + tblAdj;
                                    float w;
                                    asm("ld.shared.f32 %0, [%1];" : "=f"(w) : "r"(addr));
                                    const float4 px = __ldg(p + tj);
                                    accRe = fmaf(w, px.x, accRe);
                                    accIm = fmaf(w, px.y, accIm);
                                    accW = fmaf(w, px.z, accW);
                                    if (kTwoW) accW2 = fmaf(w, px.w, accW2);
                                }
                            }
                            p += side;
                        }
                    }
                }
            }
            // one coalesced read-modify-write of the brick (blocked layout: 32 consecutive slots)
            if (owned && (accW != 0.f || accRe != 0.f || accIm != 0.f)) {
                const size_t o = (size_t)tileId * kTileVox + (size_t)brick * 32 + lane;
                float2 v = a.Vb[o];
                v.x += accRe;
                v.y += accIm;
                a.Vb[o] = v;
                a.Wb[o] += accW;
                if (kTwoW) a.Wb2[o] += accW2;
            }
        }
    }
}

// ================================================================== K2e
struct EdgeArgs {
    Geometry geo;
    const EdgeItem* items;       // sorted by target voxel
    const int32_t* groupStart;   // nGroups+1 offsets: items of one group share the target
    int nGroups;
    const PlaneD* planesD;
    const int* planeImg;
    int nPlanes;
    const float* blobTable;
    const float4* slices;
    const float4* col0;
    size_t sliceStride;
    float2* Vb;
    float* Wb;
    float* Wb2;                  // may be nullptr
    double iDeltaD;
};

// One thread per edge TARGET voxel (it walks the lattice points aliased onto that voxel), brute force
// over the planes of the chunk, double precision positions.  These points are few: caps of the reach
// sphere poking through the Nyquist faces, the x = Z/2 plane and one row of the x = 0 plane.
__global__ void __launch_bounds__(128) k_edge(const __grid_constant__ EdgeArgs a) {
    const Geometry& c_geo = a.geo;
    int grp = blockIdx.x * blockDim.x + threadIdx.x;
    if (grp >= a.nGroups) return;
    const double r2 = (double)c_geo.r * (double)c_geo.r, rho = c_geo.rho, s2 = c_geo.s2;
    const double lim = c_geo.inplane_reach;
    const int Rp = c_geo.Rp, side = c_geo.side, K = c_geo.K;
    double accRe = 0, accIm = 0, accW = 0, accW2 = 0;
    const int i0 = a.groupStart[grp], i1 = a.groupStart[grp + 1];
    const int64_t store = a.items[i0].store;
    for (int it = i0; it < i1; ++it) {
        const EdgeItem e = a.items[it];
        const double ux = e.ux, uy = e.uy, uz = e.uz;
        for (int k = 0; k < a.nPlanes; ++k) {
            const PlaneD& pl = a.planesD[k];
            double h = ux * pl.n[0] + uy * pl.n[1] + uz * pl.n[2];
            double h2 = h * h;
            if (h2 > r2) continue;
            double al = ux * pl.e1[0] + uy * pl.e1[1] + uz * pl.e1[2];
            double be = ux * pl.e2[0] + uy * pl.e2[1] + uz * pl.e2[2];
            if (fabs(al) > lim || fabs(be) > lim) continue;
            int jw = (int)ceil(al - rho), iw = (int)ceil(be - rho);
            const int img = a.planeImg[k];
            const float4* S = a.slices + (size_t)img * a.sliceStride;
            const float4* C0 = a.col0 + (size_t)img * side;
            for (int ti = 0; ti < K; ++ti) {
                int ip = iw + ti;
                double db = be - ip;
                double rowd2 = h2 + s2 * db * db;
                if (rowd2 > r2) continue;
                for (int tj = 0; tj < K; ++tj) {
                    int j = jw + tj;
                    double da = al - j;
                    double d2 = rowd2 + s2 * da * da;
                    if (d2 > r2) continue;
                    if (e.mode == 1 && j < 0) continue;                 // originals only
                    int idx = (int)(d2 * a.iDeltaD + 0.5);              // RF.cpp:725
                    float w = __ldg(a.blobTable + idx);
                    float4 px = (e.mode == 1 && j == 0) ? __ldg(C0 + (ip + Rp)) : __ldg(S + (size_t)(ip + Rp) * side + (j + Rp));
                    accRe += (double)w * px.x;
                    accIm += (double)w * px.y;
                    accW += (double)w * px.z;
                    accW2 += (double)w * px.w;
                }
            }
        }
    }
    if (accW != 0 || accRe != 0 || accIm != 0 || accW2 != 0) {
        float2 v = a.Vb[store];
        v.x += (float)accRe;
        v.y += (float)accIm;
        a.Vb[store] = v;
        a.Wb[store] += (float)accW;
        if (a.Wb2) a.Wb2[store] += (float)accW2;
    }
}

// ================================================================== K3a
struct NormArgs {
    Geometry geo;
    const float2* Vb;
    const float* Wb;
    const float* Wb2;   // un-modulated weights for --iter > 1 (== Wb when no CTF modulators exist)
    float2* out;        // natural layout [z][y][x], Z*Z*X, input of the C2R transform
    float corr;         // corr2D_3D (RF.cpp:457-458)
    int nIterWeight;
};

// value of V*corr*Winv at natural stored index (z,y,x) following RF.cpp:1073-1078, 463-477
__device__ __forceinline__ float2 d_norm_value(const NormArgs& a, int z, int y, int x) {
    const Geometry& c_geo = a.geo;
    const int Z = c_geo.Z;
    int uy = y <= Z / 2 ? y : y - Z, uz = z <= Z / 2 ? z : z - Z;
    int64_t b = d_blocked_index(c_geo, x, uy, uz);
    float2 v = a.Vb[b];
    float w = a.Wb[b];
    if (x == 0 && uy <= c_geo.yHalf) {   // stored as orig+mirr = twice the pair average of RF.cpp:1188-1221
        v.x *= 0.5f; v.y *= 0.5f; w *= 0.5f;
    }
    if (a.nIterWeight == 0) return make_float2(v.x * a.corr, v.y * a.corr);
    float winv = (fabsf(w) > 1e-3f) ? 1.0f / w : v.x;          // RF.cpp:1076-1077 (incl. its quirk)
    if (a.nIterWeight > 1) {
        // weight refinement passes, RF.cpp:1080-1092: a re-insertion adds w*slot[target] for every pair, i.e.
        // slot * (un-modulated weight sum of the voxel); where that exceeds 1e-3 the slot is divided by it
        float w2 = a.Wb2[b];
        if (x == 0 && uy <= c_geo.yHalf) w2 *= 0.5f;
        for (int it = 1; it < a.nIterWeight; ++it) {
            float wn = winv * w2;
            if (fabsf(wn) > 1e-3f) winv /= wn;
        }
    }
    if (1.0f / winv > 1e-3f) {                                 // RF.cpp:472-473
        float s = a.corr * winv;
        return make_float2(v.x * s, v.y * s);
    }
    return make_float2(0.f, 0.f);
}

// one thread per element of the natural half volume
__global__ void __launch_bounds__(256) k_normalize(const __grid_constant__ NormArgs a) {
    const Geometry& c_geo = a.geo;
    const int Z = c_geo.Z, X = c_geo.X;
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)Z * Z * X;
    if (idx >= total) return;
    int x = (int)(idx % X);
    size_t t = idx / X;
    int y = (int)(t % Z), z = (int)(t / Z);
    float2 g = d_norm_value(a, z, y, x);
    if (x == 0 || (Z % 2 == 0 && x == Z / 2)) {
        // The C2R transform of FFTW only sees the Hermitian part of these two planes; make it explicit
        // so that the result does not depend on how cuFFT treats non-Hermitian input.
        float2 m = d_norm_value(a, d_wrap(-z, Z), d_wrap(-y, Z), x);
        g.x = 0.5f * (g.x + m.x);
        g.y = 0.5f * (g.y - m.y);
    }
    a.out[idx] = g;
}

// blocked accumulators -> natural layout for inspection (x = 0 plane halved, see header)
__global__ void __launch_bounds__(256) k_export(const __grid_constant__ Geometry c_geo, const float2* __restrict__ Vb,
                                                const float* __restrict__ Wb, float2* __restrict__ V, float* __restrict__ W) {
    const int Z = c_geo.Z, X = c_geo.X;
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)Z * Z * X;
    if (idx >= total) return;
    int x = (int)(idx % X);
    size_t t = idx / X;
    int y = (int)(t % Z), z = (int)(t / Z);
    int uy = y <= Z / 2 ? y : y - Z, uz = z <= Z / 2 ? z : z - Z;
    int64_t b = d_blocked_index(c_geo, x, uy, uz);
    float2 v = Vb[b];
    float w = Wb[b];
    if (x == 0 && uy <= c_geo.yHalf) { v.x *= 0.5f; v.y *= 0.5f; w *= 0.5f; }
    V[idx] = v;
    W[idx] = w;
}

// Parameter upload by the SMs: reads the chunk's (small) parameter blocks straight from pinned, mapped host
// memory.  A DMA copy would queue behind the 268 MB particle transfers of the copy stream in the single
// host-to-device engine and stall the compute stream for milliseconds (seen in the stage timeline).
__global__ void __launch_bounds__(256) k_fetch_params(uint4* __restrict__ dst, const uint4* __restrict__ srcHost, size_t n16) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = srcHost[i];
}

// y += x (merging a saved half-set into the current accumulators)
__global__ void __launch_bounds__(256) k_axpy(float* __restrict__ y, const float* __restrict__ x, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] += x[i];
}

// deterministic two-stage FP64 sum of the weight accumulator (fixed grid, fixed tree)
__global__ void __launch_bounds__(256) k_weight_sum_partial(const float* __restrict__ Wb, int64_t n, double* __restrict__ partial) {
    __shared__ double sh[256];
    double acc = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) acc += (double)Wb[i];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
__global__ void __launch_bounds__(256) k_weight_sum_final(const double* __restrict__ partial, int n, double* __restrict__ out) {
    __shared__ double sh[256];
    double acc = 0;
    for (int i = threadIdx.x; i < n; i += 256) acc += partial[i];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sh[0];
}

// ================================================================== K3b
// out[kk][ii][jj] (N^3) = vol[k mod Z][i mod Z][j mod Z] * G[k^2+i^2+j^2], logical k = kk - N/2
__global__ void __launch_bounds__(256) k_crop_correct(const float* __restrict__ vol, const float* __restrict__ G,
                                                      float* __restrict__ out, int N, int Z) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)N * N * N;
    if (idx >= total) return;
    int jj = (int)(idx % N);
    size_t t = idx / N;
    int ii = (int)(t % N), kk = (int)(t / N);
    int first = -(N / 2);
    int k = first + kk, i = first + ii, j = first + jj;
    float v = __ldg(vol + ((size_t)d_wrap(k, Z) * Z + d_wrap(i, Z)) * Z + d_wrap(j, Z));
    out[idx] = v * __ldg(G + (k * k + i * i + j * j));
}

}  // namespace rfb200
