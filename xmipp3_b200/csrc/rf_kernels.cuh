// rf_kernels.cuh — hand-written sm_100a kernels of the direct Fourier reconstruction (everything except the
// insertion itself, which lives in rf_sticks.cuh).
//
//   K1a k_pad_images     raw N x N particle -> zero-padded, centred, shifted P x P image  (RF.cpp:388-402)
//   (cuFFT R2C batched)                                                                  (RF.cpp:405-407)
//   CTF device functions per-pixel wCTF / wModulator (RF.cpp:600-625, data/ctf.h:452-502) used by K1b'
//   K3a k_normalize      weight symmetrisation + normalisation (RF.cpp:1056-1101, 453-479, 1188-1221)
//   (cuFFT C2R 3-D)                                                                      (RF.cpp:1145)
//   K3b k_crop_correct   CenterFFT + crop + gridding correction (RF.cpp:1146-1178)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rf_types.h"

namespace rfb200 {

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
    // stored-offset coordinates are non-negative: unsigned arithmetic turns the divisions into plain shifts and masks
    const unsigned x = (unsigned)ux, y = (unsigned)(uy - geo.lo), z = (unsigned)(uz - geo.lo);
    const unsigned tile = ((z / kTileZ) * (unsigned)geo.ty + (y / kTileY)) * (unsigned)geo.tx + (x / kTileX);
    return (int64_t)tile * kTileVox + d_tile_slot((int)(x % kTileX), (int)(y % kTileY), (int)(z % kTileZ));
}
// natural lattice point owned by the tile gather? (host twin: host::main_owns)
__device__ __forceinline__ bool d_main_owns(const Geometry& geo, int ux, int uy) {
    const int Z = geo.Z;
    if (ux == 0) return uy <= geo.yHalf;
    if (ux > Z / 2) return false;
    return d_wrap(-ux, Z) > Z / 2;
}

// ================================================================== K1a
// Cubic B-spline coefficients for the images whose shift is fractional: xmippCore's readApplyGeo interpolates with
// BSPLINE3.  produceSplineCoefficients = bilib direct transform with the MIRROR-OFF-BOUNDS (half-sample
// symmetric) extension; convention pinned by the reference's rotate() known answer (test_transformation_main.cpp:76-95,
// tests/test_oracle_kat.py).  One thread filters one line; `stride` selects rows (1) or columns (N).  Recursion in double.
__device__ __forceinline__ void d_bspline_line(float* c, int n, long stride) {
    if (n == 1) return;
    const double z1 = -0.26794919243112270647, lambda = 6.0;   // sqrt(3)-2, (1-z1)(1-1/z1)
    // causal initialisation: c+[0] = s[0] + z1 * sum_{m>=0} z1^m s~[m] over the mirrored, 2n-periodic signal (z1^80 ~ 1e-46)
    double sum = 0.0, zm = 1.0;
    for (int m = 0; m <= 80; ++m) {
        const int q = m % (2 * n);
        const int idx = q < n ? q : 2 * n - 1 - q;
        sum += zm * lambda * (double)c[idx * stride];
        zm *= z1;
    }
    // causal pass; the running value is kept in double, the line holds float
    double prev = lambda * (double)c[0] + z1 * sum;
    c[0] = (float)prev;
    for (int k = 1; k < n; ++k) {
        prev = lambda * (double)c[k * stride] + z1 * prev;
        c[k * stride] = (float)prev;
    }
    // anticausal pass
    double nxt = (z1 / (z1 - 1.0)) * prev;
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
        // applyGeometry: the source coordinate (j + mx) + ux is wrapped into [-0.5, N - 0.5) (realWRAP), then
        // interpolatedElementBSpline2D reflects out-of-range neighbours (l < 0 -> -l-1, l >= N -> 2N-l-1)
        float wx[4], wy[4];
        d_bspline3_weights(q.ux, wx);
        d_bspline3_weights(q.uy, wy);
        int bj = j + q.mx, bi = i + q.my;                       // floor of the source coordinate
        {
            const float xs = (float)bj + q.ux, ys = (float)bi + q.uy;
            if (xs < -1e-6f || xs > (float)(N - 1) + 1e-6f) bj -= (int)floorf((xs + 0.5f) / (float)N) * N;
            if (ys < -1e-6f || ys > (float)(N - 1) + 1e-6f) bi -= (int)floorf((ys + 0.5f) / (float)N) * N;
        }
        const float* c = coef + (size_t)img * N * N;
        v = 0.f;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            int ri = bi - 1 + a;
            ri = ri < 0 ? -ri - 1 : (ri >= N ? 2 * N - ri - 1 : ri);
            const float* row = c + (size_t)ri * N;
            float r = 0.f;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                int rj = bj - 1 + b;
                rj = rj < 0 ? -rj - 1 : (rj >= N ? 2 * N - rj - 1 : rj);
                r = fmaf(wx[b], __ldg(row + rj), r);
            }
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
    } else {
        if constexpr (sizeof(T) == 4) wc = __frcp_rn(wc);   // correctly rounded reciprocal, no slow path
        else wc = T(1) / wc;
    }
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

// Per-image constants of the per-pixel evaluation, prepared once per CTA: FP32 copies of the amplitude terms and
// the phase coefficients as 0.64 fixed-point TURNS.  With integer pixel frequencies (j, ip) the phase of the pure
// CTF (no envelope / phase plate) is
//   arg / 2pi = c1 * r2 + c2 * r2^2 + cA * (j^2 - ip^2) + cB * (2 j ip),   r2 = j^2 + ip^2,
//   c1 = a2 K1 dAvg / 2pi,  c2 = a2^2 K2 / 2pi,  cA = a2 K1 dDev cos(2 az) / 2pi,  cB = a2 K1 dDev sin(2 az) / 2pi
// (the astigmatism term cos(2(atan2(Y,X) - az)) * r2 is a polynomial in j, ip), and only its fractional part
// matters: 64-bit integer multiply-adds wrap exactly like the angle does, so the range reduction is exact and the
// per-pixel path needs neither FP64 nor a division.
struct CtfFloat {
    float KK, Kcos, Ksin, cos2az, sin2az;
    int fast;                       // 1: fixed-point phase path valid (no envelope, no phase plate)
    long long fix1, fix2, fixA, fixB;
};
__device__ __forceinline__ long long d_turns_to_fix(double turns) {
    const double fr = turns - rint(turns);                     // [-0.5, 0.5]
    return (long long)((unsigned long long)llrint(fr * 9223372036854775808.0) << 1);   // * 2^64, modulo 2^64
}
__device__ __forceinline__ void d_ctf_prepare(const CtfConsts& c, const SliceParams& sp, CtfFloat& f) {
    f.KK = (float)(c.K * c.K);
    f.Kcos = (float)c.Kcos;
    f.Ksin = (float)c.Ksin;
    f.cos2az = (float)c.cos2az;
    f.sin2az = (float)c.sin2az;
    const double i2pi = 0.15915494309189533577;
    f.fix1 = d_turns_to_fix(sp.a2 * c.K1 * c.defocus_average * i2pi);
    f.fix2 = d_turns_to_fix(sp.a2 * sp.a2 * c.K2 * i2pi);
    f.fixA = d_turns_to_fix(sp.a2 * c.K1 * c.defocus_deviation * c.cos2az * i2pi);
    f.fixB = d_turns_to_fix(sp.a2 * c.K1 * c.defocus_deviation * c.sin2az * i2pi);
    // precomputeValues zeroes deltaf only where |X|,|Y| < 1e-6: with a >= 1e-6 that is the origin, where arg = 0 anyway
    f.fast = (!c.has_vpp && !c.has_envelope && sp.a >= 1e-6f) ? 1 : 0;
}

// sin and cos of (q quarter turns + x), x in [-pi/4, pi/4]: the classic single-precision minimax kernels (~1 ulp)
__device__ __forceinline__ void d_sincos_quadrant(float x, int q, float& sn, float& cs) {
    const float x2 = x * x;
    const float sp_ = fmaf(x * x2, fmaf(x2, fmaf(x2, -1.9515295891e-4f, 8.3321608736e-3f), -1.6666654611e-1f), x);
    const float cp_ = fmaf(x2 * x2, fmaf(x2, fmaf(x2, 2.443315711809948e-5f, -1.388731625493765e-3f), 4.166664568298827e-2f),
                           fmaf(x2, -0.5f, 1.0f));
    sn = (q & 1) ? cp_ : sp_;
    cs = (q & 1) ? sp_ : cp_;
    if (q & 2) sn = -sn;
    if ((q + 1) & 2) cs = -cs;
}

// wCTF / wModulator of half-plane pixel (j, ip) — RF.cpp:600-625.  Fast path: exact fixed-point phase (see CtfFloat).
// General path (envelope or phase plate): the phase argument is formed in double from exact integer frequencies and
// reduced to [-pi/4, pi/4] in double.  sin/cos, the amplitude and the minCTF rules run in FP32.
// kGeneral = false: the caller guarantees f.fast (no image of the launch has envelope or phase-plate terms), and the
// double-precision general path is not even compiled into the kernel (registers).
template <bool kGeneral = true>
__device__ __forceinline__ void d_ctf_weights(const CtfConsts& c, const CtfFloat& f, const SliceParams& sp, int j, int ip, float& wCTF, float& wMod) {
    float sn, cs, E = 1.0f;
    if (!kGeneral || f.fast) {
        const int jj = j * j, ii = ip * ip;
        typedef unsigned long long u64;                      // all products wrap modulo 2^64 = one turn
        const u64 r2 = (u64)(jj + ii);
        const u64 ph = ((u64)f.fix1 + (u64)f.fix2 * r2) * r2 + (u64)f.fixA * (u64)(long long)(jj - ii) + (u64)f.fixB * (u64)(long long)(2 * j * ip);
        const unsigned long long rq = ph + (1ull << 61);                     // round to the nearest quarter turn
        const int q = (int)(rq >> 62);
        const int res = (int)((long long)(ph - ((rq >> 62) << 62)) >> 32);    // residual in 2^-32 turns, |res| <= 2^29
        const float x = (float)res * 1.4629180792671596e-9f;                 // * 2 pi / 2^32
        d_sincos_quadrant(x, q, sn, cs);
    } else {
        const int r2i = j * j + ip * ip;                      // exact: |freq|^2 in units of (1/(P*Ts))^2
        const double u2 = (double)r2i * sp.a2;
        double deltaf = 0.0;
        const float ax = fabsf((float)j) * sp.a, ay = fabsf((float)ip) * sp.a;
        if (!(ax < 1e-6f && ay < 1e-6f)) {                    // precomputeValues(X,Y): deltaf = 0 at the origin
            const float inv = __frcp_rn((float)r2i);
            const float c2 = (float)(j * j - ip * ip) * inv, s2 = (float)(2 * j * ip) * inv;
            deltaf = c.defocus_average + c.defocus_deviation * (double)(c2 * f.cos2az + s2 * f.sin2az);
        }
        double arg = u2 * fma(c.K1, deltaf, c.K2 * u2);
        if (c.has_vpp) arg += -c.phase_shift * (1.0 - exp(-u2 / (2.0 * c.vpp_radius * c.vpp_radius)));
        const double two_over_pi = 0.63661977236758134, pio2_hi = 1.5707963267948966, pio2_lo = 6.123233995736766e-17;
        const double kq = rint(arg * two_over_pi);
        double red = fma(-kq, pio2_hi, arg);
        red = fma(-kq, pio2_lo, red);
        d_sincos_quadrant((float)red, (int)kq, sn, cs);
        if (c.has_envelope) E = (float)d_ctf_envelope(c, u2, deltaf);
    }
    const float v = f.KK * (f.Kcos * cs - f.Ksin * sn) * E;
    if (fabsf(fabsf(v) - sp.minCtfF) < 4e-6f) {
        d_ctf_weights_exact(c, sp, j, ip, wCTF, wMod);
        return;
    }
    d_ctf_rules<float>(v, sp.minCtfF, sp.phaseFlipped, j, ip, wCTF, wMod);
}

constexpr int kSliceRowsPerThread = 4;   // rows handled by one thread of the slice kernel

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

// ---- peer-memory reduce over NVLink (rfb200_reduce_p2p).  Every rank owns the slice [lo, hi) of the accumulators: it
// reads that slice from the memory of ALL ranks (its own locally, the others through their IPC-mapped pointers, 16-byte
// loads), adds them in rank order (so the sum does not depend on who computes it or on timing) and stores the result
// into the root's memory.  In NCCL terms a reduce-scatter and a gather to the root in one kernel, with every link of the
// switch carrying 1/N of the data instead of the root's links carrying all of it.
constexpr int kMaxP2PRanks = 16;
struct P2PArgs {
    const float4* src[kMaxP2PRanks];   // the same array of every rank (as float4)
    float4* dst;                       // the root's array
    long long lo, hi;                  // float4 range this rank reduces
};
template <int NR>
__global__ void __launch_bounds__(256) k_reduce_p2p(const __grid_constant__ P2PArgs a, int nRanksRt) {
    const int nr = NR > 0 ? NR : nRanksRt;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = a.lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.hi; i += stride) {
        float4 v[NR > 0 ? NR : 1];
        float4 acc;
        if (NR > 0) {
#pragma unroll
            for (int k = 0; k < NR; ++k) v[k] = a.src[k][i];          // all loads in flight before the first add
            acc = v[0];
#pragma unroll
            for (int k = 1; k < NR; ++k) { acc.x += v[k].x; acc.y += v[k].y; acc.z += v[k].z; acc.w += v[k].w; }
        } else {
            acc = a.src[0][i];
            for (int k = 1; k < nr; ++k) {
                const float4 t = a.src[k][i];
                acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
            }
        }
        a.dst[i] = acc;
    }
}

// y += x (merging a saved half-set into the current accumulators)
__global__ void __launch_bounds__(256) k_axpy(float* __restrict__ y, const float* __restrict__ x, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] += x[i];
}

// deterministic two-stage FP64 sum of the weight accumulator (fixed grid, fixed tree)
// partial sums of Wb[i] (+ extra[4 i + 2] when `extra` is given: the W lane of the --fast scratch accumulators)
__global__ void __launch_bounds__(256) k_weight_sum_partial(const float* __restrict__ Wb, int64_t n, double* __restrict__ partial,
                                                            const float* __restrict__ extra = nullptr) {
    __shared__ double sh[256];
    double acc = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        acc += (double)Wb[i];
        if (extra) acc += (double)extra[4 * i + 2];
    }
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
