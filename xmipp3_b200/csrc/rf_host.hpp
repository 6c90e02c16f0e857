// rf_host.hpp — host-side (double precision) precomputation for the B200 direct
// Fourier reconstruction: blob tables, gridding-correction table, Euler/plane
// parameters, CTF constants, active-tile list and edge-item list.
//
// Everything here follows the reference's CPU program; citations are to
//   RF.cpp   = src/xmipp/libraries/reconstruction/reconstruct_fourier.cpp
//   blobs    = src/xmipp/libraries/data/blobs.cpp
//   ctf      = src/xmipp/libraries/data/ctf.{h,cpp}
// The Bessel polynomials are the Numerical-Recipes ones xmippCore uses (FP32 twins are
// visible in reconstruction_cuda/cuda_gpu_reconstruct_fourier.cpp:85-130).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "rf_types.h"

namespace rfb200 {
namespace host {

constexpr double kPi = 3.14159265358979323846;

// ---------------------------------------------------------------- Bessel / Kaiser
inline double bessi0(double x) {
    double ax = std::fabs(x);
    if (ax < 3.75) {
        double y = (x / 3.75) * (x / 3.75);
        return 1.0 + y * (3.5156229 + y * (3.0899424 + y * (1.2067492 + y * (0.2659732 + y * (0.360768e-1 + y * 0.45813e-2)))));
    }
    double y = 3.75 / ax;
    double poly = 0.39894228 + y * (0.1328592e-1 + y * (0.225319e-2 + y * (-0.157565e-2 + y * (0.916281e-2 +
                  y * (-0.2057706e-1 + y * (0.2635537e-1 + y * (-0.1647633e-1 + y * 0.392377e-2)))))));
    return (std::exp(ax) / std::sqrt(ax)) * poly;
}
inline double bessi1(double x) {
    double ax = std::fabs(x), ans;
    if (ax < 3.75) {
        double y = (x / 3.75) * (x / 3.75);
        ans = ax * (0.5 + y * (0.87890594 + y * (0.51498869 + y * (0.15084934 + y * (0.2658733e-1 + y * (0.301532e-2 + y * 0.32411e-3))))));
    } else {
        double y = 3.75 / ax;
        ans = 0.2282967e-1 + y * (-0.2895312e-1 + y * (0.1787654e-1 - y * 0.420059e-2));
        ans = 0.39894228 + y * (-0.3988024e-1 + y * (-0.362018e-2 + y * (0.163801e-2 + y * (-0.1031555e-1 + y * ans))));
        ans *= std::exp(ax) / std::sqrt(ax);
    }
    return x < 0 ? -ans : ans;
}
inline double bessi2(double x) { return x == 0 ? 0 : bessi0(x) - (2.0 / x) * bessi1(x); }
inline double bessi3(double x) { return x == 0 ? 0 : bessi1(x) - (4.0 / x) * bessi2(x); }
inline double bessi4(double x) { return x == 0 ? 0 : bessi2(x) - (6.0 / x) * bessi3(x); }
inline double bessi0_5(double x) { return x == 0 ? 0 : std::sqrt(2 / (kPi * x)) * std::sinh(x); }
inline double bessi1_5(double x) { return x == 0 ? 0 : std::sqrt(2 / (kPi * x)) * (std::cosh(x) - std::sinh(x) / x); }
inline double bessi2_5(double x) { return x == 0 ? 0 : bessi0_5(x) - 3.0 / x * bessi1_5(x); }
inline double bessi3_5(double x) { return x == 0 ? 0 : bessi1_5(x) - 5.0 / x * bessi2_5(x); }
inline double bessj1_5(double x) {
    if (x == 0) return 0;
    double rx = 1.0 / x;
    return std::sqrt(rx * 2 / kPi) * (std::sin(x) * rx - std::cos(x));
}
inline double bessj3_5(double x) {
    if (x == 0) return 0;
    double rx = 1.0 / x, rxs = rx * rx;
    return std::sqrt(rx * 2 / kPi) * ((15 * rxs * rx - 6 * rx) * std::sin(x) - (15 * rxs - 1) * std::cos(x));
}

// blobs.cpp:37-88
inline double kaiser_value(double r, double a, double alpha, int m) {
    double rda = r / a;
    if (rda > 1.0) return 0.0;
    double rdas = rda * rda, root = std::sqrt(1.0 - rdas), arg = alpha * root;
    switch (m) {
        case 0: return bessi0(arg) / bessi0(alpha);
        case 1: return alpha != 0.0 ? root * bessi1(arg) / bessi1(alpha) : root;
        case 2: return alpha != 0.0 ? root * root * bessi2(arg) / bessi2(alpha) : root * root;
        case 3: return alpha != 0.0 ? root * root * root * bessi3(arg) / bessi3(alpha) : root * root * root;
        case 4: return alpha != 0.0 ? root * root * root * root * bessi4(arg) / bessi4(alpha) : root * root * root * root;
    }
    return std::nan("");
}
// blobs.cpp:144-169
inline double kaiser_fourier_value(double w, double a, double alpha, int m) {
    double t = 2.0 * kPi * a * w;
    double sigma = std::sqrt(std::fabs(alpha * alpha - t * t));
    double lead = std::pow(2.0 * kPi, 1.5) * std::pow(a, 3.0);
    if (m == 2) {
        double b = (t > alpha) ? bessj3_5(sigma) : bessi3_5(sigma);
        return lead * std::pow(alpha, 2.0) * b / (bessi0(alpha) * std::pow(sigma, 3.5));
    }
    double b = (t > alpha) ? bessj1_5(sigma) : bessi1_5(sigma);
    return lead * b / (bessi0(alpha) * std::pow(sigma, 1.5));
}

struct Tables {
    std::vector<double> blobSqrt;      // blobTableSqrt     (uniform in d^2)
    std::vector<double> fourierBlob;   // Fourier_blob_table (uniform in radius)
    double iDeltaSqrt, iDeltaFourier;
};
// RF.cpp:224-269
inline Tables build_tables(int N, double padProj, double padVol, double r, int order, double alpha) {
    Tables t;
    const int T = kBlobTable;
    t.blobSqrt.resize(T);
    t.fourierBlob.resize(T);
    double rFourier = r / (padVol * N);
    double rNormalized = r / (padProj / padVol);
    double deltaSqrt = (r * r) / (T - 1);
    double deltaFourier = (std::sqrt(3.0) * N / 2.0) / (T - 1);
    double iw0 = 1.0 / kaiser_fourier_value(0.0, rNormalized, alpha, order);
    double padXdim3 = padVol * N;
    padXdim3 = padXdim3 * padXdim3 * padXdim3;
    double step = r * std::sqrt(1.0 / (T - 1));
    for (int i = 0; i < T; ++i) {
        t.blobSqrt[i] = kaiser_value(step * std::sqrt((double)i), r, order == 0 ? alpha : alpha, order) * iw0;
        t.fourierBlob[i] = kaiser_fourier_value(deltaFourier * i, rFourier, alpha, order) * padXdim3 * iw0;
    }
    t.iDeltaSqrt = 1.0 / deltaSqrt;
    t.iDeltaFourier = 1.0 / deltaFourier;
    return t;
}

inline double sinc_pi(double x) {
    if (std::fabs(x) < 0.0001) return 1.0;
    return std::sin(kPi * x) / (kPi * x);
}

// Gridding-correction table indexed by n = k^2+i^2+j^2 (RF.cpp:1153-1178).  The reference
// divides every voxel by (pad_vol/pad_proj)^3 * sinc^2(R/2N) * Fourier_blob_table[ROUND(R*iDeltaFourier)]
// and multiplies by the mean of sinc^2 over the N^3 crop; all of it depends on R^2 only.
inline std::vector<float> build_gridding_table(int N, double padProj, double padVol, const Tables& t, int nIterWeight,
                                               double* meanFactor2Out) {
    const int first = -(N / 2), last = first + N - 1;
    const int maxC = std::max(first * first, last * last);
    const size_t nmax = (size_t)3 * maxC;
    // histogram of k^2+i^2 then of k^2+i^2+j^2
    std::vector<int64_t> c2((size_t)2 * maxC + 1, 0), c3(nmax + 1, 0);
    for (int a = first; a <= last; ++a)
        for (int b = first; b <= last; ++b) c2[(size_t)a * a + (size_t)b * b]++;
    for (size_t m = 0; m < c2.size(); ++m) {
        if (!c2[m]) continue;
        for (int c = first; c <= last; ++c) c3[m + (size_t)c * c] += c2[m];
    }
    double padRel = padProj / padVol;
    padRel = padRel * padRel * padRel;
    double ipad = 1.0 / padRel;
    std::vector<double> f2(nmax + 1), fac(nmax + 1);
    double sum = 0;
    for (size_t n = 0; n <= nmax; ++n) {
        double radius = std::sqrt((double)n);
        long idx = std::lround(radius * t.iDeltaFourier);
        if (idx > kBlobTable - 1) idx = kBlobTable - 1;
        fac[n] = t.fourierBlob[idx];
        double s = sinc_pi(radius / (2.0 * N));
        f2[n] = s * s;
        sum += f2[n] * (double)c3[n];
    }
    double mean = sum / ((double)N * N * N);
    if (meanFactor2Out) *meanFactor2Out = mean;
    std::vector<float> g(nmax + 1);
    for (size_t n = 0; n <= nmax; ++n) {
        double v = (nIterWeight != 0) ? mean / (ipad * f2[n] * fac[n]) : 1.0 / (ipad * fac[n]);
        g[n] = (float)v;
    }
    return g;
}

// ---------------------------------------------------------------- geometry
inline int wrapi(int x, int n) { int r = x % n; return r < 0 ? r + n : r; }

// Euler_angles2matrix (xmippCore; pinned by test_binding.py:59-69), row-major 3x3
inline void euler_matrix(double rot, double tilt, double psi, double A[9]) {
    double a = rot * kPi / 180.0, b = tilt * kPi / 180.0, g = psi * kPi / 180.0;
    double ca = std::cos(a), cb = std::cos(b), cg = std::cos(g), sa = std::sin(a), sb = std::sin(b), sg = std::sin(g);
    double cc = cb * ca, cs = cb * sa, sc = sb * ca, ss = sb * sa;
    A[0] = cg * cc - sg * sa;  A[1] = cg * cs + sg * ca;  A[2] = -cg * sb;
    A[3] = -sg * cc - cg * sa; A[4] = -sg * cs + cg * ca; A[5] = sg * sb;
    A[6] = sc;                 A[7] = ss;                 A[8] = cb;
}

// Plane of (image, symmetry): M = R * A^T (RF.cpp:411-412, 936); e1, e2 scaled to pixel units.
inline void make_plane(const double R[9], double rot, double tilt, double psi, double pixPerVox, PlaneD& pd) {
    double A[9], M[9];
    euler_matrix(rot, tilt, psi, A);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += R[i * 3 + k] * A[j * 3 + k];   // (A^T)[k][j] = A[j][k]
            M[i * 3 + j] = s;
        }
    for (int c = 0; c < 3; ++c) {
        pd.e1[c] = M[c * 3 + 0] * pixPerVox;
        pd.e2[c] = M[c * 3 + 1] * pixPerVox;
        // M[:,2] is never used by the reference (freq z = 0); we need the plane normal, which is
        // e1 x e2 for a proper rotation and -(e1 x e2) for an improper one — the sign is irrelevant
        // because only |h| enters.  Use the third column directly.
        pd.n[c] = M[c * 3 + 2];
    }
}

// ctf.cpp:645-680 (produceSideInfo) and :1392-1404
inline CtfConsts make_ctf(double kV, double defocusU, double defocusV, double azimuth, double Cs, double Ca, double espr,
                          double ispr, double alpha, double DeltaF, double DeltaR, double Q0, double K, double envR0,
                          double envR1, double envR2, double phaseShift, double vppRadius) {
    CtfConsts c{};
    double local_Cs = Cs * 1e7, local_Ca = Ca * 1e7, local_kV = kV * 1e3, local_ispr = ispr * 1e6;
    double lambda = 12.2643247 / std::sqrt(local_kV * (1. + 0.978466e-6 * local_kV));
    c.K1 = kPi * lambda;
    c.K2 = kPi / 2 * local_Cs * lambda * lambda * lambda;
    c.K3 = std::pow(0.25 * kPi * local_Ca * lambda * (espr / kV + 2 * local_ispr), 2) / std::log(2.0);
    c.K5 = kPi * DeltaF * lambda;
    c.K6 = kPi * kPi * alpha * alpha;
    c.K7 = local_Cs * lambda * lambda;
    c.Ksin = std::sqrt(1 - Q0 * Q0);
    c.Kcos = Q0;
    c.K = K;
    c.DeltaR = DeltaR;
    c.envR0 = envR0; c.envR1 = envR1; c.envR2 = envR2;
    c.phase_shift = phaseShift;
    c.vpp_radius = vppRadius;
    double az = azimuth * kPi / 180.0;
    c.cos2az = std::cos(2 * az);
    c.sin2az = std::sin(2 * az);
    c.defocus_average = -(defocusU + defocusV) * 0.5;
    c.defocus_deviation = -(defocusU - defocusV) * 0.5;
    c.has_envelope = (c.K3 != 0 || c.K5 != 0 || c.K6 != 0 || DeltaR != 0 || envR0 != 0 || envR1 != 0 || envR2 != 0) ? 1 : 0;
    c.has_vpp = (std::round(vppRadius * 1000) != 0) ? 1 : 0;
    return c;
}

// jmax[i' - iLo] = largest j in [0, P/2] with (j/P)^2 + (i'/P)^2 <= maxRes^2, or -1.  Evaluated exactly
// like the reference's test (RF.cpp:594-598) so that the pass/fail decisions are identical.
inline void build_cutoff(int P, double maxRes, std::vector<int>& jmax, int& iLo, int& iHi, int& R) {
    iHi = P / 2;
    iLo = -(P - 1 - P / 2);
    jmax.assign(iHi - iLo + 1, -1);
    double maxRes2 = maxRes * maxRes;
    R = 0;
    for (int ip = iLo; ip <= iHi; ++ip) {
        double fy = (double)ip / P;
        for (int j = 0; j <= P / 2; ++j) {
            double fx = (double)j / P;
            if (fx * fx + fy * fy > maxRes2) continue;   // not monotone-safe to break on FP, scan all
            jmax[ip - iLo] = j;
            R = std::max(R, std::max(j, std::abs(ip)));
        }
    }
}

inline Geometry make_geometry(int N, double padProj, double padVol, double maxRes, double r, int R) {
    Geometry g{};
    g.N = N;
    g.P = (int)(N * padProj);
    g.Z = (int)(N * padVol);
    g.X = g.Z / 2 + 1;
    g.hi = g.Z / 2;
    g.lo = -(g.Z - 1 - g.Z / 2);
    g.tx = (g.Z / 2 + 1 + kTileX - 1) / kTileX;
    g.ty = (g.Z + kTileY - 1) / kTileY;
    g.tz = (g.Z + kTileZ - 1) / kTileZ;
    g.yHalf = (g.Z % 2 == 0) ? g.Z / 2 - 1 : g.Z / 2;
    double rho = r * g.P / (double)g.Z;
    g.rho = (float)rho;
    g.K = (int)std::floor(2 * rho) + 1;
    g.R = R;
    g.Rp = R + g.K + 1;
    g.side = 2 * g.Rp + 1;
    g.s2 = (float)(((double)g.Z / g.P) * ((double)g.Z / g.P));
    g.r = (float)r;
    g.r2 = (float)(r * r);
    g.iDelta = (float)((kBlobTable - 1) / (r * r));
    g.sMax = g.r2 * g.iDelta;
    g.reach = (float)(maxRes * g.Z + r);
    g.inplane_reach = (float)(R + rho);
    g.colOff = g.K;
    g.pitch = (g.Rp + g.colOff + 2 + 1) & ~1;
    g.planeStride = g.side * g.pitch;
    g.xOwnMax = (g.Z % 2 == 0) ? g.Z / 2 - 1 : g.Z / 2;
    g.rimIn2 = -1.f;   // set by build_rim_table
    return g;
}

// Validity / multiplicity of the FULL-plane slice pixels, one packed entry per centred row i in [-Rp, Rp]:
//   (jPos+1) | (jNeg+1) << 14 | m0 << 28,  jPos = jmax of original row i (columns j > 0), jNeg = jmax of original
//   row -i (columns j < 0 are Hermitian mirrors), m0 = multiplicity of column 0 (original (0,i) + mirror of (0,-i)).
// Also sets g.rimIn2: inside that pixel radius (minus the window reach) every pixel is valid, so the gather can
// count weights without looking at the table.
inline std::vector<int32_t> build_rim_table(Geometry& g, const std::vector<int>& jmax, int iLo, int iHi) {
    auto jm = [&](int i) { return (i >= iLo && i <= iHi) ? jmax[i - iLo] : -1; };
    std::vector<int32_t> t(g.side);
    double minInvalid2 = 1e300;
    for (int i = -g.Rp; i <= g.Rp; ++i) {
        int jp = jm(i), jn = jm(-i);
        int m0 = (jp >= 0 ? 1 : 0) + (jn >= 0 ? 1 : 0);
        t[i + g.Rp] = (jp + 1) | ((jn + 1) << 14) | (m0 << 28);
        // first invalid pixel of the row on either side (column 0 is handled by the column-0 test of the gather,
        // but a row where it is not doubly valid must take the slow path too)
        double i2 = (double)i * i;
        minInvalid2 = std::min(minInvalid2, i2 + (double)(jp + 1) * (jp + 1));
        minInvalid2 = std::min(minInvalid2, i2 + (double)(jn + 1) * (jn + 1));
        if (m0 != 2) minInvalid2 = std::min(minInvalid2, i2);
    }
    // only ACCEPTED candidates count, and those lie within the blob radius rho (pixel units) of the projected voxel
    double rfree = std::sqrt(minInvalid2) - 1e-2 - (double)g.rho;
    g.rimIn2 = rfree > 0 ? (float)(rfree * rfree) : -1.f;
    return t;
}

// Sticks of class cls (0: d = x, 1: d = y, 2: d = z) whose lattice box comes within `reach` of the origin and holds
// at least one voxel owned by the main gather; origins in stored-offset coordinates permuted to (a,b,d); sorted by
// distance so that concurrently running warps work on neighbouring sticks (slice rings stay in L1/L2) and the
// heavy central sticks start first.
inline std::vector<StickUnit> build_stick_units(const Geometry& g, int cls) {
    struct T { float d; StickUnit u; };
    std::vector<T> v;
    const int ext[3] = {g.tx * kTileX, g.ty * kTileY, g.tz * kTileZ};       // allocated extents x, y-lo, z-lo
    const int ax[3][3] = {{1, 2, 0}, {0, 2, 1}, {0, 1, 2}};                 // (a,b,d) -> natural axis
    const int A = ax[cls][0], B = ax[cls][1], D = ax[cls][2];
    const int off[3] = {0, g.lo, g.lo};
    const int maxc[3] = {g.xOwnMax, g.hi, g.hi};                            // largest owned centred coordinate
    auto axis = [](int a, int b) { return (a > 0) ? (double)a : (b < 0 ? (double)-b : 0.0); };
    for (int t0 = 0; t0 < ext[D]; t0 += kStickL)
        for (int b0 = 0; b0 < ext[B]; b0 += kStickB)
            for (int a0 = 0; a0 < ext[A]; a0 += kStickA) {
                int lo3[3], hi3[3];
                lo3[A] = a0 + off[A]; hi3[A] = std::min(a0 + kStickA - 1 + off[A], maxc[A]);
                lo3[B] = b0 + off[B]; hi3[B] = std::min(b0 + kStickB - 1 + off[B], maxc[B]);
                lo3[D] = t0 + off[D]; hi3[D] = std::min(t0 + kStickL - 1 + off[D], maxc[D]);
                if (hi3[0] < lo3[0] || hi3[1] < lo3[1] || hi3[2] < lo3[2]) continue;
                double dx = axis(lo3[0], hi3[0]), dy = axis(lo3[1], hi3[1]), dz = axis(lo3[2], hi3[2]);
                double d = std::sqrt(dx * dx + dy * dy + dz * dz);
                if (d > g.reach + 1e-3) continue;
                v.push_back({(float)d, StickUnit{a0, b0, t0, 0}});
            }
    std::stable_sort(v.begin(), v.end(), [](const T& a, const T& b) { return a.d < b.d; });
    std::vector<StickUnit> out(v.size());
    for (size_t i = 0; i < v.size(); ++i) out[i] = v[i].u;
    return out;
}

// class of a plane = axis dominating its normal (ties -> lowest axis)
inline int plane_class(const PlaneD& pd) {
    double ax = std::fabs(pd.n[0]), ay = std::fabs(pd.n[1]), az = std::fabs(pd.n[2]);
    if (ax >= ay && ax >= az) return 0;
    if (ay >= az) return 1;
    return 2;
}
// permute the components of a plane to the (a,b,d) order of its class
inline void permute_plane(const PlaneD& pd, int cls, int img, float weight, PlaneD& pdp, PlaneS& ps) {
    const int ax[3][3] = {{1, 2, 0}, {0, 2, 1}, {0, 1, 2}};
    for (int c = 0; c < 3; ++c) {
        pdp.e1[c] = pd.e1[ax[cls][c]];
        pdp.e2[c] = pd.e2[ax[cls][c]];
        pdp.n[c] = pd.n[ax[cls][c]];
    }
    ps.e1a = (float)pdp.e1[0]; ps.e1b = (float)pdp.e1[1]; ps.e1d = (float)pdp.e1[2];
    ps.e2a = (float)pdp.e2[0]; ps.e2b = (float)pdp.e2[1]; ps.e2d = (float)pdp.e2[2];
    ps.na = (float)pdp.n[0]; ps.nb = (float)pdp.n[1]; ps.nd = (float)pdp.n[2];
    ps.img = img;
    ps.weight = weight;
    ps.invNd = (float)(1.0 / pdp.n[2]);
}

// local voxel (vx,vy,vz) in [0,16)x[0,16)x[0,8) -> slot inside a tile: brick * 32 + lane (4x4x2 bricks whose 32
// slots are contiguous in memory)
inline int tile_slot(int vx, int vy, int vz) {
    int brick = (vx >> 2) | ((vy >> 2) << 2) | ((vz >> 1) << 4);
    int lane = (vx & 3) | ((vy & 3) << 2) | ((vz & 1) << 4);
    return brick * 32 + lane;
}
// blocked index of centred lattice point (ux in [0,Z/2], uy,uz in [lo,hi])
inline int64_t blocked_index(const Geometry& g, int ux, int uy, int uz) {
    int x = ux, y = uy - g.lo, z = uz - g.lo;
    int64_t tile = ((int64_t)(z / kTileZ) * g.ty + (y / kTileY)) * g.tx + (x / kTileX);
    return tile * kTileVox + tile_slot(x % kTileX, y % kTileY, z % kTileZ);
}
inline int centred(const Geometry& g, int stored) { return stored <= g.Z / 2 ? stored : stored - g.Z; }

inline bool cond_orig(const Geometry& g, int ux) { return wrapi(ux, g.Z) <= g.Z / 2; }          // RF.cpp:748 (not conjugate)
inline bool cond_mirr(const Geometry& g, int ux) { return wrapi(-ux, g.Z) > g.Z / 2; }           // mirrored pixel, conjugate case
// Does the main gather own this natural lattice point (both originals and mirrors contribute, or the
// x = 0 plane whose stored value is kept as orig+mirr = 2x the pair average)?
inline bool main_owns(const Geometry& g, int ux, int uy, int /*uz*/) {
    if (ux == 0) return uy <= g.yHalf;
    return cond_orig(g, ux) && cond_mirr(g, ux);
}

// Edge items: natural lattice points the main gather does not own, and the
// wrap-around aliases at the Nyquist faces (RF.cpp:662-693 wraps every index mod Z).
inline std::vector<EdgeItem> build_edge_items(const Geometry& g) {
    std::vector<EdgeItem> items;
    const int Z = g.Z;
    const int m = (int)std::ceil(g.r);
    const double reach2 = (double)g.reach * g.reach + 1e-6;
    auto natural = [&](int ux, int uy, int uz) { return ux >= 0 && ux <= Z / 2 && uy >= g.lo && uy <= g.hi && uz >= g.lo && uz <= g.hi; };
    auto consider = [&](int ux, int uy, int uz) {
        double d2 = (double)ux * ux + (double)uy * uy + (double)uz * uz;
        if (d2 > reach2) return;
        bool o = cond_orig(g, ux), mi = cond_mirr(g, ux);
        if (!o && !mi) return;
        if (natural(ux, uy, uz) && main_owns(g, ux, uy, uz)) return;
        int sx = wrapi(ux, Z), sy = wrapi(uy, Z), sz = wrapi(uz, Z);
        if (sx > Z / 2) return;   // cannot happen when o is true; mirrors imply o here
        int cy = centred(g, sy), cz = centred(g, sz);
        EdgeItem it;
        it.ux = ux; it.uy = uy; it.uz = uz;
        it.mode = (o && mi) ? 0 : 1;
        // stored voxel on the x = 0 plane that the main gather keeps in "orig+mirr" form: aliases follow suit
        if (sx == 0 && cy <= g.yHalf) it.mode = 0;
        it.store = blocked_index(g, sx, cy, cz);
        items.push_back(it);
    };
    // A) wrapped x (ux <= -Z/2): any y, z
    for (int ux = -(Z / 2) - m; ux <= -(Z + 1) / 2; ++ux)
        for (int uy = g.lo - m; uy <= g.hi + m; ++uy)
            for (int uz = g.lo - m; uz <= g.hi + m; ++uz) consider(ux, uy, uz);
    // B) natural x, y or z wrapped
    for (int ux = 0; ux <= Z / 2; ++ux) {
        for (int uy = g.lo - m; uy <= g.hi + m; ++uy) {
            bool yNat = (uy >= g.lo && uy <= g.hi);
            if (!yNat) {
                for (int uz = g.lo - m; uz <= g.hi + m; ++uz) consider(ux, uy, uz);
            } else {
                for (int uz = g.lo - m; uz < g.lo; ++uz) consider(ux, uy, uz);
                for (int uz = g.hi + 1; uz <= g.hi + m; ++uz) consider(ux, uy, uz);
            }
        }
    }
    // C) natural points not owned by the main gather (x = Z/2 plane for even Z, x = 0 exception row)
    for (int ux = 0; ux <= Z / 2; ++ux) {
        bool anyUnowned = (ux == 0) ? (g.yHalf < g.hi) : !(cond_orig(g, ux) && cond_mirr(g, ux));
        if (!anyUnowned) continue;
        for (int uy = g.lo; uy <= g.hi; ++uy) {
            if (main_owns(g, ux, uy, 0)) continue;
            for (int uz = g.lo; uz <= g.hi; ++uz) consider(ux, uy, uz);
        }
    }
    // Several aliases can add into the same stored voxel; sort by target so that one thread owns each
    // target (no atomics, deterministic order).
    std::stable_sort(items.begin(), items.end(), [](const EdgeItem& a, const EdgeItem& b) { return a.store < b.store; });
    return items;
}

// start offsets of the runs of equal `store` in a sorted item list (+ sentinel)
inline std::vector<int32_t> edge_group_starts(const std::vector<EdgeItem>& items) {
    std::vector<int32_t> starts;
    for (size_t i = 0; i < items.size(); ++i)
        if (i == 0 || items[i].store != items[i - 1].store) starts.push_back((int32_t)i);
    starts.push_back((int32_t)items.size());
    return starts;
}

}  // namespace host
}  // namespace rfb200
