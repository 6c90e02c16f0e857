// =============================================================================
// oracle/recfourier_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Double-precision CPU restatement of Xmipp's direct Fourier reconstruction
// (ProgRecFourier, reference file
//   src/xmipp/libraries/reconstruction/reconstruct_fourier.cpp  == "RF.cpp").
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load this library; the product path (xmipp3_b200/csrc)
// never links or calls it.
//
// PARITY STATUS: the reference cannot be compiled here (xmippCore, FFTW, HDF5,
// SQLite absent) and its only end-to-end test uses an external gold volume, so
// the END-TO-END result of this oracle is "parity unpinned".  The sub-steps are
// pinned by the reference's own known-answer tests (see tests/test_oracle_kat.py):
//   forward FFT scale+sign   test_fftw_main.cpp:35-51
//   FFT_IDX2DIGFREQ          test_fftw_main.cpp:80-108
//   Euler_angles2matrix      test_binding.py:59-69, data/euler.cpp (ZYZ)
// and the scatter loop is cross-checked against an independent numpy
// restatement (oracle/mini_oracle.py).
//
// Every function cites the reference lines it restates.  Nothing here is copied
// from the reference: containers, FFT, threading and control flow are our own.
// =============================================================================
#include <algorithm>
#include <atomic>
#include <cmath>
#include <complex>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "oracle_abi.h"

namespace {

using cd = std::complex<double>;
constexpr double kPi = 3.14159265358979323846;
constexpr int kTable = 10000;           // BLOB_TABLE_SIZE_SQRT, reconstruct_fourier.h:41-44
constexpr double kAccuracy = 0.001;     // ACCURACY, reconstruct_fourier.h:41-44
constexpr double kEqualAccuracy = 1e-6; // XMIPP_EQUAL_ACCURACY (xmippCore)

// ---------------------------------------------------------------------------
// Numerical-Recipes polynomial Bessel functions, as used by xmippCore
// (FP32 twins visible in reconstruction_cuda/cuda_gpu_reconstruct_fourier.cpp:85-130).
// ---------------------------------------------------------------------------
double bessi0(double x) {
    double ax = std::fabs(x);
    if (ax < 3.75) {
        double y = x / 3.75;
        y *= y;
        return 1.0 + y * (3.5156229 + y * (3.0899424 + y * (1.2067492 + y * (0.2659732 + y * (0.360768e-1 + y * 0.45813e-2)))));
    }
    double y = 3.75 / ax;
    return (std::exp(ax) / std::sqrt(ax)) *
           (0.39894228 + y * (0.1328592e-1 + y * (0.225319e-2 + y * (-0.157565e-2 + y * (0.916281e-2 + y * (-0.2057706e-1 + y * (0.2635537e-1 + y * (-0.1647633e-1 + y * 0.392377e-2))))))));
}
double bessi1(double x) {
    double ax = std::fabs(x), ans;
    if (ax < 3.75) {
        double y = x / 3.75;
        y *= y;
        ans = ax * (0.5 + y * (0.87890594 + y * (0.51498869 + y * (0.15084934 + y * (0.2658733e-1 + y * (0.301532e-2 + y * 0.32411e-3))))));
    } else {
        double y = 3.75 / ax;
        ans = 0.2282967e-1 + y * (-0.2895312e-1 + y * (0.1787654e-1 - y * 0.420059e-2));
        ans = 0.39894228 + y * (-0.3988024e-1 + y * (-0.362018e-2 + y * (0.163801e-2 + y * (-0.1031555e-1 + y * ans))));
        ans *= (std::exp(ax) / std::sqrt(ax));
    }
    return x < 0.0 ? -ans : ans;
}
double bessi2(double x) { return x == 0 ? 0 : bessi0(x) - (2.0 / x) * bessi1(x); }
double bessi3(double x) { return x == 0 ? 0 : bessi1(x) - (4.0 / x) * bessi2(x); }
double bessi4(double x) { return x == 0 ? 0 : bessi2(x) - (6.0 / x) * bessi3(x); }
double bessi0_5(double x) { return x == 0 ? 0 : std::sqrt(2 / (kPi * x)) * std::sinh(x); }
double bessi1_5(double x) { return x == 0 ? 0 : std::sqrt(2 / (kPi * x)) * (std::cosh(x) - std::sinh(x) / x); }
double bessi2_5(double x) { return x == 0 ? 0 : bessi0_5(x) - 3.0 / x * bessi1_5(x); }
double bessi3_5(double x) { return x == 0 ? 0 : bessi1_5(x) - 5.0 / x * bessi2_5(x); }
double bessj1_5(double x) {
    if (x == 0) return 0;
    double rx = 1.0 / x;
    return std::sqrt(rx * 2 / kPi) * (std::sin(x) * rx - std::cos(x));
}
double bessj3_5(double x) {
    if (x == 0) return 0;
    double rx = 1.0 / x, rxs = rx * rx;
    return std::sqrt(rx * 2 / kPi) * ((15 * rxs * rx - 6 * rx) * std::sin(x) - (15 * rxs - 1) * std::cos(x));
}
double bessj0(double x) {
    double ax = std::fabs(x);
    if (ax < 8.0) {
        double y = x * x;
        double a1 = 57568490574.0 + y * (-13362590354.0 + y * (651619640.7 + y * (-11214424.18 + y * (77392.33017 + y * (-184.9052456)))));
        double a2 = 57568490411.0 + y * (1029532985.0 + y * (9494680.718 + y * (59272.64853 + y * (267.8532712 + y * 1.0))));
        return a1 / a2;
    }
    double z = 8.0 / ax, y = z * z, xx = ax - 0.785398164;
    double a1 = 1.0 + y * (-0.1098628627e-2 + y * (0.2734510407e-4 + y * (-0.2073370639e-5 + y * 0.2093887211e-6)));
    double a2 = -0.1562499995e-1 + y * (0.1430488765e-3 + y * (-0.6911147651e-5 + y * (0.7621095161e-6 - y * 0.934935152e-7)));
    return std::sqrt(0.636619772 / ax) * (std::cos(xx) * a1 - z * std::sin(xx) * a2);
}

// kaiser_value — data/blobs.cpp:37-88
double kaiser_value(double r, double a, double alpha, int m) {
    double rda = r / a;
    if (rda > 1.0) return 0.0;
    double rdas = rda * rda;
    double arg = alpha * std::sqrt(1.0 - rdas);
    double w;
    switch (m) {
        case 0: w = bessi0(arg) / bessi0(alpha); break;
        case 1: w = std::sqrt(1.0 - rdas); if (alpha != 0.0) w *= bessi1(arg) / bessi1(alpha); break;
        case 2: w = std::sqrt(1.0 - rdas); w = w * w; if (alpha != 0.0) w *= bessi2(arg) / bessi2(alpha); break;
        case 3: w = std::sqrt(1.0 - rdas); w = w * w * w; if (alpha != 0.0) w *= bessi3(arg) / bessi3(alpha); break;
        case 4: w = std::sqrt(1.0 - rdas); w = w * w * w * w; if (alpha != 0.0) w *= bessi4(arg) / bessi4(alpha); break;
        default: w = std::nan(""); break;
    }
    return w;
}
// kaiser_Fourier_value — data/blobs.cpp:144-169 (orders 0 and 2 only)
double kaiser_Fourier_value(double w, double a, double alpha, int m) {
    double t = 2. * kPi * a * w;
    double sigma = std::sqrt(std::fabs(alpha * alpha - t * t));
    if (m == 2) {
        double b = (t > alpha) ? bessj3_5(sigma) : bessi3_5(sigma);
        return std::pow(2. * kPi, 1.5) * std::pow(a, 3.) * std::pow(alpha, 2.) * b / (bessi0(alpha) * std::pow(sigma, 3.5));
    }
    if (m == 0) {
        double b = (t > alpha) ? bessj1_5(sigma) : bessi1_5(sigma);
        return std::pow(2. * kPi, 1.5) * std::pow(a, 3.) * b / (bessi0(alpha) * std::pow(sigma, 1.5));
    }
    return std::nan("");
}

// ---------------------------------------------------------------------------
// Mixed-radix complex FFT (our own; stands in for FFTW behind xmippCore's
// FourierTransformer).  sign=-1 forward, +1 backward, both unnormalised.
// ---------------------------------------------------------------------------
struct Fft {
    int n = 0;
    std::vector<cd> tw_f;   // exp(-2 pi i k / n)
    std::vector<cd> scratch;
    explicit Fft(int n_) : n(n_), tw_f(n_), scratch(n_) {
        for (int k = 0; k < n; ++k) {
            double a = -2.0 * kPi * double(k) / double(n);
            tw_f[k] = cd(std::cos(a), std::sin(a));
        }
    }
    static int smallest_factor(int m) {
        if (m % 4 == 0) return 4;
        if (m % 2 == 0) return 2;
        for (int p = 3; p * p <= m; p += 2)
            if (m % p == 0) return p;
        return m;
    }
    inline cd tw(long idx, int sign) const {
        idx %= n;
        cd w = tw_f[idx];
        return sign < 0 ? w : std::conj(w);
    }
    void rec(int m, int stride, const cd* in, cd* out, int sign) const {
        if (m == 1) { out[0] = in[0]; return; }
        int p = smallest_factor(m);
        int q = m / p;
        for (int k = 0; k < p; ++k) rec(q, stride * p, in + (long)k * stride, out + (long)k * q, sign);
        long step = n / m;  // w_m^x = w_n^(x*step)
        if (p == 2) {
            for (int j = 0; j < q; ++j) {
                cd a = out[j], b = out[q + j] * tw((long)j * step, sign);
                out[j] = a + b;
                out[q + j] = a - b;
            }
        } else if (p == 4) {
            const cd I = (sign < 0) ? cd(0, -1) : cd(0, 1);
            for (int j = 0; j < q; ++j) {
                cd a0 = out[j];
                cd a1 = out[q + j] * tw((long)j * step, sign);
                cd a2 = out[2 * q + j] * tw(2L * j * step, sign);
                cd a3 = out[3 * q + j] * tw(3L * j * step, sign);
                cd s02 = a0 + a2, d02 = a0 - a2, s13 = a1 + a3, d13 = (a1 - a3) * I;
                out[j] = s02 + s13;
                out[q + j] = d02 + d13;
                out[2 * q + j] = s02 - s13;
                out[3 * q + j] = d02 - d13;
            }
        } else {
            std::vector<cd> t(p);
            for (int j = 0; j < q; ++j) {
                for (int k = 0; k < p; ++k) t[k] = out[(long)k * q + j] * tw((long)j * k * step, sign);
                for (int r = 0; r < p; ++r) {
                    cd acc = 0;
                    for (int k = 0; k < p; ++k) acc += t[k] * tw(((long)k * r % p) * (n / p), sign);
                    out[(long)r * q + j] = acc;
                }
            }
        }
    }
    // in-place on a strided vector
    void run(cd* data, long stride, int sign, cd* tmp_in, cd* tmp_out) const {
        for (int i = 0; i < n; ++i) tmp_in[i] = data[i * stride];
        rec(n, 1, tmp_in, tmp_out, sign);
        for (int i = 0; i < n; ++i) data[i * stride] = tmp_out[i];
    }
};

// ---------------------------------------------------------------------------
// small 3x3 helpers
// ---------------------------------------------------------------------------
struct M3 { double m[9]; };

// Euler_angles2matrix (xmippCore; convention pinned by test_binding.py:59-69 and
// data/euler.cpp ZYZ; rows listed in SURVEY App. A.2)
M3 euler_matrix(double rot, double tilt, double psi) {
    double a = rot * kPi / 180.0, b = tilt * kPi / 180.0, g = psi * kPi / 180.0;
    double ca = std::cos(a), cb = std::cos(b), cg = std::cos(g);
    double sa = std::sin(a), sb = std::sin(b), sg = std::sin(g);
    double cc = cb * ca, cs = cb * sa, sc = sb * ca, ss = sb * sa;
    M3 A;
    A.m[0] = cg * cc - sg * sa;  A.m[1] = cg * cs + sg * ca;  A.m[2] = -cg * sb;
    A.m[3] = -sg * cc - cg * sa; A.m[4] = -sg * cs + cg * ca; A.m[5] = sg * sb;
    A.m[6] = sc;                 A.m[7] = ss;                 A.m[8] = cb;
    return A;
}
M3 transpose(const M3& a) {
    M3 t;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) t.m[i * 3 + j] = a.m[j * 3 + i];
    return t;
}
M3 matmul(const M3& a, const M3& b) {
    M3 c;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        double s = 0;
        for (int k = 0; k < 3; ++k) s += a.m[i * 3 + k] * b.m[k * 3 + j];
        c.m[i * 3 + j] = s;
    }
    return c;
}

inline int wrap(int x, int n) { int r = x % n; return r < 0 ? r + n : r; }  // intWRAP(x,0,n-1)
inline double sgn(double x) { return x >= 0 ? 1.0 : -1.0; }                 // SGN(0)=+1
inline double sinc_pi(double x) {                                           // SINC / Sinc: sin(pi x)/(pi x)
    if (std::fabs(x) < 0.0001) return 1.0;
    return std::sin(kPi * x) / (kPi * x);
}
// FFT_IDX2DIGFREQ — KAT test_fftw_main.cpp:80-108
inline double idx2digfreq(int idx, int size) { return (idx <= size / 2) ? double(idx) / size : double(idx - size) / size; }

// ---------------------------------------------------------------------------
// CTF (pure part) — data/ctf.cpp:645-680 (produceSideInfo), :1392-1404,
// data/ctf.h:1002-1029 (precomputeValues), :452-502 (getValuePureAt / NoK)
// ---------------------------------------------------------------------------
struct Ctf {
    double K1, K2, K3, K5, K6, K7, Ksin, Kcos, K;
    double DeltaR, envR0, envR1, envR2, phase_shift, VPP_radius;
    double rad_azimuth, defocus_average, defocus_deviation;
    explicit Ctf(const orf_particle& p) {
        double local_Cs = p.Cs * 1e7, local_Ca = p.Ca * 1e7, local_kV = p.kV * 1e3, local_ispr = p.ispr * 1e6;
        double lambda = 12.2643247 / std::sqrt(local_kV * (1. + 0.978466e-6 * local_kV));
        K1 = kPi * lambda;
        K2 = kPi / 2 * local_Cs * lambda * lambda * lambda;
        K3 = std::pow(0.25 * kPi * local_Ca * lambda * (p.espr / p.kV + 2 * local_ispr), 2) / std::log(2.0);
        K5 = kPi * p.DeltaF * lambda;
        K6 = kPi * kPi * p.alpha * p.alpha;
        K7 = local_Cs * lambda * lambda;
        Ksin = std::sqrt(1 - p.Q0 * p.Q0);
        Kcos = p.Q0;
        K = p.K;
        DeltaR = p.DeltaR; envR0 = p.envR0; envR1 = p.envR1; envR2 = p.envR2;
        phase_shift = p.phase_shift; VPP_radius = p.vpp_radius;
        rad_azimuth = p.defocus_angle * kPi / 180.0;
        defocus_average = -(p.defocusU + p.defocusV) * 0.5;
        defocus_deviation = -(p.defocusU - p.defocusV) * 0.5;
    }
    // getValueArgument() after precomputeValues(X,Y) (ctf.h: the phase of the CTF sine)
    double argument(double X, double Y) const {
        double ang = std::atan2(Y, X);
        double u2 = X * X + Y * Y, u4 = u2 * u2;
        double deltaf;
        if (std::fabs(X) < kEqualAccuracy && std::fabs(Y) < kEqualAccuracy) deltaf = 0;
        else deltaf = defocus_average + defocus_deviation * std::cos(2 * (ang - rad_azimuth));
        double VPP = 0.0;
        if (std::round(VPP_radius * 1000) != 0)
            VPP = -phase_shift * (1 - std::exp(-u2 / (2 * VPP_radius * VPP_radius)));
        return VPP + K1 * deltaf * u2 + K2 * u4;
    }
    // getValuePureNoKAt() after precomputeValues(X,Y): K * getValuePureAt()
    double value(double X, double Y) const {
        double ang = std::atan2(Y, X);
        double u2 = X * X + Y * Y, u = std::sqrt(u2), u4 = u2 * u2;
        double deltaf;
        if (std::fabs(X) < kEqualAccuracy && std::fabs(Y) < kEqualAccuracy) deltaf = 0;
        else deltaf = defocus_average + defocus_deviation * std::cos(2 * (ang - rad_azimuth));
        double VPP = 0.0;
        if (std::round(VPP_radius * 1000) != 0)
            VPP = -phase_shift * (1 - std::exp(-u2 / (2 * VPP_radius * VPP_radius)));
        double argument = VPP + K1 * deltaf * u2 + K2 * u4;
        double Eespr = std::exp(-K3 * u4);
        double EdeltaF = bessj0(K5 * u2);
        double EdeltaR = sinc_pi(u * DeltaR);
        double aux = K7 * u2 * u + deltaf * u;
        double Ealpha = std::exp(-K6 * aux * aux);
        double E = Eespr * EdeltaF * EdeltaR * Ealpha + envR0 + envR1 * u + envR2 * u2;
        if (E < 0) E = 0;
        double pure = -K * (Ksin * std::sin(argument) - Kcos * std::cos(argument)) * E;
        return K * pure;
    }
};

// ---------------------------------------------------------------------------
// The oracle object
// ---------------------------------------------------------------------------
struct Oracle {
    orf_config cfg;
    int N, P, Z, X;   // image size, padded image, padded volume, Z/2+1
    double maxRes2;
    std::vector<double> blobTableSqrt, fourierBlobTable;
    double iDeltaSqrt, iDeltaFourier;
    std::vector<M3> R;   // R_repository: identity + symmetry matrices
    std::vector<cd> V;   // VoutFourier  [z][y][x], x fastest, x in [0,Z/2]
    std::vector<double> W;  // FourierWeights
    Fft fftP, fftZ;
    long n_inserted = 0;
    std::vector<orf_particle> history;   // every particle handed to insert(), for the --iter > 1 re-insertion passes

    explicit Oracle(const orf_config& c)
        : cfg(c), N(c.img_size), P(int(c.img_size * c.pad_proj)), Z(int(c.img_size * c.pad_vol)),
          X(Z / 2 + 1), fftP(P), fftZ(Z) {
        maxRes2 = c.max_resolution * c.max_resolution;               // RF.cpp:188
        V.assign((size_t)Z * Z * X, cd(0, 0));                        // RF.cpp:205-215
        W.assign((size_t)Z * Z * X, 0.0);
        build_tables();
        M3 I = {{1, 0, 0, 0, 1, 0, 0, 0, 1}};
        R.push_back(I);                                               // RF.cpp:272-274
        for (int s = 0; s < c.n_sym; ++s) {                           // RF.cpp:279-285
            M3 m;
            std::memcpy(m.m, c.sym_matrices + 9 * s, sizeof(m.m));
            R.push_back(m);
        }
    }

    // RF.cpp:224-269
    void build_tables() {
        blobTableSqrt.resize(kTable);
        fourierBlobTable.resize(kTable);
        double r = cfg.blob_radius, alpha = cfg.blob_alpha;
        int m = cfg.blob_order;
        double rF = r / (cfg.pad_vol * N);                            // blobFourier.radius  :231
        double rNorm = r / (cfg.pad_proj / cfg.pad_vol);              // blobnormalized.radius :233
        double deltaSqrt = (r * r) / (kTable - 1);                    // :234
        double deltaFourier = (std::sqrt(3.) * N / 2.) / (kTable - 1);  // :235
        double iw0 = 1.0 / kaiser_Fourier_value(0.0, rNorm, alpha, m);  // :238
        double padXdim3 = cfg.pad_vol * N;                            // :240-241
        padXdim3 = padXdim3 * padXdim3 * padXdim3;
        double blobTableSize = r * std::sqrt(1. / (kTable - 1));      // :242
        for (int i = 0; i < kTable; ++i) {
            blobTableSqrt[i] = kaiser_value(blobTableSize * std::sqrt((double)i), r, alpha, m) * iw0;   // :251
            fourierBlobTable[i] = kaiser_Fourier_value(deltaFourier * i, rF, alpha, m) * padXdim3 * iw0;  // :255-256
        }
        iDeltaSqrt = 1 / deltaSqrt;                                   // :268
        iDeltaFourier = 1 / deltaFourier;                             // :269
    }

    // PRELOAD_IMAGE — RF.cpp:348-450.  img: N*N float32, row-major.
    // Returns the half-plane FFT P x (P/2+1) normalised by 1/P^2 and Ainv.
    void preprocess(const float* img, const orf_particle& p, std::vector<cd>& F, M3& Ainv, double& weight) const {
        // readApplyGeo with only_apply_shifts (xmippCore; App. B): content moves by +shift, wrap.
        // Integer shifts are an exact circular shift; fractional parts are handled by the
        // cubic B-spline path below (best knowledge of xmippCore, unpinned in-tree).
        std::vector<double> shifted((size_t)N * N);
        apply_shift(img, p.shift_x, p.shift_y, shifted);
        weight = cfg.use_weights ? p.weight : 1.0;                    // RF.cpp:374-381
        // pad + CenterFFT: logical index l in [-N/2, N/2-1] -> physical (l mod P)   :390-402
        int Xh = P / 2 + 1;
        std::vector<cd> full((size_t)P * P, cd(0, 0));
        int first = -(N / 2);
        std::vector<char> rowNonZero(P, 0);
        for (int i = 0; i < N; ++i) {
            int pi = wrap(first + i, P);
            rowNonZero[pi] = 1;
            for (int j = 0; j < N; ++j) full[(size_t)pi * P + wrap(first + j, P)] = cd(shifted[(size_t)i * N + j], 0.0);
        }
        // forward FFT normalised by 1/size — RF.cpp:405-407, KAT test_fftw_main.cpp:35-51
        std::vector<cd> tin(P), tout(P);
        for (int i = 0; i < P; ++i) {
            if (!rowNonZero[i]) continue;   // FFT of an all-zero row is zero
            fftP.run(&full[(size_t)i * P], 1, -1, tin.data(), tout.data());
        }
        F.assign((size_t)P * Xh, cd(0, 0));
        double inv = 1.0 / (double(P) * double(P));
        for (int j = 0; j < Xh; ++j) {
            for (int i = 0; i < P; ++i) tin[i] = full[(size_t)i * P + j];
            fftP.rec(P, 1, tin.data(), tout.data(), -1);
            for (int i = 0; i < P; ++i) F[(size_t)i * Xh + j] = tout[i] * inv;
        }
        M3 A = euler_matrix(p.rot, p.tilt, p.psi);                    // :411
        Ainv = transpose(A);                                          // :412
    }

    // cubic B-spline helpers for fractional shifts (xmippCore applyGeometry with BSPLINE3 and wrap).  Conventions pinned by
    // the reference's known answer TransformationTest.rotate (test_transformation_main.cpp:76-95, reproduced to 1e-7 in
    // tests/test_oracle_kat.py): produceSplineCoefficients = bilib direct transform with MIRROR-OFF-BOUNDS (half-sample
    // symmetric) extension, interpolatedElementBSpline2D reflects out-of-range neighbours (l < 0 -> -l-1,
    // l >= n -> 2n-l-1), and applyGeometry wraps the source COORDINATE into [-0.5, n-0.5) (realWRAP) before interpolating.
    static void bspline_prefilter_1d(double* c, int n, long stride) {
        const double z1 = std::sqrt(3.0) - 2.0, lambda = (1.0 - z1) * (1.0 - 1.0 / z1);
        if (n == 1) return;
        for (int k = 0; k < n; ++k) c[k * stride] *= lambda;
        // causal initialisation: c+[0] = s[0] + z1 * sum_{m>=0} z1^m s~[m], s~ the half-sample mirrored (2n-periodic) signal
        double sum = 0, zm = 1.0;
        const int horizon = std::max(80, 2 * n);
        for (int m = 0; m < horizon; ++m) {
            int q = m % (2 * n);
            int idx = q < n ? q : 2 * n - 1 - q;
            sum += zm * c[idx * stride];
            zm *= z1;
            if (m >= 80) break;
        }
        c[0] = c[0] + z1 * sum;
        for (int k = 1; k < n; ++k) c[k * stride] += z1 * c[(k - 1) * stride];
        c[(n - 1) * stride] = (z1 / (z1 - 1.0)) * c[(n - 1) * stride];
        for (int k = n - 2; k >= 0; --k) c[k * stride] = z1 * (c[(k + 1) * stride] - c[k * stride]);
    }
    static inline int reflect(int l, int n) { return l < 0 ? -l - 1 : (l >= n ? 2 * n - l - 1 : l); }
    // interpolatedElementBSpline2D at physical (x, y) of an n x n coefficient array
    static double bspline_interp_2d(const double* c, int n, double x, double y) {
        int l1 = (int)std::ceil(x - 2), m1 = (int)std::ceil(y - 2);
        double columns = 0;
        for (int m = m1; m <= m1 + 3; ++m) {
            const double* row = c + (size_t)reflect(m, n) * n;
            double rows = 0;
            for (int l = l1; l <= l1 + 3; ++l) rows += row[reflect(l, n)] * ProjectorBspline03(x - (double)l);
            columns += rows * ProjectorBspline03(y - (double)m);
        }
        return columns;
    }
    static inline double ProjectorBspline03(double x) {      // BSPLINE03
        double a = std::fabs(x);
        if (a < 1.0) return a * a * (a - 2.0) * 0.5 + 2.0 / 3.0;
        if (a < 2.0) { a -= 2.0; return a * a * a / -6.0; }
        return 0.0;
    }
    void apply_shift(const float* img, double sx, double sy, std::vector<double>& out) const {
        double rx = std::round(sx), ry = std::round(sy);
        bool integer = std::fabs(sx - rx) < 1e-9 && std::fabs(sy - ry) < 1e-9;
        if (integer) {
            int isx = (int)rx, isy = (int)ry;
            for (int i = 0; i < N; ++i)
                for (int j = 0; j < N; ++j)
                    out[(size_t)i * N + j] = img[(size_t)wrap(i - isy, N) * N + wrap(j - isx, N)];
            return;
        }
        // out(x) = in(x - shift): coefficients, then per pixel the wrapped source coordinate and the reflected neighbours
        std::vector<double> c((size_t)N * N);
        for (size_t k = 0; k < c.size(); ++k) c[k] = img[k];
        for (int i = 0; i < N; ++i) bspline_prefilter_1d(&c[(size_t)i * N], N, 1);
        for (int j = 0; j < N; ++j) bspline_prefilter_1d(&c[j], N, N);
        auto wrapc = [&](double v) {       // realWRAP(v, -0.5, N - 0.5) applied when v leaves [0, N-1] (applyGeometry)
            if (v < -kEqualAccuracy || v > (N - 1) + kEqualAccuracy) v -= std::floor((v + 0.5) / N) * N;
            return v;
        };
        for (int i = 0; i < N; ++i) {
            const double y = wrapc(i - sy);
            for (int j = 0; j < N; ++j) out[(size_t)i * N + j] = bspline_interp_2d(c.data(), N, wrapc(j - sx), y);
        }
    }

    bool has_ctf() const { return cfg.use_ctf != 0; }

    // per-pixel CTF handling — RF.cpp:600-625
    inline void ctf_weights(const Ctf& ctf, int i, int j, double fx, double fy, double& wCTF, double& wModulator) const {
        double iTs = 1.0 / cfg.sampling;                                // :495
        wModulator = 1.0;
        wCTF = ctf.value(fx * iTs, fy * iTs);                           // :602-606
        if (std::isnan(wCTF)) {                                         // :609-615
            if (i == 0 && j == 0) wModulator = wCTF = 1.0;
            else wModulator = wCTF = 0.0;
        }
        if (std::fabs(wCTF) < cfg.min_ctf) {                            // :616-622
            wModulator = std::fabs(wCTF);
            wCTF = sgn(wCTF);
        } else
            wCTF = 1.0 / wCTF;
        if (cfg.phase_flipped) wCTF = std::fabs(wCTF);                  // :623-624
    }

    // One row of the insertion scatter — RF.cpp:586-792 (body of the i loop)
    void scatter_row(const std::vector<cd>& F, int i, const M3& A_SL, double weight, const Ctf* ctf) {
        const int Xh = P / 2 + 1;
        const double r = cfg.blob_radius, r2 = r * r;
        const int xsize_1 = X - 1;   // :580
        double fy = idx2digfreq(i, P);
        for (int j = 0; j < Xh; ++j) {
            double fx = idx2digfreq(j, P);                              // :594
            if (fx * fx + fy * fy > maxRes2) continue;                  // :597
            double wCTF = 1, wModulator = 1.0;
            if (ctf) ctf_weights(*ctf, i, j, fx, fy, wCTF, wModulator);
            // freq3 = A_SL * (fx, fy, 0)   :628
            double qx = A_SL.m[0] * fx + A_SL.m[1] * fy;
            double qy = A_SL.m[3] * fx + A_SL.m[4] * fy;
            double qz = A_SL.m[6] * fx + A_SL.m[7] * fy;
            double px = qx * Z, py = qy * Z, pz = qz * Z;               // :631-633
            int x0 = (int)std::ceil(px - r), x1 = (int)std::floor(px + r);   // :636-641
            int y0 = (int)std::ceil(py - r), y1 = (int)std::floor(py + r);
            int z0 = (int)std::ceil(pz - r), z1 = (int)std::floor(pz + r);
            const cd in = F[(size_t)i * Xh + j];
            for (int iz = z0; iz <= z1; ++iz) {                         // :699
                double dz = iz - pz, z2 = dz * dz;
                int wz = wrap(iz, Z), wzn = wrap(-wz, Z);               // :662-666
                for (int iy = y0; iy <= y1; ++iy) {
                    double dy = iy - py, y2z2 = dy * dy + z2;
                    if (y2z2 > r2) continue;                            // :708
                    int wy = wrap(iy, Z), wyn = wrap(-wy, Z);
                    for (int ix = x0; ix <= x1; ++ix) {
                        double dx = ix - px, d2 = dx * dx + y2z2;
                        if (d2 > r2) continue;                          // :723
                        int aux = (int)(d2 * iDeltaSqrt + 0.5);         // :725
                        double w = blobTableSqrt[aux] * weight * wModulator;   // :726
                        int wx = wrap(ix, Z);
                        bool conj = wx > xsize_1;                       // :748
                        size_t idx;
                        if (conj) idx = ((size_t)wzn * Z + wyn) * X + wrap(-wx, Z);   // :750-754
                        else idx = ((size_t)wz * Z + wy) * X + wx;                    // :758-761
                        double wEff = w * wCTF;                         // :778
                        double re = wEff * in.real(), im = wEff * in.imag();
                        V[idx] += cd(re, conj ? -im : im);              // :781-787
                        W[idx] += w;                                    // :782
                    }
                }
            }
        }
    }

    // Re-insertion pass of the weight refinement (reprocessFlag, RF.cpp:770-775): same geometry as the
    // insertion, no image data, wModulator = 1; every pair adds w * slot[target] to the new weights.
    void reprocess_all(const std::vector<double>& slot, std::vector<double>& Wn) const {
        const int Xh = P / 2 + 1;
        const double r = cfg.blob_radius, r2 = r * r;
        const int xsize_1 = X - 1;
        size_t conserveRows = (size_t)std::ceil((double)P * cfg.max_resolution * 2.0);
        conserveRows = (size_t)std::ceil((double)conserveRows / 2.0);
        for (const orf_particle& p : history) {
            double weight = cfg.use_weights ? p.weight : 1.0;
            if (weight == 0.0) continue;
            M3 Ainv = transpose(euler_matrix(p.rot, p.tilt, p.psi));
            for (size_t isym = 0; isym < R.size(); ++isym) {
                M3 A_SL = matmul(R[isym], Ainv);
                for (int i = 0; i < P; ++i) {
                    if ((size_t)i >= conserveRows && (size_t)i < (size_t)P - conserveRows) continue;
                    double fy = idx2digfreq(i, P);
                    for (int j = 0; j < Xh; ++j) {
                        double fx = idx2digfreq(j, P);
                        if (fx * fx + fy * fy > maxRes2) continue;
                        double px = (A_SL.m[0] * fx + A_SL.m[1] * fy) * Z, py = (A_SL.m[3] * fx + A_SL.m[4] * fy) * Z,
                               pz = (A_SL.m[6] * fx + A_SL.m[7] * fy) * Z;
                        int x0 = (int)std::ceil(px - r), x1 = (int)std::floor(px + r);
                        int y0 = (int)std::ceil(py - r), y1 = (int)std::floor(py + r);
                        int z0 = (int)std::ceil(pz - r), z1 = (int)std::floor(pz + r);
                        for (int iz = z0; iz <= z1; ++iz) {
                            double dz = iz - pz, z2 = dz * dz;
                            int wz = wrap(iz, Z), wzn = wrap(-wz, Z);
                            for (int iy = y0; iy <= y1; ++iy) {
                                double dy = iy - py, y2z2 = dy * dy + z2;
                                if (y2z2 > r2) continue;
                                int wy = wrap(iy, Z), wyn = wrap(-wy, Z);
                                for (int ix = x0; ix <= x1; ++ix) {
                                    double dx = ix - px, d2 = dx * dx + y2z2;
                                    if (d2 > r2) continue;
                                    double w = blobTableSqrt[(int)(d2 * iDeltaSqrt + 0.5)] * weight;
                                    int wx = wrap(ix, Z);
                                    size_t idx = (wx > xsize_1) ? ((size_t)wzn * Z + wyn) * X + wrap(-wx, Z) : ((size_t)wz * Z + wy) * X + wx;
                                    Wn[idx] += w * slot[idx];                      // RF.cpp:773-774
                                }
                            }
                        }
                    }
                }
            }
        }
    }

    // processImages — RF.cpp:835-1013 with the thread scheme of :137-151, :344, :829:
    // T persistent workers + the coordinating caller meet at a barrier before and after
    // every operation.  PRELOAD: each worker loads + FFTs one image.  PROCESS: every
    // (image, symmetry) is inserted by all workers, rows being handed out under a mutex
    // with +-minSeparation exclusion (:497-562, :795-819).
    // N-party barrier in the style of xmippCore's barrier_t (mutex + condition variable)
    struct SpinBarrier {
        explicit SpinBarrier(int n) : n_(n) {}
        void wait() {
            std::unique_lock<std::mutex> lk(m_);
            int gen = gen_;
            if (++count_ == n_) {
                count_ = 0;
                ++gen_;
                cv_.notify_all();
            } else {
                cv_.wait(lk, [&] { return gen_ != gen; });
            }
        }
        int n_, count_ = 0, gen_ = 0;
        std::mutex m_;
        std::condition_variable cv_;
    };
    struct Loaded { std::vector<cd> F; M3 Ainv; double weight = 0; orf_particle p; bool read = false; };

    void insert(const float* imgs, const orf_particle* meta, int n, int T) {
        history.insert(history.end(), meta, meta + n);
        if (T < 1) T = 1;
        const int minSep = std::max((int)std::ceil(cfg.blob_radius), 1) + 1;   // :296-303 with thrWidth=1
        // conserveRows — :927-928
        size_t conserveRows = (size_t)std::ceil((double)P * cfg.max_resolution * 2.0);
        conserveRows = (size_t)std::ceil((double)conserveRows / 2.0);

        enum Op { PRELOAD, PROCESS, EXIT };
        Op op = PRELOAD;
        SpinBarrier barrier(T + 1);
        std::vector<Loaded> L(T);
        std::vector<int> imageIndex(T, -1);
        // state of the (image, symmetry) being inserted
        const std::vector<cd>* curF = nullptr;
        M3 curA;
        double curWeight = 1;
        const Ctf* curCtf = nullptr;
        std::vector<int> status(P);   // 0 free, -1 taken, -2 discarded, >0 blocked (:949-961)
        int remaining = 0;
        std::mutex mtx;

        auto process_rows = [&] {
            std::vector<int> mine;
            for (;;) {
                int lo = -1, hi = -1;
                mine.clear();
                {
                    std::lock_guard<std::mutex> g(mtx);
                    if (remaining == 0) return;
                    for (int w = 0; w < P; ++w) {
                        if (status[w] != 0) continue;
                        lo = w;
                        hi = std::min(w + minSep - 1, P - 1);
                        for (int k = lo - minSep; k < lo; ++k) if (k >= 0 && status[k] > -1) status[k]++;
                        for (int k = lo; k <= hi; ++k) if (status[k] == 0) { status[k] = -1; --remaining; mine.push_back(k); }
                        for (int k = hi + 1; k <= hi + minSep; ++k) if (k < P && status[k] > -1) status[k]++;
                        break;
                    }
                }
                if (lo < 0) continue;   // every free row is currently blocked: poll again (as the reference does)
                for (int i : mine) scatter_row(*curF, i, curA, curWeight, curCtf);
                {
                    std::lock_guard<std::mutex> g(mtx);
                    for (int k = lo - minSep; k < lo; ++k) if (k >= 0 && status[k] > 0) status[k]--;
                    for (int k = hi + 1; k <= hi + minSep; ++k) if (k < P && status[k] > 0) status[k]--;
                }
            }
        };
        auto worker = [&](int t) {
            for (;;) {
                barrier.wait();
                if (op == EXIT) return;
                if (op == PRELOAD) {
                    L[t].read = false;
                    if (imageIndex[t] >= 0) {
                        L[t].p = meta[imageIndex[t]];
                        preprocess(imgs + (size_t)imageIndex[t] * N * N, L[t].p, L[t].F, L[t].Ainv, L[t].weight);
                        L[t].read = true;
                    }
                } else {
                    process_rows();
                }
                barrier.wait();
            }
        };
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t) th.emplace_back(worker, t);

        int next = 0;
        while (next < n) {
            op = PRELOAD;
            for (int t = 0; t < T; ++t) imageIndex[t] = (next < n) ? next++ : -1;
            barrier.wait();
            barrier.wait();
            op = PROCESS;
            for (int t = 0; t < T; ++t) {
                if (!L[t].read) continue;
                if (L[t].weight == 0.0) continue;                      // :483-484
                Ctf ctfObj(L[t].p);
                curCtf = has_ctf() ? &ctfObj : nullptr;
                curF = &L[t].F;
                curWeight = L[t].weight;
                for (size_t isym = 0; isym < R.size(); ++isym) {       // :931
                    curA = matmul(R[isym], L[t].Ainv);                 // :936
                    remaining = 0;
                    for (int i = 0; i < P; ++i) {
                        if ((size_t)i >= conserveRows && (size_t)i < (size_t)P - conserveRows) status[i] = -2;
                        else { status[i] = 0; ++remaining; }
                    }
                    barrier.wait();
                    barrier.wait();
                }
                ++n_inserted;
            }
        }
        op = EXIT;
        barrier.wait();
        for (auto& x : th) x.join();
    }

    // ---- race-free parallel insertion (test infrastructure for large cases; NOT the reference's thread scheme).
    // The reference's row scheduler loses updates with >= 3 threads, so big parity runs use this
    // mode instead: the stored volume is cut into slabs of z indices, every slab is owned by exactly one task, and a
    // task walks all (image, symmetry, row, pixel) in the single-thread order but only applies the updates that land in
    // its slab.  Every voxel therefore receives the same additions in the same order as with threads = 1: the
    // accumulators are BIT-IDENTICAL to the single-thread result (tests/test_oracle_kat.py).  Per row the pixels that can
    // reach the slab form at most four column intervals (pz is linear in j), so a task only visits those.
    void scatter_row_slab(const std::vector<cd>& F, int i, const M3& A_SL, double weight, const double* ctfW, int zLo, int zHi,
                          const double (*Q)[2], int nQ) {
        const int Xh = P / 2 + 1;
        const double r = cfg.blob_radius, r2 = r * r;
        const int xsize_1 = X - 1;
        const double fy = idx2digfreq(i, P);
        // columns j whose pz (or -pz: conjugate targets sit at -z) comes within r of one of the centred z intervals Q
        const double c0 = (A_SL.m[7] * fy) * Z, g = A_SL.m[6] * Z / (double)P;
        int iv[8][2];
        int nIv = 0;
        for (int q = 0; q < nQ; ++q)
            for (int sgnz = 0; sgnz < 2; ++sgnz) {
                // sgnz = 0: pz in [Q0 - r, Q1 + r];  sgnz = 1: -pz in that range
                double lo = Q[q][0] - r - 1e-6, hi = Q[q][1] + r + 1e-6;
                if (sgnz) { double t = -hi; hi = -lo; lo = t; }
                int jl, jh;
                if (std::fabs(g) < 1e-12) {
                    if (c0 < lo - 1e-6 || c0 > hi + 1e-6) continue;
                    jl = 0; jh = Xh - 1;
                } else {
                    double a = (lo - c0) / g, b = (hi - c0) / g;
                    if (a > b) std::swap(a, b);
                    if (b < -1 || a > Xh) continue;
                    jl = std::max(0, (int)std::floor(a) - 1);
                    jh = std::min(Xh - 1, (int)std::ceil(b) + 1);
                    if (jl > jh) continue;
                }
                iv[nIv][0] = jl; iv[nIv][1] = jh; ++nIv;
            }
        if (!nIv) return;
        // sort + merge so that every pixel is visited once
        for (int a = 1; a < nIv; ++a)
            for (int b = a; b > 0 && iv[b][0] < iv[b - 1][0]; --b) { std::swap(iv[b][0], iv[b - 1][0]); std::swap(iv[b][1], iv[b - 1][1]); }
        int m = 0;
        for (int a = 1; a < nIv; ++a) {
            if (iv[a][0] <= iv[m][1] + 1) iv[m][1] = std::max(iv[m][1], iv[a][1]);
            else { ++m; iv[m][0] = iv[a][0]; iv[m][1] = iv[a][1]; }
        }
        nIv = m + 1;
        for (int q = 0; q < nIv; ++q)
        for (int j = iv[q][0]; j <= iv[q][1]; ++j) {
            double fx = idx2digfreq(j, P);                              // :594
            if (fx * fx + fy * fy > maxRes2) continue;                  // :597
            double qx = A_SL.m[0] * fx + A_SL.m[1] * fy;
            double qy = A_SL.m[3] * fx + A_SL.m[4] * fy;
            double qz = A_SL.m[6] * fx + A_SL.m[7] * fy;
            double px = qx * Z, py = qy * Z, pz = qz * Z;               // :631-633
            int z0 = (int)std::ceil(pz - r), z1 = (int)std::floor(pz + r);
            bool any = false;
            for (int iz = z0; iz <= z1 && !any; ++iz) {
                int wz = wrap(iz, Z), wzn = wrap(-wz, Z);
                any = (wz >= zLo && wz < zHi) || (wzn >= zLo && wzn < zHi);
            }
            if (!any) continue;
            double wCTF = 1, wModulator = 1.0;
            if (ctfW) { wCTF = ctfW[2 * ((size_t)i * Xh + j)]; wModulator = ctfW[2 * ((size_t)i * Xh + j) + 1]; }   // ctf_weights(), hoisted
            int x0 = (int)std::ceil(px - r), x1 = (int)std::floor(px + r);   // :636-641
            int y0 = (int)std::ceil(py - r), y1 = (int)std::floor(py + r);
            const cd in = F[(size_t)i * Xh + j];
            for (int iz = z0; iz <= z1; ++iz) {                         // :699
                double dz = iz - pz, z2 = dz * dz;
                int wz = wrap(iz, Z), wzn = wrap(-wz, Z);               // :662-666
                const bool okN = wz >= zLo && wz < zHi, okC = wzn >= zLo && wzn < zHi;
                if (!okN && !okC) continue;
                for (int iy = y0; iy <= y1; ++iy) {
                    double dy = iy - py, y2z2 = dy * dy + z2;
                    if (y2z2 > r2) continue;                            // :708
                    int wy = wrap(iy, Z), wyn = wrap(-wy, Z);
                    for (int ix = x0; ix <= x1; ++ix) {
                        double dx = ix - px, d2 = dx * dx + y2z2;
                        if (d2 > r2) continue;                          // :723
                        int wx = wrap(ix, Z);
                        bool conj = wx > xsize_1;                       // :748
                        if (conj ? !okC : !okN) continue;               // not this slab's voxel
                        int aux = (int)(d2 * iDeltaSqrt + 0.5);         // :725
                        double w = blobTableSqrt[aux] * weight * wModulator;   // :726
                        size_t idx;
                        if (conj) idx = ((size_t)wzn * Z + wyn) * X + wrap(-wx, Z);   // :750-754
                        else idx = ((size_t)wz * Z + wy) * X + wx;                    // :758-761
                        double wEff = w * wCTF;                         // :778
                        double re = wEff * in.real(), im = wEff * in.imag();
                        V[idx] += cd(re, conj ? -im : im);              // :781-787
                        W[idx] += w;                                    // :782
                    }
                }
            }
        }
    }

    void insert_slabs(const float* imgs, const orf_particle* meta, int n, int T) {
        history.insert(history.end(), meta, meta + n);
        if (T < 1) T = 1;
        size_t conserveRows = (size_t)std::ceil((double)P * cfg.max_resolution * 2.0);   // :927-928
        conserveRows = (size_t)std::ceil((double)conserveRows / 2.0);
        const size_t bytesPerImage = sizeof(cd) * (size_t)P * (P / 2 + 1);
        const int B = (int)std::max<size_t>(T, std::min<size_t>(256, ((size_t)1 << 30) / (2 * bytesPerImage)));   // <= 1 GiB of transforms
        // slabs of stored z indices, several per thread, central (heaviest) ones first
        const int nSlab = std::min(Z, 3 * T);
        std::vector<std::pair<int, int>> slabs;
        for (int s = 0; s < nSlab; ++s) {
            int lo = (int)((long)Z * s / nSlab), hi = (int)((long)Z * (s + 1) / nSlab);
            if (hi > lo) slabs.push_back({lo, hi});
        }
        auto centreDist = [&](const std::pair<int, int>& s) { int c = (s.first + s.second) / 2; return std::min(c, Z - c); };
        std::stable_sort(slabs.begin(), slabs.end(), [&](const auto& a, const auto& b) { return centreDist(a) < centreDist(b); });
        const int R1 = (int)std::ceil(cfg.blob_radius) + 1;

        std::vector<Loaded> L(B);
        std::vector<std::vector<double>> ctfW(B);    // (wCTF, wModulator) per half-plane pixel: the reference evaluates them per
                                                     // symmetry (RF.cpp:600-625); they only depend on the pixel
        const int Xh = P / 2 + 1;
        for (int b0 = 0; b0 < n; b0 += B) {
            const int nb = std::min(B, n - b0);
            {   // phase 1: preprocess the batch, one image per task
                std::atomic<int> next(0);
                auto work = [&] {
                    for (;;) {
                        int k = next.fetch_add(1);
                        if (k >= nb) return;
                        L[k].p = meta[b0 + k];
                        preprocess(imgs + (size_t)(b0 + k) * N * N, L[k].p, L[k].F, L[k].Ainv, L[k].weight);
                        if (has_ctf()) {
                            Ctf ctf(L[k].p);
                            ctfW[k].assign((size_t)2 * P * Xh, 1.0);
                            for (int i = 0; i < P; ++i) {
                                const double fy = idx2digfreq(i, P);
                                for (int j = 0; j < Xh; ++j) {
                                    const double fx = idx2digfreq(j, P);
                                    if (fx * fx + fy * fy > maxRes2) continue;
                                    ctf_weights(ctf, i, j, fx, fy, ctfW[k][2 * ((size_t)i * Xh + j)], ctfW[k][2 * ((size_t)i * Xh + j) + 1]);
                                }
                            }
                        }
                    }
                };
                std::vector<std::thread> th;
                for (int t = 0; t < T; ++t) th.emplace_back(work);
                for (auto& x : th) x.join();
            }
            {   // phase 2: every slab task inserts the whole batch in the single-thread order
                std::atomic<int> next(0);
                auto work = [&] {
                    for (;;) {
                        int s = next.fetch_add(1);
                        if (s >= (int)slabs.size()) return;
                        const int zLo = slabs[s].first, zHi = slabs[s].second;
                        // centred z intervals whose wrapped index falls into [zLo, zHi): q = t (t <= Z/2 + R1) and q = t - Z
                        double Q[2][2];
                        int nQ = 0;
                        if (zLo <= Z / 2 + R1) { Q[nQ][0] = zLo; Q[nQ][1] = std::min(zHi - 1, Z / 2 + R1); ++nQ; }
                        if (zHi - 1 >= Z / 2 - R1) { Q[nQ][0] = std::max(zLo, Z / 2 - R1) - Z; Q[nQ][1] = zHi - 1 - Z; ++nQ; }
                        for (int k = 0; k < nb; ++k) {
                            if (L[k].weight == 0.0) continue;                      // :483-484
                            const double* cw = has_ctf() ? ctfW[k].data() : nullptr;
                            for (size_t isym = 0; isym < R.size(); ++isym) {       // :931
                                M3 A_SL = matmul(R[isym], L[k].Ainv);              // :936
                                for (int i = 0; i < P; ++i) {
                                    if ((size_t)i >= conserveRows && (size_t)i < (size_t)P - conserveRows) continue;
                                    scatter_row_slab(L[k].F, i, A_SL, L[k].weight, cw, zLo, zHi, Q, nQ);
                                }
                            }
                        }
                    }
                };
                std::vector<std::thread> th;
                for (int t = 0; t < T; ++t) th.emplace_back(work);
                for (auto& x : th) x.join();
            }
            for (int k = 0; k < nb; ++k) if (L[k].weight != 0.0) ++n_inserted;
        }
    }

    // forceWeightSymmetry — RF.cpp:1188-1221
    void force_weight_symmetry() {
        int yHalf = Z / 2; if (Z % 2 == 0) yHalf--;
        int zHalf = Z / 2; if (Z % 2 == 0) zHalf--;
        for (int k = 0; k < Z; ++k) {
            int ks = wrap(-k, Z);
            for (int i = 1; i <= yHalf; ++i) {
                int is = wrap(-i, Z);
                size_t a = ((size_t)k * Z + i) * X, b = ((size_t)ks * Z + is) * X;
                double mean = 0.5 * (W[a] + W[b]);
                W[a] = W[b] = mean;
            }
        }
        for (int k = 1; k <= zHalf; ++k) {
            int ks = wrap(-k, Z);
            size_t a = ((size_t)k * Z) * X, b = ((size_t)ks * Z) * X;
            double mean = 0.5 * (W[a] + W[b]);
            W[a] = W[b] = mean;
        }
    }
    // FourierTransformer::enforceHermitianSymmetry (xmippCore; complex twin of the above)
    void enforce_hermitian(std::vector<cd>& v) const {
        int yHalf = Z / 2; if (Z % 2 == 0) yHalf--;
        int zHalf = Z / 2; if (Z % 2 == 0) zHalf--;
        for (int k = 0; k < Z; ++k) {
            int ks = wrap(-k, Z);
            for (int i = 1; i <= yHalf; ++i) {
                int is = wrap(-i, Z);
                size_t a = ((size_t)k * Z + i) * X, b = ((size_t)ks * Z + is) * X;
                cd mean = 0.5 * (v[a] + std::conj(v[b]));
                v[a] = mean;
                v[b] = std::conj(mean);
            }
        }
        for (int k = 1; k <= zHalf; ++k) {
            int ks = wrap(-k, Z);
            size_t a = ((size_t)k * Z) * X, b = ((size_t)ks * Z) * X;
            cd mean = 0.5 * (v[a] + std::conj(v[b]));
            v[a] = mean;
            v[b] = std::conj(mean);
        }
    }

    // correctWeight (--iter 0/1) + finishComputations — RF.cpp:1056-1101, 453-479, 1103-1180.
    // Works on copies so that the accumulators stay inspectable.  out: N^3 doubles.
    void finalize(double* out) {
        std::vector<cd> v = V;
        std::vector<double> w = W;
        std::swap(w, W);
        force_weight_symmetry();                                       // :1059
        std::swap(w, W);
        const size_t n = v.size();
        if (cfg.n_iter_weight == 0) {
            for (size_t k = 0; k < n; ++k) w[k] = 1;                    // :1060-1065
        } else {
            // :1069-1099 with NiterWeight==1: slot = 1/W where |W|>1e-3, else it keeps Re(V)
            for (size_t k = 0; k < n; ++k) w[k] = (std::fabs(w[k]) > 1e-3) ? 1.0 / w[k] : v[k].real();
            for (int it = 1; it < cfg.n_iter_weight; ++it) {        // :1080-1092
                std::vector<double> Wn(n, 0.0);
                reprocess_all(w, Wn);
                std::swap(Wn, W);
                force_weight_symmetry();
                std::swap(Wn, W);
                for (size_t k = 0; k < n; ++k)
                    if (std::fabs(Wn[k]) > 1e-3) w[k] /= Wn[k];
            }
        }
        enforce_hermitian(v);                                          // :1122
        double corr2D_3D = std::pow(cfg.pad_proj, 2.) / (N * std::pow(cfg.pad_vol, 3.));   // :457-458
        for (size_t k = 0; k < n; ++k) {                                // :463-477
            if (cfg.n_iter_weight == 0) v[k] *= corr2D_3D;
            else if (1.0 / w[k] > kAccuracy) v[k] *= corr2D_3D * w[k];
            else v[k] = 0;
        }
        finish_fourier(v, out);
    }

    // tail of finishComputations (RF.cpp:1145-1179): inverse transform, CenterFFT, crop, gridding correction.
    // Also the tail of the --fast programs (reconstruct_fourier_gpu.cpp:894-931), which is the same code.
    void finish_fourier(std::vector<cd>& v, double* out) const {
        // inverse c2r, unnormalised — :1145 (complex along z, y; c2r along x using Re of x=0, x=Z/2)
        std::vector<double> vol((size_t)Z * Z * Z);
        c2r_3d(v, vol);
        // CenterFFT(false) + window to N^3 — :1146-1152: logical (k,i,j) <-> unshifted index mod Z
        int first = -(N / 2);
        double pad_relation = cfg.pad_proj / cfg.pad_vol;
        pad_relation = pad_relation * pad_relation * pad_relation;     // :1153-1154
        double ipad = 1.0 / pad_relation;
        double meanFactor2 = 0;
        for (int kk = 0; kk < N; ++kk) {
            int k = first + kk;
            for (int ii = 0; ii < N; ++ii) {
                int i = first + ii;
                for (int jj = 0; jj < N; ++jj) {
                    int j = first + jj;
                    double val = vol[((size_t)wrap(k, Z) * Z + wrap(i, Z)) * Z + wrap(j, Z)];
                    double radius = std::sqrt((double)(k * k + i * i + j * j));     // :1161
                    double aux = radius * iDeltaFourier;
                    double factor = fourierBlobTable[(int)std::round(aux)];        // :1163 (ROUND)
                    double factor2 = std::pow(sinc_pi(radius / (2 * N)), 2);       // :1164
                    if (cfg.n_iter_weight != 0) {
                        val /= (ipad * factor2 * factor);                          // :1167
                        meanFactor2 += factor2;
                    } else
                        val /= (ipad * factor);                                    // :1171
                    out[((size_t)kk * N + ii) * N + jj] = val;
                }
            }
        }
        if (cfg.n_iter_weight != 0) {                                              // :1173-1178
            meanFactor2 /= (double)N * N * N;
            for (size_t k = 0; k < (size_t)N * N * N; ++k) out[k] *= meanFactor2;
        }
    }

    void c2r_3d(std::vector<cd>& v, std::vector<double>& vol) const {
        int T = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        // along y (stride X) for each z, x ; along z (stride Z*X) for each y, x
        auto pass = [&](int axis) {
            std::vector<std::thread> th;
            std::atomic<int> next(0);
            for (int t = 0; t < T; ++t)
                th.emplace_back([&] {
                    std::vector<cd> a(Z), b(Z);
                    for (;;) {
                        int o = next.fetch_add(1);
                        if (o >= Z) break;
                        for (int x = 0; x < X; ++x) {
                            cd* base = axis == 0 ? &v[((size_t)o * Z) * X + x] : &v[(size_t)o * X + x];
                            long stride = axis == 0 ? X : (long)Z * X;
                            fftZ.run(base, stride, +1, a.data(), b.data());
                        }
                    }
                });
            for (auto& x : th) x.join();
        };
        pass(0);
        pass(1);
        // c2r along x: hermitian completion, Re of DC and Nyquist only (FFTW hc2r semantics)
        std::vector<std::thread> th;
        std::atomic<int> next(0);
        for (int t = 0; t < T; ++t)
            th.emplace_back([&] {
                std::vector<cd> a(Z), b(Z);
                for (;;) {
                    int z = next.fetch_add(1);
                    if (z >= Z) break;
                    for (int y = 0; y < Z; ++y) {
                        const cd* row = &v[((size_t)z * Z + y) * X];
                        a[0] = cd(row[0].real(), 0);
                        for (int x = 1; x < X; ++x) a[x] = row[x];
                        if (Z % 2 == 0) a[Z / 2] = cd(row[Z / 2].real(), 0);
                        for (int x = X; x < Z; ++x) a[x] = std::conj(row[Z - x]);
                        fftZ.rec(Z, 1, a.data(), b.data(), +1);
                        double* o = &vol[((size_t)z * Z + y) * Z];
                        for (int x = 0; x < Z; ++x) o[x] = b[x].real();
                    }
                }
            });
        for (auto& x : th) x.join();
    }
};


// =============================================================================
// FourierProjector (libraries/data/fourier_projection.cpp = "FP"): central-slice projector used by
// xmipp_phantom_project (reconstruction/project.cpp:992) to make the particle sets this path consumes.
//   produceSideInfo FP:265-308, produceSideInfoProjection FP:310-333, project FP:91-262
// xmippCore pieces as used there (source absent, semantics from their call sites and headers): window() about the
// Xmipp origin, completeFourierTransform (forward, 1/size), ShiftFFT (multiply by exp(-2 pi i k.shift/size)),
// CenterFFT(forward) (cyclic shift by size/2), produceSplineCoefficients (bilib ChangeBasisVolume, cubic, mirror-off-
// bounds: the same half-sample mirror FP:189-215 applies when it reads the coefficients), interpolatedElement3D
// (trilinear, 0 outside), inverseFourierTransform (c2r, unnormalised).
// =============================================================================
struct ProjectorOracle {
    int N, P, degree;            // volume size, padded size, 0 NEAREST / 1 LINEAR / 3 BSPLINE3
    double maxFrequency;
    int start, dim;              // logical index of coefficient 0 and edge length of the (windowed) coefficient cube
    std::vector<cd> C;           // VfourierRealCoefs + i VfourierImagCoefs, [z][y][x]
    std::vector<double> phA, phB;

    static void prefilter_line(cd* c, int n, long stride) {
        // cubic B-spline direct transform, mirror-off-bounds (half-sample symmetric) extension
        if (n == 1) return;
        const double z = std::sqrt(3.0) - 2.0, lambda = (1.0 - z) * (1.0 - 1.0 / z);
        for (int k = 0; k < n; ++k) c[k * stride] *= lambda;
        // causal initialisation: c+[0] = s[0] + z * sum_{m>=0} z^m s~[m], s~ the half-sample mirrored, 2n-periodic signal
        cd sum = 0;
        double zm = 1.0;
        for (int m = 0; m < 64 * 1 + 2 * n && std::fabs(zm) > 1e-300; ++m) {
            int q = m % (2 * n);
            int idx = q < n ? q : 2 * n - 1 - q;
            sum += zm * c[idx * stride];
            zm *= z;
            if (m >= 80 && m >= 2 * n) break;
        }
        c[0] = c[0] + z * sum;
        for (int k = 1; k < n; ++k) c[k * stride] += z * c[(k - 1) * stride];
        c[(n - 1) * stride] = (z / (z - 1.0)) * c[(n - 1) * stride];
        for (int k = n - 2; k >= 0; --k) c[k * stride] = z * (c[(k + 1) * stride] - c[k * stride]);
    }

    ProjectorOracle(const float* vol, int N_, double padding, double maxFreq, int degree_)
        : N(N_), P((int)(padding * N_)), degree(degree_), maxFrequency(maxFreq) {
        const int hP = P / 2, firstN = -(N / 2);
        std::vector<cd> V((size_t)P * P * P, cd(0, 0));
        for (int k = 0; k < N; ++k)                                             // window FP:270-272
            for (int i = 0; i < N; ++i)
                for (int j = 0; j < N; ++j)
                    V[((size_t)(k + firstN + hP) * P + (i + firstN + hP)) * P + (j + firstN + hP)] = vol[((size_t)k * N + i) * N + j];
        Fft f(P);
        std::vector<cd> a(P), b(P);
        for (int z = 0; z < P; ++z)
            for (int y = 0; y < P; ++y) f.run(&V[((size_t)z * P + y) * P], 1, -1, a.data(), b.data());
        for (int z = 0; z < P; ++z)
            for (int x = 0; x < P; ++x) f.run(&V[(size_t)z * P * P + x], P, -1, a.data(), b.data());
        for (int y = 0; y < P; ++y)
            for (int x = 0; x < P; ++x) f.run(&V[(size_t)y * P + x], (long)P * P, -1, a.data(), b.data());
        // 1/size of the forward transform, ShiftFFT by FIRST_XMIPP_INDEX(P) = -P/2 (FP:278), K = P^3/N^2 (FP:283-285)
        const double xx = -2.0 * kPi * (double)(-hP) / (double)P;
        const double scale = 1.0 / ((double)N * (double)N);                    // (1/P^3) * K
        std::vector<cd> ph(P);
        for (int k = 0; k < P; ++k) ph[k] = cd(std::cos(k * xx), std::sin(k * xx));
        std::vector<cd> Cfull((size_t)P * P * P);
        for (int z = 0; z < P; ++z)
            for (int y = 0; y < P; ++y)
                for (int x = 0; x < P; ++x) {
                    cd v = V[((size_t)z * P + y) * P + x] * ph[z] * ph[y] * ph[x] * scale;
                    // CenterFFT(forward): element k moves to (k + P/2) mod P (FP:279)
                    Cfull[((size_t)((z + hP) % P) * P + (y + hP) % P) * P + (x + hP) % P] = v;
                }
        V.clear();
        V.shrink_to_fit();
        start = -hP;
        dim = P;
        if (degree == 3) {                                                      // FP:288-308
            for (int z = 0; z < P; ++z)
                for (int y = 0; y < P; ++y) prefilter_line(&Cfull[((size_t)z * P + y) * P], P, 1);
            for (int z = 0; z < P; ++z)
                for (int x = 0; x < P; ++x) prefilter_line(&Cfull[(size_t)z * P * P + x], P, P);
            for (int y = 0; y < P; ++y)
                for (int x = 0; x < P; ++x) prefilter_line(&Cfull[(size_t)y * P + x], P, (long)P * P);
            int idxMax = (int)(maxFrequency * P + 10);
            idxMax = std::min(P - 1 - hP, idxMax);
            int idxMin = std::max(-idxMax, -hP);
            dim = idxMax - idxMin + 1;
            start = idxMin;
            C.resize((size_t)dim * dim * dim);
            for (int z = 0; z < dim; ++z)
                for (int y = 0; y < dim; ++y)
                    for (int x = 0; x < dim; ++x)
                        C[((size_t)z * dim + y) * dim + x] = Cfull[((size_t)(z + idxMin + hP) * P + (y + idxMin + hP)) * P + (x + idxMin + hP)];
        } else {
            C.swap(Cfull);
        }
        // produceSideInfoProjection FP:310-333
        const int Xh = N / 2 + 1;
        phA.resize((size_t)N * Xh);
        phB.resize((size_t)N * Xh);
        const double shift = (double)(N / 2);                                   // -FIRST_XMIPP_INDEX(N)
        const double xxs = -2 * kPi * shift / N;
        for (int i = 0; i < N; ++i) {
            double phasey = (double)i * xxs;
            for (int j = 0; j < Xh; ++j) {
                double dotp = (double)j * xxs + phasey;
                phB[(size_t)i * Xh + j] = std::sin(dotp);
                phA[(size_t)i * Xh + j] = std::cos(dotp);
            }
        }
    }

    inline cd at(int k, int i, int j) const {      // logical indices; 0 outside
        k -= start; i -= start; j -= start;
        if (k < 0 || i < 0 || j < 0 || k >= dim || i >= dim || j >= dim) return cd(0, 0);
        return C[((size_t)k * dim + i) * dim + j];
    }
    static inline double bspline03(double x) {
        double a = std::fabs(x);
        if (a < 1.0) return a * a * (a - 2.0) * 0.5 + 2.0 / 3.0;
        if (a < 2.0) { a -= 2.0; return a * a * a / -6.0; }
        return 0.0;
    }

    // FP:91-262; ctf: N x (N/2+1) or nullptr; out: N x N
    void project(double rot, double tilt, double psi, const double* ctf, double* out) const {
        const int Xh = N / 2 + 1;
        M3 E = euler_matrix(rot, tilt, psi);
        std::vector<cd> PF((size_t)N * Xh, cd(0, 0));
        const double maxFreq2 = maxFrequency * maxFrequency;
        for (int i = 0; i < N; ++i) {
            double freqy = idx2digfreq(i, N), freqy2 = freqy * freqy;
            double fYX = E.m[3] * freqy, fYY = E.m[4] * freqy, fYZ = E.m[5] * freqy;
            for (int j = 0; j < Xh; ++j) {
                double freqx = idx2digfreq(j, N);
                if (freqy2 + freqx * freqx > maxFreq2) continue;
                double fX = fYX + E.m[0] * freqx, fY = fYY + E.m[1] * freqx, fZ = fYZ + E.m[2] * freqx;
                double c, d;
                if (degree == 0) {
                    cd v = at((int)std::round(fZ * P), (int)std::round(fY * P), (int)std::round(fX * P));   // outside: 0 (the reference reads out of bounds there)
                    c = v.real(); d = v.imag();
                } else if (degree == 1) {
                    double z = fZ * P, y = fY * P, x = fX * P;
                    int x0 = (int)std::floor(x), y0 = (int)std::floor(y), z0 = (int)std::floor(z);
                    double fx = x - x0, fy = y - y0, fz = z - z0;
                    cd d000 = at(z0, y0, x0), d001 = at(z0, y0, x0 + 1), d010 = at(z0, y0 + 1, x0), d011 = at(z0, y0 + 1, x0 + 1);
                    cd d100 = at(z0 + 1, y0, x0), d101 = at(z0 + 1, y0, x0 + 1), d110 = at(z0 + 1, y0 + 1, x0), d111 = at(z0 + 1, y0 + 1, x0 + 1);
                    auto lin = [](double a, cd l, cd h) { return l + (h - l) * a; };      // LIN_INTERP
                    cd dx00 = lin(fx, d000, d001), dx01 = lin(fx, d100, d101), dx10 = lin(fx, d010, d011), dx11 = lin(fx, d110, d111);
                    cd dxy0 = lin(fy, dx00, dx10), dxy1 = lin(fy, dx01, dx11);
                    cd v = lin(fz, dxy0, dxy1);
                    c = v.real(); d = v.imag();
                } else {
                    double z = fZ * P - start, y = fY * P - start, x = fX * P - start;
                    int l1 = (int)std::ceil(x - 2), m1 = (int)std::ceil(y - 2), n1 = (int)std::ceil(z - 2);
                    cd acc = 0;
                    for (int nn = n1; nn <= n1 + 3; ++nn) {
                        int en = nn < 0 ? -nn - 1 : (nn >= dim ? 2 * dim - nn - 1 : nn);
                        cd yx = 0;
                        for (int m = m1; m <= m1 + 3; ++m) {
                            int em = m < 0 ? -m - 1 : (m >= dim ? 2 * dim - m - 1 : m);
                            cd xs = 0;
                            for (int l = l1; l <= l1 + 3; ++l) {
                                int el = l < 0 ? -l - 1 : (l >= dim ? 2 * dim - l - 1 : l);
                                xs += C[((size_t)en * dim + em) * dim + el] * bspline03(x - (double)l);
                            }
                            yx += xs * bspline03(y - (double)m);
                        }
                        acc += yx * bspline03(z - (double)nn);
                    }
                    c = acc.real(); d = acc.imag();
                }
                double a = phA[(size_t)i * Xh + j], b = phB[(size_t)i * Xh + j];
                if (ctf) { a *= ctf[(size_t)i * Xh + j]; b *= ctf[(size_t)i * Xh + j]; }
                PF[(size_t)i * Xh + j] = cd(a * c - b * d, a * d + b * c);
            }
        }
        // inverseFourierTransform: complex along y for every stored column, then c2r along x (unnormalised)
        Fft f(N);
        std::vector<cd> ta(N), tb(N);
        for (int j = 0; j < Xh; ++j) f.run(&PF[j], Xh, +1, ta.data(), tb.data());
        for (int i = 0; i < N; ++i) {
            const cd* row = &PF[(size_t)i * Xh];
            ta[0] = cd(row[0].real(), 0);
            for (int x = 1; x < Xh; ++x) ta[x] = row[x];
            if (N % 2 == 0) ta[N / 2] = cd(row[N / 2].real(), 0);
            for (int x = Xh; x < N; ++x) ta[x] = std::conj(row[N - x]);
            f.rec(N, 1, ta.data(), tb.data(), +1);
            for (int x = 0; x < N; ++x) out[(size_t)i * N + x] = tb[x].real();
        }
    }
};

}  // namespace

// =============================================================================
// C ABI (see oracle_abi.h)
// =============================================================================
extern "C" {

void* orf_create(const orf_config* cfg) {
    try { return new Oracle(*cfg); } catch (...) { return nullptr; }
}
void orf_destroy(void* h) { delete static_cast<Oracle*>(h); }
void orf_dims(void* h, int* N, int* P, int* Z) {
    auto* o = static_cast<Oracle*>(h);
    *N = o->N; *P = o->P; *Z = o->Z;
}
void orf_insert_slabs(void* h, const float* imgs, const orf_particle* meta, int n, int threads) {
    static_cast<Oracle*>(h)->insert_slabs(imgs, meta, n, threads);
}
void orf_insert(void* h, const float* imgs, const orf_particle* meta, int n, int threads) {
    static_cast<Oracle*>(h)->insert(imgs, meta, n, threads);
}
// V as interleaved (re,im) doubles, W doubles; both Z*Z*(Z/2+1), [z][y][x]
void orf_get_accumulators(void* h, double* V, double* W) {
    auto* o = static_cast<Oracle*>(h);
    std::memcpy(V, o->V.data(), o->V.size() * sizeof(cd));
    std::memcpy(W, o->W.data(), o->W.size() * sizeof(double));
}
void orf_add_accumulators(void* h, const double* V, const double* W) {
    auto* o = static_cast<Oracle*>(h);
    for (size_t k = 0; k < o->V.size(); ++k) { o->V[k] += cd(V[2 * k], V[2 * k + 1]); o->W[k] += W[k]; }
}
void orf_finalize(void* h, double* out) { static_cast<Oracle*>(h)->finalize(out); }
void* orf_projector_create(const float* vol, int N, double padding, double max_freq, int degree) {
    if (degree != 0 && degree != 1 && degree != 3) return nullptr;
    try { return new ProjectorOracle(vol, N, padding, max_freq, degree); } catch (...) { return nullptr; }
}
void orf_projector_destroy(void* h) { delete static_cast<ProjectorOracle*>(h); }
void orf_projector_project(void* h, double rot, double tilt, double psi, const double* ctf, double* out) {
    static_cast<ProjectorOracle*>(h)->project(rot, tilt, psi, ctf, out);
}
void orf_finish_fourier(void* h, const double* Vri, double* out) {
    Oracle* o = static_cast<Oracle*>(h);
    std::vector<cd> v((size_t)o->Z * o->Z * o->X);
    for (size_t k = 0; k < v.size(); ++k) v[k] = cd(Vri[2 * k], Vri[2 * k + 1]);
    o->finish_fourier(v, out);
}
void orf_tables(void* h, double* blobTableSqrt, double* fourierBlobTable, double* iDeltaSqrt, double* iDeltaFourier) {
    auto* o = static_cast<Oracle*>(h);
    std::memcpy(blobTableSqrt, o->blobTableSqrt.data(), kTable * sizeof(double));
    std::memcpy(fourierBlobTable, o->fourierBlobTable.data(), kTable * sizeof(double));
    *iDeltaSqrt = o->iDeltaSqrt;
    *iDeltaFourier = o->iDeltaFourier;
}
// half-plane FFT of one image after shift/pad/CenterFFT: P*(P/2+1) interleaved doubles
void orf_preprocess(void* h, const float* img, const orf_particle* p, double* F, double* Ainv) {
    auto* o = static_cast<Oracle*>(h);
    std::vector<cd> f; M3 a; double w;
    o->preprocess(img, *p, f, a, w);
    std::memcpy(F, f.data(), f.size() * sizeof(cd));
    std::memcpy(Ainv, a.m, sizeof(a.m));
}
void orf_apply_shift(void* h, const float* img, double sx, double sy, double* out) {
    auto* o = static_cast<Oracle*>(h);
    std::vector<double> v((size_t)o->N * o->N);
    o->apply_shift(img, sx, sy, v);
    std::memcpy(out, v.data(), v.size() * sizeof(double));
}
// wCTF / wModulator for pixel (i,j) of the padded half-plane
void orf_ctf_weights(void* h, const orf_particle* p, int i, int j, double* wCTF, double* wMod) {
    auto* o = static_cast<Oracle*>(h);
    Ctf c(*p);
    o->ctf_weights(c, i, j, idx2digfreq(j, o->P), idx2digfreq(i, o->P), *wCTF, *wMod);
}
// spline primitives for the known-answer test: in-place coefficients of an n x n array, interpolation at physical (x, y)
void orf_bspline_coeffs_2d(double* a, int n) {
    for (int i = 0; i < n; ++i) Oracle::bspline_prefilter_1d(a + (size_t)i * n, n, 1);
    for (int j = 0; j < n; ++j) Oracle::bspline_prefilter_1d(a + j, n, n);
}
double orf_bspline_interp_2d(const double* c, int n, double x, double y) { return Oracle::bspline_interp_2d(c, n, x, y); }
double orf_ctf_value(const orf_particle* p, double X, double Y) { Ctf c(*p); return c.value(X, Y); }
double orf_ctf_argument(const orf_particle* p, double X, double Y) { Ctf c(*p); return c.argument(X, Y); }
double orf_ctf_K1(const orf_particle* p) { Ctf c(*p); return c.K1; }
// values (what = 0) or arguments (what = 1) on the n x n FFT grid of digital frequencies scaled by 1/Tm: the loop of
// errorBetween2CTFs / errorMaxFreqCTFs2D (data/ctf.cpp:150-165, 262-275)
void orf_ctf_grid(const orf_particle* p, int n, double Tm, int what, double* out) {
    Ctf c(*p);
    const double iTm = 1.0 / Tm;
    for (int i = 0; i < n; ++i) {
        double fy = idx2digfreq(i, n) * iTm;
        for (int j = 0; j < n; ++j) {
            double fx = idx2digfreq(j, n) * iTm;
            out[(size_t)i * n + j] = what ? c.argument(fx, fy) : c.value(fx, fy);
        }
    }
}

// --- known-answer-test entry points ---
void orf_euler(double rot, double tilt, double psi, double* m9) { M3 a = euler_matrix(rot, tilt, psi); std::memcpy(m9, a.m, sizeof(a.m)); }
double orf_idx2digfreq(int idx, int size) { return idx2digfreq(idx, size); }
double orf_kaiser_value(double r, double a, double alpha, int m) { return kaiser_value(r, a, alpha, m); }
double orf_kaiser_fourier_value(double w, double a, double alpha, int m) { return kaiser_Fourier_value(w, a, alpha, m); }
double orf_bessi0(double x) { return bessi0(x); }
double orf_bessj0(double x) { return bessj0(x); }
// FourierTransformer::FourierTransform on a real ny x nx array: out ny*(nx/2+1) interleaved, scaled 1/size
void orf_fft2_r2c(const double* in, int ny, int nx, double* out) {
    Fft fx(nx), fy(ny);
    int xh = nx / 2 + 1;
    std::vector<cd> rows((size_t)ny * nx), a(std::max(nx, ny)), b(std::max(nx, ny));
    for (int i = 0; i < ny; ++i) {
        for (int j = 0; j < nx; ++j) a[j] = in[(size_t)i * nx + j];
        fx.rec(nx, 1, a.data(), b.data(), -1);
        for (int j = 0; j < nx; ++j) rows[(size_t)i * nx + j] = b[j];
    }
    double inv = 1.0 / ((double)nx * ny);
    for (int j = 0; j < xh; ++j) {
        for (int i = 0; i < ny; ++i) a[i] = rows[(size_t)i * nx + j];
        fy.rec(ny, 1, a.data(), b.data(), -1);
        for (int i = 0; i < ny; ++i) { out[2 * ((size_t)i * xh + j)] = b[i].real() * inv; out[2 * ((size_t)i * xh + j) + 1] = b[i].imag() * inv; }
    }
}
// complex 1-D FFT for cross-checks against numpy
void orf_fft1(const double* in, int n, int sign, double* out) {
    Fft f(n);
    std::vector<cd> a(n), b(n);
    for (int i = 0; i < n; ++i) a[i] = cd(in[2 * i], in[2 * i + 1]);
    f.rec(n, 1, a.data(), b.data(), sign);
    for (int i = 0; i < n; ++i) { out[2 * i] = b[i].real(); out[2 * i + 1] = b[i].imag(); }
}

}  // extern "C"
