// oracle/recfourier_fast_oracle.cpp — TEST INFRASTRUCTURE ONLY (see recfourier_oracle.cpp).
//
// CPU restatement of the `--fast` arithmetic of Xmipp's Fourier reconstruction ("Do the blobing at the end of
// the computation", reconstruct_fourier_gpu.cpp:71-72): every voxel column crossing a projection plane takes the
// NEAREST pixel of the cropped, centred half-plane transform, and the Kaiser-Bessel blob is applied once, at the
// end, as a 3-D convolution of the accumulated volume and weights.  Followed here, function by function:
//   ProgRecFourierGPU (libraries/reconstruction_adapt_cuda/reconstruct_fourier_gpu.cpp = "G")
//     produceSideinfo G:207-290, cropAndShift G:292-321, prepareBuffer G:323-415, cuboid/AABB G:475-530,
//     computeCTFCorrection G:552-593, applyBlob G:623-661, convertToExpectedSpace G:664-681,
//     mirrorAndCrop G:697-730, forceHermitianSymmetry G:732-749, processWeights G:751-767,
//     computeTraverseSpace G:769-814, finishComputations G:879-932
//   device side (libraries/reconstruction_cuda/cuda_gpu_reconstruct_fourier.cpp = "D")
//     getZ/getY/getX D:391-413, multiply D:417-424, processVoxel D:455-503, processProjection<useFast> D:655-735
//   (CPU twin of the same arithmetic: ProgRecFourierAccel, reconstruct_fourier_accel.cpp:595-625, 717-745, 792-831)
// The reference computes all of this in single precision; so does this file, with every float operation written
// as its own statement and the translation unit compiled with -ffp-contract=off, so that the nearest-pixel and
// nearest-voxel decisions are reproducible.  (The reference's own device build lets nvcc contract a*b+c, so its
// last-bit behaviour is compiler dependent; decisions that hinge on it are ties by construction.)
// Parity status: unpinned by the reference's tests (no golden volume in-tree), like the exact path.
//
// Double-precision pieces (image shift, padding, forward FFT, Euler matrix, CTF value, final inverse FFT and
// gridding correction) are the ones of recfourier_oracle.cpp, reached through its C ABI.
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstring>
#include <limits>
#include <vector>

#include "oracle_abi.h"

namespace {

using cf = std::complex<float>;
constexpr int kTable = 10000;   // BLOB_TABLE_SIZE_SQRT

struct P3 { float x, y, z; };

// D:417-424 / G:475-482: left-to-right float evaluation
inline void multiply(const float t[9], P3& p) {
    float a0 = t[0] * p.x, a1 = t[1] * p.y, a2 = t[2] * p.z;
    float b0 = t[3] * p.x, b1 = t[4] * p.y, b2 = t[5] * p.z;
    float c0 = t[6] * p.x, c1 = t[7] * p.y, c2 = t[8] * p.z;
    float s0 = a0 + a1, s1 = b0 + b1, s2 = c0 + c1;
    p.x = s0 + a2;
    p.y = s1 + b2;
    p.z = s2 + c2;
}

struct TraverseSpace {           // reconstruct_fourier_projection_traverse_space.h:37-59
    int minX, minY, minZ, maxX, maxY, maxZ;
    int dir;                     // 0 = XY, 1 = XZ, 2 = YZ
    P3 unitNormal, bottomOrigin, topOrigin;
    float maxDistanceSqr;
    float transformInv[9];
    float weight;
};

struct FastOracle {
    orf_config cfg;
    int N, Pv, S, sx, sy;        // image, padded size (= N*pad_vol), maxVolumeIndexYZ, fftSizeX, fftSizeY
    void* geo = nullptr;         // base oracle with pad_proj := pad_vol (preprocessing: both programs pad to N*pad_vol)
    void* fin = nullptr;         // base oracle with the true paddings (tables, final transform and correction)
    std::vector<double> sym;     // identity + symmetry matrices
    std::vector<float> blobTableSqrt;
    float iDeltaSqrt;
    std::vector<cf> tempVolume;  // (S+1)^3 [z][y][x]
    std::vector<float> tempWeights;
    bool useFast = true;         // false: the GPU program WITHOUT --fast (blob applied per voxel by the device code, which
                                 // is not restated here: the temporary spaces then come from the compiled reference kernel,
                                 // oracle/ref_harness.cu); this object supplies its host side (buffers before, finish after)

    explicit FastOracle(const orf_config& c, bool fast = true) : cfg(c), useFast(fast) {
        N = c.img_size;
        Pv = (int)(N * c.pad_vol);                                              // G:229
        size_t conserveRows = (size_t)std::ceil((double)Pv * c.max_resolution * 2.0);   // G:230-232
        conserveRows = (size_t)std::ceil((double)conserveRows / 2.0);
        S = 2 * (int)conserveRows;
        sx = S / 2;                                                             // G:434
        sy = S;
        orf_config g = c;
        g.pad_proj = c.pad_vol;
        g.n_sym = 0;
        g.sym_matrices = nullptr;
        g.use_weights = 0;
        geo = orf_create(&g);
        orf_config f = c;
        f.n_sym = 0;
        f.sym_matrices = nullptr;
        f.n_iter_weight = 1;
        fin = orf_create(&f);
        std::vector<double> bt(kTable), ft(kTable);
        double ids, idf;
        orf_tables(fin, bt.data(), ft.data(), &ids, &idf);                      // G:235-272
        blobTableSqrt.resize(kTable);
        for (int i = 0; i < kTable; ++i) blobTableSqrt[i] = (float)bt[i];       // float blobTableSqrt[] (reconstruct_fourier_gpu.h)
        iDeltaSqrt = (float)ids;
        sym.assign(9, 0.0);
        sym[0] = sym[4] = sym[8] = 1.0;                                         // G:275-277
        for (int s = 0; s < c.n_sym; ++s) sym.insert(sym.end(), c.sym_matrices + 9 * s, c.sym_matrices + 9 * s + 9);
        const size_t n = (size_t)(S + 1) * (S + 1) * (S + 1);                   // G:838-843
        tempVolume.assign(n, cf(0, 0));
        tempWeights.assign(n, 0.f);
    }
    ~FastOracle() {
        orf_destroy(geo);
        orf_destroy(fin);
    }

    // ---- G:484-530
    static void createProjectionCuboid(P3* c, float sizeX, float sizeY, float blobSize) {
        float halfY = sizeY / 2.0f;
        c[3].x = c[2].x = c[7].x = c[6].x = 0.f - blobSize;
        c[0].x = c[1].x = c[4].x = c[5].x = sizeX + blobSize;
        c[3].y = c[0].y = c[7].y = c[4].y = -(halfY + blobSize);
        c[1].y = c[2].y = c[5].y = c[6].y = halfY + blobSize;
        c[3].z = c[0].z = c[1].z = c[2].z = 0.f + blobSize;
        c[7].z = c[4].z = c[5].z = c[6].z = 0.f - blobSize;
    }
    static void computeAABB(P3* AABB, const P3* c, float minX, float minY, float minZ, float maxX, float maxY, float maxZ) {
        AABB[0].x = AABB[0].y = AABB[0].z = std::numeric_limits<float>::max();
        AABB[1].x = AABB[1].y = AABB[1].z = std::numeric_limits<float>::min();   // sic (G:509)
        for (int i = 0; i < 8; ++i) {
            P3 t = c[i];
            if (AABB[0].x > t.x) AABB[0].x = t.x;
            if (AABB[0].y > t.y) AABB[0].y = t.y;
            if (AABB[0].z > t.z) AABB[0].z = t.z;
            if (AABB[1].x < t.x) AABB[1].x = t.x;
            if (AABB[1].y < t.y) AABB[1].y = t.y;
            if (AABB[1].z < t.z) AABB[1].z = t.z;
        }
        if (AABB[0].x < minX) AABB[0].x = minX;
        if (AABB[0].y < minY) AABB[0].y = minY;
        if (AABB[0].z < minZ) AABB[0].z = minZ;
        if (AABB[1].x > maxX) AABB[1].x = maxX;
        if (AABB[1].y > maxY) AABB[1].y = maxY;
        if (AABB[1].z > maxZ) AABB[1].z = maxZ;
    }

    // ---- G:769-814
    void computeTraverseSpace(const float transform[9], const float transformInv[9], TraverseSpace& sp) const {
        P3 cuboid[8], AABB[2];
        P3 origin = {S / 2.f, S / 2.f, S / 2.f};        // maxVolumeIndexX == maxVolumeIndexYZ while inserting
        const float blobSize = useFast ? 0.f : (float)cfg.blob_radius;              // G:775
        createProjectionCuboid(cuboid, (float)sx, (float)sy, blobSize);
        for (int i = 0; i < 8; ++i) multiply(transform, cuboid[i]);
        for (int i = 0; i < 8; ++i) {
            cuboid[i].x += origin.x;
            cuboid[i].y += origin.y;
            cuboid[i].z += origin.z;
        }
        computeAABB(AABB, cuboid, 0, 0, 0, (float)S, (float)S, (float)S);
        sp.minZ = (int)std::floor(AABB[0].z);
        sp.minY = (int)std::floor(AABB[0].y);
        sp.minX = (int)std::floor(AABB[0].x);
        sp.maxZ = (int)std::ceil(AABB[1].z);
        sp.maxY = (int)std::ceil(AABB[1].y);
        sp.maxX = (int)std::ceil(AABB[1].x);
        sp.topOrigin = cuboid[4];
        sp.bottomOrigin = cuboid[0];
        float e = (float)sx + blobSize;
        sp.maxDistanceSqr = e * e;
        std::memcpy(sp.transformInv, transformInv, sizeof(float) * 9);
        sp.unitNormal = {0.f, 0.f, 1.f};
        multiply(transform, sp.unitNormal);
        float nX = std::fabs(sp.unitNormal.x), nY = std::fabs(sp.unitNormal.y), nZ = std::fabs(sp.unitNormal.z);
        if (nX >= nY && nX >= nZ) sp.dir = 2;           // iterate the YZ plane
        else if (nY >= nX && nY >= nZ) sp.dir = 1;      // XZ
        else sp.dir = 0;                                // XY
    }

    // ---- D:391-413
    static float getZ(float x, float y, const P3& n, const P3& p0) {
        float a = -n.x, dx = x - p0.x, dy = y - p0.y;
        float t0 = a * dx, t1 = n.y * dy;
        float num = t0 - t1;
        float q = num / n.z;
        return q + p0.z;
    }
    static float getY(float x, float z, const P3& n, const P3& p0) {
        float a = -n.x, dx = x - p0.x, dz = z - p0.z;
        float t0 = a * dx, t1 = n.z * dz;
        float num = t0 - t1;
        float q = num / n.y;
        return q + p0.y;
    }
    static float getX(float y, float z, const P3& n, const P3& p0) {
        float a = -n.y, dy = y - p0.y, dz = z - p0.z;
        float t0 = a * dy, t1 = n.z * dz;
        float num = t0 - t1;
        float q = num / n.x;
        return q + p0.x;
    }
    static int clampi(int v, int lo, int hi) { return v > hi ? hi : (v < lo ? lo : v); }

    // ---- D:455-503
    void processVoxel(int x, int y, int z, const cf* img, const float* CTF, const float* mod, const TraverseSpace& sp) {
        P3 p;
        p.x = (float)(x - S / 2);
        p.y = (float)(y - S / 2);
        p.z = (float)(z - S / 2);
        float xx = p.x * p.x, yy = p.y * p.y, zz = p.z * p.z;
        float d = xx + yy;
        d = d + zz;
        if (d > sp.maxDistanceSqr) return;
        multiply(sp.transformInv, p);
        if (p.x < 0.f) return;
        float fx = p.x + 0.5f;
        float fy = p.y + 0.5f;
        fy = fy + (float)(S / 2);
        int imgX = clampi((int)fx, 0, sx - 1);
        int imgY = clampi((int)fy, 0, sy - 1);
        const size_t i3 = ((size_t)z * (S + 1) + y) * (S + 1) + x;
        const int i2 = imgY * sx + imgX;
        float wCTF = 1.f, wMod = 1.f;
        if (CTF) {
            wCTF = CTF[i2];
            wMod = mod[i2];
        }
        float weight = 1.f * wMod;
        weight = weight * sp.weight;
        float re = img[i2].real() * weight, im = img[i2].imag() * weight;
        re = re * wCTF;
        im = im * wCTF;
        tempVolume[i3] += cf(re, im);
        tempWeights[i3] += weight;
    }

    // ---- D:655-735 with useFast
    void processProjection(const cf* img, const float* CTF, const float* mod, const TraverseSpace& sp) {
        for (int idy = 0; idy <= S; ++idy)
            for (int idx = 0; idx <= S; ++idx) {
                if (sp.dir == 0) {
                    if (idy >= sp.minY && idy <= sp.maxY && idx >= sp.minX && idx <= sp.maxX) {
                        float hit = getZ((float)idx, (float)idy, sp.unitNormal, sp.bottomOrigin);
                        float r = hit + 0.5f;
                        processVoxel(idx, idy, (int)r, img, CTF, mod, sp);
                    }
                } else if (sp.dir == 1) {
                    if (idy >= sp.minZ && idy <= sp.maxZ && idx >= sp.minX && idx <= sp.maxX) {
                        float hit = getY((float)idx, (float)idy, sp.unitNormal, sp.bottomOrigin);
                        float r = hit + 0.5f;
                        processVoxel(idx, (int)r, idy, img, CTF, mod, sp);
                    }
                } else {
                    if (idy >= sp.minZ && idy <= sp.maxZ && idx >= sp.minY && idx <= sp.maxY) {
                        float hit = getX((float)idx, (float)idy, sp.unitNormal, sp.bottomOrigin);
                        float r = hit + 0.5f;
                        processVoxel((int)r, idx, idy, img, CTF, mod, sp);
                    }
                }
            }
    }

    // A_SL.inv() (xmippCore Matrix2D::inv, 3x3: adjugate over determinant), double
    static void inv3(const double m[9], double o[9]) {
        double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
        double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
        double id = 1.0 / det;
        o[0] = c00 * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
        o[3] = c01 * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
        o[6] = c02 * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
    }

    // ---- prepareBuffer G:323-415 for one image: cropped + centred transform, CTF tables, traverse spaces.
    // Returns false when the image is skipped (zero weight, G:351-353).
    bool prepare(const float* image, const orf_particle& p, std::vector<double>& F, cf* img, float* CTF, float* mod,
                 std::vector<TraverseSpace>& spaces) const {
        const int Xh = Pv / 2 + 1;
        const float maxResolutionSqr = (float)(cfg.max_resolution * cfg.max_resolution);   // float member (reconstruct_fourier_gpu.h)
        const double iTs = 1.0 / cfg.sampling;
        if (cfg.use_weights && (float)p.weight == 0.f) return false;                        // G:351-353
        double Ainv[9];
        orf_preprocess(geo, image, &p, F.data(), Ainv);                                     // G:384-394 (double FFT, 1/size)
        // cropAndShift G:292-321
        std::fill(img, img + (size_t)sx * sy, cf(0, 0));
        const int halfY = Pv / 2;
        for (int i = 0; i < Pv; ++i)
            for (int j = 0; j < sx; ++j) {
                if (!(i < sx || i >= Pv - sx)) continue;
                double re = F[2 * ((size_t)i * Xh + j)], im = F[2 * ((size_t)i * Xh + j) + 1];
                double f0 = orf_idx2digfreq(j, Pv), f1 = orf_idx2digfreq(i, Pv);
                if (f0 * f0 + f1 * f1 > maxResolutionSqr) re = im = 0.0;
                int myPadI = (i < halfY) ? i + sx : i - Pv + sx;
                img[(size_t)myPadI * sx + j] = cf((float)re, (float)im);
            }
        // computeCTFCorrection G:552-593 (its x loop runs to fftSizeY and spills into the following rows, which are
        // rewritten afterwards: the surviving values are the ones of x < fftSizeX)
        if (cfg.use_ctf) {
            for (int y = 0; y < sy; ++y) {
                float freqY = (y - (Pv / 2.f)) / (float)Pv;
                for (int x = 0; x < sx; ++x) {
                    float freqX = (float)orf_idx2digfreq(x, Pv);
                    float CTFVal, modulatorVal = 1.f;
                    CTFVal = (float)orf_ctf_value(&p, freqX * iTs, freqY * iTs);
                    if (std::isnan(CTFVal)) {
                        if (x == 0 && y == 0) modulatorVal = CTFVal = 1.0f;
                        else modulatorVal = CTFVal = 0.0f;
                    }
                    if (std::fabs(CTFVal) < cfg.min_ctf) {
                        modulatorVal = std::fabs(CTFVal);
                        CTFVal = (CTFVal >= 0) ? 1.f : -1.f;                             // SGN
                    } else {
                        CTFVal = (float)(1.0 / CTFVal);
                    }
                    if (cfg.phase_flipped) CTFVal = std::fabs(CTFVal);
                    CTF[(size_t)y * sx + x] = CTFVal;
                    mod[(size_t)y * sx + x] = modulatorVal;
                }
            }
        }
        const int nSym = (int)(sym.size() / 9);
        spaces.resize(nSym);
        for (int s = 0; s < nSym; ++s) {                                                 // G:363-377
            double A_SL[9], A_SLInv[9];
            const double* R = &sym[9 * s];
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) {
                    double t = 0;
                    for (int c = 0; c < 3; ++c) t += R[a * 3 + c] * Ainv[c * 3 + b];
                    A_SL[a * 3 + b] = t;
                }
            inv3(A_SL, A_SLInv);
            float transf[9], transfInv[9];
            for (int a = 0; a < 9; ++a) {
                transf[a] = (float)A_SL[a];
                transfInv[a] = (float)A_SLInv[a];
            }
            computeTraverseSpace(transf, transfInv, spaces[s]);
            spaces[s].weight = cfg.use_weights ? (float)p.weight : 1.0f;
        }
        return true;
    }

    // ---- prepareBuffer + processBufferKernel D:900-950 for n images (the --fast device arithmetic, restated)
    void insert(const float* imgs, const orf_particle* meta, int n) {
        const int Xh = Pv / 2 + 1;
        std::vector<double> F((size_t)Pv * Xh * 2);
        std::vector<cf> img((size_t)sx * sy);
        std::vector<float> CTF((size_t)sx * sy), mod((size_t)sx * sy);
        std::vector<TraverseSpace> spaces;
        for (int k = 0; k < n; ++k) {
            if (!prepare(imgs + (size_t)k * N * N, meta[k], F, img.data(), CTF.data(), mod.data(), spaces)) continue;
            for (const TraverseSpace& sp : spaces)
                processProjection(img.data(), cfg.use_ctf ? CTF.data() : nullptr, cfg.use_ctf ? mod.data() : nullptr, sp);
        }
    }

    // ---- the buffer ProgRecFourierGPU hands to processBufferGPU (RecFourierBufferData: FFTs, CTFs, modulators, spaces) for
    // n images; returns the number of images kept.  FFTs: kept x sy x sx complex, CTFs / mods: kept x sy x sx (untouched
    // without CTF), spaces: kept x nSym
    int export_buffer(const float* imgs, const orf_particle* meta, int n, float* FFTs, float* CTFs, float* mods, orf_space* out) const {
        const int Xh = Pv / 2 + 1;
        std::vector<double> F((size_t)Pv * Xh * 2);
        std::vector<float> CTF((size_t)sx * sy), mod((size_t)sx * sy);
        std::vector<TraverseSpace> spaces;
        const size_t px = (size_t)sx * sy;
        const int nSym = (int)(sym.size() / 9);
        int kept = 0;
        for (int k = 0; k < n; ++k) {
            cf* img = reinterpret_cast<cf*>(FFTs) + px * kept;
            if (!prepare(imgs + (size_t)k * N * N, meta[k], F, img, CTF.data(), mod.data(), spaces)) continue;
            if (cfg.use_ctf) {
                std::memcpy(CTFs + px * kept, CTF.data(), sizeof(float) * px);
                std::memcpy(mods + px * kept, mod.data(), sizeof(float) * px);
            }
            for (int s = 0; s < nSym; ++s) {
                const TraverseSpace& sp = spaces[s];
                orf_space& o = out[(size_t)kept * nSym + s];
                o.minX = sp.minX; o.minY = sp.minY; o.minZ = sp.minZ; o.maxX = sp.maxX; o.maxY = sp.maxY; o.maxZ = sp.maxZ;
                o.dir = sp.dir;
                o.projectionIndex = kept;
                o.maxDistanceSqr = sp.maxDistanceSqr;
                o.unitNormal[0] = sp.unitNormal.x; o.unitNormal[1] = sp.unitNormal.y; o.unitNormal[2] = sp.unitNormal.z;
                o.topOrigin[0] = sp.topOrigin.x; o.topOrigin[1] = sp.topOrigin.y; o.topOrigin[2] = sp.topOrigin.z;
                o.bottomOrigin[0] = sp.bottomOrigin.x; o.bottomOrigin[1] = sp.bottomOrigin.y; o.bottomOrigin[2] = sp.bottomOrigin.z;
                std::memcpy(o.transformInv, sp.transformInv, sizeof(float) * 9);
                o.weight = sp.weight;
            }
            ++kept;
        }
        return kept;
    }

    // ---- applyBlob G:623-661 on the half space [S+1][S+1][X+1]
    template <typename T>
    void applyBlob(std::vector<T>& a, int X) const {
        const float blobSize = (float)cfg.blob_radius;
        const float blobSizeSqr = blobSize * blobSize;
        const int blob = (int)std::floor(blobSize);
        std::vector<T> out(a.size());
        const size_t sy_ = (size_t)(X + 1), sz_ = (size_t)(S + 1) * (X + 1);
        for (int i = 0; i <= S; ++i)
            for (int j = 0; j <= S; ++j)
                for (int k = 0; k <= X; ++k) {
                    T tmp = (T)0;
                    for (int z = std::max(0, i - blob); z <= std::min(S, i + blob); ++z) {
                        float dZSqr = (float)((i - z) * (i - z));
                        for (int y = std::max(0, j - blob); y <= std::min(S, j + blob); ++y) {
                            float dYSqr = (float)((j - y) * (j - y));
                            for (int x = std::max(0, k - blob); x <= std::min(X, k + blob); ++x) {
                                float dXSqr = (float)((k - x) * (k - x));
                                float distanceSqr = dZSqr + dYSqr;
                                distanceSqr = distanceSqr + dXSqr;
                                if (distanceSqr > blobSizeSqr) continue;
                                float t = distanceSqr * iDeltaSqrt;
                                t = t + 0.5f;
                                int aux = (int)t;
                                float w = blobTableSqrt[aux];
                                T prod = w * a[z * sz_ + y * sy_ + x];
                                tmp += prod;
                            }
                        }
                    }
                    out[i * sz_ + j * sy_ + k] = tmp;
                }
        a.swap(out);
    }

    // ---- mirrorAndCrop G:704-730
    template <typename T, typename F>
    std::vector<T> mirrorAndCrop(const std::vector<T>& in, int X, F f) const {
        std::vector<T> out((size_t)(S + 1) * (S + 1) * (X + 1), (T)0);
        const size_t iy = (size_t)(S + 1), iz = (size_t)(S + 1) * (S + 1);
        const size_t oy = (size_t)(X + 1), oz = (size_t)(S + 1) * (X + 1);
        for (int z = 0; z <= S; ++z)
            for (int y = 0; y <= S; ++y)
                for (int x = 0; x <= S; ++x) {
                    if (x < X) out[(S - z) * oz + (S - y) * oy + (S - x - X)] += f(in[z * iz + y * iy + x]);
                    else out[z * oz + y * oy + (x - X)] += in[z * iz + y * iy + x];
                }
        return out;
    }

    void finalize(double* out) const {
        std::vector<double> VF = fourier();
        orf_finish_fourier(fin, VF.data(), out);                                             // G:894-931
    }

    // everything of finishComputations before the inverse transform: the Pv x Pv x (Pv/2+1) volume (interleaved re, im)
    std::vector<double> fourier() const {
        const int X = S / 2;                                                                  // G:698
        std::vector<float> W = mirrorAndCrop(tempWeights, X, [](float v) { return v; });
        std::vector<cf> V = mirrorAndCrop(tempVolume, X, [](cf v) { return std::conj(v); });
        if (useFast) {                                                                        // G:881-884
            applyBlob(V, X);
            applyBlob(W, X);
        }
        const size_t oy = (size_t)(X + 1), oz = (size_t)(S + 1) * (X + 1);
        for (int z = 0; z <= S; ++z)                                                          // forceHermitianSymmetry G:732-749
            for (int y = 0; y <= S / 2; ++y) {
                const size_t a = z * oz + y * oy, b = (size_t)(S - z) * oz + (size_t)(S - y) * oy;
                cf t1 = 0.5f * (V[b] + std::conj(V[a]));
                float t2 = 0.5f * (W[b] + W[a]);
                V[b] = t1;
                V[a] = std::conj(t1);
                W[b] = W[a] = t2;
            }
        const float corr2D_3D = (float)(std::pow(cfg.pad_proj, 2.) / (N * std::pow(cfg.pad_vol, 3.)));   // processWeights G:751-767
        for (size_t k = 0; k < V.size(); ++k) {
            float w = W[k];
            if (w > 0.001f) {
                float s = corr2D_3D / w;
                V[k] = cf(V[k].real() * s, V[k].imag() * s);
            } else
                V[k] = cf(0, 0);
        }
        // convertToExpectedSpace G:664-681 into the Pv x Pv x (Pv/2+1) double volume
        const int Xf = Pv / 2 + 1, half = S / 2;
        std::vector<double> VF((size_t)Pv * Pv * Xf * 2, 0.0);
        for (int z = 0; z <= S; ++z)
            for (int y = 0; y <= S; ++y)
                for (int x = 0; x <= half; ++x) {
                    int ny = (y < half) ? Pv - half + y : y - half;
                    int nz = (z < half) ? Pv - half + z : z - half;
                    size_t o = ((size_t)nz * Pv + ny) * Xf + x;
                    VF[2 * o] += V[z * oz + y * oy + x].real();
                    VF[2 * o + 1] += V[z * oz + y * oy + x].imag();
                }
        return VF;
    }
};

}  // namespace

extern "C" {

void* orf_fast_create(const orf_config* cfg) {
    try { return new FastOracle(*cfg); } catch (...) { return nullptr; }
}
void* orf_fast_create2(const orf_config* cfg, int use_fast) {
    try { return new FastOracle(*cfg, use_fast != 0); } catch (...) { return nullptr; }
}
int orf_fast_export_buffer(void* h, const float* imgs, const orf_particle* meta, int n, float* FFTs, float* CTFs, float* mods,
                           orf_space* spaces) {
    return static_cast<FastOracle*>(h)->export_buffer(imgs, meta, n, FFTs, CTFs, mods, spaces);
}
void orf_fast_tables(void* h, float* blobTableSqrt, float* iDeltaSqrt, float* iw0) {
    FastOracle* o = static_cast<FastOracle*>(h);
    std::memcpy(blobTableSqrt, o->blobTableSqrt.data(), sizeof(float) * kTable);
    *iDeltaSqrt = o->iDeltaSqrt;
    // iw0 = 1 / kaiser_Fourier_value(0, blobnormalized.radius, alpha, order)  (G:241-243); the table already carries it
    *iw0 = (float)(1.0 / orf_kaiser_fourier_value(0.0, o->cfg.blob_radius / (o->cfg.pad_proj / o->cfg.pad_vol), o->cfg.blob_alpha, o->cfg.blob_order));
}
void orf_fast_destroy(void* h) { delete static_cast<FastOracle*>(h); }
void orf_fast_dims(void* h, int* S, int* sx, int* sy, int* Pv) {
    FastOracle* o = static_cast<FastOracle*>(h);
    *S = o->S; *sx = o->sx; *sy = o->sy; *Pv = o->Pv;
}
void orf_fast_insert(void* h, const float* imgs, const orf_particle* meta, int n) { static_cast<FastOracle*>(h)->insert(imgs, meta, n); }
void orf_fast_get_temp(void* h, float* Vri, float* W) {
    FastOracle* o = static_cast<FastOracle*>(h);
    std::memcpy(Vri, o->tempVolume.data(), sizeof(cf) * o->tempVolume.size());
    std::memcpy(W, o->tempWeights.data(), sizeof(float) * o->tempWeights.size());
}
void orf_fast_set_temp(void* h, const float* Vri, const float* W) {
    FastOracle* o = static_cast<FastOracle*>(h);
    std::memcpy(o->tempVolume.data(), Vri, sizeof(cf) * o->tempVolume.size());
    std::memcpy(o->tempWeights.data(), W, sizeof(float) * o->tempWeights.size());
}
/* mirrorAndCrop (G:697-730) of the temporary spaces only: the accumulated half space [S+1][S+1][S/2+1] (interleaved complex /
 * real), index 0 of the last axis = centred x 0, before symmetrisation and weighting */
void orf_fast_half_spaces(void* h, float* Vri, float* W) {
    FastOracle* o = static_cast<FastOracle*>(h);
    const int X = o->S / 2;
    std::vector<float> w = o->mirrorAndCrop(o->tempWeights, X, [](float v) { return v; });
    std::vector<cf> v = o->mirrorAndCrop(o->tempVolume, X, [](cf c) { return std::conj(c); });
    std::memcpy(Vri, v.data(), sizeof(cf) * v.size());
    std::memcpy(W, w.data(), sizeof(float) * w.size());
}
void orf_fast_fourier(void* h, double* VFri) {
    std::vector<double> v = static_cast<FastOracle*>(h)->fourier();
    std::memcpy(VFri, v.data(), sizeof(double) * v.size());
}
void orf_fast_finalize(void* h, double* out) { static_cast<FastOracle*>(h)->finalize(out); }

}  // extern "C"
