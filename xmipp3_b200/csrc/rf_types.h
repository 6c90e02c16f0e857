// rf_types.h — POD types shared by the host code and the CUDA kernels of the
// B200 direct-Fourier reconstruction path.
#pragma once
#include <stdint.h>

namespace rfb200 {

// Blocked accumulator layout: tiles of 16 x 16 x 8 voxels, each made of 64 bricks of 4 x 4 x 2 voxels whose 32
// accumulators are contiguous (one coalesced warp access per brick).
constexpr int kTileX = 16, kTileY = 16, kTileZ = 8;
constexpr int kTileVox = kTileX * kTileY * kTileZ;   // 2048
constexpr int kLaunchPlanes = 264;           // (image, symmetry) planes of ONE class per gather launch: their PlaneS + PlaneD
                                             // tables travel as kernel parameters (31 KB of the 32,764-byte parameter space,
                                             // i.e. constant bank 0), so every launch carries its own tables: handles and
                                             // streams on one device are independent
constexpr int kBlobTable = 10000;            // BLOB_TABLE_SIZE_SQRT (reconstruct_fourier.h:41-44)
constexpr int kMaxWin = 8;                   // largest candidate window edge supported by the gather

// One projection plane = one (image, symmetry operator) pair.  M = R_sym * A^T
// (RF.cpp:936).  A voxel u (centred lattice coords, voxel units) sees the plane at
//   alpha = u . e1   (pixel units along image x)
//   beta  = u . e2   (pixel units along image y)
//   h     = u . n    (voxel units, signed distance to the plane)
// with e1 = M[:,0]*(P/Z), e2 = M[:,1]*(P/Z), n = M[:,2].
struct PlaneD {            // 72 B: double-precision plane (stick-origin projections, edge and damped-weight kernels)
    double e1[3];
    double e2[3];
    double n[3];
};

// Per-image CTF constants (data/ctf.cpp:645-680, 1392-1404), computed on the host in double.
struct CtfConsts {
    double K1, K2, K3, K5, K6, K7, Ksin, Kcos, K;
    double DeltaR, envR0, envR1, envR2, phase_shift, vpp_radius;
    double cos2az, sin2az;             // cos/sin of 2*rad_azimuth
    double defocus_average, defocus_deviation;
    int32_t has_envelope;              // any of K3,K5,K6,DeltaR,envR* non-zero
    int32_t has_vpp;                   // round(VPP_radius*1000) != 0
};

struct ImgParams {         // per image of a chunk
    float weight;          // 1 or the metadata weight (RF.cpp:374-381)
    int32_t mx, my;        // floor(-shiftX), floor(-shiftY): source pixel of destination j is j + mx (+ fraction)
    float ux, uy;          // fractional parts of -shift in [0,1); both 0 on the exact (integer) path
    int32_t spline;        // 1: cubic B-spline interpolation (a shift is fractional), 0: exact circular shift
    int32_t skip;          // weight == 0 -> image not inserted (RF.cpp:483-484)
    int32_t pad0;
};

// Edge work item: a lattice point that the main gather does not own (orig-only planes
// x = 0 exception row, x = Z/2, and wrap-around aliases at the Nyquist faces).
struct EdgeItem {
    int32_t ux, uy, uz;    // unwrapped lattice point
    int32_t mode;          // 0: originals + mirrors, 1: originals only
    int64_t store;         // blocked index of the stored voxel it adds into
};

// ---- stick gather (rf_sticks.cuh) ------------------------------------------------------------------------
// A stick is the unit of work of one warp: 4 x 4 columns running kStickL voxels along the axis d that dominates
// the normal of the planes it processes; two lanes share a column (even / odd depth), so the 32 voxels a warp
// handles at a time form a compact 4 x 4 x 2 sheet hugging the plane.  Class 0: d = x, (a,b) = (y,z); class 1:
// d = y, (a,b) = (x,z); class 2: d = z, (a,b) = (x,y).
constexpr int kStickA = 4, kStickB = 4;
constexpr int kStickCols = kStickA * kStickB;   // 16
#ifndef RF_STICK_L
#define RF_STICK_L 32
#endif
constexpr int kStickL = RF_STICK_L;
struct PlaneS {            // 48 B, __constant__: plane of one (image, symmetry) with components permuted to (a,b,d)
    float e1a, e1b, e1d;
    float e2a, e2b, e2d;
    float na, nb, nd;
    int32_t img;           // image index inside the chunk
    float weight;          // image weight (RF.cpp:374-381)
    float invNd;           // 1 / nd
};
struct StickUnit {         // 16 B: stick origin in stored-offset coordinates (x, y - lo, z - lo), permuted to (a,b,d)
    int32_t a0, b0, t0;
    int32_t pad;
};

struct Geometry {
    int32_t N, P, Z, X;            // image, padded image, padded volume, Z/2+1
    int32_t lo, hi;                // centred range of y,z: [lo, hi]
    int32_t tx, ty, tz;            // tiles per axis
    int32_t yHalf;                 // rows 1..yHalf of the x=0 plane are pair-averaged (RF.cpp:1190-1195)
    int32_t R;                     // largest |pixel index| passing the resolution cut-off
    int32_t Rp;                    // R + apron: slices are (2Rp+1)^2
    int32_t side;                  // 2Rp+1
    int32_t K;                     // candidate window edge = floor(2*rho)+1
    float rho;                     // blob radius in pixel units = r*P/Z
    float s2;                      // (Z/P)^2: pixel^2 -> voxel^2
    float r2;                      // blob radius^2 (voxel units)
    float r;                       // blob radius
    float iDelta;                  // (T-1)/r^2
    float sMax;                    // r^2 * iDelta: largest scaled squared distance inside the blob
    float reach;                   // maxRes*Z + r: no lattice point farther than this is touched
    float inplane_reach;           // R + rho (pixel units)
    // slice format v3 (stick gather): HALF-plane slices of overlapping pixel pairs.  Per image `side` rows (centred row i
    // at i + Rp) of `pitch` float4 entries; entry (i, j) = (pixel(i,j), pixel(i,j+1)) sits at column j + colOff for
    // j in [-colOff, Rp].  Pixels with j < -colOff are not stored: the full-plane slice F is Hermitian, F(-i,-j) =
    // conj F(i,j) (originals for j > 0, mirrors for j < 0, their sum on j = 0), so a voxel projecting to alpha < 0 is
    // gathered at (-alpha, -beta) and its sum conjugated.  Half the bytes of the full-plane format for K1c to write.
    int32_t colOff;                // stored columns left of j = 0 (= K: enough for any window of a voxel with alpha >= 0)
    int32_t pitch;                 // row pitch in entries
    int32_t planeStride;           // side * pitch entries per image
    int32_t xOwnMax;               // largest ux owned by the main gather (originals and mirrors both land there)
    float rimIn2;                  // (pixel radius)^2 inside which every candidate of a window is a valid pixel
                                   // with multiplicity 1 unless the window touches column j = 0
};

}  // namespace rfb200
