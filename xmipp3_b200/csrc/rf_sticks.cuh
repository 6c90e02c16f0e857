// rf_sticks.cuh — second-generation insertion kernels (sm_100a).
//
//   K1b' k_make_slices2     half-plane FFT -> half-plane slice of overlapping pixel PAIRS: entry (i,j) is the float4
//                           (pixel(i,j), pixel(i,j+1)), so a candidate-window row is two 16-byte loads whatever the
//                           parity of its origin, and neighbouring lanes share cache lines; columns j >= -colOff only
//                           (the slice is Hermitian: a voxel at alpha < 0 reads (-alpha, -beta) and conjugates) + the
//                           true weights of the CTF-damped pixels (RF.cpp:600-625 hoisted)
//   K2'  k_gather_sticks    voxel-centric gather, "column walk": a warp owns a stick of 4 x 4 columns that run
//                           along the axis dominating the plane normal; two lanes per column (even / odd depth);
//                           at every step the lane's voxel is inside the blob slab of the plane, so (nearly) all
//                           lanes do useful work, and the 32 voxels of a step are compact (few image rows per
//                           load instruction).  Accumulators of the stick live in shared memory ([depth][column]:
//                           conflict-free for any per-column depth), ownership is exclusive -> bit-reproducible.
//                           Replaces the scatter loop RF.cpp:586-792.
//   K2e' k_edge2            lattice points the stick gather does not own (as k_edge, new slice format)
//   K2r  k_damped_scatter   W of the (rare) pixels whose CTF is below --minCTF: their weight is |CTF| instead of 1
//                           (RF.cpp:616-622).  They carry a flag in the LSB of `re`; the gather counts every other
//                           valid pixel with the image weight (no per-pixel weight fetch in the hot loop) and skips
//                           the flagged ones, this kernel adds their true weights with 64-bit fixed-point atomics
//                           (order independent -> deterministic; only positive terms -> no cancellation)
//        k_fold_damped      W += damped weights, once before W is consumed
#pragma once
#include <cuda.h>          // CUtensorMap (type only: the map is encoded on the host through the driver entry point)

#include "rf_kernels.cuh"

namespace rfb200 {

// The planes of one launch (one class, components permuted to (a,b,d)) are kernel PARAMETERS: FP32 for the per-voxel
// arithmetic, FP64 for the projection of the stick origins.  Parameters live in constant bank 0, so the (warp-uniform)
// accesses are the same LDC / LDCU instructions a __constant__ array gives, but every launch owns its tables: two handles
// or streams on one device never share state (the reference runs N host threads on N streams against one volume,
// reconstruct_fourier_gpu.cpp:417-473).

#ifndef RF_STICK_WARPS
#define RF_STICK_WARPS 16
#endif
#ifndef RF_STICK_LANEMAP
#define RF_STICK_LANEMAP 1     // lane <-> (column, depth parity) map of the gather, see k_gather_sticks (measured: 0: 29.62, 1: 29.27, 2: 29.43 ms)
#endif
constexpr int kStickWarps = RF_STICK_WARPS;
constexpr int kStickThreads = kStickWarps * 32;
constexpr size_t kStickSmem = (size_t)kStickWarps * kStickL * kStickCols * (sizeof(float2) + sizeof(float));
static_assert(kStickL <= 64 && kStickL % 4 == 0, "touched mask: one bit per two depths; bricks are 4 deep along x and y");
constexpr double kFixedScale = 4294967296.0;   // 2^32 fixed point

// rimTab entry of centred slice row i: (jPos+1) | (jNeg+1) << 14 | m0 << 28 with
//   jPos = largest valid original column of row i (-1: none), jNeg = same for row -i (serves columns j < 0 through
//   the Hermitian mirror), m0 = multiplicity of column 0 (original (0,i) + mirror of original (0,-i)).
__device__ __forceinline__ float d_rim_mult(int rt, int j) {
    const int jPos = (rt & 0x3fff) - 1, jNeg = ((rt >> 14) & 0x3fff) - 1;
    const float side = (j > 0) ? (j <= jPos ? 1.f : 0.f) : (-j <= jNeg ? 1.f : 0.f);
    return j == 0 ? (float)(rt >> 28) : side;
}

// ================================================================== K1b'
// contribution of original half-plane pixel (j >= 0, ip): (x, y) = weight*wMod*wCTF*F/P^2, z = weight*wMod
// (0 for a pixel outside the cut-off), w = 1 if the CTF damps the pixel (wMod != 1)
__device__ __forceinline__ float4 d_pixel_contrib2(const float2* __restrict__ fft, const int* __restrict__ jmax,
                                                   const SliceParams& sp, const CtfConsts* ctf, const CtfFloat& cf, float weight, int j, int ip) {
    float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ip < sp.iLo || ip > sp.iHi) return out;
    if (j > jmax[ip - sp.iLo]) return out;            // resolution cut-off, RF.cpp:597
    int row = ip < 0 ? ip + sp.P : ip;
    float2 F = __ldg(fft + (size_t)row * sp.Xh + j);
    float wc = 1.f, wm = 1.f;
    if (sp.useCtf) d_ctf_weights(*ctf, cf, sp, j, ip, wc, wm);
    float s = weight * wm * wc * sp.invP2;
    out.x = F.x * s;
    out.y = F.y * s;
    out.z = weight * wm;
    out.w = (wm != 1.0f) ? 1.f : 0.f;
    return out;
}
__device__ __forceinline__ float d_set_flag(float re, bool flag) {
    return __uint_as_float((__float_as_uint(re) & ~1u) | (flag ? 1u : 0u));
}

struct Slice2Args {
    SliceParams sp;
    int pitch, planeStride, colOff;
    const float2* fft;
    float4* slices;       // per image side x pitch entries (pixel(i,j), pixel(i,j+1))
    float2* col0;         // per image `side` originals-only entries of column j = 0
    float* damped;        // per image (2R+1) x (R+1): true weight of a flagged pixel, -1 if not flagged; nullptr without CTF
    float* damped2;       // same shape: un-modulated weight of a flagged pixel (only for --iter > 1), else nullptr
    uint32_t* dampedMask; // per image (2R+1) x ceil((R+1)/32) words: bit b of word w of a row <-> column 32w+b is flagged
    const ImgParams* ip;
    const CtfConsts* ctfs;
    const int* jmax;
};

// same thread mapping as k_make_slices: grid (ceil((R+1)/32), ceil((2R+1)/32), nImg), block (32, 8)
__global__ void __launch_bounds__(256, 3) k_make_slices2(const __grid_constant__ Slice2Args a) {
    const SliceParams& sp = a.sp;
    const int j = blockIdx.x * 32 + threadIdx.x;
    const bool active = j <= sp.R;          // a warp = 32 consecutive columns of one row; no early exit (ballots below)
    const int img = blockIdx.z;
    const float2* f = a.fft + (size_t)img * sp.P * sp.Xh;
    // the image's CTF constants are staged in shared memory once per CTA (19 doubles that every pixel needs)
    __shared__ CtfConsts sCtf;
    __shared__ CtfFloat sCtfF;
    if (sp.useCtf) {
        const int t = threadIdx.y * 32 + threadIdx.x;
        if (t < (int)(sizeof(CtfConsts) / 8)) reinterpret_cast<double*>(&sCtf)[t] = reinterpret_cast<const double*>(a.ctfs + img)[t];
        if (t == 32) d_ctf_prepare(a.ctfs[img], sp, sCtfF);
        __syncthreads();
    }
    const CtfConsts* ctf = sp.useCtf ? &sCtf : nullptr;
    const float weight = a.ip[img].weight;
    // as float2: entry (i,j) = elements 2*(i*pitch+j) [pixel (i,j)] and +1 [pixel (i,j+1)]; a pixel is written to
    // its own entry and to the second half of the entry on its left: two adjacent 8-byte stores
    float2* S2 = reinterpret_cast<float2*>(a.slices + (size_t)img * a.planeStride);
    const size_t dOff = (size_t)img * (2 * sp.R + 1) * (sp.R + 1);
    const int wordsPerRow = (sp.R + 1 + 31) / 32;
    const int rowBase = blockIdx.y * (8 * kSliceRowsPerThread) + threadIdx.y;
#pragma unroll 2
    for (int q = 0; q < kSliceRowsPerThread; ++q) {
        const int r = rowBase + 8 * q;
        if (r > 2 * sp.R) break;
        const int ipx = r - sp.R;
        bool flag = false;
        if (active) {
            float4 c = d_pixel_contrib2(f, a.jmax, sp, ctf, sCtfF, weight, j, ipx);
            const size_t o1 = (size_t)(ipx + sp.Rp) * a.pitch + (j + a.colOff);
            flag = c.w != 0.f;
            float unmod = (c.z != 0.f || flag) ? weight : 0.f;       // weight of a valid pixel without the CTF modulator
            if (j > 0) {
                const float re = d_set_flag(c.x, flag);
                const float2 v = make_float2(re, c.y), vm = make_float2(re, -c.y);
                S2[2 * o1] = v;  S2[2 * o1 - 1] = v;
                // only the first colOff mirrored columns are stored (half-plane format)
                const size_t o2 = (size_t)(-ipx + sp.Rp) * a.pitch + (-j + a.colOff);
                if (j <= a.colOff) S2[2 * o2] = vm;
                if (j < a.colOff) S2[2 * o2 - 1] = vm;
            } else {
                // column j = 0 holds original (0,ip) plus the mirror of original (0,-ip): the reference inserts this
                // column twice for x > 0 voxels.  The combined entry is flagged if either part is
                // damped; the damped-weight pass then supplies the weights of both parts.
                float4 m = d_pixel_contrib2(f, a.jmax, sp, ctf, sCtfF, weight, 0, -ipx);
                flag = flag || (m.w != 0.f);
                const float2 v = make_float2(d_set_flag(c.x + m.x, flag), c.y - m.y);
                S2[2 * o1] = v; S2[2 * o1 - 1] = v;
                a.col0[(size_t)img * sp.side + (ipx + sp.Rp)] = make_float2(d_set_flag(c.x, flag), c.y);
            }
            if (a.damped && flag) a.damped[dOff + (size_t)r * (sp.R + 1) + j] = c.z;
            if (a.damped2 && flag) a.damped2[dOff + (size_t)r * (sp.R + 1) + j] = unmod;
        }
        if (a.dampedMask) {
            const unsigned mw = __ballot_sync(0xffffffffu, flag);
            if (threadIdx.x == 0) a.dampedMask[((size_t)img * (2 * sp.R + 1) + r) * wordsPerRow + blockIdx.x] = mw;
        }
    }
}

// ================================================================== K2'
struct StickArgs {
    Geometry geo;
    const StickUnit* units;      // sticks of this class, heaviest (closest to the origin) first
    int nUnits;
    int* counter;
    int cls;                     // 0: d = x, 1: d = y, 2: d = z
    int nPlanes;                 // planes of this launch: StickLaunch::ps / pd [0, nPlanes)
    const float* blobTable;
    const float* planesSoA;      // 9 x kLaunchPlanes floats, same planes as SoA (culling phase, lane <-> plane)
    const float4* slices;        // slice format v2: overlapping pixel pairs
    const int* rimTab;           // already offset by +Rp: index with the centred row
    float2* Vb;
    float* Wb;
    float* Wb2;                  // un-modulated weight sum for --iter > 1 with CTF (else nullptr)
};

// ---- packed single precision (sm_100: FADD2 / FMUL2 / FFMA2 operate on an aligned register pair and take a scalar
// broadcast operand, so {x, x} pairs are free).  The gather is bound by instruction issue, not by the FP32 pipe: one
// packed instruction does the work of two issue slots.
typedef unsigned long long u64;
__device__ __forceinline__ u64 d_pk(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void d_upk(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 d_add2(u64 a, u64 b) {
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 d_mul2(u64 a, u64 b) {
    u64 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 d_fma2(u64 a, u64 b, u64 c) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// Address of 16-byte entry `off` of a (warp-uniform) slice: one IMAD.WIDE.U32 instead of the sign extension + 64-bit
// shift/add chain a signed index costs.
__device__ __forceinline__ u64 d_entry_addr(const float4* base, unsigned off) {
    u64 r;
    asm("mad.wide.u32 %0, %1, 16, %2;" : "=l"(r) : "r"(off), "l"(base));
    return r;
}
// Pixel pair at row + kByteOff, fetched only if dy <= lim, zero otherwise.  (The zeroing costs two CS2R per pair.  Declaring the
// registers as plain outputs of a predicated load would save them - a skipped pair is only read by predicated-off
// candidates - but then the 32 pixel registers are live across the whole step loop and the kernel spills: measured at
// compile time, 350 bytes of spill traffic in the hot loop.)
template <int kByteOff>
__device__ __forceinline__ float4 d_load_pair_if(u64 row, float dy, float lim) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    asm("{\n\t"
        ".reg .pred p;\n\t"
        "setp.le.f32 p, %5, %6;\n\t"
        "@p ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4+%7];\n\t"
        "}"
        : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)
        : "l"(row), "f"(dy), "f"(lim), "n"(kByteOff));
    return v;
}

// Two adjacent candidates (tj = 2q, 2q+1) of one window row, fully predicated:
//   S = dy + dx (one FADD2 with dy broadcast); idx = round(S) by the 2^23 trick (one FADD2); per candidate: accept test,
//   table address (LEA), table load, V += w * (re, im), flag test fused into the predicate (LOP3.PAND), W += w (* multiplicity
//   on the checked path).  (A predicated FFMA2 is lowered to FFMA2 + 2 SEL, so the complex update stays two scalar FFMAs.)
//   Even candidates go to accumulator set A, odd ones to set B.
struct PairAcc { float reA, imA, wA, reB, imB, wB; };
template <bool kSlow, bool kFlags>
__device__ __forceinline__ void d_pair(const float dy, const u64 dx2, const float sMax, const uint32_t tblAdj, const float4 px,
                                       const float m0, const float m1, PairAcc& A) {
    const u64 dy2 = d_pk(dy, dy);
    const u64 magic2 = d_pk(8388608.0f, 8388608.0f);
#define RF_PAIR_HEAD                                   \
    "{\n\t"                                            \
    ".reg .b64 s2, m2;\n\t"                            \
    ".reg .f32 s0, s1, w0, w1;\n\t"                    \
    ".reg .b32 i0, i1, t0, t1;\n\t"                    \
    ".reg .pred p0, p1, q0, q1;\n\t"                   \
    "add.rn.f32x2 s2, %6, %7;\n\t"                     \
    "add.rn.f32x2 m2, s2, %8;\n\t"                     \
    "mov.b64 {s0, s1}, s2;\n\t"                        \
    "mov.b64 {i0, i1}, m2;\n\t"                        \
    "setp.le.f32 p0, s0, %9;\n\t"                      \
    "setp.le.f32 p1, s1, %9;\n\t"                      \
    "shl.b32 i0, i0, 2;\n\t"                           \
    "add.u32 i0, i0, %10;\n\t"                         \
    "shl.b32 i1, i1, 2;\n\t"                           \
    "add.u32 i1, i1, %10;\n\t"                         \
    "@p0 ld.shared.f32 w0, [i0];\n\t"                  \
    "@p1 ld.shared.f32 w1, [i1];\n\t"                  \
    "@p0 fma.rn.f32 %0, w0, %11, %0;\n\t"              \
    "@p0 fma.rn.f32 %1, w0, %12, %1;\n\t"              \
    "@p1 fma.rn.f32 %3, w1, %13, %3;\n\t"              \
    "@p1 fma.rn.f32 %4, w1, %14, %4;\n\t"
#define RF_PAIR_OUT "+f"(A.reA), "+f"(A.imA), "+f"(A.wA), "+f"(A.reB), "+f"(A.imB), "+f"(A.wB)
#define RF_PAIR_IN "l"(dy2), "l"(dx2), "l"(magic2), "f"(sMax), "r"(tblAdj), "f"(px.x), "f"(px.y), "f"(px.z), "f"(px.w)
    if (kFlags) {
        if (kSlow)
            asm(RF_PAIR_HEAD
                "lop3.and.b32 t0|q0, %15, 1, 0, 0x0C, p0;\n\t"
                "lop3.and.b32 t1|q1, %16, 1, 0, 0x0C, p1;\n\t"
                "@q0 fma.rn.f32 %2, w0, %17, %2;\n\t"
                "@q1 fma.rn.f32 %5, w1, %18, %5;\n\t"
                "}"
                : RF_PAIR_OUT
                : RF_PAIR_IN, "r"(__float_as_uint(px.x)), "r"(__float_as_uint(px.z)), "f"(m0), "f"(m1));
        else
            asm(RF_PAIR_HEAD
                "lop3.and.b32 t0|q0, %15, 1, 0, 0x0C, p0;\n\t"
                "lop3.and.b32 t1|q1, %16, 1, 0, 0x0C, p1;\n\t"
                "@q0 add.rn.f32 %2, %2, w0;\n\t"
                "@q1 add.rn.f32 %5, %5, w1;\n\t"
                "}"
                : RF_PAIR_OUT
                : RF_PAIR_IN, "r"(__float_as_uint(px.x)), "r"(__float_as_uint(px.z)));
    } else {
        if (kSlow)
            asm(RF_PAIR_HEAD
                "@p0 fma.rn.f32 %2, w0, %15, %2;\n\t"
                "@p1 fma.rn.f32 %5, w1, %16, %5;\n\t"
                "}"
                : RF_PAIR_OUT
                : RF_PAIR_IN, "f"(m0), "f"(m1));
        else
            asm(RF_PAIR_HEAD
                "@p0 add.rn.f32 %2, %2, w0;\n\t"
                "@p1 add.rn.f32 %5, %5, w1;\n\t"
                "}"
                : RF_PAIR_OUT
                : RF_PAIR_IN);
    }
#undef RF_PAIR_HEAD
#undef RF_PAIR_OUT
#undef RF_PAIR_IN
}

// One step of one column, second generation: the K x K candidate window of the lane's voxel, two candidates per asm
// block.  `off` is the entry index of the window origin inside the image's slice `sl` (warp-uniform base, 32-bit per-lane
// offset: a row address is one IMAD.WIDE.U32); (da0, db0) is the offset of the projected voxel from the window origin,
// h2s = h^2 * iDelta.  kSlow weighs every candidate with its multiplicity (0 outside the resolution disc, 2 on column
// j = 0), looked up per window row.
template <int K, bool kSlow, bool kFlags>
__device__ __forceinline__ void d_stick_window2(const float4* __restrict__ sl, const unsigned off, const unsigned pitch, const float da0, const float db0,
                                                const float h2s, const float kI, const float sMax, const uint32_t tblAdj, const int jc, const int ic,
                                                const int* __restrict__ rimTab, float& accRe, float& accIm, float& accW) {
    constexpr int NP = (K + 1) / 2;
    const u64 kI2 = d_pk(kI, kI);
    // squared row offsets + h^2, two per register pair: dys[q] = (kI * db) * db + h2s with db = db0 - q
    float dys[2 * NP];
    {
        const u64 db2 = d_pk(db0, db0), h2 = d_pk(h2s, h2s);
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            const u64 d = d_add2(db2, d_pk(-(float)(2 * q), -(float)(2 * q + 1)));
            d_upk(d_fma2(d_mul2(kI2, d), d, h2), dys[2 * q], dys[2 * q + 1]);
        }
    }
    // squared column offsets, two per register pair; lim[q]: a pair is needed iff dys <= sMax - min(dxs of the pair).  The
    // load predicate carries a slack of 0.02 table steps, so that it can never be false for an accepted candidate whatever
    // the rounding of dy + dx (S < 10^4, one ulp = 10^-3)
    u64 dxs2[NP];
    float lim[NP];
    const u64 da2 = d_pk(da0, da0);
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        const u64 d = d_add2(da2, d_pk(-(float)(2 * q), -(float)(2 * q + 1)));
        u64 v = d_mul2(d_mul2(kI2, d), d);
        float lo, hi;
        d_upk(v, lo, hi);
        if (2 * q + 1 >= K) {          // odd window edge: the second candidate of the last pair does not exist
            hi = 3.0e38f;
            v = d_pk(lo, hi);
        }
        dxs2[q] = v;
        lim[q] = (sMax + 0.02f) - fminf(lo, hi);
    }
    // The window is processed in groups of RF_WINDOW_ROWS rows (loads of a group, then its candidates): the live pixel
    // registers are the dominant part of the register budget (16 warps per SM leave 128 registers per thread)
#ifndef RF_WINDOW_ROWS
#define RF_WINDOW_ROWS 2
#endif
    constexpr int G = RF_WINDOW_ROWS;
    PairAcc A = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int t0 = 0; t0 < K; t0 += G) {
        float4 px[G][NP];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            if (t0 + g < K) {
                const u64 row = d_entry_addr(sl, off + (unsigned)(t0 + g) * pitch);
#pragma unroll
                for (int q = 0; q < NP; ++q) {
#ifdef RF_STICK_UNCOND_LOADS
                    px[g][q] = __ldg(reinterpret_cast<const float4*>(row) + 2 * q);
#else
                    // A pixel pair is fetched only if one of its two candidates can be accepted: fewer lanes per load
                    // instruction means fewer cache lines (L1 wavefronts) per instruction.
                    if (q == 0) px[g][q] = d_load_pair_if<0>(row, dys[t0 + g], lim[q]);
                    else if (q == 1) px[g][q] = d_load_pair_if<32>(row, dys[t0 + g], lim[q]);
                    else if (q == 2) px[g][q] = d_load_pair_if<64>(row, dys[t0 + g], lim[q]);
                    else px[g][q] = d_load_pair_if<96>(row, dys[t0 + g], lim[q]);
#endif
                }
            }
        }
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int ti = t0 + g;
            if (ti < K) {
                int rt = 0;
                if (kSlow) rt = __ldg(rimTab + ic + ti);
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    float m0 = 1.f, m1 = 1.f;
                    if (kSlow) {
                        m0 = d_rim_mult(rt, jc + 2 * q);
                        m1 = d_rim_mult(rt, jc + 2 * q + 1);
                    }
                    d_pair<kSlow, kFlags>(dys[ti], dxs2[q], sMax, tblAdj, px[g][q], m0, m1, A);
                }
            }
        }
    }
    accRe = A.reA + A.reB;
    accIm = A.imA + A.imB;
    accW = A.wA + A.wB;
}

// fire-and-forget adds to the accumulators: one writer per address per launch (exclusive stick ownership) and
// launches are ordered on the stream, so the result is still bit-reproducible; no load latency in the write-out
__device__ __forceinline__ void d_red_add2(float2* addr, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void d_red_add(float* addr, float a) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}

struct StickLaunch {             // the whole parameter block of one launch (< 32,764 bytes)
    StickArgs a;
    PlaneS ps[kLaunchPlanes];
    PlaneD pd[kLaunchPlanes];
#ifdef RF_L2_PREFETCH
    // (measured variant) TMA descriptor of the chunk's slices as a 3-D tensor (floats of a row, rows, images), box = 16
    // entries x 16 rows: the patch of the next plane is requested into L2 with one UTMAPF per box
    alignas(64) CUtensorMap sliceMap;
#endif
};
static_assert(sizeof(StickLaunch) <= 32764, "kernel parameter space");

// A task = one plane crossing one stick.  Per-lane state of the column walk; `k` indexes StickLaunch::ps.
// Half-plane slices: a voxel projecting to alpha < 0 is gathered at (-alpha, -beta) and its sum conjugated.  On the
// unchecked path the sign is constant along a lane's walk and is folded into the task state once (sg, and the signs of
// ja0, jb0, arL, brL, e1d, e2d); on the checked path the state is unflipped (sg = 1) and every step decides for itself.
struct StickTask {
    int k;
    int ja0, jb0;          // integer pixel of the stick origin's projection
    float arL, brL, hL;    // lane's column at tau = 0: fractional in-plane position and height above the plane
    float e1d, e2d;        // in-plane step per unit depth
    float sg;              // +1 / -1: sign already applied to the in-plane quantities above (unchecked path)
    int tauLo, tauHi;      // the lane's first depth (column start + parity) and the column's last depth inside the slab
};

// geometry constants of the kernel that the step functions need
struct StickConsts {
    int Rp, side, pitch, planeStride, colOff;
    float rho, iDelta, sMax, kI, rimIn2, rhoCol0;
    uint32_t tblAdj;
};

// State of one step (one voxel of the lane's column) in the coordinates of the STORED half plane.
struct StickStep {
    int ja, jb;            // integer pixel of the stick origin's projection, sign applied
    int jw, iw;            // window origin relative to (ja, jb)
    float ar, br, h;       // in-plane position relative to (ja, jb), sign applied; height above the plane
    float sg;              // the sign: the voxel's sum is conjugated when sg < 0
    bool special, inb;     // needs the multiplicity path / window inside the stored slice
};
template <int K, bool kChecked>
__device__ __forceinline__ StickStep d_step(const StickConsts& c, const StickTask& t, const float nd, const int tau) {
    StickStep q;
    const float ft = (float)tau;
    q.ar = fmaf(ft, t.e1d, t.arL);
    q.br = fmaf(ft, t.e2d, t.brL);
    q.h = fmaf(ft, nd, t.hL);
    q.ja = t.ja0;
    q.jb = t.jb0;
    q.sg = t.sg;
    q.special = false;
    q.inb = true;
    if (kChecked) {
        // Fast path needs every ACCEPTED candidate (in-plane distance <= rho) to be a valid pixel of multiplicity 1:
        // inside the all-valid disc and away from column 0.
        const float Aabs = (float)t.ja0 + q.ar, Babs = (float)t.jb0 + q.br;
        q.special = (Aabs * Aabs + Babs * Babs > c.rimIn2) || (fabsf(Aabs) <= c.rhoCol0);
        if (Aabs < 0.f) {          // mirrored half: F(-i, -j) = conj F(i, j)
            q.sg = -1.f;
            q.ar = -q.ar; q.br = -q.br;
            q.ja = -q.ja; q.jb = -q.jb;
        }
    }
    q.jw = __float2int_ru(q.ar - c.rho);
    q.iw = __float2int_ru(q.br - c.rho);
    if (kChecked) {
        const int jAbs = q.ja + q.jw + c.colOff, iAbs = q.jb + q.iw + c.Rp;
        q.inb = (unsigned)jAbs <= (unsigned)(c.colOff + c.Rp + 1 - K) && (unsigned)iAbs <= (unsigned)(c.side - K);
    }
    return q;
}

// Slot of (depth tau, column col) in the warp's complex accumulators.  With lane maps that put both lanes of a column
// into the same half warp the two depth parities of a column must sit in different banks.
__device__ __forceinline__ int d_accv_slot(int tau, int col) {
#if RF_STICK_LANEMAP != 0
    return (tau >> 1) * (2 * kStickCols) + 2 * col + (tau & 1);
#else
    return tau * kStickCols + col;
#endif
}

// Walk the columns of one task, two depths per column and iteration.  kChecked = false: every step of every
// active lane is known to be in bounds, on one side of column 0 and to need no multiplicity handling (both ends of
// each column were tested; the conditions are convex along it).
template <int K, bool kChecked, bool kFlags>
__device__ __forceinline__ uint32_t d_task_run(const StickConsts& c, const StickTask& t, const PlaneS& pl, const int nIter, const float4* __restrict__ slices,
                                               const int imgStride, const int* __restrict__ rimTab, float2* accV, float* accW, const int col) {
    const float4* sl = slices + (size_t)pl.img * imgStride;
    const float weight = pl.weight;
    uint32_t touched = 0;          // bit b <-> depths 2b, 2b + 1 of the stick hold something
    if (!kChecked && t.tauLo <= t.tauHi)       // every step of the lane's walk is taken: the bits are known up front
        touched = ((2u << ((t.tauHi - t.tauLo) >> 1)) - 1u) << (t.tauLo >> 1);
    for (int s = 0; s < nIter; ++s) {
        const int tau = t.tauLo + 2 * s;
        const StickStep q = d_step<K, kChecked>(c, t, pl.nd, tau);
        const bool ok = q.inb && tau <= t.tauHi;
        bool anySlow = false;
        if (kChecked) anySlow = __any_sync(0xffffffffu, ok && q.special);
        if (ok) {
            const float da0 = q.ar - __int2float_rn(q.jw), db0 = q.br - __int2float_rn(q.iw);
            const float h2s = q.h * q.h * c.iDelta;
            const int jc = q.ja + q.jw, ic = q.jb + q.iw;
            const unsigned off = (unsigned)((ic + c.Rp) * c.pitch + (jc + c.colOff));       // in bounds: >= 0
            float accRe = 0.f, accIm = 0.f, accWt = 0.f;
            if (kChecked && anySlow)
                d_stick_window2<K, true, kFlags>(sl, off, (unsigned)c.pitch, da0, db0, h2s, c.kI, c.sMax, c.tblAdj, jc, ic, rimTab, accRe, accIm, accWt);
            else
                d_stick_window2<K, false, kFlags>(sl, off, (unsigned)c.pitch, da0, db0, h2s, c.kI, c.sMax, c.tblAdj, jc, ic, rimTab, accRe, accIm, accWt);
            const int o = tau * kStickCols + col, ov = d_accv_slot(tau, col);
            float2 v = accV[ov];
            v.x += accRe;
            v.y = fmaf(q.sg, accIm, v.y);
            accV[ov] = v;
            accW[o] = fmaf(weight, accWt, accW[o]);
            if (kChecked) touched |= 1u << (tau >> 1);
        }
    }
    return touched;
}

// L2 prefetch of the slice patch a plane projects the stick onto: bounding box of the stick's slab crossing around the
// projection of its centre (cA, cB, cD), in the coordinates of the stored half plane, as TMA box prefetches (16 entries x 16
// rows each; everything here is warp-uniform, one lane issues).
constexpr int kPfBoxCols = 16, kPfBoxRows = 16;
__device__ __forceinline__ void d_prefetch_patch(const StickConsts& c, const PlaneS& pn, const CUtensorMap* map,
                                                 const float cA, const float cB, const float cD, const float rSlab, const int lane) {
    const float hc = cA * pn.na + cB * pn.nb + cD * pn.nd;            // height of the stick centre above the plane
    const float dc = cD - hc * pn.invNd;                              // depth at which the central column crosses it
    float al = cA * pn.e1a + cB * pn.e1b + dc * pn.e1d, be = cA * pn.e2a + cB * pn.e2b + dc * pn.e2d;
#if RF_L2_PREFETCH == 2
    // cheap form: ONE box centred on the crossing point (covers the patch of all but the steepest planes)
    if (al < 0.f) { al = -al; be = -be; }
    if (lane == 0) {
        const int x = max(__float2int_rn(al) - kPfBoxCols / 2 + c.colOff, 0), y = __float2int_rn(be) - kPfBoxRows / 2 + c.Rp;
        asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(4 * x), "r"(y), "r"(pn.img) : "memory");
    }
    return;
#endif
    // half extents: the 4 x 4 footprint, the shift of the crossing depth across the footprint, the slab thickness, the blob
    const float slope = 1.5f * (fabsf(pn.na) + fabsf(pn.nb)) * fabsf(pn.invNd) + rSlab * fabsf(pn.invNd);
    const float ea = 1.5f * (fabsf(pn.e1a) + fabsf(pn.e1b)) + slope * fabsf(pn.e1d) + c.rho + 1.0f;
    const float eb = 1.5f * (fabsf(pn.e2a) + fabsf(pn.e2b)) + slope * fabsf(pn.e2d) + c.rho + 1.0f;
    if (al < 0.f) { al = -al; be = -be; }                             // the mirrored half is read at (-alpha, -beta)
    const int j0 = max(__float2int_rd(al - ea), -c.colOff), j1 = min(__float2int_ru(al + ea), c.Rp);
    const int i0 = max(__float2int_rd(be - eb), -c.Rp), i1 = min(__float2int_ru(be + eb), c.Rp);
    if (lane == 0) {
        for (int y = i0 + c.Rp; y <= i1 + c.Rp; y += kPfBoxRows)
            for (int x = j0 + c.colOff; x <= j1 + c.colOff + 1; x += kPfBoxCols)
                asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(4 * x), "r"(y), "r"(pn.img)
                             : "memory");
    }
}

#ifdef RF_STICK_MAXREG
#define RF_STICK_BOUNDS __maxnreg__(RF_STICK_MAXREG)
#else
#define RF_STICK_BOUNDS __launch_bounds__(kStickThreads, 1)
#endif
template <int K, int CLS, bool kFlags>
__global__ void RF_STICK_BOUNDS k_gather_sticks(const __grid_constant__ StickLaunch L) {
    const StickArgs& a = L.a;
    const Geometry& geo = a.geo;
    __shared__ __align__(16) float tbl[kBlobTable];   // static: its shared address is a compile-time constant
    __shared__ __align__(8) unsigned long long tblBar;    // mbarrier of the table's bulk copy
    __shared__ int sAdj;
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kAccN = kStickL * kStickCols;
    float2* accV = reinterpret_cast<float2*>(smem) + (size_t)warp * kAccN;
    float* accW = reinterpret_cast<float*>(smem + sizeof(float2) * kStickWarps * kAccN) + (size_t)warp * kAccN;

    // The blob table is the one tile-shaped operand every warp of the CTA reuses: one TMA bulk copy (cp.async.bulk,
    // 40 000 B, completion on an mbarrier) brings it in while the warps zero their accumulators.  (The image windows are
    // not staged: they are per-lane 4 x 4 windows at data-dependent positions, and re-reading a staged tile would cost the
    // same L1 wavefronts that bound this kernel.)
    static_assert((kBlobTable * sizeof(float)) % 16 == 0, "bulk copies move multiples of 16 bytes");
    const uint32_t barAddr = (uint32_t)__cvta_generic_to_shared(&tblBar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(barAddr));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t bytes = kBlobTable * sizeof(float);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barAddr), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((uint32_t)__cvta_generic_to_shared(tbl)), "l"(a.blobTable), "r"(bytes), "r"(barAddr)
                     : "memory");
    }
    for (int i = lane; i < kAccN; i += 32) {
        accV[i] = make_float2(0.f, 0.f);
        accW[i] = 0.f;
    }
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "TBL_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
        "@p bra TBL_DONE;\n\t"
        "bra TBL_WAIT;\n\t"
        "TBL_DONE:\n\t"
        "}" ::"r"(barAddr)
        : "memory");
    // Shared byte address of tbl[0] minus (bits(2^23) << 2), modulo 2^32.  Routed through shared memory so that
    // the compiler treats it as an opaque value: the lookup address is then one LEA.
    if (tid == 0) sAdj = (int)((uint32_t)__cvta_generic_to_shared(tbl) - (0x4B000000u << 2));
    __syncthreads();

    constexpr int cls = CLS;
    StickConsts c;
    c.Rp = geo.Rp; c.side = geo.side; c.pitch = geo.pitch; c.planeStride = geo.planeStride; c.colOff = geo.colOff;
    c.rho = geo.rho; c.iDelta = geo.iDelta; c.sMax = geo.sMax; c.kI = geo.s2 * geo.iDelta;
    c.rimIn2 = geo.rimIn2; c.rhoCol0 = geo.rho + 1e-2f;
    c.tblAdj = (uint32_t)(*(volatile int*)&sAdj);
    const int lo = geo.lo, hi = geo.hi;
    const float reach2 = geo.reach * geo.reach + 1.0f;
    const float rSlab = geo.r + 1e-3f;
    // two lanes per column: even / odd depth.  Which 8 lanes form a quarter warp matters: a 128-bit warp load is served
    // in four quarter-warp passes, each costing one L1 tag request per cache line its 8 lanes touch (measured 1.0 - 1.5).
#if RF_STICK_LANEMAP == 1      // quarter = 2 x 2 columns x both depths
    const int la = (lane & 1) | (((lane >> 3) & 1) << 1), lb = ((lane >> 1) & 1) | (((lane >> 4) & 1) << 1), par = (lane >> 2) & 1;
#elif RF_STICK_LANEMAP == 2    // quarter = 4 columns along a x both depths
    const int la = lane & 3, par = (lane >> 2) & 1, lb = lane >> 3;
#else                          // quarter = 4 x 2 columns of one depth parity
    const int la = lane & 3, lb = (lane >> 2) & 3, par = lane >> 4;
#endif
    static_assert(kStickA == 4 && kStickB == 4, "lane maps are written for 4 x 4 columns");
    const int col = la + kStickA * lb;
    const float laf = (float)la, lbf = (float)lb;
    const int offA = (cls == 0) ? lo : 0, offB = lo, offD = (cls == 0) ? 0 : lo;
    // culling: half extents of the stick's lattice box around its centre, in (a,b,d) order
    const float hA = 0.5f * (kStickA - 1), hB = 0.5f * (kStickB - 1), hD = 0.5f * (kStickL - 1);
    const float inLim = geo.inplane_reach + sqrtf(hA * hA + hB * hB + hD * hD) * sqrtf(1.0f / geo.s2) + 1.0f;
    const float inLim2 = inLim * inLim;
    const float pxPerVox = sqrtf(1.0f / geo.s2), rimInR = geo.rimIn2 > 0.f ? sqrtf(geo.rimIn2) : -1.f;
    const int imgStride = geo.planeStride;

    for (;;) {
        int u = 0;
        if (lane == 0) u = atomicAdd(a.counter, 1);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= a.nUnits) break;
        const int4 su = __ldg(reinterpret_cast<const int4*>(a.units) + u);
        const int A0c = su.x + offA, B0c = su.y + offB, T0c = su.z + offD;   // centred lattice coordinates of the origin

        // ---- the lane's column: which depths does the main gather own there? (ownership rules: host::main_owns)
        const int ca = A0c + la, cb = B0c + lb;
        int tMin, tMax;
        bool colOk;
        if (cls == 2) {            // a = x, b = y, d = z
            colOk = ca <= geo.xOwnMax && cb <= hi && (ca > 0 || cb <= geo.yHalf);
            tMin = lo; tMax = hi;
        } else if (cls == 1) {     // a = x, b = z, d = y
            colOk = ca <= geo.xOwnMax && cb <= hi;
            tMin = lo; tMax = (ca == 0) ? min(geo.yHalf, hi) : hi;
        } else {                   // a = y, b = z, d = x
            colOk = ca <= hi && cb <= hi;
            tMin = (ca <= geo.yHalf) ? 0 : 1; tMax = geo.xOwnMax;
        }
        {
            const float rem = reach2 - (float)ca * (float)ca - (float)cb * (float)cb;
            if (rem < 0.f) colOk = false;
            else {
                const int tr = (int)sqrtf(rem) + 1;
                tMin = max(tMin, -tr);
                tMax = min(tMax, tr);
            }
        }
        int tauMin = max(tMin - T0c, 0), tauMax = min(tMax - T0c, kStickL - 1);
        if (!colOk) tauMax = -1;
        uint32_t touched = 0;

        // ---- planes of this class: lane <-> plane culling, then the warp walks the hits in plane order
        const float cA = (float)A0c + hA, cB = (float)B0c + hB, cD = (float)T0c + hD;
        for (int kb = 0; kb < a.nPlanes; kb += 32) {
            const int kk = kb + lane;
            bool hit = false, hint = false;
            float alc = 0.f;
            if (kk < a.nPlanes) {
                const float* s = a.planesSoA + kk;
                constexpr int kMaxPlanes = kLaunchPlanes;
                const float na = __ldg(s + 6 * kMaxPlanes), nb = __ldg(s + 7 * kMaxPlanes), nd = __ldg(s + 8 * kMaxPlanes);
                const float hc = cA * na + cB * nb + cD * nd;
                const float supp = hA * fabsf(na) + hB * fabsf(nb) + hD * fabsf(nd);
                if (fabsf(hc) <= rSlab + supp + 1e-2f) {
                    const float e1d = __ldg(s + 2 * kMaxPlanes), e2d = __ldg(s + 5 * kMaxPlanes);
                    const float ac = cA * __ldg(s) + cB * __ldg(s + kMaxPlanes) + cD * e1d;
                    const float bc = cA * __ldg(s + 3 * kMaxPlanes) + cB * __ldg(s + 4 * kMaxPlanes) + cD * e2d;
                    hit = ac * ac + bc * bc <= inLim2;
                    // Plain-task hint, decided here for 32 planes at once: every voxel this stick can have inside the slab
                    // of the plane lies within F pixels (in the image) of the point where the stick's centre column crosses
                    // the plane.  If that disc is inside the all-valid radius and on one side of the strip around column 0,
                    // every step of every lane takes the unchecked path with the same mirror sign, and the per-lane test
                    // of both walk ends (two checked step evaluations per task) is skipped.
                    const float inv = __frcp_rn(nd), sh = hc * inv;
                    alc = ac - sh * e1d;
                    const float bec = bc - sh * e2d;
                    const float dd = (hA * fabsf(na) + hB * fabsf(nb) + rSlab) * fabsf(inv);
                    const float F = sqrtf(hA * hA + hB * hB + dd * dd) * pxPerVox + 0.5f;
                    const float lim = rimInR - F;
                    hint = hit && lim > 0.f && alc * alc + bec * bec <= lim * lim && fabsf(alc) > c.rhoCol0 + F;
                }
            }
            unsigned m = __ballot_sync(0xffffffffu, hit);
            const unsigned mHint = __ballot_sync(0xffffffffu, hint), mNeg = __ballot_sync(0xffffffffu, alc < 0.f);
            while (m) {
                const int k = kb + __ffs(m) - 1;
                m &= m - 1;
#ifdef RF_L2_PREFETCH
                // The patch of the NEXT plane that crosses this stick is requested into L2 now (one bulk-prefetch
                // instruction, lane <-> slice row): first touches of slice data otherwise cost a DRAM round trip in the middle
                // of a step, and a step waits for the slowest of its ~40 cache lines.
                if (m) d_prefetch_patch(c, L.ps[kb + __ffs(m) - 1], &L.sliceMap, cA, cB, cD, rSlab, lane);
#endif
                // ---- set up task k
                const PlaneS& pl = L.ps[k];
                const PlaneD& pd = L.pd[k];
                const double a0 = A0c * pd.e1[0] + B0c * pd.e1[1] + T0c * pd.e1[2];
                const double b0 = A0c * pd.e2[0] + B0c * pd.e2[1] + T0c * pd.e2[2];
                const double h0 = A0c * pd.n[0] + B0c * pd.n[1] + T0c * pd.n[2];
                const double ja = rint(a0), jb = rint(b0);
                StickTask t;
                t.k = k;
                t.ja0 = (int)ja;
                t.jb0 = (int)jb;
                // per-voxel FP32 arithmetic only sees offsets < kStickL from the (double precision) stick origin
                t.arL = fmaf(lbf, pl.e1b, fmaf(laf, pl.e1a, (float)(a0 - ja)));
                t.brL = fmaf(lbf, pl.e2b, fmaf(laf, pl.e2a, (float)(b0 - jb)));
                t.hL = fmaf(lbf, pl.nb, fmaf(laf, pl.na, (float)h0));
                t.e1d = pl.e1d;
                t.e2d = pl.e2d;
                t.sg = 1.f;
                // segment of the column inside the slab |h| <= r:  h(tau) = hL + tau * nd
                const float c0 = -t.hL * pl.invNd, hw = rSlab * fabsf(pl.invNd);
                const int colLo = max(__float2int_ru(c0 - hw), tauMin);
                t.tauHi = min(__float2int_rd(c0 + hw), tauMax);
                t.tauLo = colLo + par;
                const int nIter = __reduce_max_sync(0xffffffffu, max((t.tauHi - colLo + 2) >> 1, 0));
                if (nIter == 0) continue;
                // both ends of the lane's walk in bounds and free of special pixels -> unchecked loop
                bool allPlain = true;
                float aFirst = ((mNeg >> (k - kb)) & 1u) ? -1.f : 1.f;
                if (!((mHint >> (k - kb)) & 1u)) {         // warp-uniform: no hint from the culling pass, test the lane's walk
                    bool plain = true;
                    aFirst = 1.f;
                    if (t.tauLo <= t.tauHi) {
                        const int last = t.tauLo + ((t.tauHi - t.tauLo) & ~1);
                        const StickStep q0 = d_step<K, true>(c, t, pl.nd, t.tauLo), q1 = d_step<K, true>(c, t, pl.nd, last);
                        // "away from column 0" is the union of two half planes: both ends must lie in the SAME one (a steep
                        // column can enter the strip |alpha| <= rho between two ends that are outside it on opposite sides)
                        plain = q0.inb && q1.inb && !q0.special && !q1.special && (q0.sg == q1.sg);
                        aFirst = q0.sg;
                    }
                    allPlain = __all_sync(0xffffffffu, plain);
                }
                // which of a column's two lanes owns an accumulator depends on the parity of the column's first depth in
                // THIS plane, so consecutive tasks may touch the same shared address from different lanes: order them
                __syncwarp();
                if (allPlain) {
                    if (aFirst < 0.f) {        // the lane's whole walk lies in the mirrored half: flip it once
                        t.sg = -1.f;
                        t.ja0 = -t.ja0; t.jb0 = -t.jb0;
                        t.arL = -t.arL; t.brL = -t.brL;
                        t.e1d = -t.e1d; t.e2d = -t.e2d;
                    }
                    touched |= d_task_run<K, false, kFlags>(c, t, pl, nIter, a.slices, imgStride, a.rimTab, accV, accW, col);
                } else
                    touched |= d_task_run<K, true, kFlags>(c, t, pl, nIter, a.slices, imgStride, a.rimTab, accV, accW, col);
            }
        }

        // ---- write-out: one coalesced reduction per touched brick of the stick (blocked layout)
        touched = __reduce_or_sync(0xffffffffu, touched);
        if (touched) {
            __syncwarp();
            const int lx = lane & 3, ly = (lane >> 2) & 3, lz = lane >> 4;
            // brick extents in (a, b, tau) and the lane's offsets inside the brick
            constexpr int spanB = (cls == 2) ? 4 : 2, spanT = (cls == 2) ? 2 : 4;
            const int da = (cls == 0) ? ly : lx, db = (cls == 2) ? ly : lz, dt = (cls == 2) ? lz : (cls == 1 ? ly : lx);
            constexpr int nBa = kStickA / 4, nBb = kStickB / spanB, nBt = kStickL / spanT;
#pragma unroll 1
            for (int bt = 0; bt < nBt; ++bt) {
                const uint32_t rows = ((1u << (spanT / 2)) - 1u) << (bt * (spanT / 2));
                if (!(touched & rows)) continue;
#pragma unroll
                for (int bb = 0; bb < nBb; ++bb) {
#pragma unroll
                    for (int ba = 0; ba < nBa; ++ba) {
                        const int av = ba * 4 + da, bv = bb * spanB + db, tau = bt * spanT + dt;
                        const int o = tau * kStickCols + av + kStickA * bv, ov = d_accv_slot(tau, av + kStickA * bv);
                        const float2 v = accV[ov];
                        const float w = accW[o];
                        if (w != 0.f || v.x != 0.f || v.y != 0.f) {
                            int X, Y, Zc;   // stored-offset coordinates
                            if (cls == 2) { X = su.x + av; Y = su.y + bv; Zc = su.z + tau; }
                            else if (cls == 1) { X = su.x + av; Zc = su.y + bv; Y = su.z + tau; }
                            else { Y = su.x + av; Zc = su.y + bv; X = su.z + tau; }
                            const size_t tile = ((size_t)(Zc / kTileZ) * geo.ty + (Y / kTileY)) * geo.tx + (X / kTileX);
                            const size_t g = tile * kTileVox + d_tile_slot(X % kTileX, Y % kTileY, Zc % kTileZ);
                            d_red_add2(a.Vb + g, v.x, v.y);
                            d_red_add(a.Wb + g, w);
                            if (a.Wb2) d_red_add(a.Wb2 + g, w);
                            accV[ov] = make_float2(0.f, 0.f);
                            accW[o] = 0.f;
                        }
                    }
                }
            }
            __syncwarp();
        }
    }
}

// ================================================================== K2e'
struct Edge2Args {
    Geometry geo;
    const EdgeItem* items;       // sorted by target voxel
    const int32_t* groupStart;   // nGroups+1 offsets: items of one group share the target
    int nGroups;
    const PlaneD* planesD;       // natural (x,y,z) component order
    const int* planeImg;
    const ImgParams* img;
    int nPlanes;
    const float* blobTable;
    const float4* slices;        // format v2 (the first pixel of an entry is read)
    const float2* col0;
    const int* rimTab;           // offset by +Rp
    float2* Vb;
    float* Wb;
    float* Wb2;                  // may be nullptr
    double iDeltaD;
};

// One WARP per edge target voxel (it walks the lattice points aliased onto that voxel); the lanes split the planes of
// the launch (brute force, double precision positions) and the partial sums are combined with a fixed shuffle tree,
// so the result does not depend on scheduling.
__global__ void __launch_bounds__(128) k_edge2(const __grid_constant__ Edge2Args a) {
    const Geometry& geo = a.geo;
    const int grp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (grp >= a.nGroups) return;
    const double r2 = (double)geo.r * (double)geo.r, rho = geo.rho, s2 = geo.s2;
    const double lim = geo.inplane_reach;
    const int Rp = geo.Rp, side = geo.side, K = geo.K, pitch = geo.pitch;
    const size_t imgStride = (size_t)geo.planeStride;
    double accRe = 0, accIm = 0, accW = 0;
    const int i0 = a.groupStart[grp], i1 = a.groupStart[grp + 1];
    const int64_t store = a.items[i0].store;
    for (int it = i0; it < i1; ++it) {
        const EdgeItem e = a.items[it];
        const double ux = e.ux, uy = e.uy, uz = e.uz;
        for (int k = lane; k < a.nPlanes; k += 32) {
            const PlaneD& pl = a.planesD[k];
            double h = ux * pl.n[0] + uy * pl.n[1] + uz * pl.n[2];
            double h2 = h * h;
            if (h2 > r2) continue;
            double al = ux * pl.e1[0] + uy * pl.e1[1] + uz * pl.e1[2];
            double be = ux * pl.e2[0] + uy * pl.e2[1] + uz * pl.e2[2];
            if (fabs(al) > lim || fabs(be) > lim) continue;
            int jw = (int)ceil(al - rho), iw = (int)ceil(be - rho);
            const int img = a.planeImg[k];
            const double weight = a.img[img].weight;
            const float4* S = a.slices + (size_t)img * imgStride;
            const float2* C0 = a.col0 + (size_t)img * side;
            double wsum = 0;
            for (int ti = 0; ti < K; ++ti) {
                int ip = iw + ti;
                double db = be - ip;
                double rowd2 = h2 + s2 * db * db;
                if (rowd2 > r2) continue;
                if ((unsigned)(ip + Rp) >= (unsigned)side) continue;
                const int rt = __ldg(a.rimTab + ip);
                for (int tj = 0; tj < K; ++tj) {
                    int j = jw + tj;
                    double da = al - j;
                    double d2 = rowd2 + s2 * da * da;
                    if (d2 > r2) continue;
                    if (e.mode == 1 && j < 0) continue;                 // originals only
                    if ((unsigned)(j + Rp) >= (unsigned)side) continue;
                    int idx = (int)(d2 * a.iDeltaD + 0.5);              // RF.cpp:725
                    float w = __ldg(a.blobTable + idx);
                    float2 px;
                    float mult;
                    if (e.mode == 1 && j == 0) {
                        px = __ldg(C0 + (ip + Rp));
                        mult = ((rt & 0x3fff) - 1) >= 0 ? 1.f : 0.f;   // original (0, ip) valid?
                    } else {
                        // half-plane slices: columns j < -colOff are the conjugates of the stored (-ip, -j)
                        if (j >= -geo.colOff) {
                            px = __ldg(reinterpret_cast<const float2*>(S + (size_t)(ip + Rp) * pitch + (j + geo.colOff)));
                        } else {
                            px = __ldg(reinterpret_cast<const float2*>(S + (size_t)(-ip + Rp) * pitch + (-j + geo.colOff)));
                            px.y = -px.y;
                        }
                        mult = d_rim_mult(rt, j);
                    }
                    accRe += (double)w * px.x;
                    accIm += (double)w * px.y;
                    if (!(__float_as_uint(px.x) & 1u)) wsum += (double)w * mult;
                }
            }
            accW += weight * wsum;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        accRe += __shfl_xor_sync(0xffffffffu, accRe, o);
        accIm += __shfl_xor_sync(0xffffffffu, accIm, o);
        accW += __shfl_xor_sync(0xffffffffu, accW, o);
    }
    if (lane == 0 && (accW != 0 || accRe != 0 || accIm != 0)) {
        float2 v = a.Vb[store];
        v.x += (float)accRe;
        v.y += (float)accIm;
        a.Vb[store] = v;
        a.Wb[store] += (float)accW;
        if (a.Wb2) a.Wb2[store] += (float)accW;
    }
}

// ================================================================== K2r
struct DampedArgs {
    Geometry geo;
    const uint32_t* mask;        // per image (2R+1) x ceil((R+1)/32) words of flagged columns
    const float* damped;         // per image (2R+1) x (R+1): weight of a flagged pixel (only flagged entries are valid)
    const float* damped2;        // un-modulated weights (with D2), or nullptr
    int nImg;
    const int* imgPlane0;        // first plane of image i in planesD (-1: image skipped)
    int nSym;                    // planes per image
    const PlaneD* planesD;       // natural component order
    const float* blobTable;
    double iDeltaD;
    unsigned long long* D;       // blocked layout, 2^32 fixed point
    unsigned long long* D2;      // same for the un-modulated weights (--iter > 1), or nullptr
};

// grid (ceil(words/256), nImg).  Every thread reads one word of the flag mask (32 pixels); the warp then walks the
// flagged pixels one at a time (they are rare).  The (floor(2r)+1)^3 candidate lattice points of a pixel are tested 64 at
// a time (two per lane) and the ACCEPTED ones (about 45 %) are compacted through shared memory, so the weight / index /
// atomic part runs once per 32 accepted candidates instead of once per 32 candidates.  The pixel position and the window
// origin are double precision (the reference's decisions, RF.cpp:628-650); from the origin on, offsets are < 2r + 1 and
// the distance, the table index and the weight are single precision like everything in the main gather.  This is the
// reference's scatter (RF.cpp:628-792) restricted to W of those pixels.
__global__ void __launch_bounds__(256) k_damped_scatter(const __grid_constant__ DampedArgs a) {
    const Geometry& geo = a.geo;
    const int R = geo.R, Z = geo.Z;
    const int img = blockIdx.y;
    const int cols = R + 1, total = cols * (2 * R + 1);
    const int wordsPerRow = (cols + 31) / 32, nWords = wordsPerRow * (2 * R + 1);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int widx0 = blockIdx.x * blockDim.x + threadIdx.x;
    __shared__ uint32_t sList[8][64];      // per warp: packed offsets of the accepted candidates of one group of 64
    uint32_t word = 0;
    if (widx0 < nWords) word = __ldg(a.mask + (size_t)img * nWords + widx0);
    unsigned any = __ballot_sync(0xffffffffu, word != 0);
    if (!any) return;
    const int p0 = a.imgPlane0[img];
    if (p0 < 0) return;
    const double r = geo.r;
    const float r2F = geo.r * geo.r, iDeltaF = (float)a.iDeltaD;
    const double voxPerPix = (double)Z / (double)geo.P;
    const double sc = voxPerPix * voxPerPix;
    // candidate lattice points of a pixel: a cube of edge E = floor(2r)+1 from ceil(p - r)
    const int E = (int)floor(2.0 * r) + 1, E3 = E * E * E;
    const unsigned ltMask = (1u << lane) - 1u;
    // |u| <= Z/2 + r + 1, so one conditional add wraps an index into [0, Z)
    auto wrap1 = [Z](int x) { return x < 0 ? x + Z : (x >= Z ? x - Z : x); };
    while (any) {
      const int srcW = __ffs(any) - 1;
      any &= any - 1;
      uint32_t w = __shfl_sync(0xffffffffu, word, srcW);
      const int widx = (widx0 - lane) + srcW;
      const int row = widx / wordsPerRow, ip = row - R, jBase = 32 * (widx - row * wordsPerRow);
      while (w) {
        const int j = jBase + __ffs(w) - 1;
        w &= w - 1;
        const size_t e = (size_t)img * total + (size_t)row * cols + j;
        const float dd = __ldg(a.damped + e) * 4294967296.0f;
        const float dd2 = a.damped2 ? __ldg(a.damped2 + e) * 4294967296.0f : 0.f;
        for (int s = 0; s < a.nSym; ++s) {
            const PlaneD& pl = a.planesD[p0 + s];
            // position of the pixel in voxel units: e1/e2 carry pixel-per-voxel, so p = (j*e1 + ip*e2) * (Z/P)^2
            // (= Z * M * (fx, fy, 0), RF.cpp:628-633)
            const double px = (j * pl.e1[0] + ip * pl.e2[0]) * sc;
            const double py = (j * pl.e1[1] + ip * pl.e2[1]) * sc;
            const double pz = (j * pl.e1[2] + ip * pl.e2[2]) * sc;
            const int x0 = (int)ceil(px - r), y0 = (int)ceil(py - r), z0 = (int)ceil(pz - r);
            const float fx = (float)((double)x0 - px), fy = (float)((double)y0 - py), fz = (float)((double)z0 - pz);
            for (int g0 = 0; g0 < E3; g0 += 64) {
                bool acc[2];
                uint32_t pk[2];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int c = g0 + 32 * i + lane;
                    int cx, cy, cz;
                    if (E == 4) { cx = c & 3; cy = (c >> 2) & 3; cz = c >> 4; }
                    else { cx = c % E; cy = (c / E) % E; cz = c / (E * E); }
                    const float dx = fx + (float)cx, dy = fy + (float)cy, dz = fz + (float)cz;
                    acc[i] = c < E3 && dx * dx + dy * dy + dz * dz <= r2F;
                    pk[i] = (uint32_t)cx | ((uint32_t)cy << 8) | ((uint32_t)cz << 16);
                }
                const unsigned m0 = __ballot_sync(0xffffffffu, acc[0]), m1 = __ballot_sync(0xffffffffu, acc[1]);
                const int n0 = __popc(m0), n = n0 + __popc(m1);
                if (acc[0]) sList[warp][__popc(m0 & ltMask)] = pk[0];
                if (acc[1]) sList[warp][n0 + __popc(m1 & ltMask)] = pk[1];
                __syncwarp();
                for (int k = lane; k < n; k += 32) {
                    const uint32_t q3 = sList[warp][k];
                    const int cx = q3 & 255, cy = (q3 >> 8) & 255, cz = q3 >> 16;
                    const float dx = fx + (float)cx, dy = fy + (float)cy, dz = fz + (float)cz;
                    const float dist2 = dx * dx + dy * dy + dz * dz;
                    const int ti = (int)(dist2 * iDeltaF + 0.5f);                     // RF.cpp:725
                    const float tw = __ldg(a.blobTable + ti);
                    const unsigned long long q = (unsigned long long)__float2ll_rn(tw * dd);      // 2^32 fixed point
                    const unsigned long long q2 = (unsigned long long)__float2ll_rn(tw * dd2);
                    const int wx = wrap1(x0 + cx), wy = wrap1(y0 + cy), wz = wrap1(z0 + cz);
                    // original at u
                    if (wx <= Z / 2) {
                        const int oy = wy <= Z / 2 ? wy : wy - Z, oz = wz <= Z / 2 ? wz : wz - Z;
                        const int64_t o = d_blocked_index(geo, wx, oy, oz);
                        if (q) atomicAdd(a.D + o, q);
                        if (a.D2 && q2) atomicAdd(a.D2 + o, q2);
                    }
                    // Hermitian mirror at -u
                    const int sx = wx ? Z - wx : 0, sy = wy ? Z - wy : 0, sz = wz ? Z - wz : 0;
                    const int my = sy <= Z / 2 ? sy : sy - Z, mz = sz <= Z / 2 ? sz : sz - Z;
                    const bool mirr = wx > Z / 2;                                       // cond_mirr(-u)
                    if (sx <= Z / 2 && (mirr || (sx == 0 && my <= geo.yHalf))) {
                        const int64_t o = d_blocked_index(geo, sx, my, mz);
                        if (q) atomicAdd(a.D + o, q);
                        if (a.D2 && q2) atomicAdd(a.D2 + o, q2);
                    }
                }
                __syncwarp();
            }
        }
      }
    }
}

// W += damped-pixel weights (and clear them)
__global__ void __launch_bounds__(256) k_fold_damped(float* __restrict__ Wb, unsigned long long* __restrict__ D, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long q = D[i];
        if (q != 0) {
            Wb[i] += (float)((double)q * (1.0 / kFixedScale));
            D[i] = 0ull;
        }
    }
}

}  // namespace rfb200
