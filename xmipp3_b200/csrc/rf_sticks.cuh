// rf_sticks.cuh — second-generation insertion kernels (sm_100a).
//
//   K1b' k_make_slices2     half-plane FFT -> two float2 full-plane slices (B is A shifted by one pixel, so every
//                           candidate-window row is a run of 16-byte aligned pixel PAIRS in one of them) + the
//                           true weights of the CTF-damped pixels (RF.cpp:600-625 hoisted)
//   K2'  k_gather_sticks    voxel-centric gather, "column walk": a warp owns a stick of 8 x 4 columns that run
//                           along the axis dominating the plane normal; lane <-> column; at every step the lane's
//                           voxel is inside the blob slab of the plane, so (nearly) all lanes do useful work.
//                           Accumulators of the stick live in shared memory ([depth][lane]: conflict-free for any
//                           per-lane depth), ownership is exclusive -> no atomics, bit-reproducible.
//                           Replaces the scatter loop RF.cpp:586-792.
//   K2e' k_edge2            lattice points the stick gather does not own (as k_edge, new slice format)
//   K2r  k_damped_scatter   W of the (rare) pixels whose CTF is below --minCTF: their weight is |CTF| instead of 1
//                           (RF.cpp:616-622).  They carry a flag in the LSB of `re`; the gather counts every other
//                           valid pixel with the image weight (no per-pixel weight fetch in the hot loop) and skips
//                           the flagged ones, this kernel adds their true weights with 64-bit fixed-point atomics
//                           (order independent -> deterministic; only positive terms -> no cancellation)
//        k_fold_damped      W += damped weights, once before W is consumed
#pragma once
#include "rf_kernels.cuh"

namespace rfb200 {

__constant__ PlaneS c_planesS[kMaxPlanes];   // class-sorted planes of one launch, components permuted to (a,b,d)

#ifndef RF_STICK_WARPS
#define RF_STICK_WARPS 15
#endif
constexpr int kStickWarps = RF_STICK_WARPS;
constexpr int kStickThreads = kStickWarps * 32;
constexpr size_t kStickSmem = (size_t)kStickWarps * kStickL * 32 * (sizeof(float2) + sizeof(float));
static_assert(kStickL <= 32 && kStickL % 4 == 0, "touched-row mask is 32 bits; bricks are 4 deep along x and y");
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
                                                   const SliceParams& sp, const CtfConsts* ctf, float weight, int j, int ip) {
    float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ip < sp.iLo || ip > sp.iHi) return out;
    if (j > jmax[ip - sp.iLo]) return out;            // resolution cut-off, RF.cpp:597
    int row = ip < 0 ? ip + sp.P : ip;
    float2 F = __ldg(fft + (size_t)row * sp.Xh + j);
    float wc = 1.f, wm = 1.f;
    if (sp.useCtf) d_ctf_weights(*ctf, sp, j, ip, wc, wm);
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
    int pitch, planeStride;
    const float2* fft;
    float2* slices;       // per image: plane A then plane B
    float2* col0;         // per image `side` originals-only entries of column j = 0
    float* damped;        // per image (2R+1) x (R+1): true weight of a flagged pixel, -1 if not flagged; nullptr without CTF
    float* damped2;       // same shape: un-modulated weight of a flagged pixel (only for --iter > 1), else nullptr
    const ImgParams* ip;
    const CtfConsts* ctfs;
    const int* jmax;
};

// same thread mapping as k_make_slices: grid (ceil((R+1)/32), ceil((2R+1)/32), nImg), block (32, 8)
__global__ void __launch_bounds__(256, 4) k_make_slices2(const __grid_constant__ Slice2Args a) {
    const SliceParams& sp = a.sp;
    const int j = blockIdx.x * 32 + threadIdx.x;
    if (j > sp.R) return;
    const int img = blockIdx.z;
    const float2* f = a.fft + (size_t)img * sp.P * sp.Xh;
    const CtfConsts* ctf = sp.useCtf ? a.ctfs + img : nullptr;
    const float weight = a.ip[img].weight;
    float2* SA = a.slices + (size_t)img * 2 * a.planeStride;
    float2* SB = SA + a.planeStride;
    const size_t dOff = (size_t)img * (2 * sp.R + 1) * (sp.R + 1);
    const int rowBase = blockIdx.y * (8 * kSliceRowsPerThread) + threadIdx.y;
#pragma unroll 2
    for (int q = 0; q < kSliceRowsPerThread; ++q) {
        const int r = rowBase + 8 * q;
        if (r > 2 * sp.R) break;
        const int ipx = r - sp.R;
        float4 c = d_pixel_contrib2(f, a.jmax, sp, ctf, weight, j, ipx);
        const size_t o1 = (size_t)(ipx + sp.Rp) * a.pitch + (j + sp.Rp);
        bool flag = c.w != 0.f;
        float unmod = (c.z != 0.f || flag) ? weight : 0.f;       // weight of a valid pixel without the CTF modulator
        if (j > 0) {
            const size_t o2 = (size_t)(-ipx + sp.Rp) * a.pitch + (-j + sp.Rp);
            const float re = d_set_flag(c.x, flag);
            const float2 v = make_float2(re, c.y), vm = make_float2(re, -c.y);
            SA[o1] = v;  SB[o1 - 1] = v;
            SA[o2] = vm; SB[o2 - 1] = vm;
        } else {
            // column j = 0 holds original (0,ip) plus the mirror of original (0,-ip): the reference inserts this
            // column twice for x > 0 voxels.  The combined entry is flagged if either part is damped;
            // the damped-weight pass then supplies the weights of both parts.
            float4 m = d_pixel_contrib2(f, a.jmax, sp, ctf, weight, 0, -ipx);
            flag = flag || (m.w != 0.f);
            const float2 v = make_float2(d_set_flag(c.x + m.x, flag), c.y - m.y);
            SA[o1] = v; SB[o1 - 1] = v;
            a.col0[(size_t)img * sp.side + (ipx + sp.Rp)] = make_float2(d_set_flag(c.x, flag), c.y);
        }
        if (a.damped) a.damped[dOff + (size_t)r * (sp.R + 1) + j] = flag ? c.z : -1.f;
        if (a.damped2) a.damped2[dOff + (size_t)r * (sp.R + 1) + j] = flag ? unmod : -1.f;
    }
}

// ================================================================== K2'
struct StickArgs {
    Geometry geo;
    const StickUnit* units;      // sticks of this class, heaviest (closest to the origin) first
    int nUnits;
    int* counter;
    int cls;                     // 0: d = x, 1: d = y, 2: d = z
    int kBegin, kEnd;            // planes [kBegin, kEnd) of c_planesS belong to this class
    const float* blobTable;
    const PlaneD* planesDp;      // same order and permutation as c_planesS, double precision
    const float* planesSoA;      // 9 x kMaxPlanes floats, same order and permutation (culling phase, lane <-> plane)
    const float2* slices;        // slice format v2
    const int* rimTab;           // already offset by +Rp: index with the centred row
    float2* Vb;
    float* Wb;
    float* Wb2;                  // un-modulated weight sum for --iter > 1 with CTF (else nullptr)
};

// 16-byte read-only load of a pixel pair under a predicate (no branch: all loads of a window issue back to back)
__device__ __forceinline__ float4 d_ldg_pair_pred(const float2* p, bool pred) {
    float4 v;
    asm("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t@q ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
        : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
        : "l"(p), "r"((int)pred));
    return v;
}

// One step of one column: evaluate the K x K candidate window of the lane's voxel.  p points at the window origin
// (16-byte aligned pixel pair).  kSlow additionally weighs every candidate with its multiplicity (0 outside the
// resolution disc, 2 on column j = 0), looked up per window row.
template <int K, bool kSlow>
__device__ __forceinline__ void d_stick_window(const float2* __restrict__ p, const int pitch, const float (&dxs)[K], const float (&dys)[K],
                                               const float sMax, const uint32_t tblAdj, const int jc, const int ic,
                                               const int* __restrict__ rimTab, float& accRe, float& accIm, float& accW) {
    constexpr int NP = (K + 1) / 2;
    float4 px[K][NP];
#pragma unroll
    for (int ti = 0; ti < K; ++ti) {
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            const int t0 = 2 * q, t1 = 2 * q + 1;
            bool in = dys[ti] + dxs[t0] <= sMax;
            if (t1 < K) in = in || (dys[ti] + dxs[t1] <= sMax);
            px[ti][q] = d_ldg_pair_pred(p + (size_t)ti * pitch + t0, in);
        }
    }
#pragma unroll
    for (int ti = 0; ti < K; ++ti) {
        int rt = 0;
        if (kSlow) rt = __ldg(rimTab + ic + ti);
#pragma unroll
        for (int q = 0; q < NP; ++q) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int tj = 2 * q + e;
                if (tj < K) {
                    const float S = dys[ti] + dxs[tj];
                    if (S <= sMax) {
                        // (int)(d2*iDelta + 0.5) of RF.cpp:725: adding 2^23 rounds S to the nearest integer in the
                        // mantissa; (bits << 2) + tblAdj is then the shared-memory byte address of the table entry
                        const uint32_t addr = (__float_as_uint(S + 8388608.0f) << 2) + tblAdj;
                        float w;
                        asm("ld.shared.f32 %0, [%1];" : "=f"(w) : "r"(addr));
                        const float re = e ? px[ti][q].z : px[ti][q].x;
                        accRe = fmaf(w, re, accRe);
                        accIm = fmaf(w, e ? px[ti][q].w : px[ti][q].y, accIm);
                        if (!(__float_as_uint(re) & 1u)) {       // CTF-damped pixels get their weight from k_damped_scatter
                            if (kSlow) accW = fmaf(w, d_rim_mult(rt, jc + tj), accW);
                            else accW += w;
                        }
                    }
                }
            }
        }
    }
}

template <int K>
__global__ void __launch_bounds__(kStickThreads, 1) k_gather_sticks(const __grid_constant__ StickArgs a) {
    const Geometry& geo = a.geo;
    __shared__ float tbl[kBlobTable];          // static: its shared address is a compile-time constant
    __shared__ int sAdj;
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float2* accV = reinterpret_cast<float2*>(smem) + (size_t)warp * kStickL * 32;
    float* accW = reinterpret_cast<float*>(smem + sizeof(float2) * kStickWarps * kStickL * 32) + (size_t)warp * kStickL * 32;

    for (int i = tid; i < kBlobTable; i += kStickThreads) tbl[i] = __ldg(a.blobTable + i);
    for (int i = lane; i < kStickL * 32; i += 32) {
        accV[i] = make_float2(0.f, 0.f);
        accW[i] = 0.f;
    }
    // Shared byte address of tbl[0] minus (bits(2^23) << 2), modulo 2^32.  Routed through shared memory so that
    // the compiler treats it as an opaque value: the lookup address is then one LEA.
    if (tid == 0) sAdj = (int)((uint32_t)__cvta_generic_to_shared(tbl) - (0x4B000000u << 2));
    __syncthreads();
    const uint32_t tblAdj = (uint32_t)(*(volatile int*)&sAdj);

    const int cls = a.cls;
    const int lo = geo.lo, hi = geo.hi, Rp = geo.Rp, side = geo.side, pitch = geo.pitch;
    const float rho = geo.rho, iDelta = geo.iDelta, sMax = geo.sMax, kI = geo.s2 * geo.iDelta;
    const float reach2 = geo.reach * geo.reach + 1.0f, rimIn2 = geo.rimIn2;
    const float rSlab = geo.r + 1e-3f;
    const int la = lane & (kStickA - 1), lb = lane >> 3;
    const float laf = (float)la, lbf = (float)lb;
    const int offA = (cls == 0) ? lo : 0, offB = lo, offD = (cls == 0) ? 0 : lo;
    // culling: half extents of the stick's lattice box around its centre, in (a,b,d) order
    const float hA = 0.5f * (kStickA - 1), hB = 0.5f * (kStickB - 1), hD = 0.5f * (kStickL - 1);
    const float inLim = geo.inplane_reach + sqrtf(hA * hA + hB * hB + hD * hD) * sqrtf(1.0f / geo.s2) + 1.0f;
    const float inLim2 = inLim * inLim;
    const size_t imgStride = 2 * (size_t)geo.planeStride;

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
        for (int kb = a.kBegin; kb < a.kEnd; kb += 32) {
            const int kk = kb + lane;
            bool hit = false;
            if (kk < a.kEnd) {
                const float* s = a.planesSoA + kk;
                const float na = __ldg(s + 6 * kMaxPlanes), nb = __ldg(s + 7 * kMaxPlanes), nd = __ldg(s + 8 * kMaxPlanes);
                const float hc = cA * na + cB * nb + cD * nd;
                const float supp = hA * fabsf(na) + hB * fabsf(nb) + hD * fabsf(nd);
                if (fabsf(hc) <= rSlab + supp + 1e-2f) {
                    const float ac = cA * __ldg(s) + cB * __ldg(s + kMaxPlanes) + cD * __ldg(s + 2 * kMaxPlanes);
                    const float bc = cA * __ldg(s + 3 * kMaxPlanes) + cB * __ldg(s + 4 * kMaxPlanes) + cD * __ldg(s + 5 * kMaxPlanes);
                    hit = ac * ac + bc * bc <= inLim2;
                }
            }
            unsigned m = __ballot_sync(0xffffffffu, hit);
            while (m) {
                const int k = kb + __ffs(m) - 1;
                m &= m - 1;
                const PlaneS& pl = c_planesS[k];
                // segment of the column inside the slab |h| <= r:  h(tau) = hL + tau * nd
                const PlaneD pd = a.planesDp[k];
                const double a0 = A0c * pd.e1[0] + B0c * pd.e1[1] + T0c * pd.e1[2];
                const double b0 = A0c * pd.e2[0] + B0c * pd.e2[1] + T0c * pd.e2[2];
                const double h0 = A0c * pd.n[0] + B0c * pd.n[1] + T0c * pd.n[2];
                const double ja = rint(a0), jb = rint(b0);
                const int ja0 = (int)ja, jb0 = (int)jb;
                // per-voxel FP32 arithmetic only sees offsets < 32 from the (double precision) stick origin
                const float arL = fmaf(lbf, pl.e1b, fmaf(laf, pl.e1a, (float)(a0 - ja)));
                const float brL = fmaf(lbf, pl.e2b, fmaf(laf, pl.e2a, (float)(b0 - jb)));
                const float hL = fmaf(lbf, pl.nb, fmaf(laf, pl.na, (float)h0));
                const float c0 = -hL * pl.invNd, hw = rSlab * fabsf(pl.invNd);
                const int tauLo = max(__float2int_ru(c0 - hw), tauMin);
                const int tauHi = min(__float2int_rd(c0 + hw), tauMax);
                const int nSteps = __reduce_max_sync(0xffffffffu, max(tauHi - tauLo + 1, 0));
                if (nSteps == 0) continue;
                const float2* sl = a.slices + (size_t)pl.img * imgStride;
                const float e1d = pl.e1d, e2d = pl.e2d, nd = pl.nd, weight = pl.weight;
                const float jaf = (float)ja0, jbf = (float)jb0;
                for (int s = 0; s < nSteps; ++s) {
                    const int tau = tauLo + s;
                    const float ft = (float)tau;
                    const float ar = fmaf(ft, e1d, arL), br = fmaf(ft, e2d, brL), h = fmaf(ft, nd, hL);
                    const int jw = __float2int_ru(ar - rho), iw = __float2int_ru(br - rho);
                    const int jc = ja0 + jw, ic = jb0 + iw;                 // centred pixel of the window origin
                    const int jAbs = jc + Rp, iAbs = ic + Rp;
                    const bool ok = tau <= tauHi && (unsigned)jAbs <= (unsigned)(side - K) && (unsigned)iAbs <= (unsigned)(side - K);
                    // fast path needs: every candidate a valid pixel of multiplicity 1
                    const float Aabs = jaf + ar, Babs = jbf + br;
                    const bool special = (Aabs * Aabs + Babs * Babs > rimIn2) || ((unsigned)(-jc) <= (unsigned)(K - 1));
                    const bool anySlow = __any_sync(0xffffffffu, ok && special);
                    if (ok) {
                        float dxs[K], dys[K];
                        const float da0 = ar - __int2float_rn(jw), db0 = br - __int2float_rn(iw);
                        const float h2s = h * h * iDelta;
#pragma unroll
                        for (int q = 0; q < K; ++q) {
                            const float da = da0 - (float)q, db = db0 - (float)q;
                            dxs[q] = kI * da * da;
                            dys[q] = fmaf(kI * db, db, h2s);
                        }
                        // odd window origin: read plane B (B[j] = A[j+1]) at jAbs-1 so that pairs stay 16-byte aligned
                        const int odd = jAbs & 1;
                        const float2* p = sl + ((size_t)(iAbs * pitch + jAbs - odd) + (odd ? (size_t)geo.planeStride : 0));
                        float accRe = 0.f, accIm = 0.f, accWt = 0.f;
                        if (anySlow) d_stick_window<K, true>(p, pitch, dxs, dys, sMax, tblAdj, jc, ic, a.rimTab, accRe, accIm, accWt);
                        else d_stick_window<K, false>(p, pitch, dxs, dys, sMax, tblAdj, jc, ic, a.rimTab, accRe, accIm, accWt);
                        const int o = tau * 32 + lane;
                        float2 v = accV[o];
                        v.x += accRe;
                        v.y += accIm;
                        accV[o] = v;
                        accW[o] = fmaf(weight, accWt, accW[o]);
                        touched |= 1u << tau;
                    }
                }
            }
        }

        // ---- write-out: one coalesced read-modify-write per touched brick of the stick (blocked layout)
        touched = __reduce_or_sync(0xffffffffu, touched);
        if (touched) {
            __syncwarp();
            const int lx = lane & 3, ly = (lane >> 2) & 3, lz = lane >> 4;
            // brick extents in (a, b, tau) and the lane's offsets inside the brick
            int spanB, spanT, da, db, dt;
            if (cls == 2) { spanB = 4; spanT = 2; da = lx; db = ly; dt = lz; }
            else if (cls == 1) { spanB = 2; spanT = 4; da = lx; db = lz; dt = ly; }
            else { spanB = 2; spanT = 4; da = ly; db = lz; dt = lx; }
            const int nBa = kStickA / 4, nBb = kStickB / spanB, nBt = kStickL / spanT;
            for (int q = 0; q < nBa * nBb * nBt; ++q) {
                const int ba = q % nBa, bb = (q / nBa) % nBb, bt = q / (nBa * nBb);
                const uint32_t rows = ((1u << spanT) - 1u) << (bt * spanT);
                if (!(touched & rows)) continue;
                const int av = ba * 4 + da, bv = bb * spanB + db, tau = bt * spanT + dt;
                const int o = tau * 32 + av + kStickA * bv;
                const float2 v = accV[o];
                const float w = accW[o];
                if (w != 0.f || v.x != 0.f || v.y != 0.f) {
                    int X, Y, Zc;   // stored-offset coordinates
                    if (cls == 2) { X = su.x + av; Y = su.y + bv; Zc = su.z + tau; }
                    else if (cls == 1) { X = su.x + av; Zc = su.y + bv; Y = su.z + tau; }
                    else { Y = su.x + av; Zc = su.y + bv; X = su.z + tau; }
                    const size_t tile = ((size_t)(Zc / kTileZ) * geo.ty + (Y / kTileY)) * geo.tx + (X / kTileX);
                    const size_t g = tile * kTileVox + d_tile_slot(X % kTileX, Y % kTileY, Zc % kTileZ);
                    float2 gv = a.Vb[g];
                    gv.x += v.x;
                    gv.y += v.y;
                    a.Vb[g] = gv;
                    a.Wb[g] += w;
                    if (a.Wb2) a.Wb2[g] += w;
                    accV[o] = make_float2(0.f, 0.f);
                    accW[o] = 0.f;
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
    const float2* slices;        // format v2 (plane A is read)
    const float2* col0;
    const int* rimTab;           // offset by +Rp
    float2* Vb;
    float* Wb;
    float* Wb2;                  // may be nullptr
    double iDeltaD;
};

// One thread per edge TARGET voxel (it walks the lattice points aliased onto that voxel), brute force over the
// planes of the chunk, double precision positions.
__global__ void __launch_bounds__(128) k_edge2(const __grid_constant__ Edge2Args a) {
    const Geometry& geo = a.geo;
    int grp = blockIdx.x * blockDim.x + threadIdx.x;
    if (grp >= a.nGroups) return;
    const double r2 = (double)geo.r * (double)geo.r, rho = geo.rho, s2 = geo.s2;
    const double lim = geo.inplane_reach;
    const int Rp = geo.Rp, side = geo.side, K = geo.K, pitch = geo.pitch;
    const size_t imgStride = 2 * (size_t)geo.planeStride;
    double accRe = 0, accIm = 0, accW = 0;
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
            const double weight = a.img[img].weight;
            const float2* S = a.slices + (size_t)img * imgStride;
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
                        px = __ldg(S + (size_t)(ip + Rp) * pitch + (j + Rp));
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
    if (accW != 0 || accRe != 0 || accIm != 0) {
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
    const float* damped;         // per image (2R+1) x (R+1): weight of a flagged pixel, -1 if not flagged
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

// grid (ceil((R+1)*(2R+1)/256), nImg).  Every thread reads one entry; the warp then walks its flagged entries one
// at a time (they are rare), the 32 lanes sharing the (2*ceil(r)+1)^3 candidate lattice points of the pixel.
// This is the reference's scatter (RF.cpp:628-792) restricted to W of those pixels.
__global__ void __launch_bounds__(256) k_damped_scatter(const __grid_constant__ DampedArgs a) {
    const Geometry& geo = a.geo;
    const int R = geo.R, Z = geo.Z;
    const int img = blockIdx.y;
    const int cols = R + 1, total = cols * (2 * R + 1);
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    float d = -1.f, d2 = 0.f;
    if (idx < total) {
        d = __ldg(a.damped + (size_t)img * total + idx);
        if (a.damped2) d2 = fmaxf(__ldg(a.damped2 + (size_t)img * total + idx), 0.f);
    }
    unsigned m = __ballot_sync(0xffffffffu, d > 0.f || (d >= 0.f && d2 > 0.f));
    if (!m) return;
    const int p0 = a.imgPlane0[img];
    if (p0 < 0) return;
    const double r = geo.r, r2 = r * r;
    const double voxPerPix = (double)Z / (double)geo.P;
    const double sc = voxPerPix * voxPerPix;
    while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const float dd = __shfl_sync(0xffffffffu, d, src), dd2 = __shfl_sync(0xffffffffu, d2, src);
        const int pidx = (idx - lane) + src;
        const int row = pidx / cols, j = pidx - row * cols, ip = row - R;
        for (int s = 0; s < a.nSym; ++s) {
            const PlaneD& pl = a.planesD[p0 + s];
            // position of the pixel in voxel units: e1/e2 carry pixel-per-voxel, so p = (j*e1 + ip*e2) * (Z/P)^2
            // (= Z * M * (fx, fy, 0), RF.cpp:628-633)
            const double px = (j * pl.e1[0] + ip * pl.e2[0]) * sc;
            const double py = (j * pl.e1[1] + ip * pl.e2[1]) * sc;
            const double pz = (j * pl.e1[2] + ip * pl.e2[2]) * sc;
            const int x0 = (int)ceil(px - r), x1 = (int)floor(px + r);
            const int y0 = (int)ceil(py - r), y1 = (int)floor(py + r);
            const int z0 = (int)ceil(pz - r), z1 = (int)floor(pz + r);
            const int nx = x1 - x0 + 1, ny = y1 - y0 + 1, nz = z1 - z0 + 1;
            if (nx <= 0 || ny <= 0 || nz <= 0) continue;
            const int nc = nx * ny * nz;
            for (int c = lane; c < nc; c += 32) {
                const int ux = x0 + c % nx, uy = y0 + (c / nx) % ny, uz = z0 + c / (nx * ny);
                const double dx = ux - px, dy = uy - py, dz = uz - pz;
                const double dist2 = dx * dx + dy * dy + dz * dz;
                if (dist2 > r2) continue;
                const int ti = (int)(dist2 * a.iDeltaD + 0.5);                 // RF.cpp:725
                const double tw = (double)__ldg(a.blobTable + ti);
                const unsigned long long q = (unsigned long long)__double2ll_rn(tw * (double)dd * kFixedScale);
                const unsigned long long q2 = (unsigned long long)__double2ll_rn(tw * (double)dd2 * kFixedScale);
                // original at u
                {
                    const int sx = d_wrap(ux, Z);
                    if (sx <= Z / 2) {
                        int sy = d_wrap(uy, Z), sz = d_wrap(uz, Z);
                        int cy = sy <= Z / 2 ? sy : sy - Z, cz = sz <= Z / 2 ? sz : sz - Z;
                        const int64_t o = d_blocked_index(geo, sx, cy, cz);
                        if (q) atomicAdd(a.D + o, q);
                        if (a.D2 && q2) atomicAdd(a.D2 + o, q2);
                    }
                }
                // Hermitian mirror at -u
                {
                    const int sx = d_wrap(-ux, Z);
                    int sy = d_wrap(-uy, Z), sz = d_wrap(-uz, Z);
                    int cy = sy <= Z / 2 ? sy : sy - Z, cz = sz <= Z / 2 ? sz : sz - Z;
                    const bool mirr = d_wrap(ux, Z) > Z / 2;                        // cond_mirr(-u)
                    if (sx <= Z / 2 && (mirr || (sx == 0 && cy <= geo.yHalf))) {
                        const int64_t o = d_blocked_index(geo, sx, cy, cz);
                        if (q) atomicAdd(a.D + o, q);
                        if (a.D2 && q2) atomicAdd(a.D2 + o, q2);
                    }
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
