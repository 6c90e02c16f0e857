// rf_fft.cuh — fused preprocessing chain for power-of-two padded sizes (sm_100a):
//
//   K1r k_fft_rows<P>          raw N x N particle -> (integer shift, zero-pad, CenterFFT as one index map,
//                              RF.cpp:388-402) -> P-point FFT of the N non-zero rows, two real rows per complex
//                              transform -> intermediate T[img][kx][row] (kx = 0..P/2), 8 B x (P/2+1) x N per image
//   K1c k_fft_cols_slices<P>   P-point FFT of every needed column (only N of its P inputs are non-zero), then, on
//                              the transform still in shared memory, everything k_make_slices2 does: 1/P^2,
//                              resolution cut-off, CTF weights, flags, pixel-pair slice entries (RF.cpp:405-407,
//                              594-625)
//
// Compared with k_pad_images -> cuFFT R2C -> k_make_slices2 the padded image (4 B x P^2) and the half-plane
// transform (8 B x P x (P/2+1)) never touch HBM, and the zero rows are never transformed: 0.26 MB read + 0.53 MB
// intermediate (written and read once) + the slice write per particle instead of ~9 MB.  The transforms are our
// own shared-memory Stockham FFTs (radix 8/4/2, 8 points per thread and stage, one padded buffer per sequence);
// cuFFT stays the general path (any size, fractional shifts).
#pragma once
#include "rf_sticks.cuh"

namespace rfb200 {

__host__ __device__ constexpr int fft_phys(int i) { return i + (i >> 3); }   // padding: conflict-free strided stores
template <int P> constexpr int kFftBuf = P + P / 8 + 2;   // float2 elements per sequence buffer (+2: sequences land in different banks)

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ void bf2(float2& a, float2& b) {
    const float2 t = a;
    a = make_float2(t.x + b.x, t.y + b.y);
    b = make_float2(t.x - b.x, t.y - b.y);
}
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }   // * exp(-i pi/2)

// DFT of size R on v[0..R), natural order output in x[0..R)
template <int R>
__device__ __forceinline__ void dft_small(float2* v, float2* x);
template <>
__device__ __forceinline__ void dft_small<2>(float2* v, float2* x) {
    bf2(v[0], v[1]);
    x[0] = v[0]; x[1] = v[1];
}
template <>
__device__ __forceinline__ void dft_small<4>(float2* v, float2* x) {
    bf2(v[0], v[2]); bf2(v[1], v[3]);
    v[3] = mul_mi(v[3]);
    bf2(v[0], v[1]); bf2(v[2], v[3]);
    x[0] = v[0]; x[1] = v[2]; x[2] = v[1]; x[3] = v[3];
}
template <>
__device__ __forceinline__ void dft_small<8>(float2* v, float2* x) {
    const float s = 0.70710678118654752440f;
    bf2(v[0], v[4]); bf2(v[1], v[5]); bf2(v[2], v[6]); bf2(v[3], v[7]);
    v[5] = make_float2(s * (v[5].x + v[5].y), s * (v[5].y - v[5].x));      // * exp(-i pi/4)
    v[6] = mul_mi(v[6]);
    v[7] = make_float2(s * (v[7].y - v[7].x), -s * (v[7].x + v[7].y));     // * exp(-3i pi/4)
    bf2(v[0], v[2]); bf2(v[1], v[3]); bf2(v[4], v[6]); bf2(v[5], v[7]);
    v[3] = mul_mi(v[3]); v[7] = mul_mi(v[7]);
    bf2(v[0], v[1]); bf2(v[2], v[3]); bf2(v[4], v[5]); bf2(v[6], v[7]);
    x[0] = v[0]; x[1] = v[4]; x[2] = v[2]; x[3] = v[6]; x[4] = v[1]; x[5] = v[5]; x[6] = v[3]; x[7] = v[7];
}

// Barrier among the P/8 threads that transform ONE sequence (`seq` = index of the sequence inside the CTA).  Sequences are
// independent, so a CTA-wide barrier would only make them wait for each other: sequences of >= 64 threads use their own
// named barrier (ids 1..15; id 0 is __syncthreads), a sequence of <= 32 threads lives inside one warp.
template <int P>
__device__ __forceinline__ void fft_seq_barrier(int seq) {
    constexpr int TPS = P / 8;
    if (TPS >= 64) asm volatile("bar.sync %0, %1;" ::"r"(seq + 1), "n"(TPS) : "memory");
    else __syncwarp();
}

// One Stockham stage of radix R over a sequence of P points held in the padded shared buffer `buf`; the P/8 threads
// of the sequence (index t) each handle 8 points = 8/R butterflies.  W[k] = exp(-2 pi i k / P).  All threads of
// the sequence call this together (two sequence barriers).
template <int P, int R, int Ns>
__device__ __forceinline__ void fft_stage(float2* buf, const float2* __restrict__ W, int t, int seq) {
    constexpr int NB = 8 / R;               // butterflies per thread
    float2 v[8];
#pragma unroll
    for (int q = 0; q < NB; ++q) {
        const int j = t * NB + q, k = j % Ns;
        // twiddles W^(r k P/(Ns R)), r = 1..R-1: one table load, the powers by multiplication (the stages are bound by the
        // shared-memory instruction queue, the FP32 pipe has room; 6 complex products replace 6 loads for radix 8)
        float2 tw[8];
        if (Ns > 1) {
            tw[1] = W[(k * (P / (Ns * R))) & (P - 1)];
            if (R > 2) { tw[2] = cmul(tw[1], tw[1]); tw[3] = cmul(tw[2], tw[1]); }
            if (R > 4) { tw[4] = cmul(tw[2], tw[2]); tw[5] = cmul(tw[4], tw[1]); tw[6] = cmul(tw[3], tw[3]); tw[7] = cmul(tw[4], tw[3]); }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float2 a = buf[fft_phys(j + r * (P / R))];
            if (Ns > 1 && r > 0) a = cmul(a, tw[r]);
            v[q * R + r] = a;
        }
    }
    fft_seq_barrier<P>(seq);
#pragma unroll
    for (int q = 0; q < NB; ++q) {
        const int j = t * NB + q, k = j % Ns;
        float2 x[R];
        dft_small<R>(v + q * R, x);
        const int j0 = (j / Ns) * Ns * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) buf[fft_phys(j0 + r * Ns)] = x[r];
    }
    fft_seq_barrier<P>(seq);
}

// First stage (radix 8, Ns = 1) with the inputs already in registers: thread t of the sequence holds points t + r P/8,
// r = 0..7 - exactly what the kernels' load loops read, so the loaded values never make the round trip through shared
// memory (8 stores, 8 loads and one barrier less per thread; the transforms are bound by the shared-memory queue).
template <int P>
__device__ __forceinline__ void fft_first_stage(float2* v, float2* buf, int t, int seq) {
    float2 x[8];
    dft_small<8>(v, x);
#pragma unroll
    for (int r = 0; r < 8; ++r) buf[fft_phys(8 * t + r)] = x[r];
    fft_seq_barrier<P>(seq);
}

// forward FFT (e^{-i...}, unnormalised) of P points, natural order; v[r] = point t + r P/8 on entry, result in buf
template <int P>
__device__ __forceinline__ void fft_block(float2* v, float2* buf, const float2* __restrict__ W, int t, int seq);
template <>
__device__ __forceinline__ void fft_block<64>(float2* v, float2* buf, const float2* __restrict__ W, int t, int seq) {
    fft_first_stage<64>(v, buf, t, seq);
    fft_stage<64, 8, 8>(buf, W, t, seq);
}
template <>
__device__ __forceinline__ void fft_block<128>(float2* v, float2* buf, const float2* __restrict__ W, int t, int seq) {
    fft_first_stage<128>(v, buf, t, seq);
    fft_stage<128, 8, 8>(buf, W, t, seq);
    fft_stage<128, 2, 64>(buf, W, t, seq);
}
template <>
__device__ __forceinline__ void fft_block<256>(float2* v, float2* buf, const float2* __restrict__ W, int t, int seq) {
    fft_first_stage<256>(v, buf, t, seq);
    fft_stage<256, 8, 8>(buf, W, t, seq);
    fft_stage<256, 4, 64>(buf, W, t, seq);
}
template <>
__device__ __forceinline__ void fft_block<512>(float2* v, float2* buf, const float2* __restrict__ W, int t, int seq) {
    fft_first_stage<512>(v, buf, t, seq);
    fft_stage<512, 8, 8>(buf, W, t, seq);
    fft_stage<512, 8, 64>(buf, W, t, seq);
}
template <>
__device__ __forceinline__ void fft_block<1024>(float2* v, float2* buf, const float2* __restrict__ W, int t, int seq) {
    fft_first_stage<1024>(v, buf, t, seq);
    fft_stage<1024, 8, 8>(buf, W, t, seq);
    fft_stage<1024, 8, 64>(buf, W, t, seq);
    fft_stage<1024, 2, 512>(buf, W, t, seq);
}

#ifndef RF_FFT_SEQS
#define RF_FFT_SEQS 8
#endif
constexpr int kFftSeqs = RF_FFT_SEQS;   // sequences transformed side by side by one CTA (P/8 threads each)

// ================================================================== K1r
struct FftRowsArgs {
    const float* raw;            // nImg x N x N
    const ImgParams* ip;
    const float2* twiddle;       // P entries
    float2* T;                   // nImg x (P/2+1) x N: T[img][kx][row]
    int N;
};

// grid (ceil(N / 16), nImg), block kFftSeqs * P/8 threads: 16 image rows = 8 complex sequences
template <int P>
__global__ void __launch_bounds__(kFftSeqs* P / 8) k_fft_rows(const __grid_constant__ FftRowsArgs a) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    float2* W = reinterpret_cast<float2*>(smemRaw);
    float2* bufs = W + P;
    constexpr int TPS = P / 8, Xh = P / 2 + 1;
    const int N = a.N, img = blockIdx.y;
    const int tid = threadIdx.x, seq = tid / TPS, t = tid % TPS;
    for (int i = tid; i < P; i += kFftSeqs * TPS) W[i] = __ldg(a.twiddle + i);
    const ImgParams q = a.ip[img];
    const float* src = a.raw + (size_t)img * N * N;
    float2* buf = bufs + seq * kFftBuf<P>;
    // rows 2*seq and 2*seq+1 of this CTA's 16 (destination rows before padding; the shift moves the source)
    const int rowA = blockIdx.x * 2 * kFftSeqs + 2 * seq, rowB = rowA + 1;
    const float* ra = rowA < N ? src + (size_t)d_wrap(rowA + q.my, N) * N : nullptr;
    const float* rb = rowB < N ? src + (size_t)d_wrap(rowB + q.my, N) * N : nullptr;
    float2 v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int x = t + TPS * e;                        // padded column
        const int jj = (x + N / 2) & (P - 1);             // destination column before padding: x = (jj - N/2) mod P
        float2 z = make_float2(0.f, 0.f);
        if (jj < N) {
            const int sj = d_wrap(jj + q.mx, N);
            if (ra) z.x = __ldg(ra + sj);
            if (rb) z.y = __ldg(rb + sj);
        }
        v[e] = z;
    }
    __syncthreads();              // twiddle table
    fft_block<P>(v, buf, W, t, seq);
    __syncthreads();              // the untangling reads every sequence of the CTA
    // untangle the two real transforms and write T[kx][row], 16 consecutive rows per kx
    const int nOut = Xh * 2 * kFftSeqs;
    for (int o = tid; o < nOut; o += kFftSeqs * TPS) {
        const int r16 = o & (2 * kFftSeqs - 1), kx = o / (2 * kFftSeqs);
        const int row = blockIdx.x * 2 * kFftSeqs + r16;
        if (row >= N) continue;
        const float2* b = bufs + (r16 >> 1) * kFftBuf<P>;
        const float2 zk = b[fft_phys(kx)], zm = b[fft_phys((P - kx) & (P - 1))];
        float2 out;
        if (r16 & 1) out = make_float2(0.5f * (zk.y + zm.y), 0.5f * (zm.x - zk.x));      // B = (Z[k] - conj Z[P-k]) / 2i
        else out = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));             // A = (Z[k] + conj Z[P-k]) / 2
        a.T[((size_t)img * Xh + kx) * N + row] = out;
    }
}

// ================================================================== K1c
struct FftColsArgs {
    Slice2Args s;                // slice outputs, CTF, cut-off (s.fft is unused)
    const float2* twiddle;       // P entries
    const float2* T;             // nImg x (P/2+1) x N
    int N;
};

// value and weights of pixel (j, ip) from its transform value F (as d_pixel_contrib2)
// kCtf: 0 = the run has no CTF, 1 = every image of the launch takes the fixed-point phase path, 2 = general
template <int kCtf>
__device__ __forceinline__ float4 d_contrib_from_F(float2 F, bool valid, const SliceParams& sp, const CtfConsts* ctf, const CtfFloat& cf, float weight, int j, int ip) {
    float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!valid) return out;
    float wc = 1.f, wm = 1.f;
    if (kCtf == 1) d_ctf_weights<false>(*ctf, cf, sp, j, ip, wc, wm);
    if (kCtf == 2) d_ctf_weights<true>(*ctf, cf, sp, j, ip, wc, wm);
    const float s = weight * wm * wc * sp.invP2;
    out.x = F.x * s;
    out.y = F.y * s;
    out.z = weight * wm;
    out.w = (wm != 1.0f) ? 1.f : 0.f;
    return out;
}
__device__ __forceinline__ bool d_pixel_valid(const int* __restrict__ jmax, const SliceParams& sp, int j, int ip) {
    return ip >= sp.iLo && ip <= sp.iHi && j <= jmax[ip - sp.iLo];
}

// Columns per CTA of K1c: 8 for P <= 256, 4 for P = 1024 (thread limit) and for P = 512, where a CTA of 4 x 64 threads at
// <= 128 registers lets TWO CTAs share an SM: the FFT phase of one (shared-memory queue bound) overlaps the slice phase
// of the other (ALU / latency bound).  Measured: 7.77 ms per 4096 particles with 8 columns and one CTA per SM, 7.41 ms
// with 4 columns and two (profiles/r2b_ab_k1c.txt).
#ifndef RF_K1C_COLS
#define RF_K1C_COLS 4          // columns per CTA at P = 512
#endif
template <int P> constexpr int kColsPerCta = (P >= 1024) ? 4 : (P == 512 ? RF_K1C_COLS : 8);

// grid (ceil((R+2) / NC), nImg), block (NC + halo) * P/8 threads: columns kx = NC*blockIdx.x .. +NC-1.  A slice entry is
// a pixel PAIR, E(r, j-1) = (p(r,j-1), p(r,j)) (and, for the few columns next to j = 0 that the half-plane format keeps on
// the mirrored side, E(-r, -j) = (conj p(r,j), conj p(r,j-1))): inside a CTA the left neighbour comes by shuffle and a
// lane writes the whole 16-byte entry; at the CTA's first and last column the two halves of an entry belong to two CTAs
// and are written as 8-byte stores.
// Variant -DRF_K1C_HALO (round 1 / early round 2 default): the CTA also transforms the column on its left (a ninth
// sequence) and evaluates its pixels, so that every store is a whole entry.  Measured on B200 (profiles/r2b_ab_k1c.txt):
// 9.25 ms per 4096 particles with the halo, 8.44 ms without - the extra FFT and CTF evaluations (1/8 more work, and 576
// threads do not divide the 513 x 8 pixels of a CTA as well as 512 do) cost more than the split stores.
#ifdef RF_K1C_HALO
constexpr int kK1cHalo = 1;
#else
constexpr int kK1cHalo = 0;
#endif
#ifndef RF_K1C_CTAS
#define RF_K1C_CTAS 4          // CTAs per SM the register budget of K1c is sized for at P = 512 (2 with the general CTF path)
#endif
template <int P, int kCtf>
__global__ void __launch_bounds__((kColsPerCta<P> + kK1cHalo) * P / 8, (P == 512) ? (kCtf == 2 ? 2 : RF_K1C_CTAS) : 1) k_fft_cols_slices(const __grid_constant__ FftColsArgs a) {
    constexpr int NC = kColsPerCta<P>, NS = NC + kK1cHalo;
    extern __shared__ __align__(16) unsigned char smemRaw[];
    float2* W = reinterpret_cast<float2*>(smemRaw);
    float2* bufs = W + P;
    float2* sHalo = bufs + NS * kFftBuf<P>;          // slice value of the halo pixel of every row (P + 1 entries)
    int* sJmax = reinterpret_cast<int*>(sHalo + (P + 2));   // resolution cut-off per row: read by every pixel, staged once per CTA
    __shared__ CtfConsts sCtf;
    __shared__ CtfFloat sCtfF;
    constexpr int TPS = P / 8, Xh = P / 2 + 1, NT = NS * TPS;
    const SliceParams& sp = a.s.sp;
    const int N = a.N, img = blockIdx.y;
    const int tid = threadIdx.x, seq = tid / TPS, t = tid % TPS;
    for (int i = tid; i < P; i += NT) W[i] = __ldg(a.twiddle + i);
    for (int i = tid; i <= sp.iHi - sp.iLo; i += NT) sJmax[i] = __ldg(a.s.jmax + i);
    if (kCtf && tid < (int)(sizeof(CtfConsts) / 8))
        reinterpret_cast<double*>(&sCtf)[tid] = reinterpret_cast<const double*>(a.s.ctfs + img)[tid];
    if (kCtf && tid == NT - 1) d_ctf_prepare(a.s.ctfs[img], sp, sCtfF);
    float2* buf = bufs + seq * kFftBuf<P>;
    const int j0 = blockIdx.x * NC;                   // first own column; sequence 0 transforms column j0 - 1
    const int kx = j0 + seq - kK1cHalo;
    const float2* col = (kx >= 0 && kx <= sp.R) ? a.T + ((size_t)img * Xh + kx) * N : nullptr;
    float2 v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int y = t + TPS * e;                        // padded row
        const int ii = (y + N / 2) & (P - 1);             // destination row before padding
        float2 z = make_float2(0.f, 0.f);
        if (col && ii < N) z = __ldg(col + ii);
        v[e] = z;
    }
    __syncthreads();              // twiddle, cut-off and CTF tables
    fft_block<P>(v, buf, W, t, seq);
    if (kK1cHalo) __syncthreads();   // the halo pass reads sequence 0

    const CtfConsts* ctf = kCtf ? &sCtf : nullptr;
    const float weight = a.s.ip[img].weight;
    // ---- halo pass: the slice value (flag in the LSB of re) of pixel (j0 - 1, ip) for every row; its weights, mask bits
    // and column-0 extras belong to the CTA that owns the column
    const int nRows = 2 * sp.R + 1;
    if (kK1cHalo && j0 >= 2) {
        for (int r = tid; r < nRows; r += NT) {
            const int ipx = r - sp.R, jh = j0 - 1;
            const float2 F = bufs[fft_phys(ipx & (P - 1))];
            const float4 cc = d_contrib_from_F<kCtf>(F, d_pixel_valid(sJmax, sp, jh, ipx), sp, ctf, sCtfF, weight, jh, ipx);
            sHalo[r] = make_float2(d_set_flag(cc.x, cc.w != 0.f), cc.y);
        }
    }
    __syncthreads();

    // ---- the slice kernel's work on the transform in shared memory: thread <-> (row r = ip + R, column c), c fastest
    float4* S4 = a.s.slices + (size_t)img * a.s.planeStride;
    const size_t dOff = (size_t)img * (2 * sp.R + 1) * (sp.R + 1);
    const int wordsPerRow = (sp.R + 1 + 31) / 32;
    const int nElem = nRows * NC;
    const int nIter = (nElem + NT - 1) / NT;
#ifndef RF_K1C_UNROLL
#define RF_K1C_UNROLL 1
#endif
    constexpr int kUnroll = RF_K1C_UNROLL;
#pragma unroll kUnroll
    for (int it = 0; it < nIter; ++it) {
        const int o = tid + it * NT;
        const int c = o & (NC - 1), r = o / NC;
        const int j = j0 + c, ipx = r - sp.R;
        const bool inRow = o < nElem;
        const bool own = inRow && j <= sp.R;              // a pixel of the image
        const bool act = inRow && j <= sp.R + 1;          // column R + 1 (value 0) completes the entries E(r, R), E(-r, -R-1)
        float2 pv = make_float2(0.f, 0.f);       // the pixel's slice value (flag in the LSB of re)
        bool flag = false;
        float wDamped = 0.f, wUnmod = 0.f;
        if (own) {
            const float2* b = bufs + (c + kK1cHalo) * kFftBuf<P>;
            const float2 F = b[fft_phys(ipx & (P - 1))];
            float4 cc = d_contrib_from_F<kCtf>(F, d_pixel_valid(sJmax, sp, j, ipx), sp, ctf, sCtfF, weight, j, ipx);
            flag = cc.w != 0.f;
            wDamped = cc.z;
            wUnmod = (cc.z != 0.f || flag) ? weight : 0.f;
            if (j > 0) {
                pv = make_float2(d_set_flag(cc.x, flag), cc.y);
            } else {
                // column 0: original (0,ip) plus the mirror of original (0,-ip) (see k_make_slices2)
                const float2 Fm = b[fft_phys((-ipx) & (P - 1))];
                float4 m = d_contrib_from_F<kCtf>(Fm, d_pixel_valid(sJmax, sp, 0, -ipx), sp, ctf, sCtfF, weight, 0, -ipx);
                flag = flag || (m.w != 0.f);
                pv = make_float2(d_set_flag(cc.x + m.x, flag), cc.y - m.y);
                a.s.col0[(size_t)img * sp.side + (ipx + sp.Rp)] = make_float2(d_set_flag(cc.x, flag), cc.y);
            }
        }
        // left neighbour: inside the group of NC lanes by shuffle, for the first lane from the halo pass
        float2 prv;
        prv.x = __shfl_up_sync(0xffffffffu, pv.x, 1, NC);
        prv.y = __shfl_up_sync(0xffffffffu, pv.y, 1, NC);
        if (!act) continue;
        if (kK1cHalo && c == 0) prv = (j0 >= 2) ? sHalo[r] : make_float2(0.f, 0.f);
        if (j > 0) {
            const size_t o1 = (size_t)(ipx + sp.Rp) * a.s.pitch + (j + a.s.colOff);
            if (kK1cHalo || c > 0) {
                S4[o1 - 1] = make_float4(prv.x, prv.y, pv.x, pv.y);            // E(r, j-1)
            } else {
                // first column of the CTA: the left neighbour belongs to the previous CTA, which writes its half itself
                reinterpret_cast<float2*>(S4 + (o1 - 1))[1] = pv;              // second half of E(r, j-1)
            }
            if (!kK1cHalo && c == NC - 1 && j <= sp.R) reinterpret_cast<float2*>(S4 + o1)[0] = pv;   // first half of E(r, j): the next CTA adds the second
            if (j <= a.s.colOff) {                                         // half-plane format: the first colOff mirrored columns only
                const size_t o2 = (size_t)(-ipx + sp.Rp) * a.s.pitch + (-j + a.s.colOff);
                if (kK1cHalo || c > 0) S4[o2] = make_float4(pv.x, -pv.y, prv.x, -prv.y);         // E(-r, -j)
                else reinterpret_cast<float2*>(S4 + o2)[0] = make_float2(pv.x, -pv.y);
            }
            if (!kK1cHalo && c == NC - 1 && j + 1 <= a.s.colOff)          // second half of E(-r, -(j+1)), owned by the next CTA
                reinterpret_cast<float2*>(S4 + ((size_t)(-ipx + sp.Rp) * a.s.pitch + (-(j + 1) + a.s.colOff)))[1] = make_float2(pv.x, -pv.y);
        }
        if (flag) {
            if (a.s.damped) a.s.damped[dOff + (size_t)r * (sp.R + 1) + j] = wDamped;
            if (a.s.damped2) a.s.damped2[dOff + (size_t)r * (sp.R + 1) + j] = wUnmod;
            if (a.s.dampedMask)     // the mask is zeroed before the launch; a word spans several CTAs
                atomicOr(a.s.dampedMask + ((size_t)img * (2 * sp.R + 1) + r) * wordsPerRow + (j >> 5), 1u << (j & 31));
        }
    }
}

}  // namespace rfb200
