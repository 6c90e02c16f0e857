/* oracle/oracle_abi.h — C ABI of the CPU oracle (TEST INFRASTRUCTURE ONLY).
 * Mirrors the product ABI in include/recfourier_b200.h but in double precision. */
#ifndef ORACLE_ABI_H
#define ORACLE_ABI_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int32_t img_size;        /* N (images are N x N float32) */
    int32_t n_sym;           /* symmetry matrices WITHOUT the identity (SL.trueSymsNo()) */
    double pad_proj, pad_vol;    /* --padding */
    double max_resolution;       /* --max_resolution (digital frequency, Nyquist = 0.5) */
    double blob_radius, blob_alpha;
    int32_t blob_order;
    int32_t use_ctf;             /* --useCTF and CTF columns present */
    double sampling;             /* --sampling */
    double min_ctf;              /* --minCTF */
    int32_t phase_flipped;       /* --phaseFlipped */
    int32_t use_weights;         /* --weight */
    int32_t n_iter_weight;       /* --iter (0 or 1) */
    int32_t reserved;
    const double* sym_matrices;  /* n_sym * 9, row-major 3x3 */
} orf_config;

typedef struct {
    double rot, tilt, psi, shift_x, shift_y, weight;
    /* CTF columns (MDL_CTF_*), data/ctf.cpp:365-419, 1172-1212 */
    double kV, defocusU, defocusV, defocus_angle, Cs, Ca, espr, ispr, alpha;
    double DeltaF, DeltaR, Q0, K, envR0, envR1, envR2, phase_shift, vpp_radius;
} orf_particle;

void* orf_create(const orf_config* cfg);
void orf_destroy(void* h);
void orf_dims(void* h, int* N, int* P, int* Z);
void orf_insert(void* h, const float* imgs, const orf_particle* meta, int n, int threads);
/* race-free parallel insertion (slabs of the volume owned by one task each): bit-identical to threads = 1 */
void orf_insert_slabs(void* h, const float* imgs, const orf_particle* meta, int n, int threads);
void orf_get_accumulators(void* h, double* V, double* W);
void orf_add_accumulators(void* h, const double* V, const double* W);
void orf_finalize(void* h, double* out);
/* tail of finishComputations on a given Z*Z*(Z/2+1) Fourier volume (interleaved re,im): c2r, CenterFFT, crop, correction */
void orf_finish_fourier(void* h, const double* Vri, double* out);

/* ---- FourierProjector (data/fourier_projection.cpp:91-333): degree 0 NEAREST, 1 LINEAR, 3 BSPLINE3;
 * vol N^3 float [z][y][x] about the Xmipp origin; ctf N*(N/2+1) doubles or NULL; out N*N doubles */
void* orf_projector_create(const float* vol, int N, double padding, double max_freq, int degree);
void orf_projector_destroy(void* h);
void orf_projector_project(void* h, double rot, double tilt, double psi, const double* ctf, double* out);

/* ---- --fast (recfourier_fast_oracle.cpp): nearest-pixel insertion + final blob convolution, single precision,
 * restating ProgRecFourierGPU with useFast (reconstruction_adapt_cuda/reconstruct_fourier_gpu.cpp and the device
 * functions of reconstruction_cuda/cuda_gpu_reconstruct_fourier.cpp:391-503, 655-760). */
/* one traverse space (reconstruct_fourier_projection_traverse_space.h:37-59) as plain data; same layout as refk_space in
 * oracle/ref_harness.cu */
typedef struct {
    int32_t minX, minY, minZ, maxX, maxY, maxZ;
    int32_t dir;                 /* 0 XY, 1 XZ, 2 YZ */
    int32_t projectionIndex;
    float maxDistanceSqr;
    float unitNormal[3], topOrigin[3], bottomOrigin[3];
    float transformInv[9];
    float weight;
} orf_space;

void* orf_fast_create(const orf_config* cfg);
/* use_fast = 0: host side of ProgRecFourierGPU WITHOUT --fast (traverse spaces with blob thickness, no final blob
 * convolution); its device arithmetic is not restated: the temporary spaces come from the compiled reference kernel */
void* orf_fast_create2(const orf_config* cfg, int use_fast);
/* the buffer handed to processBufferGPU for n images; returns the number of images kept (zero-weight ones are skipped) */
int orf_fast_export_buffer(void* h, const float* imgs, const orf_particle* meta, int n, float* FFTs, float* CTFs, float* mods,
                           orf_space* spaces);
void orf_fast_tables(void* h, float* blobTableSqrt, float* iDeltaSqrt, float* iw0);
void orf_fast_destroy(void* h);
void orf_fast_dims(void* h, int* S, int* sx, int* sy, int* Pv);
void orf_fast_insert(void* h, const float* imgs, const orf_particle* meta, int n);
void orf_fast_get_temp(void* h, float* Vri, float* W);     /* (S+1)^3 interleaved complex / real, [z][y][x] */
void orf_fast_set_temp(void* h, const float* Vri, const float* W);
void orf_fast_half_spaces(void* h, float* Vri, float* W);    /* mirrorAndCrop only: [S+1][S+1][S/2+1] */
void orf_fast_half_spaces(void* h, float* Vri, float* W);    /* mirrorAndCrop only: [S+1][S+1][S/2+1] */
void orf_fast_fourier(void* h, double* VFri);                /* Pv*Pv*(Pv/2+1) interleaved: input of the inverse transform */
void orf_fast_finalize(void* h, double* out);
void orf_tables(void* h, double* blobTableSqrt, double* fourierBlobTable, double* iDeltaSqrt, double* iDeltaFourier);
void orf_preprocess(void* h, const float* img, const orf_particle* p, double* F, double* Ainv);
void orf_apply_shift(void* h, const float* img, double sx, double sy, double* out);
void orf_ctf_weights(void* h, const orf_particle* p, int i, int j, double* wCTF, double* wMod);
void orf_bspline_coeffs_2d(double* a, int n);                        /* produceSplineCoefficients(BSPLINE3), in place */
double orf_bspline_interp_2d(const double* c, int n, double x, double y);   /* interpolatedElementBSpline2D, physical coords */
double orf_ctf_value(const orf_particle* p, double X, double Y);
double orf_ctf_argument(const orf_particle* p, double X, double Y);   /* getValueArgument */
double orf_ctf_K1(const orf_particle* p);                            /* pi * lambda (produceSideInfo) */
void orf_ctf_grid(const orf_particle* p, int n, double Tm, int what, double* out);   /* n x n values (0) / arguments (1) */
void orf_euler(double rot, double tilt, double psi, double* m9);
double orf_idx2digfreq(int idx, int size);
double orf_kaiser_value(double r, double a, double alpha, int m);
double orf_kaiser_fourier_value(double w, double a, double alpha, int m);
double orf_bessi0(double x);
double orf_bessj0(double x);
void orf_fft2_r2c(const double* in, int ny, int nx, double* out);
void orf_fft1(const double* in, int n, int sign, double* out);

#ifdef __cplusplus
}
#endif
#endif
