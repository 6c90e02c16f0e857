/* =============================================================================
 * recfourier_b200.h — C ABI of the B200-native direct Fourier reconstruction.
 *
 * This is the drop-in boundary for Xmipp's Fourier-reconstruction hot path.  It
 * replaces the 14 free C++ functions through which the reference host program
 * (libraries/reconstruction_adapt_cuda/reconstruct_fourier_gpu.cpp) drives its
 * CUDA module, declared in
 *   libraries/reconstruction_cuda/cuda_gpu_reconstruct_fourier.h:67-156
 * and the data struct that crosses it (RecFourierBufferData,
 *   libraries/reconstruction/reconstruct_fourier_buffer_data.h:41-201).
 *
 *   reference (file:line)                                    this ABI
 *   -------------------------------------------------------  ---------------------------
 *   createStreams :87 / allocateWrapper :75 /
 *   allocateTempVolumeGPU :100 / copyConstants :142 /
 *   copyBlobTable :126 / pinMemory :134                       rfb200_create
 *   processBufferGPU                            (:153-156)    rfb200_insert_batch[_device]
 *   waitForGPU                                  (:120)        rfb200_sync
 *   copyTempVolumes :113 (+ CPU mirrorAndCrop,
 *     forceHermitianSymmetry, processWeights, FFTW c2r,
 *     crop + gridding correction on the CPU:
 *     reconstruct_fourier_gpu.cpp:683-767,879-932)            rfb200_finalize
 *   MPI_Reduce per row (parallel_adapt_cuda/
 *     mpi_reconstruct_fourier_gpu.cpp:250-268)                rfb200_reduce_nccl (or rfb200_reduce_p2p)
 *   releaseWrapper :81 / releaseTempVolumeGPU :106 /
 *   releaseBlobTable :132 / deleteStreams :93 / unpinMemory :136   rfb200_destroy
 *
 * Numerical contract: the result equals the reference CPU program
 * (ProgRecFourier, libraries/reconstruction/reconstruct_fourier.cpp) on the same
 * inputs within rel-L2 <= 1e-4 on the real-space map and FSC >= 0.999 to Nyquist.
 *
 * Plain C types only; the handle owns all device memory; the caller owns its
 * input buffers and may reuse them as soon as a call returns.  Functions return
 * RFB200_OK (0) or a negative error code and never call exit().  There is no CPU
 * fallback: without a CUDA device rfb200_create fails with RFB200_ERR_CUDA.
 * ============================================================================= */
#ifndef RECFOURIER_B200_H
#define RECFOURIER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RFB200_ABI_VERSION 1

enum {
    RFB200_OK = 0,
    RFB200_ERR_ARG = -1,         /* invalid argument / unsupported parameter value */
    RFB200_ERR_CUDA = -2,        /* CUDA runtime / cuFFT failure (see rfb200_last_error) */
    RFB200_ERR_NCCL = -3,        /* NCCL not loadable or collective failed */
    RFB200_ERR_STATE = -4,       /* call not valid in the handle's current state */
    RFB200_ERR_UNSUPPORTED = -5  /* feature of the reference not implemented on this path yet */
};

typedef struct rfb200_handle_s* rfb200_handle;

/* Program parameters; field meanings and defaults are those of the reference CLI
 * (reconstruct_fourier.cpp:42-58, readParams :64-86). */
typedef struct {
    int32_t abi_version;      /* = RFB200_ABI_VERSION */
    int32_t img_size;         /* N: particles are N x N float32                              */
    double pad_proj;          /* --padding <proj=2>                                          */
    double pad_vol;           /* --padding <vol=2>                                           */
    double max_resolution;    /* --max_resolution <0.5> (digital frequency)                  */
    double blob_radius;       /* --blob <radius=1.9>                                         */
    double blob_alpha;        /* --blob <alpha=15>                                           */
    int32_t blob_order;       /* --blob <order=0> (0 or 2, as kaiser_Fourier_value)           */
    int32_t n_sym;            /* SL.trueSymsNo(): symmetry matrices WITHOUT the identity     */
    const double* sym_matrices; /* n_sym x 9 row-major 3x3 (SL.getMatrices R), may be NULL   */
    int32_t use_ctf;          /* --useCTF and the metadata carries CTF columns               */
    int32_t phase_flipped;    /* --phaseFlipped                                              */
    double sampling;          /* --sampling <Ts=1>                                           */
    double min_ctf;           /* --minCTF <0.01>                                             */
    int32_t use_weights;      /* --weight                                                    */
    int32_t n_iter_weight;    /* --iter <1>: 0 = no weight correction, >= 1 passes           */
    int32_t fast;             /* --fast: nearest-pixel insertion in single precision + one final 3-D blob convolution
                               * (reconstruct_fourier_gpu.cpp:71-72, 879-893; cuda_gpu_reconstruct_fourier.cpp:455-503).
                               * Images are then padded to N*pad_vol, n_iter_weight is ignored (the --fast programs
                               * have no --iter) and the accumulators are the (S+1)^3 temporary volume / weights.       */
    int32_t device;           /* CUDA device ordinal                                         */
    int32_t max_batch;        /* largest n passed to rfb200_insert_batch (0 = default 1024)  */
    int32_t reserved0;
} rfb200_config;

/* One metadata row (RF.cpp:362-381; CTF columns data/ctf.cpp:365-419,1172-1212).
 * Defaults when a column is absent: weight 1, kV 100, K 1, defocusV = defocusU, rest 0. */
typedef struct {
    double rot, tilt, psi;          /* angleRot, angleTilt, anglePsi (degrees)        */
    double shift_x, shift_y;        /* shiftX, shiftY (pixels; content moves by +shift) */
    double weight;                  /* weight (used only with use_weights)            */
    double kV, defocusU, defocusV, defocus_angle, Cs, Ca, espr, ispr, alpha;
    double DeltaF, DeltaR, Q0, K, envR0, envR1, envR2, phase_shift, vpp_radius;
} rfb200_particle;

/* Wall/device timing of the stages since create or the last reset (milliseconds,
 * CUDA events on the handle's stream). */
typedef struct {
    double h2d_ms, preprocess_ms, fft2d_ms, slice_ms, gather_ms, edge_ms, finalize_ms, reduce_ms;
    int64_t images, planes, gather_launches, kernel_launches;
} rfb200_timings;

/* Geometry of the accumulators held by the handle. */
typedef struct {
    int32_t N, P, Z, X;          /* image, padded image, padded volume, Z/2+1           */
    int32_t tiles_x, tiles_y, tiles_z, tile;  /* blocked layout: tile^3 voxels per tile  */
    int64_t n_blocked;           /* entries of the blocked V (float2) and W (float)      */
    int32_t chunk_images;        /* images per gather launch                             */
    int32_t n_tiles_active, n_edge_items;
} rfb200_info;

int rfb200_create(const rfb200_config* cfg, rfb200_handle* out);
void rfb200_destroy(rfb200_handle h);
const char* rfb200_last_error(rfb200_handle h);  /* h may be NULL: error of the last failed create */
int rfb200_get_info(rfb200_handle h, rfb200_info* info);

/* Insert n particles.  images: n*N*N float32 in HOST memory (pageable or pinned).
 * Returns after the inputs have been consumed (the caller may reuse both buffers);
 * kernels may still be running — see rfb200_sync. */
int rfb200_insert_batch(rfb200_handle h, const float* images, const rfb200_particle* meta, int32_t n);
/* Same, with the images already resident in DEVICE memory of cfg.device. */
int rfb200_insert_batch_device(rfb200_handle h, const float* d_images, const rfb200_particle* meta, int32_t n);

int rfb200_sync(rfb200_handle h);
int rfb200_reset(rfb200_handle h);   /* zero the accumulators and the timings */

/* Multi-GPU: one process per GPU, each with its own handle over a particle subset.
 * rfb200_nccl_unique_id fills a 128-byte ncclUniqueId (rank 0 calls it, the host program
 * distributes it); rfb200_nccl_init joins the communicator; rfb200_reduce_nccl sums the
 * partial V and W of all ranks onto `root` over NVLink (one ncclReduce each). */
int rfb200_nccl_unique_id(void* id128);
int rfb200_nccl_init(rfb200_handle h, const void* id128, int32_t n_ranks, int32_t rank);
int rfb200_reduce_nccl(rfb200_handle h, int32_t root);
/* Peer-memory alternative to rfb200_reduce_nccl for the ranks of ONE node (NVLink / NVSwitch): every rank exports the
 * CUDA IPC handles of its accumulators (rfb200_ipc_export fills RFB200_IPC_BYTES bytes), the host program hands every
 * rank's blob to every other rank (rfb200_ipc_import, once per peer), and rfb200_reduce_p2p then runs ONE kernel per
 * rank that reads its 1/n_ranks slice of V and W from the memory of all ranks, adds the copies in rank order and stores
 * the sums into the root's accumulators - a reduce-scatter and a gather to the root in one pass, every link carrying
 * 1/n_ranks of the data.  The communicator of rfb200_nccl_init is still required: two 4-byte all-reduces order the
 * ranks before and after the kernel.  Same call discipline as rfb200_reduce_nccl (every rank calls it, same root);
 * the accumulators of the other ranks are left unchanged.  Replaces the MPI_Reduce per row of
 * mpi_reconstruct_fourier_gpu.cpp:250-268 like rfb200_reduce_nccl does. */
#define RFB200_IPC_BYTES 256
int rfb200_ipc_export(rfb200_handle h, void* out_blob);
int rfb200_ipc_import(rfb200_handle h, int32_t rank, const void* peer_blob);
int rfb200_reduce_p2p(rfb200_handle h, int32_t root);
/* The same reduce without NCCL, for host programs that can make their ranks wait for each other themselves (MPI_Barrier, a
 * file rendezvous): rfb200_set_ranks instead of rfb200_nccl_init, then per reduce
 *     rfb200_reduce_p2p_prepare(h);  BARRIER;  rfb200_reduce_p2p_run(h, root);  BARRIER;
 * prepare returns when this rank's accumulators are final, run when its slice of the sums has been stored in the root's
 * memory.  Nothing of NCCL is loaded or set up (its connection set-up costs 0.1 - 1.5 s, more than a short run). */
int rfb200_set_ranks(rfb200_handle h, int32_t n_ranks, int32_t rank);
int rfb200_reduce_p2p_prepare(rfb200_handle h);
int rfb200_reduce_p2p_run(rfb200_handle h, int32_t root);
/* Unmaps the peers' accumulators (after the last rfb200_reduce_p2p).  A process must not destroy a handle whose
 * accumulators another process still has mapped: every rank calls this, the host program makes the ranks wait for
 * each other, then the handles may be destroyed in any order. */
int rfb200_ipc_release(rfb200_handle h);
/* Raw device pointers of the blocked accumulators (for a host program that wants to run
 * its own collective on them): V = n_blocked float2, W = n_blocked float. */
int rfb200_accumulator_ptrs(rfb200_handle h, void** d_V, void** d_W, int64_t* n_blocked);

/* Copy the accumulators to the host in the reference layout [z][y][x], x in [0,Z/2]:
 * V = Z*Z*X interleaved (re,im) float32, W = Z*Z*X float32.  On the x = 0 plane the values
 * are those AFTER forceWeightSymmetry / enforceHermitianSymmetry (RF.cpp:1188-1221). */
int rfb200_export_accumulators(rfb200_handle h, float* V, float* W);
/* With cfg.fast the call returns the temporary spaces instead (copyTempVolumes of the reference): V = (S+1)^3
 * interleaved (re,im), W = (S+1)^3, [z][y][x] with the origin at S/2, S + 1 = rfb200_info.tile. */

/* correctWeight + finishComputations (RF.cpp:1056-1180): weight normalisation, 3-D inverse
 * FFT, crop and gridding correction.  out: N*N*N float32 in HOST memory, [z][y][x].
 * The accumulators are left untouched, so more batches may follow. */
int rfb200_finalize(rfb200_handle h, float* out);
/* Optional: pay the one-off set-up costs of the END of a run while particles are still being inserted.  Creates the 3-D
 * inverse-FFT plan and the finalisation buffers and, when a communicator is attached, runs a 4-byte collective on a
 * private stream so that NCCL's connection set-up (about a second) does not land on the first rfb200_reduce_nccl.  The
 * only entry point that may run on a second thread while another thread inserts into the same handle; it must have
 * returned before rfb200_reduce_nccl / rfb200_finalize are called, and every rank must call it (or none). */
int rfb200_warmup(rfb200_handle h);

/* Half-set support for --prepare_fsc (RF.cpp:991-1053): push saves the current accumulators aside and zeroes
 * them (so the next particles form an independent half set); merge adds the saved half back. */
int rfb200_halfset_push(rfb200_handle h);
int rfb200_halfset_merge(rfb200_handle h);

int rfb200_get_timings(rfb200_handle h, rfb200_timings* t);

/* Device-side stopwatch on the handle's compute stream (CUDA events): start records an event
 * behind everything queued so far, stop records a second one, waits for it and returns the
 * elapsed milliseconds.  Work issued through this handle in between (including the H2D copies
 * it depends on) is what gets measured. */
int rfb200_timer_start(rfb200_handle h);
int rfb200_timer_stop(rfb200_handle h, double* elapsed_ms);

/* Sum of the weight accumulator (FP64 reduction on the device, 8-byte read-back): a cheap
 * per-step result that forces completion of everything inserted so far. */
int rfb200_weight_sum(rfb200_handle h, double* sum);
/* Split form for a pipelined host program: _begin enqueues the reduction behind everything inserted so far and
 * returns at once, _end waits for it and returns the value.  A host loop "insert(k+1); end(k); begin(k+1)" overlaps
 * the PCIe transfer of batch k+1 with the kernels of batch k.  One reduction may be outstanding per handle. */
int rfb200_weight_sum_begin(rfb200_handle h);
int rfb200_weight_sum_end(rfb200_handle h, double* sum);

/* Number of CUDA devices visible to the process (for host programs that start one process per GPU). */
int rfb200_device_count(int32_t* n);
/* Measured FP32 SIMT peak of a device: a register-resident FFMA chain kernel (8 independent chains per thread, 2048
 * threads per SM) timed with CUDA events; *tflops = 2 * FMAs / time.  The denominator of the FP32 roofline fraction the
 * benchmark reports, instead of the nominal SMs x 128 lanes x 2 x clock. */
int rfb200_measure_fp32_peak(int32_t device, double* tflops);

/* Page-locked host memory for the image batches handed to rfb200_insert_batch: transfers from it run at full PCIe
 * speed and asynchronously (cudaHostAlloc / cudaFreeHost; replaces pinMemory / unpinMemory of
 * cuda_gpu_reconstruct_fourier.h:134-136).  Not tied to a handle. */
int rfb200_host_alloc(void** ptr, size_t bytes);
int rfb200_host_free(void* ptr);

/* The CUDA streams the handle launches on (cudaStream_t), for host programs that want to
 * order their own work against it. */
int rfb200_get_streams(rfb200_handle h, void** compute_stream, void** copy_stream);

/* Diagnostics used by the parity tests: intermediate products of one particle. */
/* Full-plane slice of image `idx` of the last chunk: (2*Rp+1)^2 float4 (re, im, m, 0). */
int rfb200_debug_slice_dims(rfb200_handle h, int32_t* side, int32_t* apron_radius);
int rfb200_debug_get_slice(rfb200_handle h, int32_t idx, float* out4);
/* cfg.fast only: the Pv x Pv x (Pv/2+1) transform (interleaved re, im; Pv = rfb200_info.Z) that finalisation hands to
 * the inverse FFT, i.e. the temporary spaces after mirrorAndCrop, applyBlob, forceHermitianSymmetry, processWeights and
 * convertToExpectedSpace (reconstruct_fourier_gpu.cpp:879-893). */
int rfb200_debug_fast_fourier(rfb200_handle h, float* out);

/* ---------------------------------------------------------------------------------------------------------------
 * Central-slice projector: the producer of the particle sets this path consumes (xmipp_phantom_project,
 * reconstruction/project.cpp:992).  Replaces FourierProjector (data/fourier_projection.h:91-175:
 * FourierProjector(V, paddFactor, maxFreq, BSplineDegree), project(rot, tilt, psi, ctf), projection()) and its CUDA twin
 * CudaFourierProjector (reconstruction_cuda/cuda_fourier_projection.h:183-216).
 *   volume   N^3 float32 in HOST memory, [z][y][x], centred like setXmippOrigin (voxel N/2 is the origin)
 *   degree   0 NEAREST, 1 LINEAR, 3 BSPLINE3 (xmipp_transformation)
 *   angles   n x (rot, tilt, psi) in degrees
 *   ctf      optional n x N x (N/2+1) float32 multipliers of the half-plane transform (the `ctf` image of project()), or NULL
 *   images   n x N x N float32
 * Same status codes as above; no CPU fallback. */
typedef struct rfb200_projector_s* rfb200_projector;
int rfb200_projector_create(const float* volume, int32_t N, double padding, double max_freq, int32_t degree, int32_t device,
                            rfb200_projector* out);
int rfb200_projector_project(rfb200_projector p, const double* angles, const float* ctf, int32_t n, float* images);
/* same with `d_ctf` (or NULL) and `d_images` in DEVICE memory of the projector's device */
int rfb200_projector_project_device(rfb200_projector p, const double* angles, const float* d_ctf, int32_t n, float* d_images);
const char* rfb200_projector_last_error(rfb200_projector p);
void rfb200_projector_destroy(rfb200_projector p);

#ifdef __cplusplus
}
#endif
#endif /* RECFOURIER_B200_H */
