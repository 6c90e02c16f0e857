// prog_rec_fourier.h — host program of the B200 direct Fourier reconstruction.
//
// Mirrors the reference's program class for this path (same parameter names, meaning and defaults):
//   ProgRecFourier       libraries/reconstruction/reconstruct_fourier.{h,cpp}   (defineParams :36-61,
//                        readParams :64-86, show :89-121, run :124-181, setIO :1182-1186)
//   ProgRecFourierGPU    libraries/reconstruction_adapt_cuda/reconstruct_fourier_gpu.cpp:42-98
//                        (--device, --bufferSize, --thr, --fast, --fftOnGPU)
// and drives the CUDA library exclusively through the C ABI of include/recfourier_b200.h, which it
// loads with dlopen so that this host code builds and unit-tests without CUDA.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "metadata.h"
#include "recfourier_b200.h"

namespace rfhost {

struct ProgramError : public std::runtime_error {
    using std::runtime_error::runtime_error;
};

class ProgRecFourierB200 {
public:
    // ---- parameters (reconstruct_fourier.h:80-118)
    std::string fn_sel;                 // -i
    std::string fn_out = "rec_fourier.vol";   // -o
    std::string fn_sym = "c1";          // --sym
    std::string fn_fsc;                 // --prepare_fsc
    bool do_weights = false;            // --weight
    double padding_factor_proj = 2.0;   // --padding <proj> <vol>
    double padding_factor_vol = 2.0;
    double blob_radius = 1.9;           // --blob <radius> <order> <alpha>
    int blob_order = 0;
    double blob_alpha = 15.0;
    double maxResolution = 0.5;         // --max_resolution
    int numThreads = 1;                 // --thr <threads> <rows>: here, host image-loader threads
    int thrWidth = 1;
    int NiterWeight = 1;                // --iter
    bool useCTF = false;                // --useCTF
    bool phaseFlipped = false;          // --phaseFlipped
    double minCTF = 0.01;               // --minCTF
    double Ts = 1.0;                    // --sampling
    // GPU-program parameters
    int device = 0;                     // --device
    int bufferSize = 1024;              // --bufferSize: particles handed to the GPU per call
    bool fast = false;                  // --fast
    int verbose = 1;                    // -v
    // multi-GPU (replaces mpirun + xmipp_mpi_cuda_reconstruct_fourier, parallel_adapt_cuda/mpi_reconstruct_fourier_gpu.cpp):
    // --gpus N forks one process per GPU (devices device .. device+N-1); every process reconstructs a contiguous
    // shard of the particles into private accumulators, one NCCL reduce onto rank 0 precedes the finalisation.
    // An external launcher can instead export RFB200_RANK / RFB200_WORLD_SIZE / RFB200_ID_FILE (+ --device per rank).
    int gpus = 1;                       // --gpus <n | all>
    int rank = 0, worldSize = 1;        // set by --gpus or the environment
    std::string idFile;                 // rendezvous file for the 128-byte ncclUniqueId
    bool privateRendezvous = false;     // idFile lies in a fresh directory of our own launcher (no stale files possible)

    static std::string usage();
    // parse argv (throws ProgramError on unknown / malformed options)
    void readParams(int argc, const char* const* argv);
    void show() const;
    void setIO(const std::string& fnIn, const std::string& fnOut) { fn_sel = fnIn; fn_out = fnOut; }
    void run();
    // --gpus launcher: forks one rank per GPU (each runs run() on its shard) and waits for them
    void runRanks();
    // like XmippProgram::tryRun: 0 on success, non-zero after printing the error
    int tryRun();

    // Fill one ABI row from metadata row i (RF.cpp:362-381; data/ctf.cpp:365-419, 1172-1212).
    static void particleFromRow(const MetaData& md, size_t i, bool hasCtf, const std::string& mdDir, rfb200_particle& p);
    // path of the image of row i, relative names resolved against the metadata's directory
    static std::string imageOfRow(const MetaData& md, size_t i, const std::string& mdDir);
};

}  // namespace rfhost
