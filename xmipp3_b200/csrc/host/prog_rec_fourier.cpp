// prog_rec_fourier.cpp — see prog_rec_fourier.h
#include "prog_rec_fourier.h"

#include <dirent.h>
#include <dlfcn.h>
#include <fcntl.h>
#include <libgen.h>
#include <sched.h>
#include <signal.h>
#include <sys/wait.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <mutex>
#include <thread>

#include "image_io.h"
#include "symmetries.h"

namespace rfhost {

namespace {

// ---- the CUDA library, bound at run time
struct Api {
    void* lib = nullptr;
    int (*create)(const rfb200_config*, rfb200_handle*) = nullptr;
    void (*destroy)(rfb200_handle) = nullptr;
    const char* (*last_error)(rfb200_handle) = nullptr;
    int (*get_info)(rfb200_handle, rfb200_info*) = nullptr;
    int (*insert_batch)(rfb200_handle, const float*, const rfb200_particle*, int32_t) = nullptr;
    int (*finalize)(rfb200_handle, float*) = nullptr;
    int (*halfset_push)(rfb200_handle) = nullptr;
    int (*halfset_merge)(rfb200_handle) = nullptr;
    int (*get_timings)(rfb200_handle, rfb200_timings*) = nullptr;
    int (*host_alloc)(void**, size_t) = nullptr;
    int (*host_free)(void*) = nullptr;
    int (*device_count)(int32_t*) = nullptr;
    int (*nccl_unique_id)(void*) = nullptr;
    int (*nccl_init)(rfb200_handle, const void*, int32_t, int32_t) = nullptr;
    int (*reduce_nccl)(rfb200_handle, int32_t) = nullptr;
    int (*ipc_export)(rfb200_handle, void*) = nullptr;
    int (*ipc_import)(rfb200_handle, int32_t, const void*) = nullptr;
    int (*reduce_p2p)(rfb200_handle, int32_t) = nullptr;
    int (*ipc_release)(rfb200_handle) = nullptr;
    int (*set_ranks)(rfb200_handle, int32_t, int32_t) = nullptr;
    int (*reduce_p2p_prepare)(rfb200_handle) = nullptr;
    int (*reduce_p2p_run)(rfb200_handle, int32_t) = nullptr;
    int (*sync)(rfb200_handle) = nullptr;
    int (*reset)(rfb200_handle) = nullptr;
    int (*warmup)(rfb200_handle) = nullptr;
};

// small files in the private rendezvous directory: how the forked ranks hand each other a few bytes
bool publishFile(const std::string& name, const void* data, size_t bytes) {
    const std::string tmp = name + ".tmp";
    const int wfd = open(tmp.c_str(), O_WRONLY | O_CREAT | O_EXCL | O_NOFOLLOW, 0600);   // never through a pre-existing file or symlink
    const bool wrote = wfd >= 0 && write(wfd, data, bytes) == (ssize_t)bytes;
    if (wfd >= 0) close(wfd);
    return wrote && rename(tmp.c_str(), name.c_str()) == 0;
}
bool collectFile(const std::string& name, void* data, size_t bytes, int timeoutSeconds = 120) {
    auto tStart = std::chrono::steady_clock::now();
    for (;;) {
        std::ifstream f(name, std::ios::binary);
        if (f && f.read((char*)data, bytes) && f.gcount() == (std::streamsize)bytes) return true;
        if (std::chrono::steady_clock::now() - tStart > std::chrono::seconds(timeoutSeconds)) return false;
        std::this_thread::sleep_for(std::chrono::milliseconds(2));
    }
}

std::string selfDir() {
    Dl_info info;
    if (dladdr((void*)&selfDir, &info) && info.dli_fname) {
        std::string p = info.dli_fname;
        size_t s = p.rfind('/');
        if (s != std::string::npos) return p.substr(0, s);
    }
    return ".";
}

Api loadApi() {
    Api a;
    const char* env = getenv("RFB200_LIB");
    std::vector<std::string> names;
    if (env && *env) names.push_back(env);
    names.push_back(selfDir() + "/librecfourier_b200.so");
    names.push_back("librecfourier_b200.so");
    std::string tried;
    for (auto& n : names) {
        a.lib = dlopen(n.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (a.lib) break;
        tried += "\n  " + n + ": " + (dlerror() ? dlerror() : "?");
    }
    if (!a.lib) throw ProgramError("cannot load the CUDA library librecfourier_b200.so (there is no CPU fallback):" + tried);
#define BIND(field, sym)                                                     \
    a.field = (decltype(a.field))dlsym(a.lib, sym);                          \
    if (!a.field) throw ProgramError(std::string("symbol missing in librecfourier_b200.so: ") + sym);
    BIND(create, "rfb200_create")
    BIND(destroy, "rfb200_destroy")
    BIND(last_error, "rfb200_last_error")
    BIND(get_info, "rfb200_get_info")
    BIND(insert_batch, "rfb200_insert_batch")
    BIND(finalize, "rfb200_finalize")
    BIND(halfset_push, "rfb200_halfset_push")
    BIND(halfset_merge, "rfb200_halfset_merge")
    BIND(get_timings, "rfb200_get_timings")
    BIND(host_alloc, "rfb200_host_alloc")
    BIND(host_free, "rfb200_host_free")
    BIND(device_count, "rfb200_device_count")
    BIND(nccl_unique_id, "rfb200_nccl_unique_id")
    BIND(nccl_init, "rfb200_nccl_init")
    BIND(reduce_nccl, "rfb200_reduce_nccl")
    BIND(ipc_export, "rfb200_ipc_export")
    BIND(ipc_import, "rfb200_ipc_import")
    BIND(reduce_p2p, "rfb200_reduce_p2p")
    BIND(ipc_release, "rfb200_ipc_release")
    BIND(set_ranks, "rfb200_set_ranks")
    BIND(reduce_p2p_prepare, "rfb200_reduce_p2p_prepare")
    BIND(reduce_p2p_run, "rfb200_reduce_p2p_run")
    BIND(sync, "rfb200_sync")
    BIND(reset, "rfb200_reset")
    BIND(warmup, "rfb200_warmup")
#undef BIND
    return a;
}

std::string dirOf(const std::string& path) {
    size_t s = path.rfind('/');
    return s == std::string::npos ? std::string() : path.substr(0, s);
}

bool fileExists(const std::string& p) { return access(p.c_str(), R_OK) == 0; }

// CTF columns of one row (data/ctf.cpp:365-419 and 1172-1212), with the reference's defaults
void ctfFromRow(const MetaData& md, size_t i, rfb200_particle& p) {
    p.kV = md.getValueOrDefault("ctfVoltage", i, 100);
    p.defocusU = md.getValueOrDefault("ctfDefocusU", i, 0);
    p.defocusV = md.getValueOrDefault("ctfDefocusV", i, p.defocusU);
    p.defocus_angle = md.getValueOrDefault("ctfDefocusAngle", i, 0);
    p.Cs = md.getValueOrDefault("ctfSphericalAberration", i, 0);
    p.Ca = md.getValueOrDefault("ctfChromaticAberration", i, 0);
    p.espr = md.getValueOrDefault("ctfEnergyLoss", i, 0);
    p.ispr = md.getValueOrDefault("ctfLensStability", i, 0);
    p.alpha = md.getValueOrDefault("ctfConvergenceCone", i, 0);
    p.DeltaF = md.getValueOrDefault("ctfLongitudinalDisplacement", i, 0);
    p.DeltaR = md.getValueOrDefault("ctfTransversalDisplacement", i, 0);
    p.Q0 = md.getValueOrDefault("ctfQ0", i, 0);
    p.K = md.getValueOrDefault("ctfK", i, 1);
    p.envR0 = md.getValueOrDefault("ctfEnvR0", i, 0);
    p.envR1 = md.getValueOrDefault("ctfEnvR1", i, 0);
    p.envR2 = md.getValueOrDefault("ctfEnvR2", i, 0);
    p.phase_shift = md.getValueOrDefault("ctfPhaseShift", i, 0);
    p.vpp_radius = md.getValueOrDefault("ctfVPPRadius", i, 0);
}

// Columns of the particle metadata resolved once (the per-row accessors look labels up in a map and build strings).
struct ParticleCols {
    int rot, tilt, psi, sx, sy, weight, image, ctfModel;
    int ctf[18];
    bool inlineCtf = false;
    explicit ParticleCols(const MetaData& md) {
        rot = md.column("angleRot"); tilt = md.column("angleTilt"); psi = md.column("anglePsi");
        sx = md.column("shiftX"); sy = md.column("shiftY"); weight = md.column("weight");
        image = md.column("image"); ctfModel = md.column("ctfModel");
        static const char* names[18] = {"ctfVoltage", "ctfDefocusU", "ctfDefocusV", "ctfDefocusAngle", "ctfSphericalAberration",
                                        "ctfChromaticAberration", "ctfEnergyLoss", "ctfLensStability", "ctfConvergenceCone",
                                        "ctfLongitudinalDisplacement", "ctfTransversalDisplacement", "ctfQ0", "ctfK", "ctfEnvR0",
                                        "ctfEnvR1", "ctfEnvR2", "ctfPhaseShift", "ctfVPPRadius"};
        for (int k = 0; k < 18; ++k) ctf[k] = md.column(names[k]);
        inlineCtf = ctf[1] >= 0;
    }
};
// same values as ctfFromRow / particleFromRow, by column index
void particleFromCols(const MetaData& md, const ParticleCols& c, size_t i, bool hasCtf, rfb200_particle& p) {
    memset(&p, 0, sizeof p);
    p.rot = md.cellOrDefault(i, c.rot, 0);
    p.tilt = md.cellOrDefault(i, c.tilt, 0);
    p.psi = md.cellOrDefault(i, c.psi, 0);
    p.shift_x = md.cellOrDefault(i, c.sx, 0);
    p.shift_y = md.cellOrDefault(i, c.sy, 0);
    p.weight = md.cellOrDefault(i, c.weight, 1);
    p.kV = 100;
    p.K = 1;
    if (!hasCtf || !c.inlineCtf) return;
    p.kV = md.cellOrDefault(i, c.ctf[0], 100);
    p.defocusU = md.cellOrDefault(i, c.ctf[1], 0);
    p.defocusV = md.cellOrDefault(i, c.ctf[2], p.defocusU);
    p.defocus_angle = md.cellOrDefault(i, c.ctf[3], 0);
    p.Cs = md.cellOrDefault(i, c.ctf[4], 0);
    p.Ca = md.cellOrDefault(i, c.ctf[5], 0);
    p.espr = md.cellOrDefault(i, c.ctf[6], 0);
    p.ispr = md.cellOrDefault(i, c.ctf[7], 0);
    p.alpha = md.cellOrDefault(i, c.ctf[8], 0);
    p.DeltaF = md.cellOrDefault(i, c.ctf[9], 0);
    p.DeltaR = md.cellOrDefault(i, c.ctf[10], 0);
    p.Q0 = md.cellOrDefault(i, c.ctf[11], 0);
    p.K = md.cellOrDefault(i, c.ctf[12], 1);
    p.envR0 = md.cellOrDefault(i, c.ctf[13], 0);
    p.envR1 = md.cellOrDefault(i, c.ctf[14], 0);
    p.envR2 = md.cellOrDefault(i, c.ctf[15], 0);
    p.phase_shift = md.cellOrDefault(i, c.ctf[16], 0);
    p.vpp_radius = md.cellOrDefault(i, c.ctf[17], 0);
}

struct Args {
    std::vector<std::string> v;
    size_t find(const std::string& name) const {
        for (size_t i = 0; i < v.size(); ++i)
            if (v[i] == name) return i;
        return std::string::npos;
    }
    // values following `name` up to the next option; an option is a token starting with '-' that is not a number
    std::vector<std::string> values(const std::string& name) const {
        std::vector<std::string> out;
        size_t i = find(name);
        if (i == std::string::npos) return out;
        for (size_t k = i + 1; k < v.size(); ++k) {
            const std::string& t = v[k];
            bool isOpt = t.size() > 1 && t[0] == '-' && !(isdigit((unsigned char)t[1]) || t[1] == '.');
            if (isOpt) break;
            out.push_back(t);
        }
        return out;
    }
};

double toDouble(const std::string& s, const std::string& opt) {
    char* end = nullptr;
    double v = strtod(s.c_str(), &end);
    if (end == s.c_str() || *end) throw ProgramError("option " + opt + ": '" + s + "' is not a number");
    return v;
}

}  // namespace

std::string ProgRecFourierB200::usage() {
    return
        "Generate 3D reconstructions from projections using direct Fourier interpolation with arbitrary geometry.\n"
        "Kaiser-windows are used for interpolation in Fourier space.  B200-native implementation of\n"
        "xmipp_reconstruct_fourier / xmipp_cuda_reconstruct_fourier (same options).\n"
        "   -i <md_file>                      : Metadata file with input projections\n"
        "  [-o <volume_file=\"rec_fourier.vol\">] : Filename for output volume\n"
        "  [--iter <iterations=1>]            : Number of iterations for weight correction\n"
        "  [--sym <symfile=c1>]               : Enforce symmetry in projections\n"
        "  [--padding <proj=2.0> <vol=2.0>]   : Padding used for projections and volume\n"
        "  [--max_resolution <p=0.5>]         : Max resolution (Nyquist=0.5)\n"
        "  [--weight]                         : Use weights stored in the image metadata\n"
        "  [--thr <threads=all> <rows=1>]     : Number of host threads reading images; default (or \"all\"): the cores of\n"
        "                                       the machine, at most 16, as in xmipp_cuda_reconstruct_fourier (rows is\n"
        "                                       accepted and ignored)\n"
        "  [--blob <radius=1.9> <order=0> <alpha=15>] : Blob parameters\n"
        "  [--useCTF]                         : Use CTF information if present\n"
        "  [--sampling <Ts=1>]                : sampling rate of the input images in Angstroms/pixel\n"
        "  [--phaseFlipped]                   : Give this flag if images have been already phase flipped\n"
        "  [--minCTF <ctf=0.01>]              : Minimum value of the CTF that will be inverted\n"
        "  [--device <dev=0>]                 : GPU device to use\n"
        "  [--bufferSize <size=1024>]         : Number of projections handed to the GPU per call\n"
        "  [--gpus <n=1>]                     : use n GPUs (devices dev .. dev+n-1, \"all\": every visible one): one process\n"
        "                                       per GPU on a shard of the particles, one NCCL reduce before the\n"
        "                                       normalisation (replaces mpirun xmipp_mpi_cuda_reconstruct_fourier)\n"
        "  [--fftOnGPU]                       : accepted for compatibility (the FFT always runs on the GPU)\n"
        "  [-gpusPerNode <num>]               : (xmipp_mpi_cuda_reconstruct_fourier) GPUs to use on this node: same as --gpus\n"
        "  [-threadsPerGPU <num>]             : (xmipp_mpi_cuda_reconstruct_fourier) loader threads per GPU: --thr = num x gpus\n"
        "  [--mpi_job_size <size=1000>]       : (xmipp_mpi_cuda_reconstruct_fourier) accepted and ignored: the particles are\n"
        "                                       sharded statically, one contiguous range per GPU\n"
        "  [--fast]                           : Do the blobing at the end of the computation (nearest-pixel insertion,\n"
        "                                       one final blob convolution). Gives slightly different results;\n"
        "                                       --iter and --padding <proj> are then ignored like in the reference\n"
        "  [--prepare_fsc <fscfile>]          : Filename root for FSC files (<root>_1_recons.vol, <root>_2_recons.vol)\n"
        "  [-v <verbosity=1>]\n";
}

void ProgRecFourierB200::readParams(int argc, const char* const* argv) {
    Args a;
    for (int i = 1; i < argc; ++i) a.v.push_back(argv[i]);
    static const char* known[] = {"-i", "-o", "--iter", "--sym", "--padding", "--prepare_fsc", "--max_resolution", "--weight",
                                  "--thr", "--blob", "--useCTF", "--sampling", "--phaseFlipped", "--minCTF", "--device",
                                  "--bufferSize", "--fftOnGPU", "--fast", "--gpus", "-v", "-h", "--help",
                                  "--mpi_job_size", "-gpusPerNode", "-threadsPerGPU"};
    for (auto& t : a.v) {
        bool isOpt = t.size() > 1 && t[0] == '-' && !(isdigit((unsigned char)t[1]) || t[1] == '.');
        if (!isOpt) continue;
        bool ok = false;
        for (auto k : known) ok = ok || t == k;
        if (!ok) throw ProgramError("unknown option " + t + "\n" + usage());
    }
    if (a.find("-h") != std::string::npos || a.find("--help") != std::string::npos) throw ProgramError(usage());
    auto one = [&](const std::string& opt, std::string& dst) {
        auto v = a.values(opt);
        if (a.find(opt) != std::string::npos) {
            if (v.empty()) throw ProgramError("option " + opt + " needs a value");
            dst = v[0];
        }
    };
    auto num = [&](const std::string& opt, size_t idx, double& dst) {
        auto v = a.values(opt);
        if (v.size() > idx) dst = toDouble(v[idx], opt);
    };
    one("-i", fn_sel);
    if (fn_sel.empty()) throw ProgramError("-i <md_file> is required\n" + usage());
    one("-o", fn_out);
    one("--sym", fn_sym);
    one("--prepare_fsc", fn_fsc);
    do_weights = a.find("--weight") != std::string::npos;
    num("--padding", 0, padding_factor_proj);
    num("--padding", 1, padding_factor_vol);
    num("--blob", 0, blob_radius);
    double d;
    d = blob_order; num("--blob", 1, d); blob_order = (int)d;
    num("--blob", 2, blob_alpha);
    num("--max_resolution", 0, maxResolution);
    d = std::min(16u, std::max(1u, std::thread::hardware_concurrency())); {   // reconstruct_fourier_gpu.cpp:132-145: all cores
        auto v = a.values("--thr");
        if (!v.empty()) {
            if (v[0] == "all") d = std::max(1u, std::thread::hardware_concurrency());
            else d = toDouble(v[0], "--thr");
        }
    }
    numThreads = std::max(1, (int)d);
    d = thrWidth; num("--thr", 1, d); thrWidth = (int)d;
    d = NiterWeight; num("--iter", 0, d); NiterWeight = (int)d;
    useCTF = a.find("--useCTF") != std::string::npos;
    phaseFlipped = a.find("--phaseFlipped") != std::string::npos;
    num("--minCTF", 0, minCTF);
    if (useCTF) num("--sampling", 0, Ts);
    d = device; num("--device", 0, d); device = (int)d;
    d = bufferSize; num("--bufferSize", 0, d); bufferSize = std::max(1, (int)d);
    fast = a.find("--fast") != std::string::npos;
    d = verbose; num("-v", 0, d); verbose = (int)d;
    {
        auto v = a.values("--gpus");
        if (a.find("--gpus") != std::string::npos) {
            if (v.empty()) throw ProgramError("option --gpus needs a value");
            gpus = v[0] == "all" ? -1 : (int)toDouble(v[0], "--gpus");
            if (gpus == 0 || gpus < -1) throw ProgramError("--gpus must be a positive number or \"all\"");
        }
        // flags of the MPI + CUDA program (parallel_adapt_cuda/mpi_reconstruct_fourier_gpu.cpp:52-65): one node here
        d = 0; num("-gpusPerNode", 0, d);
        if (d >= 1 && a.find("--gpus") == std::string::npos) gpus = (int)d;
        d = 0; num("-threadsPerGPU", 0, d);
        if (d >= 1 && a.find("--thr") == std::string::npos) numThreads = std::max(1, (int)d * std::max(1, gpus));
    }
    // ranks started by an external launcher (one process per GPU, like the MPI program's workers)
    if (const char* e = getenv("RFB200_WORLD_SIZE")) {
        worldSize = atoi(e);
        const char* r = getenv("RFB200_RANK");
        const char* f = getenv("RFB200_ID_FILE");
        rank = r ? atoi(r) : 0;
        idFile = f ? f : "";
        if (worldSize < 1 || rank < 0 || rank >= worldSize) throw ProgramError("bad RFB200_RANK / RFB200_WORLD_SIZE");
        if (worldSize > 1 && idFile.empty()) throw ProgramError("RFB200_WORLD_SIZE > 1 needs RFB200_ID_FILE (rendezvous file of the NCCL id)");
    }
}

void ProgRecFourierB200::show() const {
    if (verbose <= 0) return;
    std::cout << " =====================================================================\n"
              << " Direct 3D reconstruction method using Kaiser windows as interpolators\n"
              << " =====================================================================\n"
              << " Input selfile             : " << fn_sel << "\n"
              << " padding_factor_proj       : " << padding_factor_proj << "\n"
              << " padding_factor_vol        : " << padding_factor_vol << "\n"
              << " Output volume             : " << fn_out << "\n";
    if (!fn_sym.empty()) std::cout << " Symmetry file for projections : " << fn_sym << "\n";
    if (!fn_fsc.empty()) std::cout << " File root for FSC files: " << fn_fsc << "\n";
    std::cout << (do_weights ? " Use weights stored in the image headers or doc file\n" : " Do NOT use weights\n");
    if (useCTF)
        std::cout << "Using CTF information\nSampling rate: " << Ts << "\nPhase flipped: " << phaseFlipped << "\nMinimum CTF: " << minCTF << "\n";
    std::cout << "\n Interpolation Function"
              << "\n   blrad                 : " << blob_radius
              << "\n   blord                 : " << blob_order
              << "\n   blalpha               : " << blob_alpha
              << "\n max_resolution          : " << maxResolution
              << "\n GPU device              : " << device
              << "\n -----------------------------------------------------------------" << std::endl;
}

std::string ProgRecFourierB200::imageOfRow(const MetaData& md, size_t i, const std::string& mdDir) {
    std::string name;
    if (!md.getValue("image", i, name)) throw ProgramError("metadata has no 'image' column");
    size_t idx;
    std::string path, fmt;
    parseImageName(name, idx, path, fmt);
    if (!path.empty() && path[0] != '/' && !fileExists(path) && !mdDir.empty()) {
        // relative to the metadata file
        std::string cand = mdDir + "/" + path;
        if (fileExists(cand)) {
            size_t at = name.find('@');
            return (at == std::string::npos ? std::string() : name.substr(0, at + 1)) + mdDir + "/" + name.substr(at == std::string::npos ? 0 : at + 1);
        }
    }
    return name;
}

void ProgRecFourierB200::particleFromRow(const MetaData& md, size_t i, bool hasCtf, const std::string& mdDir, rfb200_particle& p) {
    memset(&p, 0, sizeof p);
    p.rot = md.getValueOrDefault("angleRot", i, 0);
    p.tilt = md.getValueOrDefault("angleTilt", i, 0);
    p.psi = md.getValueOrDefault("anglePsi", i, 0);
    p.shift_x = md.getValueOrDefault("shiftX", i, 0);
    p.shift_y = md.getValueOrDefault("shiftY", i, 0);
    p.weight = md.getValueOrDefault("weight", i, 1);
    p.kV = 100;
    p.K = 1;
    if (!hasCtf) return;
    if (md.containsLabel("ctfDefocusU")) {
        ctfFromRow(md, i, p);
    } else if (md.containsLabel("ctfModel")) {
        // indirection through a .ctfparam metadata file (data/ctf.cpp:388-395)
        std::string fn;
        md.getValue("ctfModel", i, fn);
        if (!fn.empty() && fn[0] != '/' && !fileExists(fn) && !mdDir.empty() && fileExists(mdDir + "/" + fn)) fn = mdDir + "/" + fn;
        MetaData ctf;
        ctf.read(fn);
        if (ctf.size() == 0) throw ProgramError("empty CTF model file " + fn);
        ctfFromRow(ctf, 0, p);
    }
}

// --gpus: one process per GPU.  The parent never touches CUDA (so forking is safe), forks the ranks, and waits;
// a rank that fails takes the others down (they would otherwise wait for it inside NCCL).
void ProgRecFourierB200::runRanks() {
    int n = gpus;
    if (n < 0) {    // "all": ask the CUDA library from a throw-away child
        int fd[2];
        if (pipe(fd) != 0) throw ProgramError("pipe() failed");
        std::cout.flush();
        pid_t pid = fork();
        if (pid < 0) throw ProgramError("fork() failed");
        if (pid == 0) {
            int32_t c = 0;
            try {
                Api api = loadApi();
                api.device_count(&c);
            } catch (...) {
                c = 0;
            }
            ssize_t w = write(fd[1], &c, sizeof c);
            _exit(w == (ssize_t)sizeof c ? 0 : 1);
        }
        close(fd[1]);
        int32_t c = 0;
        if (read(fd[0], &c, sizeof c) != (ssize_t)sizeof c) c = 0;
        close(fd[0]);
        int st;
        waitpid(pid, &st, 0);
        if (c < 1) throw ProgramError("no CUDA device visible (there is no CPU fallback)");
        n = c - device;
        if (n < 1) throw ProgramError("--device is beyond the last visible GPU");
    }
    if (n == 1) {
        gpus = 1;
        run();
        return;
    }
    // the rendezvous file lives in a private directory (mode 0700): no other local user can plant a symlink or a bogus
    // id under its name, and a stale file of a crashed run can never be picked up (the directory name is fresh)
    char dirTmpl[] = "/tmp/rfb200_XXXXXX";
    if (!mkdtemp(dirTmpl)) throw ProgramError("cannot create the rendezvous directory in /tmp");
    const std::string idPath = std::string(dirTmpl) + "/nccl_id";
    const char* tmpl = idPath.c_str();
    std::cout.flush();
    std::cerr.flush();
    std::vector<pid_t> pids;
    for (int r = 0; r < n; ++r) {
        pid_t pid = fork();
        if (pid < 0) {
            for (pid_t p : pids) kill(p, SIGTERM);
            throw ProgramError("fork() failed");
        }
        if (pid == 0) {
            rank = r;
            worldSize = n;
            idFile = tmpl;
            privateRendezvous = true;
            // the ranks share the host: every rank gets its slice of the loader threads and of the cores (the reference's
            // MPI program leaves this to mpirun's binding)
            {
                const int cores = (int)std::max(1u, std::thread::hardware_concurrency());
                numThreads = std::max(1, numThreads / n);
                const int c0 = (int)((long)cores * r / n), c1 = std::max(c0 + 1, (int)((long)cores * (r + 1) / n));
                cpu_set_t set;
                CPU_ZERO(&set);
                for (int c = c0; c < c1 && c < CPU_SETSIZE; ++c) CPU_SET(c, &set);
                sched_setaffinity(0, sizeof set, &set);       // best effort
            }
            device += r;
            gpus = 1;
            if (r > 0) verbose = 0;
            int code = tryRun();
            std::cout.flush();
            std::cerr.flush();
            _exit(code);
        }
        pids.push_back(pid);
    }
    int failed = 0;
    size_t left = pids.size();
    while (left > 0) {
        int st = 0;
        pid_t p = wait(&st);
        if (p < 0) break;
        bool mine = false;
        for (auto& q : pids)
            if (q == p) {
                q = -1;
                mine = true;
            }
        if (!mine) continue;
        --left;
        const int code = WIFEXITED(st) ? WEXITSTATUS(st) : 128 + (WIFSIGNALED(st) ? WTERMSIG(st) : 0);
        if (code != 0 && !failed) {
            failed = code;
            for (pid_t q : pids)
                if (q > 0) kill(q, SIGTERM);
        }
    }
    if (DIR* d = opendir(dirTmpl)) {                  // the rendezvous file and the small files of the peer-memory reduce
        while (dirent* e = readdir(d)) {
            const std::string n = e->d_name;
            if (n != "." && n != "..") unlink((std::string(dirTmpl) + "/" + n).c_str());
        }
        closedir(d);
    }
    rmdir(dirTmpl);
    if (failed) throw ProgramError("a GPU rank failed (exit code " + std::to_string(failed) + ")");
}

void ProgRecFourierB200::run() {
    if (gpus != 1 && worldSize == 1) {
        runRanks();
        return;
    }
    show();
    if (NiterWeight < 0) throw ProgramError("--iter must be >= 0");

    // ---- produceSideinfo (RF.cpp:184-287)
    MetaData SF;
    try {
        SF.read(fn_sel);
    } catch (const std::exception& e) {
        throw ProgramError(e.what());
    }
    SF.removeDisabled();
    if (SF.size() == 0) throw ProgramError("no (enabled) images in " + fn_sel);
    std::string mdPath = fn_sel.substr(fn_sel.find('@') == std::string::npos ? 0 : fn_sel.find('@') + 1);
    const std::string mdDir = dirOf(mdPath);
    ImageInfo info;
    try {
        info = readImageInfo(imageOfRow(SF, 0, mdDir));
    } catch (const std::exception& e) {
        throw ProgramError(e.what());
    }
    if (info.nx != info.ny) throw ProgramError("This algorithm only works for squared images");     // RF.cpp:202-203
    const int N = info.nx;
    std::vector<Mat3> sym;
    if (!fn_sym.empty()) {
        try {
            sym = symmetryMatrices(fn_sym);
        } catch (const std::exception& e) {
            throw ProgramError(e.what());
        }
    }
    const bool hasCtf = useCTF && (SF.containsLabel("ctfModel") || SF.containsLabel("ctfDefocusU"));     // RF.cpp:335-336

    Api api = loadApi();
    rfb200_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.abi_version = RFB200_ABI_VERSION;
    cfg.img_size = N;
    cfg.pad_proj = padding_factor_proj;
    cfg.pad_vol = padding_factor_vol;
    cfg.max_resolution = maxResolution;
    cfg.blob_radius = blob_radius;
    cfg.blob_alpha = blob_alpha;
    cfg.blob_order = blob_order;
    cfg.n_sym = (int)sym.size();
    cfg.sym_matrices = sym.empty() ? nullptr : sym[0].data();
    cfg.use_ctf = hasCtf ? 1 : 0;
    cfg.phase_flipped = phaseFlipped ? 1 : 0;
    cfg.sampling = Ts;
    cfg.min_ctf = minCTF;
    cfg.use_weights = do_weights ? 1 : 0;
    cfg.n_iter_weight = NiterWeight;
    cfg.fast = fast ? 1 : 0;       // reconstruct_fourier_gpu.cpp:87
    cfg.device = device;
    cfg.max_batch = bufferSize;
    rfb200_handle h = nullptr;
    bool useP2P = false;           // peer-memory reduce instead of ncclReduce (decided by all ranks together below)
    int rc = api.create(&cfg, &h);
    if (rc != RFB200_OK) throw ProgramError(std::string("GPU initialisation failed: ") + api.last_error(nullptr));
    // RFB200_REDUCE: unset / "auto" = peer-memory reduce when every GPU can map every other one, else NCCL;
    // "p2p" = peer memory or fail; "nccl" = the two ncclReduce calls
    const char* redEnv = getenv("RFB200_REDUCE");
    const std::string redMode = redEnv && *redEnv ? redEnv : "auto";
    if (redMode != "auto" && redMode != "p2p" && redMode != "nccl") {
        api.destroy(h);
        throw ProgramError("RFB200_REDUCE must be auto, p2p or nccl");
    }
    // With an external launcher the rendezvous file may sit in a directory that outlives the run: stale barrier files of
    // an earlier run would let a rank read its peers too early, so "auto" means NCCL there and peer memory must be asked
    // for (with a fresh RFB200_ID_FILE path per run).
    const bool wantP2P = worldSize > 1 && (redMode == "p2p" || (redMode == "auto" && privateRendezvous));
    // barrier over the forked ranks through the private rendezvous directory (one small file per rank and barrier)
    int barrierSeq = 0;
    auto fileBarrier = [&] {
        const std::string base = idFile + ".b" + std::to_string(barrierSeq++) + "_";
        char one = 1;
        if (!publishFile(base + std::to_string(rank), &one, 1)) throw ProgramError("cannot write into the rendezvous directory");
        for (int k = 0; k < worldSize; ++k)
            if (!collectFile(base + std::to_string(k), &one, 1, 7 * 24 * 3600)) throw ProgramError("timed out waiting for rank " + std::to_string(k));
    };
    if (wantP2P) {
        // Peer-memory reduce (the default when it is possible): NO NCCL at all (its communicator and connection set-up cost
        // 0.1 - 1.5 s, more than a short run).  Every rank publishes the IPC handles of its accumulators next to the
        // rendezvous file and maps everybody else's, the ranks order themselves with file barriers around the kernel
        // (rfb200_reduce_p2p_prepare / _run).  On NVSwitch systems NCCL's in-switch reduction moves the bytes a little
        // faster (about a millisecond at 8 GPUs); what this path saves is the set-up.
        if (api.set_ranks(h, worldSize, rank) != RFB200_OK) {
            std::string msg = std::string("set_ranks failed: ") + api.last_error(h);
            api.destroy(h);
            throw ProgramError(msg);
        }
        char blob[RFB200_IPC_BYTES];
        char ok = api.ipc_export(h, blob) == RFB200_OK ? 1 : 0;
        if (!publishFile(idFile + ".ipc" + std::to_string(rank), blob, sizeof blob)) ok = 0;
        for (int k = 0; k < worldSize && ok; ++k) {
            if (k == rank) continue;
            char peer[RFB200_IPC_BYTES];
            if (!collectFile(idFile + ".ipc" + std::to_string(k), peer, sizeof peer) || api.ipc_import(h, k, peer) != RFB200_OK) ok = 0;
        }
        publishFile(idFile + ".p2p" + std::to_string(rank), &ok, 1);
        useP2P = true;
        for (int k = 0; k < worldSize; ++k) {
            char theirs = 0;
            if (!collectFile(idFile + ".p2p" + std::to_string(k), &theirs, 1) || !theirs) useP2P = false;
        }
        if (!useP2P) {
            if (redMode == "p2p") {
                std::string msg = std::string("RFB200_REDUCE=p2p: the GPUs cannot map each other's memory (") + api.last_error(h) + ")";
                api.destroy(h);
                throw ProgramError(msg);
            }
            api.ipc_release(h);        // whatever was mapped: nobody frees memory a peer still has open
            fileBarrier();
        }
    }
    if (worldSize > 1 && !useP2P) {
        // rendezvous: rank 0 publishes the 128-byte ncclUniqueId in idFile (written aside, then renamed), the others
        // wait for it (the MPI program broadcasts its job ranges the same way, mpi_reconstruct_fourier_gpu.cpp:150-200)
        char id[128];
        if (rank == 0) {
            if (api.nccl_unique_id(id) != RFB200_OK) {
                api.destroy(h);
                throw ProgramError("ncclGetUniqueId failed (libnccl.so.2 not loadable?)");
            }
            const std::string tmp = idFile + ".tmp";
            // O_EXCL | O_NOFOLLOW: never write through a pre-existing file or symlink
            const int wfd = open(tmp.c_str(), O_WRONLY | O_CREAT | O_EXCL | O_NOFOLLOW, 0600);
            const bool wrote = wfd >= 0 && write(wfd, id, sizeof id) == (ssize_t)sizeof id;
            if (wfd >= 0) close(wfd);
            if (!wrote || rename(tmp.c_str(), idFile.c_str()) != 0) {
                if (wfd >= 0) unlink(tmp.c_str());
                api.destroy(h);
                throw ProgramError("cannot write the rendezvous file " + idFile);
            }
        } else {
            auto tStart = std::chrono::steady_clock::now();
            for (;;) {
                std::ifstream f(idFile, std::ios::binary);
                if (f && f.read(id, sizeof id) && f.gcount() == (std::streamsize)sizeof id) break;
                if (std::chrono::steady_clock::now() - tStart > std::chrono::seconds(120)) {
                    api.destroy(h);
                    throw ProgramError("timed out waiting for the rendezvous file " + idFile);
                }
                std::this_thread::sleep_for(std::chrono::milliseconds(5));
            }
        }
        rc = api.nccl_init(h, id, worldSize, rank);
        if (rc != RFB200_OK) {
            std::string msg = std::string("NCCL initialisation failed: ") + api.last_error(h);
            api.destroy(h);
            throw ProgramError(msg);
        }
    }

    // ---- processImages (RF.cpp:835-1013): batches of bufferSize particles in two page-locked buffers; numThreads
    // loader threads fill batch k+1 (files -> pinned memory, metadata rows -> rfb200_particle) while batch k is being
    // uploaded and inserted (insert_batch returns once the batch has been copied to the GPU)
    const size_t n = SF.size();
    const size_t B = (size_t)bufferSize;
    float* buf[2] = {nullptr, nullptr};
    std::vector<rfb200_particle> meta[2];
    for (int k = 0; k < 2; ++k) {
        void* p = nullptr;
        if (api.host_alloc(&p, B * (size_t)N * N * sizeof(float)) != RFB200_OK) {
            if (buf[0]) api.host_free(buf[0]);
            api.destroy(h);
            throw ProgramError("cannot allocate the page-locked image buffers");
        }
        buf[k] = (float*)p;
        meta[k].resize(B);
    }
    auto t0 = std::chrono::steady_clock::now();
    // The one-off set-up of the END of the run (3-D FFT plan, finalisation buffers, NCCL's connection set-up: 0.6 + 1 s)
    // is paid on a second thread while the particles are loaded and inserted; joined before the first reduce / finalize.
    int warmRc = RFB200_OK;
    std::thread warm;                           // started inside the try block below: it must be joined on every path
    double tWarmWait = 0;
    auto joinWarm = [&] {
        if (warm.joinable()) {
            const auto tw = std::chrono::steady_clock::now();
            warm.join();
            tWarmWait += std::chrono::duration<double>(std::chrono::steady_clock::now() - tw).count();
        }
        if (warmRc != RFB200_OK) throw ProgramError(std::string("warm-up failed: ") + api.last_error(h));
    };
    std::string loadError;
    std::mutex errM;
    // ---- loader (the reference's prepareBuffer threads, reconstruct_fourier_gpu.cpp:323-415, minus their CPU FFT): every
    // row is resolved ONCE to (file, image index) — the path lookups and stat() calls of imageOfRow happen per distinct
    // file, not per particle — and a loader thread then needs one pread per image, straight into the page-locked batch
    // buffer: no name parsing, no lock, no intermediate copy.  Metadata cells are read by column index.
    const ParticleCols cols(SF);
    if (cols.image < 0) throw ProgramError("metadata has no 'image' column");
    struct RowRef { uint32_t file; uint32_t index; };
    std::vector<RowRef> rowRef(n);
    std::vector<std::pair<std::string, std::string>> files;       // resolved path, format
    {
        std::map<std::string, uint32_t> known;                    // "path:fmt" as written in the metadata -> files[]
        for (size_t i = 0; i < n; ++i) {
            size_t idx;
            std::string path, fmt;
            parseImageName(SF.cell(i, cols.image), idx, path, fmt);
            const std::string key = path + ":" + fmt;
            auto it = known.find(key);
            if (it == known.end()) {
                std::string resolved = path;
                if (!path.empty() && path[0] != '/' && !fileExists(path) && !mdDir.empty() && fileExists(mdDir + "/" + path))
                    resolved = mdDir + "/" + path;                // relative to the metadata file
                files.emplace_back(resolved, fmt);
                it = known.emplace(key, (uint32_t)files.size() - 1).first;
            }
            rowRef[i] = {it->second, (uint32_t)idx};
        }
    }
    const bool ctfFromModel = hasCtf && !cols.inlineCtf;          // .ctfparam indirection: the slow accessor handles it
    auto loadBatch = [&](size_t first, size_t cnt, int slot) {
        std::atomic<size_t> next(0);
        auto worker = [&] {
            std::shared_ptr<const ImageSource> src;               // the file of the previous image: stacks hit this every time
            uint32_t srcFile = UINT32_MAX;
            for (;;) {
                // blocks of 8 consecutive images per grab: neighbouring reads stay on one thread (read-ahead friendly)
                const size_t k0 = next.fetch_add(8);
                if (k0 >= cnt) return;
                const size_t k1 = std::min(cnt, k0 + 8);
                try {
                    for (size_t k = k0; k < k1; ++k) {
                        const size_t row = first + k;
                        if (ctfFromModel) particleFromRow(SF, row, hasCtf, mdDir, meta[slot][k]);
                        else particleFromCols(SF, cols, row, hasCtf, meta[slot][k]);
                        const RowRef& rr = rowRef[row];
                        if (rr.file != srcFile) {
                            src = openImageSource(files[rr.file].first, files[rr.file].second);
                            srcFile = rr.file;
                        }
                        readImage2D(*src, rr.index, &buf[slot][k * (size_t)N * N], N, N);
                    }
                } catch (const std::exception& e) {
                    std::lock_guard<std::mutex> g(errM);
                    if (loadError.empty()) loadError = e.what();
                    return;
                }
            }
        };
        std::vector<std::thread> th;
        for (int t = 1; t < numThreads; ++t) th.emplace_back(worker);
        worker();
        for (auto& x : th) x.join();
    };
    int slot = 0;
    size_t done = 0;
    try {
        warm = std::thread([&] { warmRc = api.warmup(h); });
        // --prepare_fsc (RF.cpp:846, 991-1053): images [0, FSCIndex] and (FSCIndex, n) are reconstructed as two
        // independent half sets "<root>_1_recons.vol" / "<root>_2_recons.vol" before the full map is formed
        const bool saveFSC = !fn_fsc.empty();
        const size_t FSCIndex = (n - 1) / 2;
        std::vector<float> vol((size_t)N * N * N);
        double tInsert = 0, tReduce = 0, tFinish = 0;      // wall seconds per phase (this rank)
        struct PhaseTimer {
            double& t;
            std::chrono::steady_clock::time_point a;
            ~PhaseTimer() { t += std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count(); }
        };
        auto insertRange = [&](size_t begin, size_t end) {
            if (worldSize > 1) {    // this rank's contiguous shard of [begin, end)
                // sizes differ by at most one, earlier ranks get the extra (same rule as xmipp3_b200/sharding.py)
                const size_t len = end - begin, base = len / (size_t)worldSize, extra = len % (size_t)worldSize;
                const size_t r = (size_t)rank;
                begin = begin + r * base + std::min(r, extra);
                end = begin + base + (r < extra ? 1 : 0);
            }
            if (begin >= end) return;
            PhaseTimer acc{tInsert, std::chrono::steady_clock::now()};
            loadBatch(begin, std::min(B, end - begin), slot);
            for (size_t first = begin; first < end; first += B, slot ^= 1) {
                const size_t cnt = std::min(B, end - first);
                if (!loadError.empty()) throw ProgramError(loadError);
                // prefetch the next batch into the other buffer while this one is uploaded and inserted
                const size_t nextFirst = first + B;
                std::thread prefetch;
                if (nextFirst < end) prefetch = std::thread(loadBatch, nextFirst, std::min(B, end - nextFirst), slot ^ 1);
                rc = api.insert_batch(h, buf[slot], meta[slot].data(), (int32_t)cnt);
                if (prefetch.joinable()) prefetch.join();
                if (rc != RFB200_OK) throw ProgramError(std::string("insert_batch failed: ") + api.last_error(h));
                done += cnt;
                if (verbose > 0) std::cout << "\r " << done << " / " << n << " images inserted" << std::flush;
            }
            if (!loadError.empty()) throw ProgramError(loadError);
        };
        // every rank's partial V and W are summed onto rank 0 over NVLink; the others are done with these particles
        auto reduce = [&] {
            joinWarm();
            if (worldSize == 1) return;
            PhaseTimer acc{tReduce, std::chrono::steady_clock::now()};
            if (useP2P) {
                rc = api.reduce_p2p_prepare(h);                 // this rank's accumulators are final
                if (rc == RFB200_OK) {
                    fileBarrier();
                    rc = api.reduce_p2p_run(h, 0);              // its slice of the sums is in rank 0's memory
                }
                if (rc == RFB200_OK) fileBarrier();
            } else
                rc = api.reduce_nccl(h, 0);
            if (rc != RFB200_OK) throw ProgramError(std::string("reduce failed: ") + api.last_error(h));
            if (rank != 0 && api.reset(h) != RFB200_OK) throw ProgramError(std::string("reset failed: ") + api.last_error(h));
        };
        auto finish = [&](const std::string& name) {
            joinWarm();
            if (rank != 0) return;
            PhaseTimer acc{tFinish, std::chrono::steady_clock::now()};
            rc = api.finalize(h, vol.data());                                                   // RF.cpp:1056-1180
            if (rc != RFB200_OK) throw ProgramError(std::string("finalize failed: ") + api.last_error(h));
            writeVolume(name, vol.data(), N, N, N);                                             // RF.cpp:1179
        };
        if (saveFSC) {
            insertRange(0, FSCIndex + 1);
            reduce();
            finish(fn_fsc + "_1_recons.vol");
            if (rank == 0 && api.halfset_push(h) != RFB200_OK) throw ProgramError(std::string("halfset_push failed: ") + api.last_error(h));
            insertRange(FSCIndex + 1, n);
            reduce();
            finish(fn_fsc + "_2_recons.vol");
            if (rank == 0 && api.halfset_merge(h) != RFB200_OK) throw ProgramError(std::string("halfset_merge failed: ") + api.last_error(h));
        } else {
            insertRange(0, n);
            reduce();
        }
        if (verbose > 0) std::cout << std::endl;
        finish(fn_out);
        if (verbose > 0) {
            rfb200_timings t;
            double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (api.get_timings(h, &t) == RFB200_OK)
                std::cout << " GPU time (ms): h2d " << t.h2d_ms << ", pad " << t.preprocess_ms << ", fft " << t.fft2d_ms << ", slices "
                          << t.slice_ms << ", gather " << t.gather_ms << ", edge " << t.edge_ms << ", finalize " << t.finalize_ms << "\n";
            std::cout << " wall (s): load + insert " << tInsert << ", reduce (incl. waiting for the slowest rank and, the first time, NCCL's connection set-up) "
                      << tReduce << ", finalize + write " << tFinish << ", waiting for the set-up thread (FFT plan, NCCL connections) "
                      << tWarmWait << "\n";
            std::cout << " " << n << " images in " << secs << " s (" << n / secs << " images/s including file I/O"
                      << (worldSize > 1 ? ", " + std::to_string(worldSize) + " GPUs" : std::string()) << ")" << std::endl;
        }
    } catch (...) {
        if (warm.joinable()) warm.join();
        api.destroy(h);
        api.host_free(buf[0]);
        api.host_free(buf[1]);
        closeImageCache();
        throw;
    }
    if (useP2P) {
        // nobody frees accumulators that a peer still has mapped: unmap, tell the others, wait for them
        api.ipc_release(h);
        fileBarrier();
    }
    api.destroy(h);
    api.host_free(buf[0]);
    api.host_free(buf[1]);
    closeImageCache();
}

int ProgRecFourierB200::tryRun() {
    try {
        run();
    } catch (const std::exception& e) {
        std::cerr << "XMIPP_ERROR: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}

}  // namespace rfhost
