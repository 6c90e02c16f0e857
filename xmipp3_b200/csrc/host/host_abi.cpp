// host_abi.cpp — small extern "C" hooks over the host-side helpers so that the CPU test-suite (ctypes)
// can exercise the metadata reader, the image readers/writers, the symmetry lists and the CLI parser
// without a GPU.  Not part of the drop-in boundary (that is include/recfourier_b200.h).
#include <cstring>
#include <string>
#include <vector>

#include "image_io.h"
#include "metadata.h"
#include "prog_rec_fourier.h"
#include "symmetries.h"

using namespace rfhost;

namespace {
thread_local std::string g_err;
int fail(const std::exception& e) {
    g_err = e.what();
    return -1;
}
}  // namespace

extern "C" {

const char* rfh_last_error() { return g_err.c_str(); }

// number of non-identity symmetry matrices of `name`; out (may be NULL) receives up to max*9 doubles
int rfh_symmetry_matrices(const char* name, double* out, int max) {
    try {
        std::vector<Mat3> m = symmetryMatrices(name);
        if (out)
            for (int i = 0; i < (int)m.size() && i < max; ++i) memcpy(out + 9 * i, m[i].data(), 9 * sizeof(double));
        return (int)m.size();
    } catch (const std::exception& e) {
        return fail(e);
    }
}

// Read a metadata file the way the program does (removeDisabled included) and convert every row:
// particles (n rows of rfb200_particle) and image names (n strings of at most nameLen bytes).
// Returns the number of rows, or -1.  Pass particles = NULL to only count.
int rfh_read_particles(const char* mdFile, int useCtf, rfb200_particle* particles, char* names, int nameLen, int max, int* hasCtf) {
    try {
        MetaData md;
        md.read(mdFile);
        md.removeDisabled();
        std::string spec = mdFile;
        std::string path = spec.substr(spec.find('@') == std::string::npos ? 0 : spec.find('@') + 1);
        size_t s = path.rfind('/');
        std::string dir = s == std::string::npos ? std::string() : path.substr(0, s);
        bool ctf = useCtf && (md.containsLabel("ctfModel") || md.containsLabel("ctfDefocusU"));
        if (hasCtf) *hasCtf = ctf ? 1 : 0;
        if (particles)
            for (size_t i = 0; i < md.size() && (int)i < max; ++i) {
                ProgRecFourierB200::particleFromRow(md, i, ctf, dir, particles[i]);
                if (names) {
                    std::string n = ProgRecFourierB200::imageOfRow(md, i, dir);
                    strncpy(names + (size_t)i * nameLen, n.c_str(), nameLen - 1);
                    names[(size_t)i * nameLen + nameLen - 1] = 0;
                }
            }
        return (int)md.size();
    } catch (const std::exception& e) {
        return fail(e);
    }
}

int rfh_image_info(const char* spec, int* nx, int* ny, int* nz, long* nImages) {
    try {
        ImageInfo i = readImageInfo(spec);
        *nx = i.nx; *ny = i.ny; *nz = i.nz; *nImages = (long)i.nImages;
        return 0;
    } catch (const std::exception& e) {
        return fail(e);
    }
}

int rfh_read_image(const char* spec, float* out, int nx, int ny) {
    try {
        readImage2D(spec, out, nx, ny);
        return 0;
    } catch (const std::exception& e) {
        return fail(e);
    }
}

int rfh_write_volume(const char* spec, const float* data, int nx, int ny, int nz) {
    try {
        writeVolume(spec, data, nx, ny, nz);
        return 0;
    } catch (const std::exception& e) {
        return fail(e);
    }
}

int rfh_write_stack(const char* spec, const float* data, int nx, int ny, long n) {
    try {
        writeStack(spec, data, nx, ny, (size_t)n);
        closeImageCache();
        return 0;
    } catch (const std::exception& e) {
        return fail(e);
    }
}

void rfh_close_image_cache() { closeImageCache(); }

// Parse a command line the way the CLI does; writes the parsed fields as "key=value\n" lines.
int rfh_parse_cli(int argc, const char* const* argv, char* out, int outLen) {
    try {
        ProgRecFourierB200 p;
        p.readParams(argc, argv);
        char buf[2048];
        snprintf(buf, sizeof buf,
                 "fn_sel=%s\nfn_out=%s\nfn_sym=%s\nfn_fsc=%s\ndo_weights=%d\npad_proj=%g\npad_vol=%g\nblob_radius=%g\nblob_order=%d\n"
                 "blob_alpha=%g\nmax_resolution=%g\nthreads=%d\niter=%d\nuseCTF=%d\nphaseFlipped=%d\nminCTF=%g\nsampling=%g\ndevice=%d\n"
                 "bufferSize=%d\nfast=%d\n",
                 p.fn_sel.c_str(), p.fn_out.c_str(), p.fn_sym.c_str(), p.fn_fsc.c_str(), (int)p.do_weights, p.padding_factor_proj,
                 p.padding_factor_vol, p.blob_radius, p.blob_order, p.blob_alpha, p.maxResolution, p.numThreads, p.NiterWeight,
                 (int)p.useCTF, (int)p.phaseFlipped, p.minCTF, p.Ts, p.device, p.bufferSize, (int)p.fast);
        strncpy(out, buf, outLen - 1);
        out[outLen - 1] = 0;
        return 0;
    } catch (const std::exception& e) {
        return fail(e);
    }
}

}  // extern "C"
