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
#include "../rf_host.hpp"

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
                 "bufferSize=%d\nfast=%d\ngpus=%d\nrank=%d\nworldSize=%d\n",
                 p.fn_sel.c_str(), p.fn_out.c_str(), p.fn_sym.c_str(), p.fn_fsc.c_str(), (int)p.do_weights, p.padding_factor_proj,
                 p.padding_factor_vol, p.blob_radius, p.blob_order, p.blob_alpha, p.maxResolution, p.numThreads, p.NiterWeight,
                 (int)p.useCTF, (int)p.phaseFlipped, p.minCTF, p.Ts, p.device, p.bufferSize, (int)p.fast, p.gpus, p.rank, p.worldSize);
        strncpy(out, buf, outLen - 1);
        out[outLen - 1] = 0;
        return 0;
    } catch (const std::exception& e) {
        return fail(e);
    }
}


// Consistency check of the gather's host-side plan for one geometry (no GPU): builds the geometry, the resolution
// cut-off, the validity table and the stick units exactly like rfb200_create, then verifies
//   out[0..2] number of stick units per class, out[3] number of edge items,
//   out[4] owned voxels inside the reach sphere that NO stick of some class covers      (must be 0)
//   out[5] owned voxels covered by MORE than one stick of a class                       (must be 0)
//   out[6] validity-table entries that disagree with a brute-force evaluation           (must be 0)
//   out[7] pixels inside the "all valid" radius that are invalid or have multiplicity != 1 off column 0 (must be 0)
//   out[8] lattice points inside the reach sphere that are neither owned nor an edge item although a pixel
//          could reach them through the original or the mirror rule                      (must be 0)
int rfh_gather_plan_check(int N, double padProj, double padVol, double maxRes, double r, long long* out) {
    try {
        using namespace rfb200;
        std::vector<int> jmax;
        int iLo, iHi, R;
        host::build_cutoff((int)(N * padProj), maxRes, jmax, iLo, iHi, R);
        Geometry g = host::make_geometry(N, padProj, padVol, maxRes, r, R);
        std::vector<int32_t> rim = host::build_rim_table(g, jmax, iLo, iHi);
        const int ext[3] = {g.tx * kTileX, g.ty * kTileY, g.tz * kTileZ};
        const int off[3] = {0, g.lo, g.lo};
        const int axes[3][3] = {{1, 2, 0}, {0, 2, 1}, {0, 1, 2}};
        const double reach2 = (double)g.reach * g.reach;
        long long missing = 0, dup = 0;
        for (int cls = 0; cls < 3; ++cls) {
            std::vector<StickUnit> units = host::build_stick_units(g, cls);
            out[cls] = (long long)units.size();
            std::vector<unsigned char> cover((size_t)ext[0] * ext[1] * ext[2], 0);
            const int A = axes[cls][0], B = axes[cls][1], D = axes[cls][2];
            for (const StickUnit& u : units)
                for (int t = 0; t < kStickL; ++t)
                    for (int b = 0; b < kStickB; ++b)
                        for (int a = 0; a < kStickA; ++a) {
                            int p[3];
                            p[A] = u.a0 + a; p[B] = u.b0 + b; p[D] = u.t0 + t;
                            if (p[0] >= ext[0] || p[1] >= ext[1] || p[2] >= ext[2]) continue;
                            unsigned char& c = cover[((size_t)p[2] * ext[1] + p[1]) * ext[0] + p[0]];
                            if (c < 255) ++c;
                        }
            for (int z = 0; z < ext[2]; ++z)
                for (int y = 0; y < ext[1]; ++y)
                    for (int x = 0; x < ext[0]; ++x) {
                        int ux = x + off[0], uy = y + off[1], uz = z + off[2];
                        if (ux > g.Z / 2 || uy > g.hi || uz > g.hi) continue;
                        if (!host::main_owns(g, ux, uy, uz)) continue;
                        if ((double)ux * ux + (double)uy * uy + (double)uz * uz > reach2) continue;
                        unsigned char c = cover[((size_t)z * ext[1] + y) * ext[0] + x];
                        if (c == 0) ++missing;
                        if (c > 1) ++dup;
                    }
        }
        out[4] = missing;
        out[5] = dup;
        // validity table against brute force
        auto validO = [&](int j, int i) { return i >= iLo && i <= iHi && j >= 0 && j <= jmax[i - iLo]; };
        long long badRim = 0, badIn = 0;
        const double rIn = g.rimIn2 > 0 ? std::sqrt((double)g.rimIn2) + g.rho : -1.0;
        for (int i = -g.Rp; i <= g.Rp; ++i) {
            int rt = rim[i + g.Rp];
            int jPos = (rt & 0x3fff) - 1, jNeg = ((rt >> 14) & 0x3fff) - 1, m0 = rt >> 28;
            for (int j = -g.Rp; j <= g.Rp; ++j) {
                int mult = j > 0 ? (int)validO(j, i) : (j < 0 ? (int)validO(-j, -i) : (int)validO(0, i) + (int)validO(0, -i));
                int tab = j > 0 ? (j <= jPos) : (j < 0 ? (-j <= jNeg) : m0);
                if (mult != tab) ++badRim;
                if (rIn > 0 && (double)i * i + (double)j * j <= rIn * rIn) {
                    if (j != 0 && mult != 1) ++badIn;
                    if (j == 0 && mult != 2) ++badIn;
                }
            }
        }
        out[6] = badRim;
        out[7] = badIn;
        // every reachable lattice point is either owned by the stick gather or an edge item
        std::vector<EdgeItem> edge = host::build_edge_items(g);
        out[3] = (long long)edge.size();
        long long lost = 0;
        {
            std::vector<long long> keys;
            keys.reserve(edge.size());
            const long long S = 4 * (long long)g.Z + 64;
            auto key = [&](int x, int y, int z) { return ((long long)(x + 2 * g.Z) * S + (y + 2 * g.Z)) * S + (z + 2 * g.Z); };
            for (const EdgeItem& e : edge) keys.push_back(key(e.ux, e.uy, e.uz));
            std::sort(keys.begin(), keys.end());
            const int m = (int)std::ceil(g.r);
            for (int ux = -(g.Z / 2) - m; ux <= g.Z / 2; ++ux)
                for (int uy = g.lo - m; uy <= g.hi + m; ++uy)
                    for (int uz = g.lo - m; uz <= g.hi + m; ++uz) {
                        if ((double)ux * ux + (double)uy * uy + (double)uz * uz > reach2) continue;
                        bool o = host::cond_orig(g, ux), mi = host::cond_mirr(g, ux);
                        if (!o && !mi) continue;
                        bool natural = ux >= 0 && ux <= g.Z / 2 && uy >= g.lo && uy <= g.hi && uz >= g.lo && uz <= g.hi;
                        if (natural && host::main_owns(g, ux, uy, uz)) continue;
                        if (host::wrapi(ux, g.Z) > g.Z / 2) continue;
                        if (!std::binary_search(keys.begin(), keys.end(), key(ux, uy, uz))) ++lost;
                    }
        }
        out[8] = lost;
        return 0;
    } catch (const std::exception& e) {
        return fail(e);
    }
}
}  // extern "C"
